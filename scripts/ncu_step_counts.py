#!/usr/bin/env python
"""Warp instructions of EVERY kernel of one batch of the fused last-pixel aggregation (reset, cull, setup, reserve,
fill, rasterizer, winner reset, resolve) from an ncu CSV log of
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --csv --log-file X \
      python bench.py --steps 1 --warmup 0 --skip pixel_sum,c3,c4,c5,e2e,cpu
-> key "<config>:last_pixel:<B>:step" of profiles/inst_counts.json (the sum over the kernels of the per-launch mean
over the survey's last 50 launches of each) and a readable table.  bench.py divides that sum by the live time of one
batch for the step-wide issue fraction (`roofline.step_issue`): what all of the step's kernels together make of the
SMs' issue slots, where `roofline.frac` charges the rasterizer alone with the whole launch window it shares.
Usage: python scripts/ncu_step_counts.py gpurun_out/r5_step_counts.csv profiles/r02_step_counts.txt [key]"""
import collections
import csv
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LAST = 50
STEP_KERNELS = ["k_reset_views", "k_cull_blocks", "k_setup_faces", "k_reserve_tiles", "k_fill_bins",
                "k_raster_tiles<4, float, 0>", "k_reset_winners", "k_resolve_batch<float>"]


def main(src, dst, key="c2:last_pixel:10:step"):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
    per = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in data:
        for k in STEP_KERNELS:
            if k in r[ki]:
                per[k][r[mi]].append(float(r[vi].replace(",", "")))
    lines = ["# ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none: python bench.py --steps 1 "
             "--warmup 0 --skip pixel_sum,c3,c4,c5,e2e,cpu",
             f"# per launch (= one batch of 10 views); mean over the last {LAST} launches of each kernel (one pass over the survey)"]
    total_inst, total_ns = 0.0, 0.0
    for k in STEP_KERNELS:
        if k not in per:
            lines.append(f"{k:32s} (not launched)")
            continue
        inst = statistics.mean(per[k]["smsp__inst_executed.sum"][-LAST:])
        ns = statistics.mean(per[k]["gpu__time_duration.sum"][-LAST:])
        total_inst += inst
        total_ns += ns
        lines.append(f"{k:32s} launches={len(per[k]['smsp__inst_executed.sum']):4d} warp_inst={inst:14.0f} time_ns={ns:10.0f}")
    lines.append(f"{'sum over the step kernels':32s}               warp_inst={total_inst:14.0f} time_ns={total_ns:10.0f} "
                 f"(serialised, cold cache) -> {total_inst / total_ns:6.1f} Gwarp-inst/s under ncu")
    p = ROOT / "profiles" / "inst_counts.json"
    old = json.loads(p.read_text())
    old[key] = int(total_inst)
    old["_comment_step"] = ("key ...:step = smsp__inst_executed.sum summed over ALL kernels of one batch (reset, cull, setup, "
                            "reserve, fill, rasterizer, winner reset, resolve), each the mean over its last 50 launches; "
                            "scripts/ncu_step_counts.py")
    p.write_text(json.dumps(old, indent=1) + "\n")
    Path(dst).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:])
