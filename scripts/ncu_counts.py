#!/usr/bin/env python
"""Per-launch instruction counts and DRAM traffic of the k_raster_tiles variants from an ncu CSV log
(ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
 --clock-control none --csv --log-file X python bench.py ...) -> profiles/inst_counts.json, profiles/traffic.json and a
readable summary.  Usage: python scripts/ncu_counts.py gpurun_out/counts.csv profiles/r02_raster_counts.txt "<command>" """
import collections
import csv
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LAST = 50
KEYS = {"k_raster_tiles<4, float, 0>": "c2:last_pixel:10", "k_raster_tiles<2, float, 10>": "c2:pixel_sum:10",
        "k_raster_tiles<3, unsigned char, 0>": "c4:render_flat:10"}


def main(src, dst, command):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
    per = collections.defaultdict(lambda: collections.defaultdict(list))  # kernel -> metric -> values (launch order)
    for r in data:
        name = r[ki]
        for k in KEYS:
            if k in name:
                per[k][r[mi]].append(float(r[vi].replace(",", "")))
    inst, traffic, lines = {}, {}, [f"# {command}", f"# per launch of 10 views; mean over the last {LAST} launches of each kernel = one pass over the "
                                    "500-view survey, the same launches bench.py times (warm-up launches come first and are dropped)"]
    for k, key in KEYS.items():
        if k not in per:
            continue
        m = per[k]
        n = len(m["smsp__inst_executed.sum"])
        sl = slice(-LAST, None) if n > LAST else slice(None)  # the launches of the timed pass (warm-up launches come first)
        i = statistics.mean(m["smsp__inst_executed.sum"][sl])
        rd, wr = statistics.mean(m["dram__bytes_read.sum"][sl]), statistics.mean(m["dram__bytes_write.sum"][sl])
        t = statistics.mean(m["gpu__time_duration.sum"][sl])
        # ncu prints bytes in the unit of the column; values here are already scaled by --csv to base units when
        # --print-units base is used
        inst[key], traffic[key] = int(i), int(rd + wr)
        lines.append(f"{k:42s} launches={n:3d} warp_inst={i:14.0f} dram_read_B={rd:14.0f} dram_write_B={wr:14.0f} "
                     f"time_ns={t:12.0f} -> {i / t:7.2f} Gwarp-inst/s, {(rd + wr) / t:7.1f} GB/s (serialised, cold cache)")
    note = ("smsp__inst_executed.sum per launch of k_raster_tiles (10 views), mean over the survey's 50 launches in one ncu pass of: "
            + command + ". Key = config:mode:views_per_launch.")
    old_inst = json.loads((ROOT / "profiles" / "inst_counts.json").read_text())  # keeps the ...:step keys (ncu_step_counts.py)
    old_inst.update({"_comment": note, **inst})
    (ROOT / "profiles" / "inst_counts.json").write_text(json.dumps(old_inst, indent=1) + "\n")
    old = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    old["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of k_raster_tiles, mean over the survey's 50 launches in one "
                       "ncu pass of: " + command + ". Key = config:mode:views_per_launch.")
    old.update(traffic)
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(old, indent=1) + "\n")
    Path(dst).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:4])
