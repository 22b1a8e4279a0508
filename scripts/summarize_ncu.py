#!/usr/bin/env python
"""Turn ncu outputs (brought back in gpurun_out/) into the small text summaries kept under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/rNN_launches.txt "title"
  python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/rNN_kernel_ncu_full.txt "title"
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    # atomics: global reductions (RED) at L2 and shared-memory atomics
    "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_shared_atom.sum",
    "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum.per_second", "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "lts__t_sector_hit_rate.pct",
]


def launches(src, dst, title):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:70s} n={len(v):4d} total_us={sum(v):10.1f} avg_us={sum(v)/len(v):9.1f} share={100*sum(v)/tot:5.1f}%\n")


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {title}\n")
        for D in rows[2:]:
            f.write("----\n")
            for i, h in enumerate(H):
                stall = h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
                if stall and float(D[i] or 0) < 0.1:
                    continue
                if h in KEEP or h == "Kernel Name" or stall:
                    f.write(f"{h:80s} {U[i]:16s} {D[i][:110]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
