#!/usr/bin/env python
"""Device-to-host delivery of a fresh 1 GB result, by kind of destination: torch pageable (.cpu()), a NumPy array
(numpy asks for transparent huge pages on large allocations), a NumPy array pre-faulted by a few threads, fresh and
re-used page-locked blocks.  Usage: python scripts/d2h_dest_bench.py > profiles/r02_d2h_dest.txt"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

N = 2**27  # float64 -> 1 GiB
src = torch.rand(N, dtype=torch.float64, device="cuda")
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())


def clock(label, fn, reps=3):
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{label:56s} rep {r}: {1e3 * dt:7.1f} ms  {N * 8 / dt / 1e9:6.2f} GB/s")
        del out


def numpy_dest():
    dst = np.empty(N, dtype=np.float64)
    torch.from_numpy(dst).copy_(src)
    return dst


pool = ThreadPoolExecutor(8)


def prefaulted(threads=8):
    dst = np.empty(N, dtype=np.float64)
    step = N // threads
    list(pool.map(lambda i: dst[i * step:(i + 1) * step:512].fill(0), range(threads)))  # one write per 4 KiB page
    torch.from_numpy(dst).copy_(src)
    return dst


def pinned():
    dst = torch.empty(N, dtype=torch.float64, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    return dst


clock("torch pageable: src.cpu()", lambda: src.cpu())
clock("numpy destination: from_numpy(np.empty).copy_(src)", numpy_dest)
clock("numpy destination pre-faulted by 8 threads", prefaulted)
clock("page-locked block (first rep allocates, later reps re-use)", pinned)
