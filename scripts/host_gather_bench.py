"""Host-side row gather of the pageable-image route (development aid): gg_gather_rows_host vs np.take, by thread count."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from geograypher_b200 import _lib
H, W, C = 3648, 5472, 10
imgs = [np.random.rand(H, W, C).astype(np.float32) for _ in range(8)]
rng = np.random.default_rng(0)
n = 34000
pix = [np.sort(rng.integers(0, H * W, n)).astype(np.int32) for _ in range(10)]
pairs = np.zeros((10 * n, 2), np.int32); pairs[:, 1] = np.concatenate(pix)
offs = np.arange(11) * n
out = np.empty((10 * n, C), np.float32)
ims = [imgs[v % 8] for v in range(10)]
for T in (1, 2, 4, 8, 16, 32, 0):
    best = 1e9
    for rep in range(5):
        t = time.perf_counter(); _lib.gather_rows_host(ims, pairs, offs, out, T); best = min(best, time.perf_counter() - t)
    print(f"native {T:2d} threads {best * 1e3:7.2f} ms {10 * n / best / 1e6:7.1f} M rows/s")
def job(a):
    v, lo, hi = a
    np.take(ims[v].reshape(-1, C), pix[v][lo:hi], axis=0, out=out[v * n + lo:v * n + hi], mode="clip")
jobs = [(v, lo, min(lo + 8500, n)) for v in range(10) for lo in range(0, n, 8500)]
pool = ThreadPoolExecutor(16)
best = 1e9
for rep in range(5):
    t = time.perf_counter(); list(pool.map(job, jobs)); best = min(best, time.perf_counter() - t)
print(f"np.take, 16 Python threads {best * 1e3:7.2f} ms {10 * n / best / 1e6:7.1f} M rows/s")
