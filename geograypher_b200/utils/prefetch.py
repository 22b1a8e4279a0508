"""Ordered read-ahead of per-view inputs (prediction images decoded from disk) on a small thread pool.

The reference reads one prediction per view inside its aggregation loop (``cameras.get_image_by_index``,
geograypher/meshes/meshes.py:1959-1975; ``LookUpSegmentor`` decodes a PNG per call, predictors/derived_segmentors.py:
38-51).  At 0.1 ms of GPU work per 20-Mpx view a serial 100-ms PNG decode would bound the whole aggregation, so the
views ahead of the one being aggregated are fetched concurrently: the decoders (Pillow, ``np.load``) release the GIL.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor


class OrderedPrefetcher:
    """``fetch(k)`` for k = 0 .. n-1, asked for in increasing order, computed up to ``depth`` items ahead on ``threads``
    workers.  Results come back in order and exceptions of ``fn`` surface at the ``fetch`` of their item.  The number
    of items held ahead is bounded by ``max_bytes`` once the size of an item is known (``size_of(item)``), so that a
    deep queue of 800-MB score images cannot exhaust the host."""

    def __init__(self, fn, n: int, threads: int = 8, depth: int | None = None, max_bytes: int = 2 << 30, size_of=None):
        self.fn, self.n = fn, int(n)
        self.threads = max(1, int(threads))
        self.depth = max(1, int(depth if depth is not None else 2 * self.threads))
        self.max_bytes, self.size_of = int(max_bytes), size_of
        self._pool = ThreadPoolExecutor(self.threads, thread_name_prefix="gg-prefetch")
        self._futures = {}
        self._next = 0  # first index not yet submitted
        self._sized = False

    def _submit_up_to(self, hi):
        hi = min(hi, self.n)
        while self._next < hi:
            self._futures[self._next] = self._pool.submit(self.fn, self._next)
            self._next += 1

    def fetch(self, k: int):
        if not 0 <= k < self.n:
            raise IndexError(k)
        for old in [i for i in self._futures if i < k]:  # skipped items: let go of them
            self._futures.pop(old).cancel()
        if k >= self._next:
            self._next = k
        # the first item decides how many may be held ahead; until then only a pool's worth is in flight
        self._submit_up_to(k + (self.depth if self._sized else min(self.depth, self.threads)))
        item = self._futures.pop(k).result()
        if not self._sized:
            self._sized = True
            if self.size_of is not None:
                nbytes = max(1, int(self.size_of(item)))
                self.depth = max(1, min(self.depth, self.max_bytes // nbytes))
        self._submit_up_to(k + 1 + self.depth)
        return item

    __call__ = fetch

    def close(self):
        for f in self._futures.values():
            f.cancel()
        self._futures.clear()
        self._pool.shutdown(wait=True, cancel_futures=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def default_threads() -> int:
    """Decoding scales with the cores until the memory system saturates; one core stays with the aggregation loop."""
    return max(1, min(16, (os.cpu_count() or 2) - 1))
