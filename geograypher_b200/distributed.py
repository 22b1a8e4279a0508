"""Camera-sharded multi-GPU aggregation.

Views are independent; the only coupling between them is the per-face accumulator pair
``(sum[F, C] float64, count[F] int32)``, a commutative sum over views (reference meshes.py:2057-2067; the
reference's chunked variant already merges partial sums this way, derived_meshes.py:292-302).  So: one process per
GPU, the mesh replicated, the cameras split into contiguous blocks, and ONE all-reduce of the accumulators (counts
packed behind the sums) at the end, followed by the mean / argmax epilogue.  ``render_flat`` has no coupling at all (replicas only).

The functions take an initialised ``torch.distributed`` process group (NCCL on GPUs; the host-side logic is
exercised with gloo on CPU tensors in tests/test_distributed_gloo.py).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous block of ``ceil(n / world)`` items for ``rank`` (neighbouring views share faces, which keeps a
    rank's working set of the mesh hot in L2)."""
    per = -(-n_items // world_size)
    lo = min(n_items, rank * per)
    return range(lo, min(n_items, lo + per))


def shard_cameras(cameras, rank: int, world_size: int):
    """The sub-set of a PhotogrammetryCameraSet (or of a SegmentorPhotogrammetryCameraSet) that ``rank`` owns."""
    inds = list(shard_range(len(cameras), rank, world_size))
    try:
        return cameras.get_subset_cameras(inds, deep=False)  # the shard is only read: no need to copy the cameras
    except TypeError:  # a camera-set class with the reference's signature
        return cameras.get_subset_cameras(inds)


def allreduce_accumulators(d_sum, d_count, group=None, dst_rank=None):
    """In-place sum of the per-face accumulators over all ranks with ONE collective: the int32 counts ride behind the
    float64 sums in a single float64 buffer (exact: counts are far below 2^53).  With ``dst_rank`` only that rank
    needs the total and the collective is a reduce (half the traffic of an all-reduce); the other ranks' buffers are
    then left with partial sums.  Counts (and one-hot / vote sums) are
    integers and therefore identical for any number of ranks; float64 sums of real-valued scores differ from the
    single-GPU result only by the association order of at most ``world_size`` partial sums (~1e-16 relative)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return d_sum, d_count
    n_sum = d_sum.numel()
    pack = torch.empty((n_sum + d_count.numel(),), dtype=torch.float64, device=d_sum.device)
    pack[:n_sum].copy_(d_sum.reshape(-1))
    pack[n_sum:].copy_(d_count.reshape(-1))
    if dst_rank is None:
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(pack, dst=dist.get_global_rank(group, dst_rank) if group is not None else dst_rank,
                    op=dist.ReduceOp.SUM, group=group)
    d_sum.copy_(pack[:n_sum].view(d_sum.shape))
    d_count.copy_(pack[n_sum:].view(d_count.shape))
    return d_sum, d_count


def face_slice(n_faces: int, rank: int, world_size: int):
    """Rows [lo, hi) of the per-face result that ``rank`` finishes in the sharded epilogue: ``ceil(F / world)`` each."""
    per = -(-n_faces // world_size)
    lo = min(n_faces, rank * per)
    return lo, min(n_faces, lo + per)


def reduce_scatter_accumulators(d_sum, d_count, group=None):
    """ONE reduce-scatter of the per-face accumulators: rank r receives the totals of ITS slice of the faces
    (``face_slice``) -- half the traffic of the all-reduce, and the epilogue and the device-to-host copy that follow
    then cost 1 / world_size per rank.  The int32 counts ride behind the float64 sums of the same slice in one float64
    buffer (exact below 2^53).  Returns ``(sum_slice (Fs, C) float64, count_slice (Fs,) int32)`` with
    ``Fs = ceil(F / world)`` rows, zero beyond the rank's ``hi - lo`` faces.  Backends without a reduce-scatter
    (gloo, the CPU tests) run the same packing through an all-reduce."""
    import torch
    import torch.distributed as dist

    F, C = d_sum.shape
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    Fs = -(-F // world)
    if world == 1:
        return d_sum, d_count
    chunk = Fs * (C + 1)
    pack = torch.zeros((world, chunk), dtype=torch.float64, device=d_sum.device)
    full = F // Fs  # ranks whose slice is complete
    if full:
        pack[:full, : Fs * C].copy_(d_sum[: full * Fs].reshape(full, Fs * C))
        pack[:full, Fs * C :].copy_(d_count[: full * Fs].reshape(full, Fs))
    if full < world and F > full * Fs:  # the last, shorter slice
        rest = F - full * Fs
        pack[full, : rest * C].copy_(d_sum[full * Fs :].reshape(-1))
        pack[full, Fs * C : Fs * C + rest].copy_(d_count[full * Fs :])
    mine = torch.empty((chunk,), dtype=torch.float64, device=d_sum.device)
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(mine, pack.reshape(-1), op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
        mine.copy_(pack[rank])
    return mine[: Fs * C].view(Fs, C), mine[Fs * C :].to(torch.int32)


class SharedHostResult:
    """The result arrays of a distributed aggregation -- averages (F, C), sums (F, C), counts (F,), argmax (F,), all
    float64 -- in ONE POSIX shared-memory block that every rank of the node maps and page-locks, so that each rank
    copies the slice of the faces it finished over ITS OWN PCIe link: the device-to-host copy of a 2 M-face result
    (336 MB, 5.8 ms from one GPU) takes 1 / world_size of that.  The block is created by ``dst_rank`` (which unlinks
    it at exit) and cached per (F, C): arrays returned by one call are overwritten by the next call of that shape."""

    _cache = {}

    def __init__(self, F, C, group, dst_rank):
        import atexit
        from multiprocessing import resource_tracker, shared_memory

        import torch
        import torch.distributed as dist

        self.F, self.C = int(F), int(C)
        self._buf = None
        rank = dist.get_rank(group)
        n = self.F * self.C
        self.n_bytes = (2 * n + 2 * self.F) * 8
        name = [None]
        if rank == dst_rank:
            import os

            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize < 2 * self.n_bytes:
                name[0] = ""
            else:
                self.shm = shared_memory.SharedMemory(create=True, size=self.n_bytes)
                name[0] = self.shm.name
        src = dist.get_global_rank(group, dst_rank) if group is not None else dst_rank
        dist.broadcast_object_list(name, src=src, group=group)
        if not name[0]:
            raise OSError("not enough room in /dev/shm for the shared result block")
        if rank != dst_rank:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:  # Python < 3.13 registers attached segments too and would unlink them when THIS process exits
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        buf = np.ndarray((self.n_bytes // 8,), dtype=np.float64, buffer=self.shm.buf)
        self.avg = buf[:n].reshape(self.F, self.C)
        self.sums = buf[n : 2 * n].reshape(self.F, self.C)
        self.counts = buf[2 * n : 2 * n + self.F]
        self.argmax = buf[2 * n + self.F :]
        if rank == dst_rank:
            buf[:] = 0.0  # touch the pages before they are page-locked (first touch by each rank for its own slice was
                          # measured slower at N = 2: torchrun does not bind a process to its GPU's NUMA node)
        dist.barrier(group=group)
        rc = torch.cuda.cudart().cudaHostRegister(buf.ctypes.data, self.n_bytes, 0)
        ok = torch.tensor([1.0 if int(rc) == 0 else 0.0], device=torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # all ranks succeed or all ranks give up
        if ok.item() != 1.0:
            if int(rc) == 0:
                torch.cuda.cudart().cudaHostUnregister(buf.ctypes.data)
            del buf
            self.avg = self.sums = self.counts = self.argmax = None
            self._release(rank == dst_rank)
            raise OSError("cudaHostRegister of the shared result block failed on some rank")
        self._buf = buf
        atexit.register(self._release, rank == dst_rank)

    def _release(self, unlink):
        """At interpreter exit: un-pin, drop our views, unmap (left mapped if the caller still holds arrays), and -- on
        the rank that created it -- remove the segment."""
        try:
            import torch

            if self._buf is not None:
                torch.cuda.cudart().cudaHostUnregister(self._buf.ctypes.data)
        except Exception:
            pass
        self.avg = self.sums = self.counts = self.argmax = self._buf = None
        try:
            self.shm.close()
        except BufferError:  # the caller's arrays still point into the block: the OS unmaps it with the process
            self.shm.close = lambda: None
            self.shm._mmap = None
        except Exception:
            pass
        if unlink:
            try:
                self.shm.unlink()
            except Exception:
                pass

    @classmethod
    def get(cls, F, C, group, dst_rank):
        key = (int(F), int(C), id(group), int(dst_rank))
        if key not in cls._cache:
            cls._cache[key] = cls(F, C, group, dst_rank)
        return cls._cache[key]


def finalize_sharded(ctx, d_sum, d_count, group=None, dst_rank=0, want_argmax=False):
    """Sharded epilogue of a distributed aggregation: reduce-scatter, mean / NaN marking / argmax on this rank's slice
    of the faces (gg_finalize), and this rank's copy of its slice into the node's shared result block.  Returns the
    ``SharedHostResult`` (complete on every rank once the call returns; ``dst_rank`` hands its arrays to the caller)."""
    import torch
    import torch.distributed as dist

    F, C = d_sum.shape
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    res = SharedHostResult.get(F, C, group, dst_rank)
    s_sum, s_cnt = reduce_scatter_accumulators(d_sum, d_count, group)
    s_sum = s_sum.contiguous()
    avg, argmax = ctx.finalize(s_sum, s_cnt, want_avg=True, want_argmax=want_argmax)
    lo, hi = face_slice(F, rank, world)
    k = hi - lo
    if k > 0:
        torch.from_numpy(res.avg[lo:hi]).copy_(avg[:k], non_blocking=True)
        torch.from_numpy(res.sums[lo:hi]).copy_(s_sum[:k], non_blocking=True)
        torch.from_numpy(res.counts[lo:hi]).copy_(s_cnt[:k].double(), non_blocking=True)
        if want_argmax:
            torch.from_numpy(res.argmax[lo:hi]).copy_(argmax[:k], non_blocking=True)
    ctx.sync()
    dist.barrier(group=group)  # every slice has landed
    return res


def finalize_host(summed, counts):
    """NumPy form of the epilogue of aggregate_projected_images (reference meshes.py:2069-2082), for accumulators
    that were reduced on the host (gloo)."""
    summed = np.array(summed, dtype=float, copy=True)
    counts = np.asarray(counts, dtype=float)
    summed[counts == 0] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        return summed / counts[:, None], counts, summed


def aggregate_projected_images_distributed(mesh, cameras, aggregate_img_scale: float = 1, return_argmax: bool = False,
                                           group=None, dst_rank=None, timings=None, shared_host: bool = False,
                                           **kwargs):
    """``TexturedPhotogrammetryMesh.aggregate_projected_images`` over all ranks of ``group``.

    Every rank passes the SAME full camera set; internally it only processes its own block of cameras.  By default
    every rank gets the full result back; with ``dst_rank`` only that rank copies it to the host (the others return
    ``(None, {})``), which is what a job that writes the result once wants.  ``mesh.device`` must be this rank's GPU.
    ``timings`` (a dict) receives this rank's seconds per phase (each phase ends with a device synchronisation).
    ``shared_host`` (needs ``dst_rank``; all ranks on one node): reduce-scatter instead of all-reduce, every rank
    finishes its slice of the faces and copies it over its own PCIe link into a shared host block
    (``finalize_sharded``) -- the arrays ``dst_rank`` gets back are views of that block, valid until the next call.
    """
    import time

    import torch

    def mark(name, t0):
        if timings is not None:
            torch.cuda.synchronize()
            timings[name] = time.perf_counter() - t0
        return time.perf_counter()

    import torch.distributed as dist

    from geograypher_b200 import _lib

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    t0 = time.perf_counter()
    mine = shard_cameras(cameras, rank, world)
    t0 = mark("shard", t0)
    pix2face_kwargs = {k: v for k, v in kwargs.items() if k != "check_null_image"}
    n_total = len(cameras)
    if len(mine) > 0:
        d_sum, d_count, C = mesh._accumulate_views(mine, aggregate_img_scale, _lib.MODE_LAST_PIXEL,
                                                   pix2face_kwargs=pix2face_kwargs,
                                                   single_view_total=(n_total == 1))
    else:
        import torch

        C = cameras.n_image_channels()
        dev = torch.device("cuda", mesh.device)
        d_sum = torch.zeros((mesh.faces.shape[0], C), dtype=torch.float64, device=dev)
        d_count = torch.zeros((mesh.faces.shape[0],), dtype=torch.int32, device=dev)
    mesh._get_context().drain()  # accumulators are written on the library's internal streams
    t0 = mark("accumulate", t0)
    if shared_host and dst_rank is not None and world > 1:
        try:
            res = finalize_sharded(mesh._get_context(), d_sum, d_count, group, dst_rank, want_argmax=return_argmax)
        except OSError as e:  # raised on EVERY rank (no room in /dev/shm, or it cannot be page-locked): reduce instead
            mesh.logger.warning(f"shared host result unavailable ({e}); falling back to a reduce to rank {dst_rank}")
            res = None
        if res is not None:
            mark("reduce_scatter+finalize+to_host", t0)
            if rank != dst_rank:
                return None, {}
            info = {"projection_counts": res.counts, "summed_projections": res.sums}
            if return_argmax:
                info["argmax"] = res.argmax
            return res.avg, info
    allreduce_accumulators(d_sum, d_count, group, dst_rank=dst_rank)
    t0 = mark("allreduce", t0)
    if dst_rank is not None and rank != dst_rank:
        return None, {}
    avg, argmax = mesh._get_context().finalize(d_sum, d_count, want_avg=True, want_argmax=return_argmax)
    h_avg, h_sum, h_count = mesh._to_host(avg, d_sum, d_count.double())
    info = {"projection_counts": h_count, "summed_projections": h_sum}
    if return_argmax:
        info["argmax"] = argmax.cpu().numpy()
    mark("to_host", t0)
    return h_avg, info
