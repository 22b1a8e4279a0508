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


def finalize_host(summed, counts):
    """NumPy form of the epilogue of aggregate_projected_images (reference meshes.py:2069-2082), for accumulators
    that were reduced on the host (gloo)."""
    summed = np.array(summed, dtype=float, copy=True)
    counts = np.asarray(counts, dtype=float)
    summed[counts == 0] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        return summed / counts[:, None], counts, summed


def aggregate_projected_images_distributed(mesh, cameras, aggregate_img_scale: float = 1, return_argmax: bool = False,
                                           group=None, dst_rank=None, timings=None, **kwargs):
    """``TexturedPhotogrammetryMesh.aggregate_projected_images`` over all ranks of ``group``.

    Every rank passes the SAME full camera set; internally it only processes its own block of cameras.  By default
    every rank gets the full result back; with ``dst_rank`` only that rank copies it to the host (the others return
    ``(None, {})``), which is what a job that writes the result once wants.  ``mesh.device`` must be this rank's GPU.
    ``timings`` (a dict) receives this rank's seconds per phase (each phase ends with a device synchronisation).
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
