"""world_size-2 gloo tests of the multi-GPU host logic on CPU: camera sharding and the accumulator all-reduce.
The per-rank partial accumulators are produced by the oracle here (the CUDA path needs a GPU); what is tested is
that shard + all-reduce + epilogue reproduces the single-process result of the reference's own code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from geograypher_b200 import distributed as ggd
from oracle import oracle as ora

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_shard_range_partitions_everything():
    for n in (0, 1, 3, 10, 500, 2000):
        for world in (1, 2, 3, 4, 8):
            parts = [list(ggd.shard_range(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= -(-n // world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = np.load(os.path.join(GOLDEN, "golden_scene.npz"))
        agg = np.load(os.path.join(GOLDEN, "golden_aggregate.npz"))
        p2f = scene["pix2face"].astype(np.int64)
        F, C = agg["avg2"].shape
        mine = list(ggd.shard_range(len(p2f), rank, world))
        summed = np.zeros((F, C))
        counts = np.zeros(F, dtype=np.int32)
        for k in mine:  # per-rank partial accumulators, NaN -> 0 like the device kernels
            proj = ora.project_image(p2f[k], agg["soft"][k], F)
            summed += np.nan_to_num(proj, nan=0.0)
            counts += np.any(np.isfinite(proj), axis=1).astype(np.int32)
        r_sum, r_count = torch.from_numpy(summed.copy()), torch.from_numpy(counts.copy())  # partials, for the reduce below
        d_sum, d_count = torch.from_numpy(summed), torch.from_numpy(counts)
        ggd.allreduce_accumulators(d_sum, d_count)
        avg, cnt, tot = ggd.finalize_host(d_sum.numpy(), d_count.numpy())
        np.testing.assert_array_equal(cnt, agg["counts2"])
        np.testing.assert_allclose(tot, agg["summed2"], rtol=1e-12, atol=0, equal_nan=True)
        np.testing.assert_allclose(avg, agg["avg2"], rtol=1e-12, atol=0, equal_nan=True)
        amax = ora.find_argmax_nonzero_value(avg)
        np.testing.assert_array_equal(amax, agg["argmax2"][:, 0])
        # sharded epilogue: a reduce-scatter hands every rank the totals of ITS slice of the faces (even and ragged
        # splits); gloo has no reduce-scatter, so the same packing runs through an all-reduce here
        for n_faces in (F, F - 1, 1):
            s_sum, s_cnt = ggd.reduce_scatter_accumulators(torch.from_numpy(r_sum.numpy()[:n_faces].copy()),
                                                           torch.from_numpy(r_count.numpy()[:n_faces].copy()))
            lo, hi = ggd.face_slice(n_faces, rank, world)
            assert s_sum.shape == (-(-n_faces // world), C) and s_cnt.dtype == torch.int32
            np.testing.assert_array_equal(s_cnt.numpy()[: hi - lo], d_count.numpy()[lo:hi])
            np.testing.assert_array_equal(s_sum.numpy()[: hi - lo], d_sum.numpy()[lo:hi])
            assert not s_sum.numpy()[hi - lo :].any() and not s_cnt.numpy()[hi - lo :].any()
        # result wanted on one rank only: a reduce instead of an all-reduce; the destination holds the same totals
        ggd.allreduce_accumulators(r_sum, r_count, dst_rank=1)
        if rank == 1:
            np.testing.assert_array_equal(r_count.numpy(), d_count.numpy())
            np.testing.assert_array_equal(r_sum.numpy(), d_sum.numpy())
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allreduce_of_sharded_partials_matches_reference():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
