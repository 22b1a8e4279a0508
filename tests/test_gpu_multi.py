"""Multi-GPU (NCCL) test of the camera-sharded aggregation; skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _scene_and_cams():
    import geograypher_b200 as gg

    g = dict(np.load(os.path.join(GOLDEN, "golden_scene.npz")))
    a = dict(np.load(os.path.join(GOLDEN, "golden_aggregate.npz")))
    f, cx, cy, W, H = g["intrinsics"]
    cams = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(f"/golden/{i:04d}.png", T, f, cx, cy, int(W), int(H))
                 for i, T in enumerate(g["c2ws"])])
    C = a["avg2"].shape[1]
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(list(a["soft"]), num_classes=C))
    return gg, g, a, seg


def test_distributed_api_without_process_group_equals_plain_api():
    from geograypher_b200 import distributed as ggd

    gg, g, a, seg = _scene_and_cams()
    mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), compat_negative_index=True)
    avg, info = ggd.aggregate_projected_images_distributed(mesh, seg, return_argmax=True)
    np.testing.assert_array_equal(avg, a["avg2"])
    np.testing.assert_array_equal(info["projection_counts"], a["counts2"])
    np.testing.assert_array_equal(info["argmax"], a["argmax2"][:, 0])


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    from geograypher_b200 import distributed as ggd

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        gg, g, a, seg = _scene_and_cams()
        mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), compat_negative_index=True, device=rank,
                                             log_level="WARNING")
        avg, info = ggd.aggregate_projected_images_distributed(mesh, seg, return_argmax=True)
        np.testing.assert_array_equal(info["projection_counts"], a["counts2"])
        np.testing.assert_allclose(avg, a["avg2"], rtol=1e-12, atol=0, equal_nan=True)
        np.testing.assert_array_equal(info["argmax"], a["argmax2"][:, 0])
        # result wanted on one rank only: the others skip the device-to-host copy
        avg0, info0 = ggd.aggregate_projected_images_distributed(mesh, seg, dst_rank=0)
        if rank == 0:
            np.testing.assert_allclose(avg0, a["avg2"], rtol=1e-12, atol=0, equal_nan=True)
            np.testing.assert_array_equal(info0["projection_counts"], a["counts2"])
        else:
            assert avg0 is None and info0 == {}
        # sharded epilogue: reduce-scatter, every rank finishes its slice of the faces and copies it into the node's
        # shared host block; rank 0 hands out views of that block (twice: the block is cached and reused)
        for rep in range(2):
            avg1, info1 = ggd.aggregate_projected_images_distributed(mesh, seg, dst_rank=0, shared_host=True,
                                                                     return_argmax=True)
            if rank == 0:
                np.testing.assert_allclose(avg1, a["avg2"], rtol=1e-12, atol=0, equal_nan=True)
                np.testing.assert_allclose(info1["summed_projections"], a["summed2"], rtol=1e-12, atol=0, equal_nan=True)
                np.testing.assert_array_equal(info1["projection_counts"], a["counts2"])
                np.testing.assert_array_equal(info1["argmax"], a["argmax2"][:, 0])
            else:
                assert avg1 is None and info1 == {}
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_gpus_match_reference():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}
