"""Edge cases of the path: cameras that see nothing, single-face meshes, ragged raster sizes, channel counts on both
sides of the fused dense limit, class-index images with ignore values, and the largest tile lists."""
import numpy as np
import pytest

import geograypher_b200 as gg
from geograypher_b200 import synthetic as syn
from oracle import oracle as ora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


@pytest.fixture(scope="module")
def lib():
    from geograypher_b200 import _lib

    return _lib


def _gg(lib, cam):
    out = lib.GGCamera()
    for k in range(12):
        out.m[k] = cam.m[k]
    out.f, out.px, out.py, out.W, out.H, out.znear = cam.f, cam.px, cam.py, cam.W, cam.H, cam.znear
    return out


def _ctx(torch, lib, v32, faces):
    ctx = lib.Context(0)
    ctx.set_mesh(torch.from_numpy(np.ascontiguousarray(v32, dtype=np.float32)).cuda(),
                 torch.from_numpy(np.ascontiguousarray(faces, dtype=np.int32)).cuda())
    return ctx


def test_camera_that_sees_nothing_and_single_face(torch, lib):
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
    faces = np.array([[0, 1, 2]], dtype=np.int32)
    look_away = np.eye(4)          # +Z forward from z = 5: the triangle at z = 0 is behind
    look_away[2, 3] = 5.0
    look_at = np.diag([1.0, -1.0, -1.0, 1.0])
    look_at[:3, 3] = [0.3, 0.3, 5.0]
    cams = [ora.make_camera(T, 50.0, 0, 0, 37, 29) for T in (look_away, look_at)]
    ctx = _ctx(torch, lib, verts, faces)
    p2f = ctx.rasterize([_gg(lib, c) for c in cams]).cpu().numpy()
    assert (p2f[0] == -1).all()
    np.testing.assert_array_equal(p2f[1], ora.rasterize(verts.astype(np.float32), faces, cams[1]))
    assert (p2f[1] == 0).sum() > 10
    # aggregation over both views: the empty view contributes nothing, also with the -1 -> last-face quirk
    preds = [torch.full((29, 37, 3), 0.25, device="cuda"), torch.full((29, 37, 3), 0.5, device="cuda")]
    for flags, want_count in ((0, 1), (lib.FLAG_COMPAT_NEGATIVE_INDEX, 2)):
        d_sum = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
        d_count = torch.zeros((1,), dtype=torch.int32, device="cuda")
        ctx.project_aggregate([_gg(lib, c) for c in cams], preds, lib.PRED_F32, 3, lib.MODE_LAST_PIXEL, flags, d_sum, d_count)
        ref = ora.aggregate(p2f.astype(np.int64), [p.cpu().numpy() for p in preds], 1, compat_negative_index=bool(flags))
        assert int(d_count.item()) == want_count == int(ref[1][0])
        np.testing.assert_array_equal(d_sum.cpu().numpy(), np.nan_to_num(ref[2]))


@pytest.mark.parametrize("W,H", [(1, 1), (31, 7), (33, 9), (64, 8), (65, 17), (1023, 5)])
def test_ragged_raster_sizes(torch, lib, W, H):
    verts, faces = syn.terrain_mesh(12, 1.0, seed=3, crowns=True)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    T = np.diag([1.0, -1.0, -1.0, 1.0])
    T[:3, 3] = [6.0, 6.0, 25.0]
    cam = ora.make_camera(T, 0.9 * max(W, H), 0.5, -0.25, W, H, origin=origin)
    ctx = _ctx(torch, lib, v32, faces)
    got = ctx.rasterize([_gg(lib, cam)]).cpu().numpy()[0]
    ref, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
    assert got.shape == (H, W)
    assert not ((got != ref) & (margin > 1e-5)).any()


@pytest.mark.parametrize("C", [1, 2, 7, 32, 33])
def test_dense_mode_channel_counts(torch, lib, C):
    """Fused dense epilogue for C <= 32 (compile-time and runtime channel counts), unfused fallback above."""
    v32, faces, cams = _scene()
    H, W = cams[0].H, cams[0].W
    ctx = _ctx(torch, lib, v32, faces)
    gg_c = [_gg(lib, c) for c in cams]
    p2f = ctx.rasterize(gg_c).cpu().numpy()
    rng = np.random.default_rng(C)
    host = [rng.random((H, W, C)).astype(np.float32) for _ in cams]
    F = len(faces)
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    ctx.project_aggregate(gg_c, [torch.from_numpy(h).cuda() for h in host], lib.PRED_F32, C, lib.MODE_PIXEL_SUM, 0, d_sum, d_count)
    ref = np.zeros((F, C))
    cnt = np.zeros(F, dtype=np.int64)
    for k in range(len(cams)):
        ids = p2f[k].ravel()
        keep = ids >= 0
        np.add.at(ref, ids[keep], host[k].reshape(-1, C)[keep].astype(np.float64))
        np.add.at(cnt, ids[keep], 1)
    np.testing.assert_array_equal(d_count.cpu().numpy(), cnt)
    np.testing.assert_allclose(d_sum.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


def _scene():
    verts, faces, c2ws, cfg = syn.make_survey("tiny")
    origin = 0.5 * (verts.min(0) + verts.max(0))
    W, H = cfg.image_size
    return (verts - origin).astype(np.float32), faces, [ora.make_camera(T, cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for T in c2ws[:3]]


def test_long_tile_lists(torch, lib):
    """Faces much smaller than a pixel: hundreds of faces per 32x8 tile (lists longer than one shared-memory chunk),
    for the rasters, the fused last-pixel aggregation and the fused dense mode."""
    verts, faces = syn.terrain_mesh(160, 0.05, seed=9, crowns=False)
    verts[:, 2] *= 0.02
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    T = np.diag([1.0, -1.0, -1.0, 1.0])
    T[:3, 3] = [4.0, 4.0, 20.0]
    W, H, C = 96, 40, 4
    cam = ora.make_camera(T, 200.0, 0, 0, W, H, origin=origin)   # 8 m of mesh on 80 px: ~2 faces per pixel row
    ctx = _ctx(torch, lib, v32, faces)
    got = ctx.rasterize([_gg(lib, cam)])
    stats = ctx.last_batch_stats(1)
    assert stats[0, 2] / ((W // 32) * (H // 8)) > 64  # long lists indeed
    ref, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
    g = got.cpu().numpy()[0]
    assert not ((g != ref) & (margin > 1e-5)).any()
    F = len(faces)
    soft = syn.softmax_predictions(0, H, W, C, grid=(5, 7))
    d_pred = torch.from_numpy(soft).cuda()
    # last pixel, fused vs unfused
    a_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda"); a_cnt = torch.zeros((F,), dtype=torch.int32, device="cuda")
    b_sum = torch.zeros_like(a_sum); b_cnt = torch.zeros_like(a_cnt)
    ctx.project_aggregate([_gg(lib, cam)], [d_pred], lib.PRED_F32, C, lib.MODE_LAST_PIXEL, 0, a_sum, a_cnt)
    ctx.aggregate(got[0], d_pred, lib.PRED_F32, C, lib.MODE_LAST_PIXEL, 0, b_sum, b_cnt)
    torch.cuda.synchronize()
    assert torch.equal(a_sum, b_sum) and torch.equal(a_cnt, b_cnt)
    # dense, fused vs numpy
    d_sum = torch.zeros_like(a_sum); d_cnt = torch.zeros_like(a_cnt)
    ctx.project_aggregate([_gg(lib, cam)], [d_pred], lib.PRED_F32, C, lib.MODE_PIXEL_SUM, 0, d_sum, d_cnt)
    ids = g.ravel(); keep = ids >= 0
    want = np.zeros((F, C)); np.add.at(want, ids[keep], soft.reshape(-1, C)[keep].astype(np.float64))
    np.testing.assert_allclose(d_sum.cpu().numpy(), want, rtol=1e-5, atol=1e-6)


def test_api_single_camera_and_index_images(torch):
    """API level: one camera (keeps per-channel NaNs like the reference), class-index segmentor with ignore values."""
    verts, faces, c2ws, cfg = syn.make_survey("tiny")
    W, H = cfg.image_size
    intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws, intrinsic_params_per_sensor_type=intr)
    idx = [syn.class_index_image(i, H, W, cfg.n_classes, block=8, ignore_frac=0.1) for i in range(len(cams))]
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(idx, num_classes=cfg.n_classes, one_hot=True))
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), compat_negative_index=True, log_level="WARNING")
    p2f = mesh.pix2face(cams, apply_distortion=False)
    avg, info = mesh.aggregate_projected_images(seg)
    ref = ora.aggregate(p2f, [ora.inds_to_one_hot(i, cfg.n_classes) for i in idx], len(faces))
    np.testing.assert_array_equal(avg, ref[0])
    one = mesh.aggregate_projected_images(seg.get_subset_cameras([2]))
    ref1 = ora.aggregate(p2f[2:3], [ora.inds_to_one_hot(idx[2], cfg.n_classes)], len(faces))
    np.testing.assert_array_equal(one[0], ref1[0])
    np.testing.assert_array_equal(one[1]["projection_counts"], ref1[1])


@pytest.mark.parametrize("znear", [1e-3, 0.75, 3.0])
def test_near_plane_clipping(torch, lib, znear):
    """Contract C5: faces crossing the plane z_cam = znear are clipped, not dropped.  A camera standing ON the
    terrain, looking along it: the ground under and beside the camera crosses the near plane."""
    verts, faces = syn.terrain_mesh(40, 1.0, seed=4, crowns=True)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    ground = float(verts[(np.abs(verts[:, 0] - 20.3) < 1) & (np.abs(verts[:, 1] - 6.2) < 1), 2].max())
    eye = np.array([20.3, 6.2, ground + 2.0])
    # camera looking forward (+Y) and 45 degrees down: the columns of c2w are the camera axes in world coordinates
    a = np.sqrt(0.5)
    c2w = np.eye(4)
    c2w[:3, 0] = [1, 0, 0]
    c2w[:3, 1] = [0, -a, -a]
    c2w[:3, 2] = [0, a, -a]
    c2w[:3, 3] = eye
    W, H = 320, 200
    cam = ora.make_camera(c2w, 160.0, 0, 0, W, H, origin=origin, znear=znear)
    ctx = _ctx(torch, lib, v32, faces)
    got = ctx.rasterize([_gg(lib, cam)]).cpu().numpy()[0]
    ref, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
    bad = (got != ref) & (margin > 1e-5)
    assert not bad.any(), int(bad.sum())
    assert (ref >= 0).mean() > 0.3
    # clipped faces are really there: faces with a vertex behind the near plane own pixels
    X, Y, invz, valid = ora.project(v32, cam)
    crossing = (~valid[faces]).any(axis=1) & valid[faces].any(axis=1)
    assert crossing.sum() > 0
    if znear >= 0.75:
        assert np.isin(ref[ref >= 0], np.nonzero(crossing)[0]).sum() > 50
    # a clipped face can become two triangles: it must still be aggregated (and counted) once, in every mode
    C, F = 3, len(faces)
    pred = np.random.default_rng(2).random((H, W, C)).astype(np.float32)
    for mode in (lib.MODE_LAST_PIXEL, lib.MODE_PIXEL_SUM):
        d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
        d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
        ctx.project_aggregate([_gg(lib, cam)], [torch.from_numpy(pred).cuda()], lib.PRED_F32, C, mode, 0, d_sum, d_count)
        ids = got.ravel()
        keep = ids >= 0
        if mode == lib.MODE_LAST_PIXEL:
            want = ora.aggregate(got[None].astype(np.int64), [pred], F, compat_negative_index=False)
            np.testing.assert_array_equal(d_count.cpu().numpy(), want[1].astype(np.int64))
            np.testing.assert_array_equal(d_sum.cpu().numpy(), np.nan_to_num(want[2]))
            # the same through the two-step path for pageable host images (GPU lists pixels, host gathers rows)
            pairs, counts = ctx.project_winners([_gg(lib, cam)], 0)
            ctx.sync()
            m = int(counts[0].item())
            fp = pairs[0, :m].cpu().numpy()
            assert len(np.unique(fp[:, 0])) == m  # every visible face once, also the ones split in two
            rows = torch.from_numpy(pred.reshape(-1, C)[fp[:, 1]]).cuda()
            s2 = torch.zeros_like(d_sum)
            c2 = torch.zeros_like(d_count)
            ctx.accumulate_rows(pairs[0], m, rows, lib.PRED_F32, C, mode, 0, s2, c2)
            ctx.sync()
            assert torch.equal(s2, d_sum) and torch.equal(c2, d_count)
        else:
            cnt = np.bincount(ids[keep], minlength=F)
            ref_sum = np.zeros((F, C))
            np.add.at(ref_sum, ids[keep], pred.reshape(-1, C)[keep].astype(np.float64))
            np.testing.assert_array_equal(d_count.cpu().numpy(), cnt)
            np.testing.assert_allclose(d_sum.cpu().numpy(), ref_sum, rtol=1e-5, atol=1e-5)


def test_face_order_does_not_matter(torch, lib):
    """gg_set_mesh re-orders the faces along a Z-order curve for the block culling; IDs travel with the faces, so a
    mesh with randomly shuffled faces gives the oracle's answer for THAT face numbering, and culls about as well."""
    verts, faces, cam_T, cfg = syn.make_survey("tiny", max_cameras=3)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    perm = np.random.default_rng(5).permutation(len(faces))
    shuffled = np.ascontiguousarray(faces[perm])
    W, H = cfg.image_size
    cams = [ora.make_camera(T, cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for T in cam_T]
    blocks = {}
    for name, fc in (("ordered", faces), ("shuffled", shuffled)):
        ctx = _ctx(torch, lib, v32, fc)
        p2f = ctx.rasterize([_gg(lib, c) for c in cams]).cpu().numpy()
        ctx.sync()
        blocks[name] = ctx.last_batch_stats(len(cams))[:, 0].sum()
        for i, c in enumerate(cams):
            ref = ora.rasterize(v32, fc, c)
            np.testing.assert_array_equal(p2f[i], ref)
    assert blocks["shuffled"] <= 1.25 * blocks["ordered"] + 8


@pytest.mark.parametrize("kind,C", [("f32", 3), ("f32", 10), ("f32", 16), ("f64", 4), ("f64", 8), ("u8", 10), ("u8", 5)])
def test_dense_mode_vector_path(torch, lib, kind, C):
    """GG_MODE_PIXEL_SUM on the two-channels-per-load path (even channel counts, aligned images): image rows are
    aligned here, the image has partial tiles on the right (general loop) and at the bottom (fewer rows), a block
    of null scores sends some tiles through the filtering loop, and every element type is covered."""
    verts, faces, c2ws, cfg = syn.make_survey("tiny")
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    W, H = 208, 117
    cams = [ora.make_camera(T, cfg.f * 1.3, cfg.cx, cfg.cy, W, H, origin=origin) for T in c2ws[:3]]
    ctx = _ctx(torch, lib, v32, faces)
    gg_c = [_gg(lib, c) for c in cams]
    p2f = ctx.rasterize(gg_c).cpu().numpy()
    rng = np.random.default_rng(C)
    np_dtype, pred_kind = {"f32": (np.float32, lib.PRED_F32), "f64": (np.float64, lib.PRED_F64), "u8": (np.uint8, lib.PRED_U8)}[kind]
    if kind == "u8":
        host = [rng.integers(0, 256, size=(H, W, C), dtype=np.uint8) for _ in cams]
    else:
        host = [rng.random((H, W, C)).astype(np_dtype) for _ in cams]
        host[1][5:40, 3:70, :] = np.nan  # nulls contribute nothing
    F = len(faces)
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    ctx.project_aggregate(gg_c, [torch.from_numpy(h).cuda() for h in host], pred_kind, C, lib.MODE_PIXEL_SUM, 0, d_sum, d_count)
    ref = np.zeros((F, C))
    cnt = np.zeros(F, dtype=np.int64)
    for k in range(len(cams)):
        ids = p2f[k].ravel()
        keep = ids >= 0
        np.add.at(ref, ids[keep], np.nan_to_num(host[k].reshape(-1, C)[keep].astype(np.float32).astype(np.float64)))
        np.add.at(cnt, ids[keep], 1)
    np.testing.assert_array_equal(d_count.cpu().numpy(), cnt)
    np.testing.assert_allclose(d_sum.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    assert (cnt > 0).sum() > 100


@pytest.mark.parametrize("staging", ["1", "0"])
def test_prediction_images_in_pinned_host_memory(torch, lib, staging, monkeypatch):
    """Last-pixel / vote aggregation straight from page-locked host images (rows fetched over PCIe, staged on the
    device or read in place) gives exactly what device-resident images give."""
    monkeypatch.setenv("GG_STAGE_HOST_ROWS", staging)
    v32, faces, cams = _scene()
    H, W, C, F = cams[0].H, cams[0].W, 5, len(faces)
    gg_c = [_gg(lib, c) for c in cams]
    ctx = _ctx(torch, lib, v32, faces)
    rng = np.random.default_rng(11)
    soft = [rng.random((H, W, C)).astype(np.float32) for _ in cams]
    soft[1][10:30, 20:90, 2] = np.nan
    index = [syn.class_index_image(i, H, W, C, block=8, ignore_frac=0.1) for i in range(len(cams))]
    votes = [rng.integers(0, C, size=(H, W)).astype(np.float32) for _ in cams]
    votes[0][::7, ::3] = np.nan
    cases = [(soft, lib.PRED_F32, lib.MODE_LAST_PIXEL, 0), (soft, lib.PRED_F32, lib.MODE_LAST_PIXEL, lib.FLAG_COMPAT_NEGATIVE_INDEX),
             (index, lib.PRED_INDEX_U8, lib.MODE_LAST_PIXEL, 0), (votes, lib.PRED_F32, lib.MODE_VOTE, 0)]
    for images, kind, mode, flags in cases:
        out = []
        for where in ("device", "host"):
            if where == "device":
                preds = [torch.from_numpy(a).cuda() for a in images]
            else:
                preds = [torch.from_numpy(a).pin_memory() for a in images]
                assert all(p.is_pinned() for p in preds)
            d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
            d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
            ctx.project_aggregate(gg_c, preds, kind, C, mode, flags, d_sum, d_count)
            ctx.sync()
            out.append((d_sum.cpu().numpy(), d_count.cpu().numpy()))
        np.testing.assert_array_equal(out[0][0], out[1][0])
        np.testing.assert_array_equal(out[0][1], out[1][1])
        assert out[0][1].sum() > 100


def test_dense_mode_from_pinned_host_images(torch, lib):
    """GG_MODE_PIXEL_SUM also accepts page-locked host images (every score then crosses PCIe: slow, but the same sums)."""
    v32, faces, cams = _scene()
    H, W, C, F = cams[0].H, cams[0].W, 4, len(faces)
    gg_c = [_gg(lib, c) for c in cams]
    ctx = _ctx(torch, lib, v32, faces)
    rng = np.random.default_rng(4)
    images = [rng.integers(0, 32, size=(H, W, C)).astype(np.float32) for _ in cams]
    out = []
    for preds in ([torch.from_numpy(a).cuda() for a in images], [torch.from_numpy(a).pin_memory() for a in images]):
        d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
        d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
        ctx.project_aggregate(gg_c, preds, lib.PRED_F32, C, lib.MODE_PIXEL_SUM, 0, d_sum, d_count)
        ctx.sync()
        out.append((d_sum.cpu().numpy(), d_count.cpu().numpy()))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])


def test_pageable_pointers_are_refused_by_the_fused_entry_point(torch, lib):
    """The C entry point dereferences the prediction pointers on the GPU: a pageable host array must be an error, not
    a crash (the Python API routes such arrays through gg_project_winners / gg_accumulate_rows)."""
    v32, faces, cams = _scene()
    H, W, C, F = cams[0].H, cams[0].W, 3, len(faces)
    ctx = _ctx(torch, lib, v32, faces)
    pageable = torch.zeros((H, W, C), dtype=torch.float32)  # ordinary host memory
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    with pytest.raises(lib.GeograypherB200Error, match="page-locked"):
        ctx.project_aggregate([_gg(lib, cams[0])], [pageable], lib.PRED_F32, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_count)
    ctx.sync()
    assert int(d_count.sum()) == 0


def test_guard_band_clipping_matches_oracle_and_ray_caster(torch, lib):
    """Contract C6 on the GPU: a 400 m face that reaches far behind the camera plane (its near-plane cut projects
    millions of pixels away) plus the triangle soup at the default near plane -- the CUDA setup kernel clips against
    the guard band with the oracle's float32 arithmetic (bit-identical rasters) and the result is what the
    independent float64 ray caster sees."""
    from test_oracle_raycast import EPS_EDGE
    from test_oracle_reference_pins import downward_view, plane_mesh

    ground, gfaces = plane_mesh()
    ground = ground * 10.0
    big = np.array([[-150.0, -200.0, 7.9], [150.0, -200.0, 7.9], [0.0, 200.0, 8.3]])
    verts = np.concatenate([ground, big])
    faces = np.concatenate([gfaces, [[len(ground), len(ground) + 1, len(ground) + 2]]]).astype(np.int32)
    v32 = verts.astype(np.float32)
    ctx = _ctx(torch, lib, v32, faces)
    for zc, tilt in ((9.0, 70.0), (8.2, 60.0), (8.05, 50.0)):
        T = downward_view(4, 100, 200)
        T[:3, 3] = [1.0, -2.0, zc]
        a = np.deg2rad(tilt)
        T[:3, :3] = T[:3, :3] @ np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        cam = ora.make_camera(T, 150.0, 0.0, 0.0, 240, 180)
        gpu = ctx.rasterize([_gg(lib, cam)]).cpu().numpy()[0]
        ref, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
        assert not ((gpu != ref) & (margin > 1e-5)).any()
        rc, edge_safe, rc_margin = ora.raycast(v32, faces, cam, eps_edge=EPS_EDGE)
        safe = edge_safe & (rc_margin > 1e-4)
        assert safe.mean() > 0.9 and not (safe & (gpu != rc)).any()
