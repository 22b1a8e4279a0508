"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors that
the reference's own code produced.  Run on the B200 box:  pytest -m gpu."""
import numpy as np
import pytest

from geograypher_b200 import synthetic as syn
from oracle import oracle as ora

pytestmark = pytest.mark.gpu

EPS_DEPTH = 1e-5  # contract: IDs are bit-exact on every pixel whose relative depth margin exceeds this


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def lib():
    from geograypher_b200 import _lib

    return _lib


def _to_gg(lib, cam):
    out = lib.GGCamera()
    for k in range(12):
        out.m[k] = cam.m[k]
    out.f, out.px, out.py, out.W, out.H, out.znear = cam.f, cam.px, cam.py, cam.W, cam.H, cam.znear
    return out


def _scene(name, max_cameras=None):
    verts, faces, c2ws, cfg = syn.make_survey(name, max_cameras)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    W, H = cfg.image_size
    cams = [ora.make_camera(T, cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for T in c2ws]
    return v32, faces, cams, cfg


def _context(torch, lib, v32, faces):
    ctx = lib.Context(0)
    ctx.set_mesh(torch.from_numpy(v32).cuda(), torch.from_numpy(faces.astype(np.int32)).cuda())
    return ctx


def _assert_ids_match(gpu, ref, margin, label):
    diff = gpu != ref
    unsafe = margin <= EPS_DEPTH
    bad = diff & ~unsafe
    assert not bad.any(), (
        f"{label}: {bad.sum()} pixels differ outside the depth-margin mask "
        f"(first at {np.argwhere(bad)[0]}, gpu={gpu[bad][0]}, oracle={ref[bad][0]})"
    )
    return diff.sum(), unsafe.sum()


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_projection_bit_exact(torch, lib, name):
    v32, faces, cams, _ = _scene(name, 4)
    ctx = _context(torch, lib, v32, faces)
    X, Y, invz, valid = ctx.project([_to_gg(lib, c) for c in cams])
    for k, cam in enumerate(cams):
        oX, oY, oinvz, ovalid = ora.project(v32, cam)
        np.testing.assert_array_equal(valid[k].cpu().numpy().astype(bool), ovalid)
        np.testing.assert_array_equal(X[k].cpu().numpy(), oX)
        np.testing.assert_array_equal(Y[k].cpu().numpy(), oY)
        np.testing.assert_array_equal(invz[k].cpu().numpy().view(np.uint32), oinvz.view(np.uint32))


@pytest.mark.parametrize("name,ncam", [("tiny", 4), ("c1", 10)])
def test_pix2face_matches_oracle(torch, lib, name, ncam):
    v32, faces, cams, _ = _scene(name, ncam)
    ctx = _context(torch, lib, v32, faces)
    p2f, depth = ctx.rasterize([_to_gg(lib, c) for c in cams], want_depth=True)
    p2f = p2f.cpu().numpy()
    depth = depth.cpu().numpy()
    n_diff = n_unsafe = 0
    for k, cam in enumerate(cams):
        ref, ref_w, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
        d, u = _assert_ids_match(p2f[k], ref, margin, f"{name} view {k}")
        n_diff += d
        n_unsafe += u
        hit = ref >= 0
        np.testing.assert_allclose(depth[k][hit & (p2f[k] == ref)], ref_w[hit & (p2f[k] == ref)], rtol=2e-5)
        assert (p2f[k] >= 0).sum() > 0
    # the mask must stay a vanishing fraction of the image
    assert n_unsafe <= 1e-3 * p2f.size, (n_unsafe, p2f.size)


@pytest.mark.parametrize("seed,n_faces,W,H", [(0, 300, 640, 480), (1, 40, 1500, 1100), (2, 5000, 333, 217)])
def test_triangle_soup_matches_oracle(torch, lib, seed, n_faces, W, H):
    """Random triangles of every size: edges longer than the 32-bit fast path allows, slivers, faces that cross the
    near plane or lie behind the camera, exact duplicates (tie -> lowest ID) and degenerate faces."""
    rng = np.random.default_rng(seed)
    centers = rng.uniform([-30, -30, -5], [30, 30, 60], size=(n_faces, 1, 3))
    size = rng.choice([0.3, 3.0, 40.0], size=(n_faces, 1, 1), p=[0.4, 0.4, 0.2])
    tri = centers + rng.normal(0, 1, size=(n_faces, 3, 3)) * size
    verts = tri.reshape(-1, 3)
    faces = np.arange(3 * n_faces, dtype=np.int32).reshape(-1, 3)
    faces[5] = faces[3]            # duplicate of face 3 -> exact depth tie, lowest ID must win
    faces[7] = [21, 21, 22]        # degenerate
    v32 = verts.astype(np.float32)
    c2w = np.eye(4)
    c2w[:3, 3] = [0.5, -0.25, -20.0]
    cam = ora.make_camera(c2w, 0.8 * W, 3.0, -2.0, W, H)
    ctx = _context(torch, lib, v32, faces)
    p2f = ctx.rasterize([_to_gg(lib, cam)]).cpu().numpy()[0]
    ref, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
    n_diff, n_unsafe = _assert_ids_match(p2f, ref, margin, f"soup {seed}")
    assert (ref >= 0).mean() > 0.2
    assert not (p2f == 5).any()  # the duplicate never wins over face 3
    assert n_diff <= n_unsafe


def test_fused_project_aggregate_equals_unfused(torch, lib):
    """gg_project_aggregate (winners kept in scratch, no raster in HBM) == gg_rasterize + gg_aggregate, bit for bit,
    with and without the -1 -> last-face quirk."""
    from geograypher_b200 import synthetic as syn

    v32, faces, cams, cfg = _scene("c1", 5)
    W, H = cfg.image_size
    F, C = len(faces), cfg.n_classes
    ctx = _context(torch, lib, v32, faces)
    gg = [_to_gg(lib, c) for c in cams]
    idx = [torch.from_numpy(syn.class_index_image(k, H, W, C)).cuda() for k in range(len(cams))]
    soft = [torch.from_numpy(syn.softmax_predictions(k, H, W, C, grid=(5, 7))).cuda() for k in range(len(cams))]
    for s in soft:
        s[::7, ::5, 1] = float("nan")
    p2f = ctx.rasterize(gg)
    for preds, kind in [(idx, lib.PRED_INDEX_U8), (soft, lib.PRED_F32)]:
        for flags in (0, lib.FLAG_COMPAT_NEGATIVE_INDEX):
            ref = _agg(torch, lib, ctx, p2f, preds, kind, C, lib.MODE_LAST_PIXEL, flags, F)
            d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
            d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
            out = torch.empty_like(p2f)
            ctx.project_aggregate(gg[:3], preds[:3], kind, C, lib.MODE_LAST_PIXEL, flags, d_sum, d_count, pix2face_out=out[:3])
            ctx.project_aggregate(gg[3:], preds[3:], kind, C, lib.MODE_LAST_PIXEL, flags, d_sum, d_count)
            avg, argmax = ctx.finalize(d_sum, d_count)
            assert torch.equal(out[:3], p2f[:3])
            _eq(avg.cpu().numpy(), ref[0]); _eq(d_count.cpu().numpy(), ref[1]); _eq(argmax.cpu().numpy(), ref[3])


def test_pix2face_golden_scene(torch, lib, golden_scene):
    g = golden_scene
    f, cx, cy, W, H = g["intrinsics"]
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    cams = [ora.make_camera(T, f, cx, cy, int(W), int(H), origin=g["origin"]) for T in g["c2ws"]]
    ctx = _context(torch, lib, v32, g["faces"])
    p2f = ctx.rasterize([_to_gg(lib, c) for c in cams]).cpu().numpy()
    for k, cam in enumerate(cams):
        _, _, margin = ora.rasterize(v32, g["faces"], cam, want_margin=True)
        _assert_ids_match(p2f[k], g["pix2face"][k].astype(np.int64), margin, f"golden view {k}")


def test_plane_known_answer(torch, lib):
    """The reference's render_flat known-answer test (tests/test_derived_meshes.py:23-76) on the GPU path."""
    from test_oracle_reference_pins import N, downward_view, pixel_idx, plane_mesh

    fill = np.array([[10, 20], [15, 190], [195, 5], [50, 100], [150, 120]])
    empty = np.array([[30, 40], [160, 180], [120, 40], [100, 150], [180, 100]])
    verts, faces = plane_mesh()
    colors = np.full((N * N, 3), 80, dtype=np.uint8)
    for p in fill:
        pixel_idx(colors, *p, stride=N, color=[255, 0, 0], buffer=1)
    tex = ora.vert_to_face_texture_mean(colors, faces)
    cam = ora.make_camera(downward_view(4, 100, 200), 100, 0, 0, 200, 200)
    ctx = _context(torch, lib, verts.astype(np.float32), faces)
    p2f = ctx.rasterize([_to_gg(lib, cam)])
    render = ctx.render_flat(p2f[0], torch.from_numpy(tex).cuda()).cpu().numpy()
    assert render.shape == (200, 200, 3)
    assert np.allclose(render[fill[:, 0], fill[:, 1]], [255, 0, 0])
    assert np.allclose(render[empty[:, 0], empty[:, 1]], [80, 80, 80])
    np.testing.assert_array_equal(p2f[0].cpu().numpy(), ora.rasterize(verts.astype(np.float32), faces, cam))


def _agg(torch, lib, ctx, p2f, preds, kind, C, mode, flags, F):
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    for k in range(len(preds)):
        ctx.aggregate(p2f[k], preds[k], kind, C, mode, flags, d_sum, d_count)
    avg, argmax = ctx.finalize(d_sum, d_count)
    torch.cuda.synchronize()
    return avg.cpu().numpy(), d_count.cpu().numpy(), d_sum.cpu().numpy(), argmax.cpu().numpy()


def _eq(a, b):
    np.testing.assert_array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))


def test_aggregate_golden(torch, lib, golden_scene, golden_aggregate):
    """Stage 3 + epilogue against the outputs of the reference's own aggregate_projected_images."""
    g, a = golden_scene, golden_aggregate
    F = g["faces"].shape[0]
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    ctx = _context(torch, lib, v32, g["faces"])
    p2f = torch.from_numpy(g["pix2face"]).cuda()
    C = a["avg1"].shape[1]
    compat = lib.FLAG_COMPAT_NEGATIVE_INDEX
    # (1) class-index image expanded on the fly == the reference's one-hot path
    idx = [torch.from_numpy(i).cuda() for i in a["idx_imgs"]]
    avg, cnt, summed, am = _agg(torch, lib, ctx, p2f, idx, lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, compat, F)
    _eq(avg, a["avg1"]); _eq(cnt, a["counts1"]); _eq(summed, a["summed1"]); _eq(am, a["argmax1"])
    # (1b) explicit (H,W,C) bool one-hot
    oh = [torch.from_numpy(ora.inds_to_one_hot(i, C).view(np.uint8)).cuda() for i in a["idx_imgs"]]
    avg, cnt, summed, _ = _agg(torch, lib, ctx, p2f, oh, lib.PRED_U8, C, lib.MODE_LAST_PIXEL, compat, F)
    _eq(avg, a["avg1"]); _eq(cnt, a["counts1"]); _eq(summed, a["summed1"])
    # (2) float32 scores with NaN holes, three views: bit-exact float64 sums in view order
    soft = [torch.from_numpy(s).cuda() for s in a["soft"]]
    avg, cnt, summed, am = _agg(torch, lib, ctx, p2f, soft, lib.PRED_F32, C, lib.MODE_LAST_PIXEL, compat, F)
    _eq(avg, a["avg2"]); _eq(cnt, a["counts2"]); _eq(summed, a["summed2"]); _eq(am, a["argmax2"][:, 0])
    # (2b) same as float64 input
    soft64 = [s.double() for s in soft]
    avg, cnt, summed, _ = _agg(torch, lib, ctx, p2f, soft64, lib.PRED_F64, C, lib.MODE_LAST_PIXEL, compat, F)
    _eq(avg, a["avg2"]); _eq(cnt, a["counts2"])
    # (2c) single view keeps per-channel NaNs (meshes.py:2056-2057)
    avg, cnt, summed, _ = _agg(torch, lib, ctx, p2f[:1], soft[:1], lib.PRED_F32, C, lib.MODE_LAST_PIXEL,
                               compat | lib.FLAG_KEEP_NAN, F)
    _eq(avg, a["avg2s"]); _eq(cnt, a["counts2s"]); _eq(summed, a["summed2s"])
    # (3) one-hot votes
    nv = int(a["n_vote_classes"])
    votes = [torch.from_numpy(v).cuda() for v in a["vote_imgs"]]
    avg, cnt, summed, _ = _agg(torch, lib, ctx, p2f, votes, lib.PRED_F64, nv, lib.MODE_VOTE, compat, F)
    # gg_finalize marks never-seen rows NaN (meshes.py:2070); scipy's sparse arrays leave them at 0
    _eq(cnt, a["counts3"]); _eq(np.nan_to_num(summed, nan=0.0), a["summed3"])
    _eq(np.nan_to_num(avg, nan=0.0), a["avg3"])


def test_render_flat_golden(torch, lib, golden_scene, golden_render):
    g, r = golden_scene, golden_render
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    ctx = _context(torch, lib, v32, g["faces"])
    p2f = torch.from_numpy(g["pix2face"]).cuda()
    for key_t, key_r in [("tex1", "render1"), ("tex3", "render3")]:
        out = ctx.render_flat(p2f, torch.from_numpy(r[key_t]).cuda()).cpu().numpy()
        _eq(out, r[key_r])
        u8 = ctx.render_flat(p2f, torch.from_numpy(r[key_t]).cuda(), out_dtype=lib.OUT_U8).cpu().numpy()
        for k in range(len(u8)):
            np.testing.assert_array_equal(np.squeeze(u8[k]), ora.cast_render_to_uint8(r[key_r][k]))


def test_pixel_sum_mode_matches_numpy(torch, lib, golden_scene, golden_aggregate):
    g, a = golden_scene, golden_aggregate
    F = g["faces"].shape[0]
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    ctx = _context(torch, lib, v32, g["faces"])
    p2f_np = g["pix2face"].astype(np.int64)
    p2f = torch.from_numpy(g["pix2face"]).cuda()
    soft = a["soft"].copy()
    C = soft.shape[-1]
    preds = [torch.from_numpy(s).cuda() for s in soft]
    avg, cnt, summed, _ = _agg(torch, lib, ctx, p2f, preds, lib.PRED_F32, C, lib.MODE_PIXEL_SUM, 0, F)
    ref_sum = np.zeros((F, C))
    ref_cnt = np.zeros(F, dtype=np.int64)
    for k in range(len(soft)):
        ids = p2f_np[k].ravel()
        keep = ids >= 0
        vals = np.nan_to_num(soft[k].reshape(-1, C).astype(np.float64), nan=0.0)
        np.add.at(ref_sum, ids[keep], vals[keep])
        np.add.at(ref_cnt, ids[keep], 1)
    np.testing.assert_array_equal(cnt, ref_cnt)
    seen = ref_cnt > 0
    np.testing.assert_allclose(summed[seen], ref_sum[seen], rtol=1e-9, atol=1e-12)


def test_same_bits_across_runs(torch, lib):
    v32, faces, cams, _ = _scene("c1", 3)
    ctx = _context(torch, lib, v32, faces)
    gg = [_to_gg(lib, c) for c in cams]
    a = ctx.rasterize(gg).clone()
    for _ in range(3):
        assert torch.equal(a, ctx.rasterize(gg))


def test_mixed_sizes_rejected(torch, lib):
    v32, faces, cams, _ = _scene("tiny", 2)
    ctx = _context(torch, lib, v32, faces)
    gg = [_to_gg(lib, c) for c in cams]
    gg[1].W = gg[1].W - 1
    with pytest.raises(lib.GeograypherB200Error, match="same image size"):
        ctx.rasterize(gg)


def test_scratch_overflow_is_detected_and_recovered(torch, lib):
    """A deliberately tiny (tile, face) capacity: gg_sync reports GG_ERR_OVERFLOW, the wrapper grows the scratch and
    replays; results equal the unconstrained run."""
    from geograypher_b200 import synthetic as syn

    v32, faces, cams, cfg = _scene("c1", 2)
    W, H = cfg.image_size
    F, C = len(faces), cfg.n_classes
    gg = [_to_gg(lib, c) for c in cams]
    idx = [torch.from_numpy(syn.class_index_image(k, H, W, C)).cuda() for k in range(len(cams))]
    ctx = _context(torch, lib, v32, faces)
    ref = ctx.rasterize(gg).clone()
    ref_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    ref_cnt = torch.zeros((F,), dtype=torch.int32, device="cuda")
    ctx.project_aggregate(gg, idx, lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, ref_sum, ref_cnt)

    small = _context(torch, lib, v32, faces)
    small.reserve(0, 1000)
    out = torch.empty_like(ref)
    with pytest.raises(lib.GeograypherB200Error, match="overflow"):
        small.rasterize(gg, out=out, check=False)
        small.sync()
    assert torch.equal(small.rasterize(gg), ref)  # check=True grows and replays
    small.reserve(0, 1000)
    d_sum = torch.zeros_like(ref_sum)
    d_cnt = torch.zeros_like(ref_cnt)
    small.project_aggregate(gg, idx, lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_cnt)
    assert torch.equal(d_sum, ref_sum) and torch.equal(d_cnt, ref_cnt)


@pytest.mark.parametrize("kind", ["f32", "index", "u8"])
def test_fused_pixel_sum_matches_numpy(torch, lib, kind):
    """gg_project_aggregate(GG_MODE_PIXEL_SUM): the rasterizer's dense epilogue (every pixel adds its scores) against
    a NumPy scatter-add over the oracle's rasters.  Not a reference mode; tolerance 1e-5 relative (float32 partial
    sums per tile, float64 across tiles)."""
    from geograypher_b200 import synthetic as syn

    v32, faces, cams, cfg = _scene("c1", 4)
    W, H = cfg.image_size
    F, C = len(faces), cfg.n_classes
    ctx = _context(torch, lib, v32, faces)
    gg = [_to_gg(lib, c) for c in cams]
    p2f = ctx.rasterize(gg).cpu().numpy()
    if kind == "f32":
        host = [syn.softmax_predictions(k, H, W, C, grid=(5, 7)) for k in range(len(cams))]
        for h in host:
            h[::9, ::4, 2] = np.nan
        dense = [np.nan_to_num(h.astype(np.float64), nan=0.0) for h in host]
        code = lib.PRED_F32
    elif kind == "index":
        host = [syn.class_index_image(k, H, W, C) for k in range(len(cams))]
        dense = [ora.inds_to_one_hot(h, C).astype(np.float64) for h in host]
        code = lib.PRED_INDEX_U8
    else:
        host = [ora.inds_to_one_hot(syn.class_index_image(k, H, W, C), C).view(np.uint8) for k in range(len(cams))]
        dense = [h.astype(np.float64) for h in host]
        code = lib.PRED_U8
    preds = [torch.from_numpy(h).cuda() for h in host]
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    out = torch.empty((len(cams), H, W), dtype=torch.int32, device="cuda")
    ctx.project_aggregate(gg, preds, code, C, lib.MODE_PIXEL_SUM, 0, d_sum, d_count, pix2face_out=out)
    np.testing.assert_array_equal(out.cpu().numpy(), p2f)
    ref_sum = np.zeros((F, C))
    ref_cnt = np.zeros(F, dtype=np.int64)
    for k in range(len(cams)):
        ids = p2f[k].ravel()
        keep = ids >= 0
        np.add.at(ref_sum, ids[keep], dense[k].reshape(-1, C)[keep])
        np.add.at(ref_cnt, ids[keep], 1)
    np.testing.assert_array_equal(d_count.cpu().numpy(), ref_cnt)
    np.testing.assert_allclose(d_sum.cpu().numpy(), ref_sum, rtol=1e-5, atol=1e-6)


def test_pipelined_batches_equal_serial(torch, lib):
    """Back-to-back unchecked gg_project_aggregate calls (binning of batch k+1 overlapping the rasterization of batch
    k on the internal streams) give the same bits as the same calls with the pipeline turned off."""
    from geograypher_b200 import synthetic as syn

    v32, faces, cams, cfg = _scene("c1", 10)
    W, H = cfg.image_size
    F, C = len(faces), cfg.n_classes
    gg = [_to_gg(lib, c) for c in cams]
    soft = [torch.from_numpy(syn.softmax_predictions(k, H, W, C, grid=(5, 7))).cuda() for k in range(len(cams))]
    results = []
    for pipelined in (False, True):
        ctx = _context(torch, lib, v32, faces)
        ctx.set_pipeline(pipelined)
        d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
        d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
        for rep in range(3):
            for s in range(0, len(gg), 2):
                ctx.project_aggregate(gg[s:s + 2], soft[s:s + 2], lib.PRED_F32, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_count,
                                      check=False)
        avg, argmax = ctx.finalize(d_sum, d_count)  # drains the pipeline on the caller's stream
        ctx.sync()
        results.append((avg.clone(), d_count.clone(), argmax.clone()))
    for a, b in zip(*results):
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))


def test_overflow_in_an_earlier_batch_grows_the_right_capacity(torch, lib):
    """the batch that overflows is usually not the last one before the sync.  Batch 1 (wide view: many
    face records) overflows a small record capacity, batch 2 (a camera that sees a handful of faces) does not; the
    high-water marks kept over ALL batches since the previous sync must drive the growth (records only: the tile
    capacity is not touched), and the API-level retry must end with the unconstrained result."""
    v32, faces, cams, cfg = _scene("c1", 2)
    W, H = cfg.image_size
    F, C = len(faces), cfg.n_classes
    gg = [_to_gg(lib, c) for c in cams]
    far = lib.GGCamera()  # same camera, looking at a corner: sees almost nothing
    for k in range(12):
        far.m[k] = gg[0].m[k]
    far.f, far.px, far.py, far.W, far.H, far.znear = gg[0].f * 40, gg[0].px, gg[0].py, W, H, gg[0].znear
    idx = [torch.from_numpy(syn.class_index_image(k, H, W, C)).cuda() for k in range(2)]
    ctx = _context(torch, lib, v32, faces)
    ref_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    ref_cnt = torch.zeros((F,), dtype=torch.int32, device="cuda")
    ctx.project_aggregate([gg[0]], idx[:1], lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, ref_sum, ref_cnt)
    ctx.project_aggregate([far], idx[1:], lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, ref_sum, ref_cnt)
    need = int(ctx.last_batch_stats(1)[0, 1])
    small = _context(torch, lib, v32, faces)
    small.reserve(2000, 0)
    d_sum, d_cnt = torch.zeros_like(ref_sum), torch.zeros_like(ref_cnt)
    small.project_aggregate([gg[0]], idx[:1], lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_cnt, check=False)
    small.project_aggregate([far], idx[1:], lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_cnt, check=False)
    with pytest.raises(lib.GeograypherB200Error, match="overflow"):
        small.sync()
    flags, want_recs, _ = small.overflow_info()
    assert flags == 1 and want_recs > 2000 and want_recs >= need
    bins_before = small.get_capacity()[1]
    small._grow_after_overflow()
    d_sum.zero_()
    d_cnt.zero_()
    small.project_aggregate([gg[0]], idx[:1], lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_cnt, check=False)
    small.project_aggregate([far], idx[1:], lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_cnt, check=False)
    small.sync()  # one growth step was enough
    assert small.get_capacity()[0] >= want_recs and small.get_capacity()[1] == bins_before
    assert torch.equal(d_sum, ref_sum) and torch.equal(d_cnt, ref_cnt)


def test_checked_call_reports_an_earlier_unchecked_overflow(torch, lib):
    """a checked call must not swallow (and then replay only itself after) an overflow that an earlier
    unchecked batch raised: it raises it before doing anything."""
    v32, faces, cams, cfg = _scene("c1", 1)
    gg = [_to_gg(lib, c) for c in cams]
    small = _context(torch, lib, v32, faces)
    small.reserve(0, 1000)
    out = torch.empty((1, cfg.image_size[1], cfg.image_size[0]), dtype=torch.int32, device="cuda")
    small.rasterize(gg, out=out, check=False)  # overflows, unnoticed
    with pytest.raises(lib.GeograypherB200Error, match="overflow"):
        small.rasterize(gg, out=out)
    small._grow_after_overflow()
    ref = _context(torch, lib, v32, faces).rasterize(gg)
    assert torch.equal(small.rasterize(gg), ref)


def test_set_mesh_again_with_a_larger_mesh(torch, lib):
    """the scratch slots are laid out for a mesh's block count; a second gg_set_mesh with a larger mesh
    (same record capacity class) must start the scratch over instead of writing past the visible-block lists."""
    small_v, small_f, _, _ = _scene("tiny", 1)
    v32, faces, cams, cfg = _scene("c1", 2)
    gg = [_to_gg(lib, c) for c in cams]
    ctx = _context(torch, lib, small_v, small_f)
    tcams = [_to_gg(lib, c) for c in _scene("tiny", 1)[2]]
    ctx.reserve(4096, 1 << 20)  # the same explicit capacities before and after: only the block count changes
    ctx.rasterize(tcams)
    ctx.set_mesh(torch.from_numpy(v32).cuda(), torch.from_numpy(faces.astype(np.int32)).cuda())
    ctx.reserve(1 << 17, 1 << 20)
    got = ctx.rasterize(gg)
    ref = _context(torch, lib, v32, faces).rasterize(gg)
    assert torch.equal(got, ref)


def test_gpu_against_the_independent_ray_caster(torch, lib):
    """The CUDA rasterizer against oracle_raycast.c directly (float64 ray casting, nothing shared with the
    rasterization contract) on the occlusion-rich cube / cylinder / cone scene (reference utils/example_data.py:9-112)
    from oblique cameras: equal face IDs on every edge-safe, depth-safe pixel."""
    from test_oracle_raycast import EPS_EDGE, _look_at, concept_scene

    verts, faces = concept_scene()
    v32 = verts.astype(np.float32)
    ctx = _context(torch, lib, v32, faces)
    cams = [ora.make_camera(_look_at(eye, (0, 0, 0.3)), 1000.0, 12.0, -7.0, 750, 550)
            for eye in [(-7, -7, 6), (7, -6, 5), (0, -9, 4)]]
    gpu = ctx.rasterize([_to_gg(lib, c) for c in cams]).cpu().numpy()
    for k, cam in enumerate(cams):
        rc, edge_safe, margin = ora.raycast(v32, faces, cam, eps_edge=EPS_EDGE)
        safe = edge_safe & (margin > 10 * EPS_DEPTH)
        assert safe.mean() > 0.9
        bad = safe & (gpu[k] != rc)
        assert not bad.any(), f"view {k}: {int(bad.sum())} safe pixels differ from the ray caster"
