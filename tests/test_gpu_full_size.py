"""Full-size checks (BASELINE.json configurations c2 and c5).

  * pix2face at full size against the CPU oracle (a 20-Mpx view of the 2M-face mesh takes the oracle a fraction of a
    second per view on the box's host cores, a 45-Mpx view of the 20M-face mesh a few seconds): c2 views including
    ones at the edge of the mesh, c5 nadir + oblique rig views -- every pixel, depth-margin mask as everywhere else;
  * the reference's aggregation rule restated with torch scatter ops on the full-size rasters (last pixel of every
    face, row-major: meshes.py:2001; NaN -> 0 and count of finite rows: meshes.py:2057-2067) -- bit-exact;
  * ID round trip: render_flat of the texture "face f has value f" gives back pix2face (meshes.py:1921-1937);
  * counting identity, linearity with integer-valued scores, invariance under the batch partition;
  * dense mode: a checksum over all faces equals the sum over all covered pixels.
"""
import numpy as np
import pytest

from geograypher_b200 import synthetic as syn
from oracle import oracle as ora

pytestmark = pytest.mark.gpu

N_VIEWS = 6
EPS_DEPTH = 1e-5  # the contract's depth-margin mask


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def lib():
    from geograypher_b200 import _lib

    return _lib


_HOST = {}  # name -> (float32 vertices, faces, [cam_to_world], origin) kept for the oracle comparisons


def _build(torch, lib, name, n_cams):
    verts, faces, c2ws, cfg = syn.make_survey(name, n_cams)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    W, H = cfg.image_size
    v32 = (verts - origin).astype(np.float32)
    ctx = lib.Context(0)
    ctx.set_mesh(torch.from_numpy(v32).cuda(), torch.from_numpy(faces).cuda())
    cams = [lib.make_camera(np.linalg.inv(T), cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for T in c2ws]
    _HOST[name] = (v32, faces, c2ws, origin)
    return ctx, cams, cfg, len(faces)


def _assert_matches_oracle(name, cfg, view, gpu_ids, label):
    """Every pixel of one full-size view against the oracle; IDs may differ only inside the depth-margin mask."""
    v32, faces, c2ws, origin = _HOST[name]
    W, H = cfg.image_size
    cam = ora.make_camera(c2ws[view], cfg.f, cfg.cx, cfg.cy, W, H, origin=origin)
    ref, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
    gpu = gpu_ids.cpu().numpy()
    assert gpu.shape == ref.shape == (H, W)
    diff = gpu != ref
    bad = diff & (margin > EPS_DEPTH)
    assert not bad.any(), (f"{label}: {int(bad.sum())} pixels differ outside the depth-margin mask (first at "
                           f"{np.argwhere(bad)[0]}, gpu={gpu[bad][0]}, oracle={ref[bad][0]})")
    masked = int((margin <= EPS_DEPTH).sum())
    assert masked < 1e-3 * ref.size, f"{label}: {masked} pixels inside the depth-margin mask"
    assert (ref >= 0).mean() > 0.3  # the view does look at the mesh
    return int(diff.sum()), masked


@pytest.fixture(scope="module")
def c2(torch, lib):
    ctx, cams, cfg, F = _build(torch, lib, "c2", N_VIEWS)
    W, H = cfg.image_size
    p2f = ctx.rasterize(cams)  # (n, H, W) int32
    assert tuple(p2f.shape) == (N_VIEWS, H, W) == (N_VIEWS, 3648, 5472) and F == 2_000_000
    return ctx, cams, cfg, F, p2f


@pytest.mark.parametrize("view", [0, 3, 5])
def test_c2_pix2face_matches_the_oracle_at_full_size(c2, view):
    """5472 x 3648 rasters of the 2M-face canopy mesh, every pixel: 171 x 456 tile grids, > 30 k face records per
    view, 32-bit fast-path edge values at full scale.  View 0 sits at the corner of the mesh (background pixels)."""
    _, _, cfg, _, p2f = c2
    _assert_matches_oracle("c2", cfg, view, p2f[view], f"c2 view {view}")


def _last_pixel(torch, ids, F):
    """Per face: index of its last pixel in row-major order (-1 = the view does not see it), with torch ops."""
    flat = ids.reshape(-1).long()
    keep = flat >= 0
    pix = torch.arange(flat.numel(), device=flat.device)
    last = torch.full((F,), -1, dtype=torch.long, device=flat.device)
    return last.scatter_reduce(0, flat[keep], pix[keep], reduce="amax", include_self=True)


def _reference_rule(torch, p2f, preds, F):
    """aggregate_projected_images restated with torch: per view, every seen face takes the row of its last pixel;
    sum with NaN -> 0 in view order (float64); count the views in which the row has a finite value."""
    C = preds[0].shape[-1]
    total = torch.zeros((F, C), dtype=torch.float64, device=p2f.device)
    count = torch.zeros((F,), dtype=torch.int32, device=p2f.device)
    for k, pred in enumerate(preds):
        last = _last_pixel(torch, p2f[k], F)
        seen = last >= 0
        rows = pred.reshape(-1, C)[last[seen]].double()
        total[seen] += torch.nan_to_num(rows, nan=0.0, posinf=float("inf"), neginf=float("-inf"))
        count[seen] += torch.isfinite(rows).any(dim=1).int()
    return total, count


def test_c2_last_pixel_aggregation_equals_the_reference_rule(torch, lib, c2):
    ctx, cams, cfg, F, p2f = c2
    W, H = cfg.image_size
    C = cfg.n_classes
    preds = [syn.softmax_predictions_device(k, H, W, C, torch.device("cuda", 0)) for k in range(N_VIEWS)]
    preds[2][100:900, 2000:3000, 3] = float("nan")     # partly null rows still count, their NaNs add nothing
    preds[4][1500:1700, :, :] = float("nan")           # fully null rows do not count
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    ctx.project_aggregate(cams, preds, lib.PRED_F32, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_count)
    ctx.sync()
    want_sum, want_count = _reference_rule(torch, p2f, preds, F)
    assert torch.equal(d_count, want_count)
    assert torch.equal(d_sum, want_sum)  # float64 sums in view order: bit-identical
    assert int((d_count > 0).sum()) > 30_000
    avg, amax = ctx.finalize(d_sum, d_count, want_avg=True, want_argmax=True)
    seen = d_count > 0
    assert torch.equal(avg[seen], want_sum[seen] / want_count[seen].double()[:, None])
    assert torch.isnan(avg[~seen]).all()


def test_c2_render_flat_of_face_ids_is_pix2face(torch, lib, c2):
    ctx, cams, cfg, F, p2f = c2
    tex = torch.arange(F, dtype=torch.float64, device="cuda")[:, None]
    fused_ids = torch.empty_like(p2f[:2])
    out = ctx.rasterize_render_flat(cams[:2], tex, out_dtype=lib.OUT_F64, pix2face_out=fused_ids)
    assert torch.equal(fused_ids, p2f[:2])
    hit = p2f[:2] >= 0
    assert torch.equal(out[..., 0][hit], p2f[:2][hit].double())
    assert torch.isnan(out[..., 0][~hit]).all()
    two_step = ctx.render_flat(p2f[:2], tex, out_dtype=lib.OUT_F64)
    assert torch.equal(torch.nan_to_num(two_step, nan=-1.0), torch.nan_to_num(out, nan=-1.0))
    # uint8 labels: the cast rule (values outside 0..255 -> 0, meshes.py:2323-2334) on the same rasters
    labels = (torch.arange(F, device="cuda") % 300).double()[:, None]
    u8 = ctx.rasterize_render_flat(cams[:2], labels, out_dtype=lib.OUT_U8)
    want = torch.where(hit, p2f[:2].long() % 300, torch.zeros_like(p2f[:2], dtype=torch.long))
    want = torch.where(want > 255, torch.zeros_like(want), want)
    assert torch.equal(u8[..., 0].long(), want)


def test_c2_counting_linearity_and_batch_partition(torch, lib, c2):
    ctx, cams, cfg, F, p2f = c2
    W, H = cfg.image_size
    C = 4
    gen = torch.Generator(device="cuda").manual_seed(7)
    X = [torch.randint(0, 256, (H, W, C), device="cuda", generator=gen).float() for _ in range(N_VIEWS)]
    Y = [torch.randint(0, 256, (H, W, C), device="cuda", generator=gen).float() for _ in range(N_VIEWS)]

    def run(preds, batches):
        d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
        d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
        for a, b in batches:
            ctx.project_aggregate(cams[a:b], preds[a:b], lib.PRED_F32, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_count, check=False)
        ctx.sync()
        return d_sum, d_count

    one = [(0, N_VIEWS)]
    sx, cx = run(X, one)
    sy, cy = run(Y, one)
    # counting identity: every view that sees a face counts once
    views_seeing = torch.zeros((F,), dtype=torch.int32, device="cuda")
    for k in range(N_VIEWS):
        views_seeing += (_last_pixel(torch, p2f[k], F) >= 0).int()
    assert torch.equal(cx, views_seeing) and torch.equal(cy, views_seeing)
    # linearity (integer-valued scores: exact in float32 and float64)
    Z = [2.0 * x + 3.0 * y for x, y in zip(X, Y)]
    sz, cz = run(Z, one)
    assert torch.equal(sz, 2.0 * sx + 3.0 * sy) and torch.equal(cz, cx)
    # the batch partition (and the software pipeline across batches) does not change a bit
    s3, c3 = run(X, [(0, 2), (2, 3), (3, N_VIEWS)])
    assert torch.equal(s3, sx) and torch.equal(c3, cx)


def test_c2_dense_mode_checksums(torch, lib, c2):
    ctx, cams, cfg, F, p2f = c2
    W, H = cfg.image_size
    C = cfg.n_classes
    n = 3
    gen = torch.Generator(device="cuda").manual_seed(3)
    preds = [torch.randint(0, 64, (H, W, C), device="cuda", generator=gen).float() for _ in range(n)]
    preds[1][700:720, 100:4000, 5] = float("nan")  # nulls add nothing; these tiles take the filtering loop
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    ctx.project_aggregate(cams[:n], preds, lib.PRED_F32, C, lib.MODE_PIXEL_SUM, 0, d_sum, d_count)
    ctx.sync()
    hit = p2f[:n] >= 0
    assert int(d_count.sum()) == int(hit.sum())
    want = torch.stack([torch.nan_to_num(p, nan=0.0).double()[hit[k]].sum(0) for k, p in enumerate(preds)]).sum(0)
    assert torch.equal(d_sum.sum(0), want)  # small integers: every partial sum is exact
    # per face, against a torch scatter-add of one view
    ids = p2f[0].reshape(-1).long()
    keep = ids >= 0
    one_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    one_sum.index_add_(0, ids[keep], preds[0].reshape(-1, C)[keep].double())
    s1 = torch.zeros_like(d_sum)
    c1 = torch.zeros_like(d_count)
    ctx.project_aggregate(cams[:1], preds[:1], lib.PRED_F32, C, lib.MODE_PIXEL_SUM, 0, s1, c1)
    ctx.sync()
    assert torch.equal(s1, one_sum)
    assert torch.equal(c1.long(), torch.bincount(ids[keep], minlength=F))


@pytest.fixture(scope="module")
def c5(torch, lib):
    return _build(torch, lib, "c5", 5)  # station 0 of the rig: nadir + four obliques pitched 30 degrees


@pytest.mark.parametrize("view", [0, 2])
def test_c5_pix2face_matches_the_oracle_at_full_size(torch, c5, view):
    """8192 x 5460 rasters of the 20M-face mesh: the nadir camera of a rig station and one of its obliques (long
    tile lists towards the far end of the footprint)."""
    ctx, cams, cfg, F = c5
    p2f = ctx.rasterize([cams[view]])
    _assert_matches_oracle("c5", cfg, view, p2f[0], f"c5 view {view}")


def test_c5_votes_equal_the_reference_rule(torch, lib, c5):
    """20M faces, 8192 x 5460 rig views, class-index images with ignore values: one-hot votes per face."""
    ctx, cams, cfg, F = c5
    cams = cams[:3]
    W, H = cfg.image_size
    C = cfg.n_classes
    assert (W, H) == (8192, 5460) and F > 19_000_000
    index = [torch.from_numpy(syn.class_index_image(k, H, W, C)).cuda() for k in range(3)]
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    ctx.project_aggregate(cams, index, lib.PRED_INDEX_U8, C, lib.MODE_LAST_PIXEL, 0, d_sum, d_count)
    ctx.sync()
    p2f = ctx.rasterize(cams)
    want_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda")
    want_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    for k in range(3):
        last = _last_pixel(torch, p2f[k], F)
        seen = torch.nonzero(last >= 0)[:, 0]
        cls = index[k].reshape(-1)[last[seen]].long()
        valid = cls < C  # ignored pixels (255) give an all-zero one-hot row, which is finite: the view still counts
        want_sum.index_put_((seen[valid], cls[valid]), torch.ones(int(valid.sum()), dtype=torch.float64, device="cuda"),
                            accumulate=True)
        want_count[seen] += 1
    assert torch.equal(d_count, want_count)
    assert torch.equal(d_sum, want_sum)
    assert int((d_count > 0).sum()) > 150_000
