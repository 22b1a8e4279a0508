"""The oracle's NumPy restatements against vectors produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np

from oracle import oracle as ora


def _eq(a, b):
    np.testing.assert_array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))


def test_scene_pix2face_is_reproducible(golden_scene):
    """The committed pix2face rasters are what the C oracle produces today for the stored scene."""
    g = golden_scene
    f, cx, cy, W, H = g["intrinsics"]
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    cams = [ora.make_camera(T, f, cx, cy, int(W), int(H), origin=g["origin"]) for T in g["c2ws"]]
    p2f = ora.pix2face_set(v32, g["faces"], cams)
    assert p2f.dtype == np.int64 and p2f.shape == g["pix2face"].shape
    np.testing.assert_array_equal(p2f, g["pix2face"])


def test_camera_pins(golden_scene):
    g = golden_scene
    np.testing.assert_allclose(np.linalg.inv(g["c2ws"]), g["cam_world_to_cam"], rtol=0, atol=1e-12)
    _, _, _, W, H = g["intrinsics"]
    for s, key in [(1.0, "cam_size_s1"), (0.7, "cam_size_s07"), (0.5, "cam_size_s05")]:
        assert tuple(g[key]) == ora.scaled_image_size(int(H), int(W), s)


def test_one_hot(golden_aggregate):
    a = golden_aggregate
    oh = ora.inds_to_one_hot(a["idx_imgs"][0], a["avg1"].shape[1])
    assert oh.dtype == bool
    np.testing.assert_array_equal(oh, a["onehot0"])


def test_aggregate_one_hot(golden_scene, golden_aggregate):
    a, p2f = golden_aggregate, golden_scene["pix2face"].astype(np.int64)
    F, C = a["avg1"].shape
    imgs = [ora.inds_to_one_hot(i, C) for i in a["idx_imgs"]]
    avg, counts, summed = ora.aggregate(p2f, imgs, F)
    _eq(avg, a["avg1"])
    _eq(counts, a["counts1"])
    _eq(summed, a["summed1"])


def test_aggregate_float_with_nans(golden_scene, golden_aggregate):
    a, p2f = golden_aggregate, golden_scene["pix2face"].astype(np.int64)
    F = a["avg2"].shape[0]
    avg, counts, summed = ora.aggregate(p2f, a["soft"], F)
    _eq(avg, a["avg2"])
    _eq(counts, a["counts2"])
    _eq(summed, a["summed2"])
    for k in range(len(p2f)):
        _eq(ora.project_image(p2f[k], a["soft"][k], F), a["projs2"][k])
    avg, counts, summed = ora.aggregate(p2f[:1], a["soft"][:1], F)
    _eq(avg, a["avg2s"])
    _eq(counts, a["counts2s"])
    _eq(summed, a["summed2s"])


def test_negative_index_compat_is_the_only_difference(golden_scene, golden_aggregate):
    """Without the -1 -> last-face quirk (meshes.py:2000) only face F-1 may change."""
    a, p2f = golden_aggregate, golden_scene["pix2face"].astype(np.int64)
    F = a["avg2"].shape[0]
    avg, counts, _ = ora.aggregate(p2f, a["soft"], F, compat_negative_index=False)
    _eq(avg[:-1], a["avg2"][:-1])
    _eq(counts[:-1], a["counts2"][:-1])


def test_votes(golden_scene, golden_aggregate):
    a, p2f = golden_aggregate, golden_scene["pix2face"].astype(np.int64)
    F = a["avg3"].shape[0]
    avg, counts, summed = ora.aggregate_votes(p2f, a["vote_imgs"], F, int(a["n_vote_classes"]))
    _eq(avg, a["avg3"])
    _eq(counts, a["counts3"])
    _eq(summed, a["summed3"])


def test_render_flat(golden_scene, golden_render):
    r, p2f = golden_render, golden_scene["pix2face"].astype(np.int64)
    for k in range(len(p2f)):
        _eq(ora.render_flat_gather(p2f[k], r["tex1"]), r["render1"][k])
        _eq(ora.render_flat_gather(p2f[k], r["tex3"]), r["render3"][k])


def test_argmax(golden_aggregate):
    a = golden_aggregate
    _eq(ora.find_argmax_nonzero_value(a["avg1"]), a["argmax1"])
    _eq(ora.find_argmax_nonzero_value(a["avg2"], keepdims=True), a["argmax2"])


def test_uint8_cast_rule():
    x = np.array([[-1.0, 0.0, 3.7], [255.0, 255.5, np.nan], [np.inf, 12.0, -np.inf]])[..., None]
    np.testing.assert_array_equal(
        ora.cast_render_to_uint8(x), np.array([[0, 0, 3], [255, 0, 0], [0, 12, 0]], dtype=np.uint8)
    )
