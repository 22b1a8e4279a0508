"""CPU tests of the Metashape camera-file loader and lens model against vectors
produced by the reference's own MetashapeCameraSet (tests/golden/make_golden.py::main_metashape)."""
import numpy as np
import pytest

import geograypher_b200 as gg
from conftest import GOLDEN
from oracle import oracle as ora

KEYS = ("k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN / "golden_metashape.npz"))


@pytest.fixture(scope="module")
def cams(golden, tmp_path_factory):
    path = tmp_path_factory.mktemp("ms") / "cameras.xml"
    path.write_text(str(golden["xml"]))
    return gg.MetashapeCameraSet(camera_file=path, image_folder="/mnt/images", original_image_folder="/data/survey/images")


def _small_params(golden):
    p = golden["small_params"]
    d = dict(f=p[0], cx=p[1], cy=p[2], image_width=int(p[3]), image_height=int(p[4]))
    d.update({k: p[5 + i] for i, k in enumerate(KEYS)})
    return d


def test_xml_loads_like_the_reference(golden, cams):
    assert len(cams) == int(golden["n_cameras"]) == 3  # one unaligned camera and one of another component dropped
    np.testing.assert_array_equal(np.stack([c.cam_to_world_transform for c in cams.cameras]), golden["c2w"])
    assert [str(c.image_filename) for c in cams.cameras] == list(golden["filenames"])
    np.testing.assert_array_equal(cams.get_local_to_epsg_4978_transform(), golden["local_to_epsg_4978"])
    cam = cams.cameras[0]
    assert (cam.f, cam.cx, cam.cy) == (float(golden["f"]), float(golden["cx"]), float(golden["cy"]))
    assert (cam.image_width, cam.image_height) == tuple(golden["size"])
    assert sorted(cam.distortion_params) == list(golden["dist_keys"])
    np.testing.assert_array_equal([cam.distortion_params[k] for k in sorted(cam.distortion_params)], golden["dist_vals"])
    # reference tests/test_derived_cameras.py:118-136
    expected = {"b1": 0.5262024073, "b2": -0.3058334293, "k1": -0.0919367147, "k2": -0.0762807468, "k3": 0.1162639394,
                "k4": -0.0761413904, "p1": -0.0003134847, "p2": 0.0001164035}
    for k, v in cam.distortion_params.items():
        assert np.isclose(v, expected[k])


def test_ideal_to_warped_matches_reference(golden, cams):
    xw, yw = cams.ideal_to_warped(cams.cameras[0], golden["xp"], golden["yp"])
    np.testing.assert_allclose(xw, golden["xw"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(yw, golden["yw"], rtol=0, atol=1e-9)
    cam = cams.cameras[0]
    oxw, oyw = ora.metashape_ideal_to_warped(golden["xp"], golden["yp"], f=cam.f, cx=cam.cx, cy=cam.cy,
                                             image_width=cam.image_width, image_height=cam.image_height,
                                             **cam.distortion_params)
    np.testing.assert_allclose(oxw, golden["xw"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(oyw, golden["yw"], rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        cam2 = gg.PhotogrammetryCamera(None, np.eye(4), 100, 0, 0, 10, 10, distortion_params={"k1": 0.1, "zz": 1.0})
        cams.ideal_to_warped(cam2, 1.0, 1.0)
    with pytest.raises(KeyError):  # k1 is required (derived_cameras.py:181)
        cams.ideal_to_warped(gg.PhotogrammetryCamera(None, np.eye(4), 100, 0, 0, 10, 10, distortion_params={"k2": 0.1}), 1.0, 1.0)


def test_distortion_key(golden, cams):
    p = _small_params(golden)
    assert cams.distortion_key({k: p[k] for k in KEYS}, 0.5) == str(golden["key"])
    # reference tests/test_cameras.py:240-255 style
    assert cams.distortion_key({"b": 2.0, "a": 1.0}) == "a:1.00000000|b:2.00000000|image_scale:1.00000000"


@pytest.mark.parametrize("scale,tag", [(1.0, "1_0"), (0.5, "0_5")])
def test_oracle_maps_match_reference(golden, scale, tag):
    p = _small_params(golden)
    fwd = ora.ideal_to_warped_map(p, scale)
    np.testing.assert_allclose(fwd, golden[f"i2w_{tag}"], rtol=0, atol=2e-4)  # stored as float32
    inv = ora.inverse_map_griddata(fwd, downsample=1)
    ref = golden[f"w2i_{tag}"]
    both = (inv[0] >= 0) & (ref[0] >= 0)
    assert both.mean() > 0.7
    np.testing.assert_allclose(inv[:, both], ref[:, both], rtol=0, atol=2e-3)


@pytest.mark.parametrize("scale,tag", [(1.0, "1_0"), (0.5, "0_5"), (1.0, "ds8_1_0")])
def test_exact_inverse_is_what_the_reference_interpolates(golden, scale, tag):
    """The Newton inverse used by the GPU warp vs the reference's griddata inverse (full and 8x down-sampled
    forward map): they differ by interpolation error only."""
    p = _small_params(golden)
    rows, cols = ora.exact_inverse_coordinates(p, scale)
    ref = golden[f"w2i_{tag}"]
    h, w = ref.shape[1:]
    inside = (ref[0] >= 0) & np.isfinite(rows) & (rows > 1) & (rows < h - 2) & (cols > 1) & (cols < w - 2)
    assert inside.mean() > 0.6
    tol = 0.02 if tag != "ds8_1_0" else 0.1
    assert np.abs(rows[inside] - ref[0][inside]).max() < tol
    assert np.abs(cols[inside] - ref[1][inside]).max() < tol
    ours = ora.nearest_source_index(rows, cols, h, w)
    theirs = ora.nearest_source_index(ref[0].astype(float), ref[1].astype(float), h, w)
    assert (ours[inside] == theirs[inside]).mean() > (0.99 if tag != "ds8_1_0" else 0.95)


def test_host_source_coordinates(golden, cams):
    p = _small_params(golden)
    cam = gg.PhotogrammetryCamera(None, np.eye(4), p["f"], p["cx"], p["cy"], p["image_width"], p["image_height"],
                                  distortion_params={k: p[k] for k in KEYS})
    for scale in (1.0, 0.5):
        rows, cols = cams.warp_source_coordinates(cam, scale, warped_to_ideal=False)
        erows, ecols = ora.exact_inverse_coordinates(p, scale)
        ok = np.isfinite(erows)
        np.testing.assert_allclose(rows[ok], erows[ok], rtol=0, atol=1e-6)
        np.testing.assert_allclose(cols[ok], ecols[ok], rtol=0, atol=1e-6)
        fr, fc = cams.warp_source_coordinates(cam, scale, warped_to_ideal=True)
        fwd = ora.ideal_to_warped_map(p, scale)
        np.testing.assert_allclose(fr, fwd[0], rtol=0, atol=1e-9)
        np.testing.assert_allclose(fc, fwd[1], rtol=0, atol=1e-9)


def test_sampling_rule_matches_scipy():
    """nearest_source_index == scipy.ndimage.map_coordinates(order=0, mode='grid-constant'), which is what
    skimage.transform.warp (0.21, poetry.lock) runs for coordinate-array maps (utils/image.py:108-117)."""
    from scipy.ndimage import map_coordinates

    rng = np.random.default_rng(0)
    img = np.arange(1, 7 * 9 + 1, dtype=float).reshape(7, 9)
    rows = rng.uniform(-1.5, 8.5, 4000)
    cols = rng.uniform(-1.5, 10.5, 4000)
    rows[:10] = [-0.5, -0.49, -0.51, 6.49, 6.5, 6.51, 2.5, 3.5, 0.5, 1.5]
    cols[:10] = 4.0
    want = map_coordinates(img, [rows, cols], order=0, mode="grid-constant", cval=-1.0)
    idx = ora.nearest_source_index(rows, cols, 7, 9)
    got = ora.warp_ids(img.astype(np.int64), idx, fill=-1)
    np.testing.assert_array_equal(got, want)
