"""GPU tests of the lens-distortion warp."""
from itertools import product

import numpy as np
import pytest

import geograypher_b200 as gg
from conftest import GOLDEN
from oracle import oracle as ora

pytestmark = pytest.mark.gpu
KEYS = ("k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN / "golden_metashape.npz"))


@pytest.fixture(scope="module")
def cams(golden, tmp_path_factory):
    path = tmp_path_factory.mktemp("ms") / "cameras.xml"
    path.write_text(str(golden["xml"]))
    return gg.MetashapeCameraSet(camera_file=path, image_folder="/mnt/images", original_image_folder="/data/survey/images")


def _small(golden):
    p = golden["small_params"]
    d = dict(f=p[0], cx=p[1], cy=p[2], image_width=int(p[3]), image_height=int(p[4]))
    d.update({k: p[5 + i] for i, k in enumerate(KEYS)})
    return d


@pytest.mark.parametrize("scale", [1.0, 0.5, 0.7])
def test_warp_map_matches_exact_inverse(golden, scale):
    from geograypher_b200 import _lib

    p = _small(golden)
    h, w = int(p["image_height"] * scale), int(p["image_width"] * scale)
    dist = _lib.make_distortion(p["f"], p["cx"], p["cy"], p["image_width"], p["image_height"], image_scale=scale,
                                **{k: p[k] for k in KEYS})
    src, rc = _lib.build_warp_map(dist, h, w, warped_to_ideal=False, want_coords=True)
    src, rc = src.cpu().numpy(), rc.cpu().numpy()
    rows, cols = ora.exact_inverse_coordinates(p, scale)
    want = ora.nearest_source_index(rows, cols, h, w)
    assert (src != want).sum() <= 2  # a sample that lands within 1e-9 of a .5 boundary may round the other way
    inside = (want >= 0) & (src >= 0)
    assert inside.mean() > 0.6
    np.testing.assert_allclose(rc[..., 0][inside], rows[inside], rtol=0, atol=2e-4)  # float32 storage
    np.testing.assert_allclose(rc[..., 1][inside], cols[inside], rtol=0, atol=2e-4)
    # forward direction (dewarping): exact formula, no iteration
    fsrc, frc = _lib.build_warp_map(dist, h, w, warped_to_ideal=True, want_coords=True)
    fwd = ora.ideal_to_warped_map(p, scale)
    np.testing.assert_allclose(frc.cpu().numpy()[..., 0], fwd[0], rtol=0, atol=2e-4)
    np.testing.assert_array_equal(fsrc.cpu().numpy(), ora.nearest_source_index(fwd[0], fwd[1], h, w))


def test_warp_agrees_with_reference_maps(golden):
    """IDs warped through the GPU map vs through the reference's own griddata maps (full-resolution and the default
    8x down-sampled inversion): identical except where the reference's interpolation error crosses a pixel boundary."""
    import torch

    from geograypher_b200 import _lib

    p = _small(golden)
    h, w = p["image_height"], p["image_width"]
    ids = np.arange(h * w, dtype=np.int32).reshape(h, w)
    dist = _lib.make_distortion(p["f"], p["cx"], p["cy"], w, h, **{k: p[k] for k in KEYS})
    src = _lib.build_warp_map(dist, h, w, warped_to_ideal=False)
    ours = _lib.gather_i32(torch.from_numpy(ids).cuda(), src, -1).cpu().numpy()
    for tag, floor in (("w2i_1_0", 0.99), ("w2i_ds8_1_0", 0.95)):
        ref = golden[tag].astype(float)
        theirs = ora.warp_ids(ids, ora.nearest_source_index(ref[0], ref[1], h, w), fill=-1)
        both = (ours >= 0) & (theirs >= 0)
        assert both.mean() > 0.6
        assert (ours[both] == theirs[both]).mean() > floor
        # where they differ it is by one pixel
        d = np.abs(ours[both] - theirs[both])
        assert set(np.unique(d)) <= {0, 1, w - 1, w, w + 1}


@pytest.mark.parametrize("render_img_scale", [0.5, 0.7, 0.9, 1.0])
def test_dewarp_pix2face_like_the_reference(cams, render_img_scale):
    """tests/test_derived_cameras.py:339-415 of the reference: a simplified Metashape camera (f=100, k1=-0.05, 257 px)
    parked over the plane; the warped raster pulls the corners in."""
    from test_oracle_reference_pins import downward_view, plane_mesh

    sensor = 2**8 + 1
    verts, faces = plane_mesh()
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), log_level="WARNING")
    cam = cams.cameras[0]
    cam.cx, cam.cy, cam.f = 0, 0, 100
    cam.image_height = cam.image_width = sensor
    cam.image_size = (sensor, sensor)
    cam.distortion_params = {k: 0 for k in KEYS}
    cam.distortion_params["k1"] = -0.05
    cams._local_to_epsg_4978_transform = np.eye(4)
    HT = downward_view(4, 100, sensor)
    cam.cam_to_world_transform, cam.world_to_cam_transform = HT, np.linalg.inv(HT)
    one = cams[0:1]
    kwargs = dict(cameras=one, cache_folder=None, distortion_set=cams, render_img_scale=render_img_scale)
    ideal = mesh.pix2face(**kwargs, apply_distortion=False)
    warped = mesh.pix2face(**kwargs, apply_distortion=True)
    assert len(ideal) == 1 and len(warped) == 1
    ideal, warped = ideal[0], warped[0]
    scaled = int(sensor * render_img_scale)
    for image in (ideal, warped):
        assert isinstance(image, np.ndarray) and image.dtype == np.int64 and image.shape == (scaled, scaled)
        assert image.min() >= -1 and image.max() < len(faces) and image.max() > 0.95 * len(faces)
    for corner in product([slice(None, 10), slice(-10, None)], repeat=2):
        assert len(np.unique(ideal[corner])) > 1
        assert np.all(warped[corner] == -1)
    # and exactly: the warped raster is the ideal raster gathered through the exact inverse map
    p = dict(f=100, cx=0, cy=0, image_width=sensor, image_height=sensor, **cam.distortion_params)
    rows, cols = ora.exact_inverse_coordinates(p, render_img_scale)
    want = ora.warp_ids(ideal, ora.nearest_source_index(rows, cols, scaled, scaled), fill=-1)
    assert (warped != want).sum() <= 2


def test_aggregation_with_distortion(cams, golden_scene, golden_aggregate):
    """aggregate_projected_images(distortion_set=...) == oracle raster -> exact warp -> reference aggregation."""
    g, a = golden_scene, golden_aggregate
    f, cx, cy, W, H = g["intrinsics"]
    W, H = int(W), int(H)
    dp = dict(k1=-0.08, k2=0.01, k3=0.0, k4=0.0, p1=0.002, p2=-0.001, b1=0.05, b2=-0.02)
    cam_list = [gg.PhotogrammetryCamera(f"/x/{i}.png", T, f, cx, cy, W, H, distortion_params=dict(dp))
                for i, T in enumerate(g["c2ws"])]
    cams.cameras = cam_list
    cams._local_to_epsg_4978_transform = np.eye(4)
    C = a["avg2"].shape[1]
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(list(a["soft"]), num_classes=C))
    mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), compat_negative_index=True, log_level="WARNING")
    avg, info = mesh.aggregate_projected_images(seg, distortion_set=cams, apply_distortion=True)
    p = dict(f=f, cx=cx, cy=cy, image_width=W, image_height=H, **dp)
    rows, cols = ora.exact_inverse_coordinates(p, 1.0)
    src = ora.nearest_source_index(rows, cols, H, W)
    # The lens model takes an IDEAL raster centred at (W/2, H/2) and adds cx, cy itself (reference
    # derived_cameras.py:171-208; the reference warps only the pyvista render, which has no principal point): the
    # expected value is the oracle raster with cx = cy = 0, warped through the model WITH cx, cy.  (The golden raster
    # was made with the principal point and must differ, or this test could not see a double application.)
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    ideal = [ora.rasterize(v32, g["faces"], ora.make_camera(T, f, 0.0, 0.0, W, H, origin=g["origin"])) for T in g["c2ws"]]
    assert any((ideal[k] != g["pix2face"][k]).any() for k in range(len(cam_list)))
    warped = np.stack([ora.warp_ids(ideal[k].astype(np.int64), src, fill=-1) for k in range(len(cam_list))])
    ref_avg, ref_cnt, ref_sum = ora.aggregate(warped, a["soft"], len(g["faces"]))
    np.testing.assert_array_equal(mesh.pix2face(cams, distortion_set=cams, apply_distortion=True), warped)
    np.testing.assert_array_equal(avg, ref_avg)
    np.testing.assert_array_equal(info["projection_counts"], ref_cnt)
