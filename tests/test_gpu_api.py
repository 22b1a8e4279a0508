"""GPU tests through the reference-facing Python API (TexturedPhotogrammetryMesh & co)."""
from itertools import product

import numpy as np
import pytest

import geograypher_b200 as gg
from oracle import oracle as ora

pytestmark = pytest.mark.gpu


def _eq(a, b):
    np.testing.assert_array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))


@pytest.fixture(scope="module")
def scene(golden_scene):
    g = golden_scene
    f, cx, cy, W, H = g["intrinsics"]
    cams = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(f"/golden/{i:04d}.png", T, f, cx, cy, int(W), int(H))
                 for i, T in enumerate(g["c2ws"])],
        local_to_epsg_4978_transform=np.eye(4),
    )
    return g, cams


def test_pix2face_contract(scene):
    g, cams = scene
    mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]))
    p2f = mesh.pix2face(cams, apply_distortion=False)
    assert p2f.dtype == np.int64 and p2f.shape == g["pix2face"].shape
    np.testing.assert_array_equal(p2f, g["pix2face"])
    one = mesh.pix2face(cams[1], apply_distortion=False)
    assert one.shape == g["pix2face"].shape[1:]
    np.testing.assert_array_equal(one, g["pix2face"][1])
    assert mesh.pix2face(cams[0:1], apply_distortion=False).shape == (1,) + one.shape  # length-1 set keeps the dim
    with pytest.raises(TypeError):
        mesh.pix2face([1, 2, 3])
    with pytest.raises(NotImplementedError):  # reference tests/test_derived_cameras.py:318-337
        mesh.pix2face(cameras=cams, cache_folder=None, distortion_set=cams, apply_distortion=True)


def test_principal_point_switch(scene):
    """use_principal_point=False is the base class's pyvista camera (no cx, cy: cameras.py:446-477); the default
    projects like the PyTorch3D renderer and ideal_to_warped (derived_meshes.py:772-780)."""
    g, cams = scene
    f, cx, cy, W, H = g["intrinsics"]
    assert cx != 0 or cy != 0
    centred = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(None, T, f, 0.0, 0.0, int(W), int(H)) for T in g["c2ws"]],
        local_to_epsg_4978_transform=np.eye(4),
    )
    with_pp = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]))
    without = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), use_principal_point=False)
    a = without.pix2face(cams, apply_distortion=False)
    np.testing.assert_array_equal(a, with_pp.pix2face(centred, apply_distortion=False))
    assert (a != with_pp.pix2face(cams, apply_distortion=False)).any()


def test_local_frame_transform(scene):
    """get_mesh_in_cameras_coords: an ECEF-like mesh + a similarity local->ECEF transform give the same rasters."""
    g, cams0 = scene
    f, cx, cy, W, H = g["intrinsics"]
    ang = 0.3
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    T = np.eye(4)
    T[:3, :3] = 12.354 * R
    T[:3, 3] = [-2.6e6, -4.3e6, 3.9e6]  # ECEF-sized offset
    ecef = g["verts"] @ T[:3, :3].T + T[:3, 3]
    cams = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(None, Tc, f, cx, cy, int(W), int(H)) for Tc in g["c2ws"]],
        local_to_epsg_4978_transform=T,
    )
    mesh = gg.TexturedPhotogrammetryMesh((ecef, g["faces"]), input_CRS="EPSG:4978")
    p2f = mesh.pix2face(cams, apply_distortion=False)
    # float64 round trip through a 4e6-sized offset moves vertices by ~1e-9 local units: allow a few flips at edges
    assert (p2f != g["pix2face"]).mean() < 2e-3


@pytest.mark.parametrize("sparse", [True, False])
def test_aggregate_matches_reference_outputs(scene, golden_aggregate, sparse):
    """NumPy (pageable) prediction arrays: by default the host gathers the rows the GPU asks for (sparse=True),
    else whole images are uploaded; both give the reference's numbers."""
    g, cams = scene
    a = golden_aggregate
    C = a["avg1"].shape[1]
    mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), compat_negative_index=True, views_per_batch=2,
                                         sparse_host_gather=sparse)
    # class-index segmentor -> on-the-fly one-hot on the GPU
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(list(a["idx_imgs"]), num_classes=C, one_hot=True))
    avg, info = mesh.aggregate_projected_images(seg, return_argmax=True)
    _eq(avg, a["avg1"]); _eq(info["projection_counts"], a["counts1"]); _eq(info["summed_projections"], a["summed1"])
    _eq(info["argmax"], a["argmax1"])
    assert info["projection_counts"].dtype == float
    # float predictions with NaNs
    seg2 = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(list(a["soft"]), num_classes=C))
    avg, info = mesh.aggregate_projected_images(seg2, return_all=True)
    _eq(avg, a["avg2"]); _eq(info["projection_counts"], a["counts2"]); _eq(info["summed_projections"], a["summed2"])
    for k, proj in enumerate(info["all_projections"]):
        _eq(proj, a["projs2"][k])
    _eq(gg.find_argmax_nonzero_value(avg, keepdims=True), a["argmax2"])
    # a single view keeps per-channel NaNs
    avg, info = mesh.aggregate_projected_images(seg2.get_subset_cameras([0]))
    _eq(avg, a["avg2s"]); _eq(info["summed_projections"], a["summed2s"])
    # default (bug-free) mode differs from the reference on the last face only
    mesh2 = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), sparse_host_gather=sparse)
    avg, info = mesh2.aggregate_projected_images(seg2)
    _eq(avg[:-1], a["avg2"][:-1]); _eq(info["projection_counts"][:-1], a["counts2"][:-1])


def test_lookup_segmentor_reads_ahead_and_matches_reference_outputs(scene, golden_aggregate, tmp_path):
    """Class-index images decoded from files (LookUpSegmentor, reference derived_segmentors.py:38-51): the coming views
    are read ahead on a thread pool; same numbers as the in-memory route, with and without read-ahead."""
    from geograypher_b200.utils.prefetch import OrderedPrefetcher

    g, _ = scene
    a = golden_aggregate
    C = a["avg1"].shape[1]
    f, cx, cy, W, H = g["intrinsics"]
    folder = tmp_path / "images"
    cams = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(folder / f"flight/{i:04d}.JPG", T, f, cx, cy, int(W), int(H))
                 for i, T in enumerate(g["c2ws"])], image_folder=folder)
    (tmp_path / "preds" / "flight").mkdir(parents=True)
    for i, img in enumerate(a["idx_imgs"]):
        np.save(tmp_path / "preds" / "flight" / f"{i:04d}.npy", np.asarray(img, dtype=np.uint8))
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.LookUpSegmentor(folder, tmp_path / "preds", num_classes=C))
    for threads in (None, 0, 3):
        mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), compat_negative_index=True, views_per_batch=2,
                                             prefetch_threads=threads)
        used = []
        orig = mesh._prediction_fetcher
        mesh._prediction_fetcher = lambda *args: (used.append(orig(*args)), used[-1])[1]
        avg, info = mesh.aggregate_projected_images(seg)
        assert isinstance(used[0], OrderedPrefetcher) == (threads != 0)
        _eq(avg, a["avg1"]); _eq(info["projection_counts"], a["counts1"]); _eq(info["summed_projections"], a["summed1"])


def test_pageable_route_redoes_a_batch_whose_lists_were_cut(scene, golden_aggregate):
    """The pageable-image route sizes the (face, pixel) lists after the previous batch (GG_FLAG_TRUNCATE): a batch
    that needs more room is listed again at full size -- same numbers, and no overflow is reported."""
    g, cams = scene
    a = golden_aggregate
    C = a["avg1"].shape[1]
    mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), compat_negative_index=True, views_per_batch=2)
    seg2 = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(list(a["soft"]), num_classes=C))
    mesh.__dict__["_sparse_guess"] = 3  # far fewer than the faces any view sees
    began = []
    orig = mesh._sparse_begin
    mesh._sparse_begin = lambda *args, **kw: (began.append(kw.get("full", False)), orig(*args, **kw))[1]
    avg, info = mesh.aggregate_projected_images(seg2)
    assert True in began and False in began  # the first batch was cut and redone
    _eq(avg, a["avg2"]); _eq(info["projection_counts"], a["counts2"]); _eq(info["summed_projections"], a["summed2"])
    assert mesh.__dict__["_sparse_guess"] >= 8192


@pytest.mark.parametrize("sparse", [True, False])
def test_votes_match_reference_outputs(scene, golden_aggregate, sparse):
    g, cams = scene
    a = golden_aggregate
    nv = int(a["n_vote_classes"])
    mesh = gg.TexturedPhotogrammetryMeshIndexPredictions((g["verts"], g["faces"]), compat_negative_index=True,
                                                         sparse_host_gather=sparse)
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(list(a["vote_imgs"])))
    avg, info = mesh.aggregate_projected_images(seg, n_classes=nv)
    _eq(avg.toarray(), a["avg3"])
    _eq(info["projection_counts"].toarray()[:, 0], a["counts3"])
    _eq(info["summed_projections"].toarray(), a["summed3"])


def test_render_flat_matches_reference_outputs(scene, golden_render):
    g, cams = scene
    r = golden_render
    for key_t, key_r in [("tex1", "render1"), ("tex3", "render3")]:
        # set_texture stores the array as is; passing it to the constructor would remap a one-column texture to
        # class IDs exactly like the reference's load_texture does (meshes.py:383-473)
        mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]))
        mesh.set_texture(r[key_t], is_vertex_texture=False)
        renders = list(mesh.render_flat(cams, apply_distortion=False))
        assert len(renders) == len(cams)
        for k, img in enumerate(renders):
            assert img.dtype == np.float64
            _eq(img, r[key_r][k])
        img, cam = next(mesh.render_flat(cams, return_camera=True))
        assert cam is cams[0]
        u8 = next(mesh.render_flat_device(cams, out_dtype="uint8")).cpu().numpy()
        np.testing.assert_array_equal(np.squeeze(u8[0]), ora.cast_render_to_uint8(r[key_r][0]))
        # the pipelined device-to-host path: several views per batch, a ragged last batch, images kept by the
        # caller (every image owns its host block), and the uint8 extension
        for bs in (2, len(cams)):
            kept = list(mesh.render_flat(cams, batch_size=bs, apply_distortion=False))
            assert len(kept) == len(cams)
            for k, img in enumerate(kept):
                _eq(img, r[key_r][k])
        for k, img in enumerate(mesh.render_flat(cams, batch_size=2, out_dtype="uint8")):
            assert img.dtype == np.uint8
            np.testing.assert_array_equal(np.squeeze(img), np.squeeze(ora.cast_render_to_uint8(r[key_r][k])))
    with pytest.raises(TypeError):
        list(mesh.render_flat("cameras"))


@pytest.mark.parametrize("render_img_scale", [0.5, 0.7, 0.9, 1.0])
def test_reference_structural_pins(render_img_scale):
    """tests/test_derived_cameras.py:339-415 of the reference, undistorted half."""
    from test_oracle_reference_pins import downward_view, plane_mesh

    sensor = 2**8 + 1
    verts, faces = plane_mesh()
    cam = gg.PhotogrammetryCamera(None, downward_view(4, 100, sensor), 100, 0, 0, sensor, sensor)
    cams = gg.PhotogrammetryCameraSet([cam])
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces))
    ideal = mesh.pix2face(cameras=cams, cache_folder=None, render_img_scale=render_img_scale, apply_distortion=False)
    assert len(ideal) == 1
    ideal = ideal[0]
    scaled = int(sensor * render_img_scale)
    assert isinstance(ideal, np.ndarray) and ideal.dtype == np.int64 and ideal.shape == (scaled, scaled)
    assert ideal.min() >= -1 and ideal.max() < len(faces) and ideal.max() > 0.95 * len(faces)
    for corner in product([slice(None, 10), slice(-10, None)], repeat=2):
        assert len(np.unique(ideal[corner])) > 1
    np.testing.assert_array_equal(
        ideal, ora.rasterize(verts.astype(np.float32) - 0, faces,
                             ora.make_camera(cam.cam_to_world_transform, 100, 0, 0, sensor, sensor, render_img_scale,
                                             origin=np.zeros(3))))


def test_save_renders(scene, golden_render, tmp_path):
    """save_renders writes one uint8 file per image, with the reference's cast rule, under the image's relative path."""
    g, cams0 = scene
    r = golden_render
    f, cx, cy, W, H = g["intrinsics"]
    cams = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(f"/data/imgs/flight{i % 2}/{i:04d}.JPG", T, f, cx, cy, int(W), int(H))
                 for i, T in enumerate(g["c2ws"])])
    mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), log_level="WARNING")
    mesh.set_texture(r["tex1"], is_vertex_texture=False)
    mesh.save_renders(cams, output_folder=tmp_path / "out", save_as_npy=True, apply_distortion=False)
    assert (tmp_path / "out" / "IDs_to_labels.json").exists()
    for k in range(len(cams)):
        arr = np.load(tmp_path / "out" / f"flight{k % 2}" / f"{k:04d}.npy")
        assert arr.dtype == np.uint8
        np.testing.assert_array_equal(arr, ora.cast_render_to_uint8(r["render1"][k]))
    try:
        from PIL import Image
    except ImportError:
        return
    mesh.save_renders(cams, output_folder=tmp_path / "tif", apply_distortion=False)
    img = np.asarray(Image.open(tmp_path / "tif" / "flight1" / "0001.tif"))
    np.testing.assert_array_equal(img, ora.cast_render_to_uint8(r["render1"][1]))


@pytest.mark.parametrize("order,D", [(0, 1), (1, 1), (1, 3)])
def test_resize_render_kernel_matches_skimage_semantics(order, D):
    """gg_resize_render == skimage.transform.resize(order) as save_renders uses it (reference meshes.py:2312-2321),
    i.e. scipy.ndimage.zoom(grid_mode=True, mode='mirror'): NaNs, non-integer factors, 1 and 3 channels."""
    import torch

    from geograypher_b200 import _lib

    rng = np.random.default_rng(order * 10 + D)
    img = rng.uniform(-20, 300, size=(37, 53, D))
    img[rng.random((37, 53)) < 0.05] = np.nan
    want = ora.resize_render(img, (111, 140), order)
    got = _lib.resize_render(torch.from_numpy(img).cuda(), 111, 140, order).cpu().numpy()
    if order == 0:
        np.testing.assert_array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, equal_nan=True)
    got8 = _lib.resize_render(torch.from_numpy(img).cuda(), 111, 140, order, _lib.OUT_U8).cpu().numpy()
    np.testing.assert_array_equal(got8, ora.cast_render_to_uint8(got).reshape(got8.shape))


def test_save_renders_native_resolution(scene, golden_render, tmp_path):
    """Renders at half scale written at the camera's native size: nearest neighbour for a discrete (label) texture,
    bilinear for a real-valued one, then the uint8 rule -- like the reference's host-side post-processing."""
    g, cams0 = scene
    r = golden_render
    f, cx, cy, W, H = g["intrinsics"]
    W, H = int(W), int(H)
    cams = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(f"/data/imgs/{i:04d}.JPG", T, f, cx, cy, W, H) for i, T in enumerate(g["c2ws"])])
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    for discrete in (True, False):
        tex = r["tex1"] if discrete else r["tex1"] * 0.37 + 1.5
        mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), log_level="WARNING",
                                             IDs_to_labels=({i: str(i) for i in range(10)} if discrete else None))
        mesh.set_texture(tex, is_vertex_texture=False)
        assert mesh.is_discrete_texture() == discrete
        out = tmp_path / ("nn" if discrete else "lin")
        n_bytes = mesh.save_renders(cams, render_image_scale=0.5, save_native_resolution=True, output_folder=out,
                                    save_as_npy=True, apply_distortion=False)
        assert n_bytes == len(cams) * H * W
        for k, T in enumerate(g["c2ws"]):
            cam = ora.make_camera(T, f, cx, cy, W, H, render_img_scale=0.5, origin=g["origin"])
            small = ora.render_flat_gather(ora.rasterize(v32, g["faces"], cam), tex)
            want = ora.cast_render_to_uint8(ora.resize_render(small, (H, W), 0 if discrete else 1))
            arr = np.load(out / f"{k:04d}.npy")
            assert arr.shape == (H, W) and arr.dtype == np.uint8
            np.testing.assert_array_equal(arr, np.squeeze(want))


def test_host_array_inputs_are_read_in_place(scene, golden_aggregate):
    """geograypher_b200.host_array: host memory mapped into the GPU.  Prediction images held in it go through the fused
    zero-copy route (pointer kind = managed, no upload) and give bit-identical results to every other route -- the
    reference's own aggregate on the golden scene."""
    from geograypher_b200 import _lib

    g, cams = scene
    a = golden_aggregate
    C = a["avg2"].shape[1]
    images = []
    for img in a["soft"]:
        h = gg.host_array(img.shape, img.dtype)
        assert _lib.pointer_kind(h) == _lib.POINTER_MANAGED and _lib.pointer_kind(np.zeros(4)) == _lib.POINTER_PAGEABLE
        h[...] = img
        images.append(h)
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(images, num_classes=C))
    mesh = gg.TexturedPhotogrammetryMesh((g["verts"], g["faces"]), compat_negative_index=True, log_level="WARNING")
    routes = []
    orig = mesh._to_device_or_mapped

    def spy(arr, dev, zero_copy=True):
        out = orig(arr, dev, zero_copy)
        routes.append(type(out).__name__)
        return out

    mesh._to_device_or_mapped = spy
    avg, info = mesh.aggregate_projected_images(seg)
    assert routes and set(routes) == {"_HostMapped"}
    np.testing.assert_array_equal(avg, a["avg2"])
    np.testing.assert_array_equal(info["projection_counts"], a["counts2"])
    np.testing.assert_array_equal(info["summed_projections"], a["summed2"])
    del images, seg  # the buffers are released with the arrays


def test_many_batches_through_every_host_route_equal_the_upload_route():
    """14 batches of one or two views (the software pipeline rotates through its three scratch sets several times; the
    pageable route runs two-deep): prediction images in host_array memory, page-locked memory and ordinary NumPy arrays
    give the same bits as whole images uploaded to the device, float32 scores and uint8 class indices alike."""
    import torch

    from geograypher_b200 import synthetic as syn

    verts, faces, c2ws, cfg = syn.make_survey("c1")
    W, H = cfg.image_size
    C = cfg.n_classes
    n = min(len(c2ws), 27)
    intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws[:n], intrinsic_params_per_sensor_type=intr)
    soft = [syn.softmax_predictions(k, H, W, C, grid=(5, 7)) for k in range(n)]
    for img in soft[::5]:
        img[::7, ::3, :] = np.nan  # nulls
    idx = [np.argmax(np.nan_to_num(img), axis=2).astype(np.uint8) for img in soft]

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()

    def hosted(a):
        h = gg.host_array(a.shape, a.dtype)
        h[...] = a
        return h

    results = {}
    for route, wrap in (("device", lambda a: a.copy()), ("host_array", hosted), ("pinned", pinned),
                        ("pageable", lambda a: a.copy())):
        mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=2, log_level="WARNING",
                                             sparse_host_gather=(route != "device"))
        seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([wrap(a) for a in soft], num_classes=C))
        avg, info = mesh.aggregate_projected_images(seg, return_argmax=True)
        segi = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([wrap(a) for a in idx], num_classes=C, one_hot=True))
        avgi, infoi = mesh.aggregate_projected_images(segi)
        results[route] = (avg, info["projection_counts"], info["summed_projections"], info["argmax"], avgi,
                          infoi["projection_counts"])
    assert (results["device"][1] > 0).sum() > 100
    for route in ("host_array", "pinned", "pageable"):
        for got, want in zip(results[route], results["device"]):
            np.testing.assert_array_equal(got, want, err_msg=route)
