"""CPU tests of the host-side mirror of the reference interface and of the C-ABI surface (no compute calls)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import geograypher_b200 as gg
from geograypher_b200 import _lib
from geograypher_b200 import synthetic as syn
from oracle import oracle as ora

ROOT = Path(__file__).resolve().parents[1]


def _camera_set(n=3, W=96, H=72):
    cfg = syn.SurveyConfig("t", 24, 1.0, 7, True, (1, n), (7.0, 0.0), (W, H), 70.0, 1.5, -2.25, 30.0, 17, 4)
    c2ws = syn.lawnmower_cameras(cfg, 24.0, jitter_deg=6.0)
    intr = {0: dict(f=70.0, cx=1.5, cy=-2.25, image_width=W, image_height=H, distortion_params={})}
    return gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws, intrinsic_params_per_sensor_type=intr), c2ws


def test_camera_matches_reference_pins(golden_scene):
    g = golden_scene
    f, cx, cy, W, H = g["intrinsics"]
    cams = gg.PhotogrammetryCameraSet(
        cameras=[gg.PhotogrammetryCamera(f"/golden/{i:04d}.png", T, f, cx, cy, int(W), int(H)) for i, T in enumerate(g["c2ws"])]
    )
    for k, cam in enumerate(cams.cameras):
        np.testing.assert_array_equal(cam.world_to_cam_transform, g["cam_world_to_cam"][k])
    assert cams[0].get_image_size(1.0) == tuple(g["cam_size_s1"])
    assert cams[0].get_image_size(0.7) == tuple(g["cam_size_s07"])
    assert cams[0].get_image_size(0.5) == tuple(g["cam_size_s05"])
    assert cams.n_image_channels() == int(g["cam_n_channels"])
    props = cams[0].get_camera_properties()
    assert set(props) == {"focal_length", "principal_point_x", "principal_point_y", "image_height", "image_width",
                          "distortion_params", "world_to_cam_transform"}


def test_camera_set_container_semantics():
    cams, c2ws = _camera_set(4)
    assert len(cams) == 4 and cams.n_cameras() == 4
    assert isinstance(cams[1], gg.PhotogrammetryCamera)
    sub = cams[1:3]
    assert isinstance(sub, gg.PhotogrammetryCameraSet) and len(sub) == 2
    np.testing.assert_array_equal(sub[0].cam_to_world_transform, c2ws[1])
    sub2 = cams.get_subset_cameras([3, 0])
    np.testing.assert_array_equal(sub2[0].cam_to_world_transform, c2ws[3])
    assert len(cams) == 4  # deepcopy semantics: the original is untouched
    np.testing.assert_array_equal(cams.get_local_to_epsg_4978_transform(), np.eye(4))
    with pytest.raises(ValueError):
        gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws, sensor_IDs=[0],
                                   intrinsic_params_per_sensor_type={0: {}, 1: {}})
    with pytest.raises(NotImplementedError):  # reference tests/test_derived_cameras.py:331-337
        cams.warp_dewarp_image(cams[0], np.zeros((4, 4)))


def test_segmentor_camera_set(golden_aggregate):
    a = golden_aggregate
    cams, _ = _camera_set(3)
    C = a["avg1"].shape[1]
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(list(a["idx_imgs"]), num_classes=C, one_hot=True))
    assert seg.n_image_channels() == C
    np.testing.assert_array_equal(seg.get_image_by_index(0), a["onehot0"])
    np.testing.assert_array_equal(seg.get_class_index_image_by_index(1), a["idx_imgs"][1])
    sub = seg.get_subset_cameras([2, 1])
    np.testing.assert_array_equal(sub.get_class_index_image_by_index(0), a["idx_imgs"][2])
    np.testing.assert_array_equal(gg.Segmentor.inds_to_one_hot(a["idx_imgs"][0], C), a["onehot0"])

    class RefStyle(gg.Segmentor):  # a reference-style segmentor only accepts filename / image_scale
        def segment_image(self, image, filename, image_scale):
            return np.full((2, 2), 7.0)

    assert gg.SegmentorPhotogrammetryCameraSet(cams, RefStyle()).get_image_by_index(0)[0, 0] == 7.0


def test_subsets_share_the_prediction_arrays(golden_aggregate):
    """Sub-setting a segmentor camera set deep-copies the cameras (like the reference) but never the in-memory
    prediction images -- a copy would cost gigabytes and lose their page-locking; the shards of the multi-GPU path
    do not copy the cameras either."""
    from geograypher_b200 import distributed as ggd

    a = golden_aggregate
    cams, _ = _camera_set(3)
    C = a["avg1"].shape[1]
    images = list(a["idx_imgs"])
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(images, num_classes=C, one_hot=True))
    deep = seg.get_subset_cameras([2, 0])
    assert deep.segmentor.images[2] is images[2] and deep[0] is not seg[2]
    assert deep.get_class_index_image_by_index(0) is images[2]
    for rank, want in ((0, [0, 1]), (1, [2])):
        shard = ggd.shard_cameras(seg, rank, 2)
        assert len(shard) == len(want) and shard[0] is seg[want[0]]
        for k, orig in enumerate(want):
            assert shard.get_class_index_image_by_index(k) is images[orig]
    assert len(seg) == 3


def test_find_argmax_matches_reference(golden_aggregate):
    a = golden_aggregate
    np.testing.assert_array_equal(gg.find_argmax_nonzero_value(a["avg1"]), a["argmax1"])
    np.testing.assert_array_equal(gg.find_argmax_nonzero_value(a["avg2"], keepdims=True), a["argmax2"])


def test_make_camera_equals_oracle_record():
    """The product's float32 camera record and the oracle's are built independently and must agree bit for bit."""
    cams, c2ws = _camera_set(3)
    origin = np.array([12.0, 11.5, 3.25])
    for scale in (1.0, 0.7, 0.5):
        for cam, T in zip(cams.cameras, c2ws):
            a = _lib.make_camera(cam.world_to_cam_transform, cam.f, cam.cx, cam.cy, cam.image_width,
                                 cam.image_height, render_img_scale=scale, origin=origin)
            b = ora.make_camera(T, cam.f, cam.cx, cam.cy, cam.image_width, cam.image_height, scale, origin=origin)
            assert bytes(a) == bytes(b)


def test_mesh_host_side():
    verts, faces = syn.terrain_mesh(8, 1.0, seed=1, crowns=False)
    flat = np.concatenate([np.full((len(faces), 1), 3), faces], axis=1).ravel()  # pyvista layout
    m = gg.TexturedPhotogrammetryMesh((verts, flat), texture=np.arange(len(verts), dtype=float) * 0.5)
    np.testing.assert_array_equal(m.faces, faces)
    assert m.vertex_texture.shape == (len(verts), 1) and m.face_texture is None
    ft = m.get_texture(request_vertex_texture=False)
    np.testing.assert_allclose(ft[:, 0], ora.vert_to_face_texture_mean(np.arange(len(verts)) * 0.5, faces))
    disc = m.vert_to_face_texture(np.array([0, 1, 1, 2, np.nan, 2, 2, 5.0] + [3.0] * (len(verts) - 8)), discrete=True)
    assert disc.shape == (len(faces),)
    with pytest.raises(ValueError):
        m.set_texture(np.zeros(7))
    with pytest.raises(NotImplementedError):
        gg.TexturedPhotogrammetryMesh((verts, faces), downsample_target=0.5)
    with pytest.raises(NotImplementedError):
        m.save_renders(None, make_composites=True)
    with pytest.raises(ValueError):  # reference meshes.py:1181-1185
        m.label_polygons(np.zeros((len(faces), 2)), [])
    rings = m._polygon_rings([np.array([[0, 0], [1, 0], [1, 1]]), {"exterior": [[0, 0], [4, 0], [4, 4]], "holes": [[[1, 1], [2, 1], [2, 2]]]}])
    assert [len(r) for r in rings] == [1, 2]
    assert len(m.get_mesh_hash()) == 64


def test_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/geograypher_b200.h declares."""
    header = (ROOT / "include" / "geograypher_b200.h").read_text()
    declared = set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", header)) - {"gg_context"}
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gg_abi_version() == 1
    assert ctypes.sizeof(_lib.GGCamera) == 72 == ctypes.sizeof(ora.OraCamera)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.GeograypherB200Error, match="no CPU fallback"):
        _lib.Context(0)
    h = ctypes.c_void_p()
    assert _lib.load().gg_create(0, ctypes.byref(h)) == -5
    assert b"no CPU fallback" in _lib.load().gg_last_error()


def test_product_never_imports_the_oracle():
    for path in (ROOT / "geograypher_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_gather_rows_host_matches_numpy_indexing():
    """gg_gather_rows_host (the host half of the pageable-image route; no GPU involved) == NumPy fancy indexing, for
    packed and padded pair lists, float32 rows and uint8 class indices, empty views and out-of-range pixels (clipped)."""
    from geograypher_b200 import _lib

    rng = np.random.default_rng(3)
    imgs = [rng.random((40, 50, 7)).astype(np.float32) for _ in range(5)]
    m = [300, 0, 9000, 1, 4500]
    offs = np.concatenate([[0], np.cumsum(m)])
    pix = [rng.integers(0, 2000, k).astype(np.int32) for k in m]
    pix[2][:3] = [-5, 2000, 10**6]  # np.take(mode="clip") semantics
    ref = np.concatenate([imgs[v].reshape(-1, 7)[np.clip(pix[v], 0, 1999)] for v in range(5)])
    packed = np.zeros((offs[-1], 2), np.int32)
    packed[:, 1] = np.concatenate(pix)
    for threads in (1, 3, 0):
        out = np.zeros((offs[-1], 7), np.float32)
        _lib.gather_rows_host(imgs, packed, offs, out, n_threads=threads)
        np.testing.assert_array_equal(out, ref)
    cap = 9100
    padded = np.full((5, cap, 2), -7, np.int32)
    for v in range(5):
        padded[v, : m[v], 1] = pix[v]
    out = np.zeros((offs[-1], 7), np.float32)
    _lib.gather_rows_host(imgs, padded.reshape(-1, 2), offs, out, pair_starts=np.arange(5) * cap)
    np.testing.assert_array_equal(out, ref)
    idx = [rng.integers(0, 10, (40, 50)).astype(np.uint8) for _ in range(5)]
    out8 = np.zeros((offs[-1], 1), np.uint8)
    _lib.gather_rows_host(idx, packed, offs, out8)
    np.testing.assert_array_equal(out8[:, 0], np.concatenate([idx[v].reshape(-1)[np.clip(pix[v], 0, 1999)] for v in range(5)]))
    with pytest.raises(ValueError):
        _lib.gather_rows_host(imgs, packed, offs[:-1], out)


def test_vote_accumulators_to_csr_equals_the_dense_route():
    """TexturedPhotogrammetryMeshIndexPredictions builds its CSR results (reference derived_meshes.py:527-550) where the
    accumulators live (torch ops; CPU tensors here): same data, indices, index pointers and dtypes as csr_array(dense)
    of the host route, including a face that was observed but recorded no vote."""
    import torch
    from scipy.sparse import csr_array

    from geograypher_b200.meshes.derived_meshes import TexturedPhotogrammetryMeshIndexPredictions as Mesh

    rng = np.random.default_rng(0)
    F, C = 1000, 7
    counts = rng.integers(0, 4, F).astype(np.int32)
    summed = np.zeros((F, C))
    for f in range(F):
        for _ in range(counts[f]):
            summed[f, rng.integers(0, C)] += 1
    counts[5], summed[5] = 3, 0
    got = Mesh._votes_to_csr(torch.from_numpy(summed), torch.from_numpy(counts))
    ci, si = counts.astype(np.int64), summed.astype(np.int64)
    average = np.zeros(si.shape)
    seen = ci > 0
    average[seen] = si[seen] * np.reciprocal(ci[seen].astype(float))[:, None]
    for a, b in zip(got, [csr_array(average), csr_array(ci[:, None]), csr_array(si)]):
        assert a.shape == b.shape and a.dtype == b.dtype and a.has_canonical_format
        assert a.indices.dtype == b.indices.dtype and a.indptr.dtype == b.indptr.dtype
        np.testing.assert_array_equal(a.indptr, b.indptr)
        np.testing.assert_array_equal(a.indices, b.indices)
        np.testing.assert_array_equal(a.data, b.data)
    empty = Mesh._votes_to_csr(torch.zeros((4, 3), dtype=torch.float64), torch.zeros(4, dtype=torch.int32))
    assert [m.nnz for m in empty] == [0, 0, 0] and empty[1].shape == (4, 1)


def test_to_fresh_host_delivers_owned_arrays():
    """_lib.to_fresh_host: tensors -> new NumPy arrays (pre-faulted when large), any shape / dtype / layout."""
    import torch

    from geograypher_b200 import _lib

    a = torch.arange(6_000_000, dtype=torch.float64).reshape(1000, 6000)  # large enough to be pre-faulted
    b = torch.arange(7, dtype=torch.int32)
    c = torch.zeros((0, 3), dtype=torch.uint8)
    d = torch.arange(12, dtype=torch.int64).reshape(3, 4).t()  # not contiguous
    ra, rb, rc, rd = _lib.to_fresh_host([a, b, c, d])
    for got, want in zip((ra, rb, rc, rd), (a, b, c, d)):
        assert isinstance(got, np.ndarray) and got.flags.owndata and got.flags.c_contiguous
        assert got.dtype == want.numpy().dtype and got.shape == tuple(want.shape)
        np.testing.assert_array_equal(got, want.numpy())
    assert isinstance(_lib.to_fresh_host(b), np.ndarray)
