"""TexturedPhotogrammetryMeshChunked (reference derived_meshes.py:23-317): camera clusters, sub-meshes within a buffer
of each cluster's cameras, merge by original face ID.  Chunking is not needed on a B200 (the whole mesh is resident and
every view is frustum-culled) but it CHANGES RESULTS -- far-but-visible faces are dropped, and so are far occluders --
so the class reproduces it; these tests pin the selection rule, the merge and the deliberate difference from the
unchunked path."""
import numpy as np
import pytest

import geograypher_b200 as gg
from geograypher_b200 import synthetic as syn
from oracle import oracle as ora


def _survey():
    verts, faces, c2ws, cfg = syn.make_survey("tiny")
    W, H = cfg.image_size
    intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws, intrinsic_params_per_sensor_type=intr)
    preds = [syn.softmax_predictions(i, H, W, cfg.n_classes, grid=(5, 7)) for i in range(len(cams))]
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor(preds, num_classes=cfg.n_classes))
    return verts, faces, c2ws, cfg, cams, seg, preds


def test_chunk_selection_rule():
    """A chunk = faces with at least one vertex within the buffer (planar distance) of a camera of the cluster
    (meshes.py:693-714: buffered camera points -> vertices inside -> adjacent cells); clusters partition the cameras."""
    verts, faces, c2ws, cfg, cams, seg, _ = _survey()
    mesh = gg.TexturedPhotogrammetryMeshChunked((verts, faces), log_level="WARNING")
    seen_cams = []
    for sub_mesh, sub_cams, face_IDs in mesh.get_mesh_chunks_for_cameras(cams, n_clusters=2, buffer_dist_meters=9.0):
        cam_xy = np.array([c.cam_to_world_transform[:2, 3] for c in sub_cams.cameras])
        d = np.linalg.norm(verts[:, None, :2] - cam_xy[None], axis=2).min(axis=1)
        want = np.where((d <= 9.0)[faces].any(axis=1))[0]
        np.testing.assert_array_equal(face_IDs, want)
        assert 0 < len(face_IDs) < len(faces)
        # the sub-mesh is those faces, re-indexed
        np.testing.assert_allclose(sub_mesh.points[sub_mesh.faces], verts[faces[face_IDs]])
        seen_cams += [tuple(c.cam_to_world_transform[:3, 3]) for c in sub_cams.cameras]
    assert sorted(seen_cams) == sorted(tuple(T[:3, 3]) for T in c2ws)
    with pytest.raises(ValueError):
        list(mesh.get_mesh_chunks_for_cameras(cams, n_clusters=len(c2ws) + 1))


def test_planar_coordinates_of_earth_centred_meshes():
    """ECEF input: distances are measured in the local tangent plane (the reference reprojects to a projected CRS)."""
    from geograypher_b200.meshes.derived_meshes import _planar_xy

    c = np.array([-2.7e6, -4.3e6, 3.8e6])
    up = c / np.linalg.norm(c)
    east = np.cross([0, 0, 1.0], up)
    east /= np.linalg.norm(east)
    north = np.cross(up, east)
    pts = c + np.outer([0, 30, 0, -40.0], east) + np.outer([0, 0, 50, 10.0], north) + np.outer([0, 5, -3, 2.0], up)
    xy, frame = _planar_xy(pts)
    assert frame is not None
    d = np.linalg.norm(xy[1:] - xy[0], axis=1)
    np.testing.assert_allclose(d, [30.0, 50.0, np.hypot(40, 10)], rtol=1e-6)


@pytest.mark.gpu
def test_chunked_with_a_buffer_that_covers_everything_equals_the_base_class():
    verts, faces, c2ws, cfg, cams, seg, _ = _survey()
    base = gg.TexturedPhotogrammetryMesh((verts, faces), log_level="WARNING")
    want_avg, want = base.aggregate_projected_images(seg)
    mesh = gg.TexturedPhotogrammetryMeshChunked((verts, faces), log_level="WARNING")
    avg, info = mesh.aggregate_projected_images(seg, n_clusters=2, buffer_dist_meters=1e4)
    assert info["projection_counts"].dtype.kind == "i"  # int in the reference (derived_meshes.py:264)
    np.testing.assert_array_equal(info["projection_counts"], want["projection_counts"])
    np.testing.assert_allclose(info["summed_projections"], want["summed_projections"], rtol=1e-12, atol=0, equal_nan=True)
    np.testing.assert_allclose(avg, want_avg, rtol=1e-12, atol=0, equal_nan=True)
    avg2, info2 = mesh.aggregate_projected_images(seg, chunked=False)
    np.testing.assert_array_equal(avg2, want_avg)


@pytest.mark.gpu
def test_chunked_aggregation_matches_the_reference_algorithm_and_differs_from_unchunked():
    """Small buffer: every cluster is aggregated on ITS sub-mesh (oracle rasterizer + the reference's NumPy
    aggregation, restated here chunk by chunk) and merged by face ID -- bit-identical; and the result is NOT the
    unchunked one: faces beyond the buffer are seen less often."""
    verts, faces, c2ws, cfg, cams, seg, preds = _survey()
    W, H = cfg.image_size
    mesh = gg.TexturedPhotogrammetryMeshChunked((verts, faces), log_level="WARNING")
    kw = dict(n_clusters=2, buffer_dist_meters=9.0)
    avg, info = mesh.aggregate_projected_images(seg, **kw)
    labels, cam_xy, vert_xy = mesh.camera_clusters(cams, 2)
    summed = np.zeros((len(faces), cfg.n_classes))
    counts = np.zeros(len(faces), dtype=int)
    for cluster in range(2):
        inds = np.where(labels == cluster)[0]
        d = np.linalg.norm(vert_xy[:, None] - cam_xy[inds][None], axis=2).min(axis=1)
        face_IDs = np.where((d <= 9.0)[faces].any(axis=1))[0]
        used, inverse = np.unique(faces[face_IDs].reshape(-1), return_inverse=True)
        sub_v, sub_f = verts[used], inverse.reshape(-1, 3).astype(np.int32)
        origin = 0.5 * (sub_v.min(0) + sub_v.max(0))
        v32 = (sub_v - origin).astype(np.float32)
        p2f = np.stack([ora.rasterize(v32, sub_f, ora.make_camera(c2ws[i], cfg.f, cfg.cx, cfg.cy, W, H, origin=origin))
                        for i in inds])
        _, sub_cnt, sub_sum = ora.aggregate(p2f, [preds[i] for i in inds], len(sub_f), compat_negative_index=False)
        summed[face_IDs] = np.nansum([summed[face_IDs], sub_sum], axis=0)
        counts[face_IDs] += sub_cnt.astype(int)
    summed[counts == 0] = np.nan
    np.testing.assert_array_equal(info["projection_counts"], counts)
    np.testing.assert_array_equal(info["summed_projections"], summed)
    with np.errstate(invalid="ignore", divide="ignore"):
        np.testing.assert_array_equal(avg, summed / counts[:, None])
    _, full = mesh.aggregate_projected_images(seg, chunked=False)
    assert (full["projection_counts"] >= counts).all() and (full["projection_counts"] > counts).sum() > 50


@pytest.mark.gpu
def test_chunked_render_flat_yields_one_render_per_camera():
    verts, faces, c2ws, cfg, cams, seg, _ = _survey()
    tex = syn.voronoi_face_labels(verts, faces, n_sites=10, n_classes=4)
    mesh = gg.TexturedPhotogrammetryMeshChunked((verts, faces), texture=tex, log_level="WARNING")
    renders = list(mesh.render_flat(cams, n_clusters=2, buffer_dist_meters=1e4, apply_distortion=False))
    plain = list(mesh.render_flat(cams, chunked=False, apply_distortion=False))
    assert len(renders) == len(plain) == len(c2ws)
    labels, _, _ = mesh.camera_clusters(cams, 2)
    order = [i for cluster in range(2) for i in np.where(labels == cluster)[0]]  # renders come cluster by cluster
    for r, i in zip(renders, order):
        np.testing.assert_array_equal(r, plain[i])
