"""GPU test of label_polygons against the shapely-free NumPy restatement and against
hand-checkable cases.  The reference's implementation needs geopandas / shapely (absent here) and has no test of its
own: parity unpinned, see DESIGN.md."""
import numpy as np
import pytest

import geograypher_b200 as gg
from geograypher_b200 import synthetic as syn
from oracle import oracle as ora

pytestmark = pytest.mark.gpu


def _circle(cx, cy, r, n=40):
    a = np.linspace(0, 2 * np.pi, n, endpoint=False)
    return np.stack([cx + r * np.cos(a), cy + r * np.sin(a)], axis=1)


def test_label_polygons_matches_restatement():
    verts, faces = syn.terrain_mesh(64, 1.0, seed=5, crowns=True)
    rng = np.random.default_rng(1)
    labels = syn.voronoi_face_labels(verts, faces, n_sites=25, n_classes=6, nan_frac=0.2, seed=8)[:, 0]
    weighting = rng.uniform(0.5, 2.0, len(faces))
    polys = [[_circle(*rng.uniform(8, 56, 2), rng.uniform(2, 9))] for _ in range(12)]
    polys.append([_circle(32, 32, 14), _circle(32, 32, 6)[::-1]])          # a ring with a hole
    polys.append([_circle(10, 50, 5), _circle(50, 10, 4)])                   # two-part polygon
    polys.append([np.array([[0.2, 0.2], [0.7, 0.2], [0.7, 0.6], [0.2, 0.6]])])  # smaller than a face: no vote
    polys.append([np.array([[40.0, 40.0], [47.3, 40.2], [47.1, 46.9], [40.1, 47.2], [40.0, 40.0]])])  # closed ring
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), log_level="WARNING")
    for w in (None, weighting):
        got = mesh.label_polygons(labels, _as_inputs(polys), face_weighting=w, return_class_labels=False)
        want, weights = ora.label_polygons(verts, faces, labels, polys, face_weighting=w)
        assert len(got) == len(polys)
        np.testing.assert_array_equal(np.asarray(got, dtype=float), np.asarray(want, dtype=float))
        assert np.isnan(got[14])
        assert np.isfinite(np.asarray(got[:12], dtype=float)).sum() >= 8
        from geograypher_b200 import _lib

        w_gpu = _lib.label_polygons_weights(verts, verts[:, :2], faces, labels, w, polys, weights.shape[1])
        np.testing.assert_allclose(w_gpu, weights, rtol=1e-10, atol=1e-12)


class _Ring:
    def __init__(self, pts):
        self.coords = [tuple(p) for p in pts]


class _Poly:  # shapely-like duck
    def __init__(self, rings):
        self.exterior = _Ring(rings[0])
        self.interiors = [_Ring(r) for r in rings[1:]]


class _Multi:
    def __init__(self, parts):
        self.geoms = [_Poly([p]) for p in parts]


def _as_inputs(polys):
    out = []
    for i, p in enumerate(polys):
        if i == 12:
            out.append(_Poly(p))            # hole through the shapely-like interface
        elif i == 13:
            out.append(_Multi(p))           # multi-part
        elif i % 2:
            out.append({"exterior": p[0]})  # dict form
        else:
            out.append(p[0])                # bare array
    return out


def test_label_polygons_known_answer_and_names():
    """Flat 8x8 grid, left half class 2, right half class 0 (twice the weight): a polygon over the left half is 2,
    one straddling the middle goes to the heavier side, one outside the mesh is unknown."""
    verts, faces = syn.terrain_mesh(8, 1.0, seed=0, crowns=False)
    verts[:, 2] = 0.0
    cen = verts[faces].mean(axis=1)
    labels = np.where(cen[:, 0] < 4, 2.0, 0.0)
    weighting = np.where(cen[:, 0] < 4, 1.0, 2.0)
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), IDs_to_labels={0: "bare", 1: "shrub", 2: "tree"}, log_level="WARNING")
    polys = [np.array([[0.5, 0.5], [3.5, 0.5], [3.5, 7.5], [0.5, 7.5]]),
             np.array([[1.9, 1.9], [6.1, 1.9], [6.1, 6.1], [1.9, 6.1]]),
             np.array([[20, 20], [30, 20], [30, 30], [20, 30.0]])]
    assert mesh.label_polygons(labels, polys, face_weighting=weighting) == ["tree", "bare", "unknown"]
    ids = mesh.label_polygons(labels, polys, face_weighting=weighting, return_class_labels=False)
    assert ids[:2] == [2.0, 0.0] and np.isnan(ids[2])
    # unweighted: the straddling square holds as many faces of each class; the tie goes to the lower class ID (0)
    assert mesh.label_polygons(labels, polys[1:2], return_class_labels=False) == [0.0]


def test_label_polygons_overlay_known_areas():
    """sjoin_overlay=False (reference meshes.py:1263-1276): partially covered faces vote with the area of their
    intersection.  Flat unit grid (3D/2D ratio 1), a 2.5 x 2 rectangle placed off the grid lines: the weights are the
    areas of the rectangle's parts over each class, to rounding."""
    from geograypher_b200 import _lib

    verts, faces = syn.terrain_mesh(8, 1.0, seed=0, crowns=False)
    verts[:, 2] = 0.0
    cen = verts[faces].mean(axis=1)
    labels = np.where(cen[:, 0] < 4, 1.0, 0.0)
    rect = np.array([[2.25, 1.5], [4.75, 1.5], [4.75, 3.5], [2.25, 3.5]])
    w = _lib.label_polygons_weights(verts, verts[:, :2], faces, labels, None, [[rect]], 2, overlay=True)
    np.testing.assert_allclose(w, [[0.75 * 2.0, 1.75 * 2.0]], rtol=1e-12)  # class 0: x in [4, 4.75]; class 1: [2.25, 4]
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), log_level="WARNING")
    assert mesh.label_polygons(labels, [rect], sjoin_overlay=False, return_class_labels=False) == [1.0]
    # the sjoin path only counts faces entirely inside: x in [3, 4] x y in [2, 3] for class 1, nothing for class 0
    w_in = _lib.label_polygons_weights(verts, verts[:, :2], faces, labels, None, [[rect]], 2)
    np.testing.assert_allclose(w_in, [[0.0, 1.0]], rtol=1e-12)
    # tilted mesh: the pieces carry the face's 3D / 2D area ratio
    verts[:, 2] = verts[:, 0]  # 45 degrees: ratio sqrt(2)
    w45 = _lib.label_polygons_weights(verts, verts[:, :2], faces, labels, None, [[rect]], 2, overlay=True)
    np.testing.assert_allclose(w45, np.sqrt(2.0) * w, rtol=1e-12)


def test_label_polygons_overlay_matches_restatement():
    """Random polygons (convex and concave, with a hole, multi-part, clockwise and counter-clockwise rings) on a rough
    terrain with per-face weights: the CUDA kernel (signed fan triangles) against a pure-Python clipper that cuts the
    whole rings with each triangle."""
    from geograypher_b200 import _lib

    verts, faces = syn.terrain_mesh(24, 1.0, seed=3, crowns=True)
    rng = np.random.default_rng(4)
    labels = syn.voronoi_face_labels(verts, faces, n_sites=12, n_classes=5, nan_frac=0.2, seed=2)[:, 0]
    weighting = rng.uniform(0.5, 2.0, len(faces))
    star = _circle(12, 12, 6, 14)
    star[::2] = 12 + (star[::2] - 12) * 0.45  # concave
    polys = [{"exterior": _circle(7, 7, 4.3)}, {"exterior": _circle(16, 9, 3.1)[::-1]}, {"exterior": star},
             {"exterior": _circle(15, 16, 6.2), "holes": [_circle(15, 16, 2.4)]},
             {"exterior": np.array([[0.3, 20.2], [5.6, 19.7], [4.9, 23.8], [0.6, 23.1]])},
             {"exterior": np.array([[30.0, 30.0], [31, 30], [31, 31.0]])}]  # outside the mesh
    for w in (None, weighting):
        want_labels, want = ora.label_polygons_overlay(verts, faces, labels, polys, face_weighting=w)
        rings = [[p["exterior"]] + list(p.get("holes", [])) for p in polys]
        got = _lib.label_polygons_weights(verts, verts[:, :2], faces, labels, w, rings, want.shape[1], overlay=True)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-10)
        mesh = gg.TexturedPhotogrammetryMesh((verts, faces), log_level="WARNING")
        ids = mesh.label_polygons(labels, polys, face_weighting=w, sjoin_overlay=False, return_class_labels=False)
        np.testing.assert_array_equal(np.asarray(ids, dtype=float), np.asarray(want_labels, dtype=float))
    assert np.isnan(ids[-1]) and np.isfinite(ids[:5]).all()
    # overlay weights of a polygon always reach at least its sjoin ("within") weights
    inside = _lib.label_polygons_weights(verts, verts[:, :2], faces, labels, None, rings, want.shape[1])
    over = _lib.label_polygons_weights(verts, verts[:, :2], faces, labels, None, rings, want.shape[1], overlay=True)
    assert (over >= inside - 1e-9).all() and (over.sum() > inside.sum())
