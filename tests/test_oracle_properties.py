"""Properties of the rasterization contract, checked on the CPU oracle.  The GPU tests
compare the CUDA path with the oracle; these make sure the checker itself behaves like a rasterizer should:
watertight shared edges (top-left rule), nearest face wins, lowest ID wins exact ties, the order of the faces is
irrelevant, and clipping at the near plane only removes what is behind it.  CPU only."""
import numpy as np
import pytest

from geograypher_b200.synthetic import grid_faces, terrain_mesh
from oracle import oracle as ora


def nadir(height, f=90.0, W=96, H=72, cx=0.0, cy=0.0, znear=1e-3, x=0.0, y=0.0):
    T = np.diag([1.0, -1.0, -1.0, 1.0])
    T[:3, 3] = [x, y, height]
    return ora.make_camera(T, f, cx, cy, W, H, znear=znear)


def quad(z=0.0, half=3.0, theta=0.3):
    """A square in the plane z, rotated so that no edge is axis-aligned; two triangles sharing the diagonal."""
    c, s = np.cos(theta), np.sin(theta)
    xy = np.array([[-half, -half], [half, -half], [half, half], [-half, half]]) @ np.array([[c, -s], [s, c]]).T
    verts = np.column_stack([xy, np.full(4, z)]).astype(np.float32)
    return verts, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)


def test_shared_edges_are_watertight_and_never_double_covered():
    """Top-left rule (C3): pixels on the shared diagonal belong to exactly one of the two triangles."""
    verts, faces = quad()
    cam = nadir(8.0)
    both = ora.rasterize(verts, faces, cam)
    alone = [ora.rasterize(verts, faces[k : k + 1], cam) >= 0 for k in range(2)]
    assert not (alone[0] & alone[1]).any()                      # no pixel is claimed twice
    np.testing.assert_array_equal(alone[0] | alone[1], both >= 0)  # and none falls through the crack
    np.testing.assert_array_equal(both == 0, alone[0])
    np.testing.assert_array_equal(both == 1, alone[1])
    # the same on a grid of 2 x 24 x 24 triangles, where every interior edge is shared
    g = np.linspace(-3, 3, 25)
    X, Y = np.meshgrid(g, g, indexing="xy")
    gv = np.column_stack([X.ravel(), Y.ravel(), 0.1 * np.sin(X.ravel()) * np.cos(Y.ravel())]).astype(np.float32)
    gf = grid_faces(24, 24)
    p2f = ora.rasterize(gv, gf, cam)
    covered = np.zeros(p2f.shape, dtype=int)
    for k in range(0, len(gf), 97):  # a sample of single faces: their pixels are exactly the ones they win
        own = ora.rasterize(gv, gf[k : k + 1], cam) >= 0
        np.testing.assert_array_equal(own, p2f == k)
        covered += own
    assert covered.max() <= 1
    inside = (p2f >= 0)
    assert inside[20:50, 30:60].all()  # no holes in the interior of the sheet


def test_nearest_face_wins_and_exact_ties_go_to_the_lowest_id():
    near_v, f = quad(z=1.0, half=1.5)
    far_v, _ = quad(z=0.0, half=3.0)
    verts = np.vstack([far_v, near_v])
    faces = np.vstack([f, f + 4]).astype(np.int32)  # IDs 0, 1 far; 2, 3 near
    cam = nadir(8.0)
    p2f, depth, margin = ora.rasterize(verts, faces, cam, want_depth=True, want_margin=True)
    near_only = ora.rasterize(near_v, f, cam) >= 0
    assert np.isin(p2f[near_only], [2, 3]).all()
    assert np.isin(p2f[~near_only & (p2f >= 0)], [0, 1]).all()
    assert (margin[near_only] > 0.1).all()  # the far plane is 1/7 further away
    # the face list in the opposite order (IDs swapped accordingly) gives the same picture
    swapped = ora.rasterize(verts, np.vstack([f + 4, f]).astype(np.int32), cam)
    np.testing.assert_array_equal(np.where(swapped >= 0, (swapped + 2) % 4, -1), p2f)
    # an exact duplicate of the near quad: depth ties everywhere, the lower IDs keep every pixel
    dup = np.vstack([faces, f + 4]).astype(np.int32)  # IDs 4, 5 duplicate 2, 3
    p2f_dup, _, margin_dup = ora.rasterize(verts, dup, cam, want_depth=True, want_margin=True)
    np.testing.assert_array_equal(p2f_dup, p2f)
    assert (margin_dup[near_only] == 0).all()  # and the margin mask knows these pixels are ties


def test_face_order_is_irrelevant():
    verts, faces = terrain_mesh(20, 1.0, seed=6, crowns=True)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    cam = nadir(30.0, f=110.0, W=120, H=90, cx=2.5, cy=-1.5)
    ref, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
    perm = np.random.default_rng(0).permutation(len(faces))
    got = ora.rasterize(v32, np.ascontiguousarray(faces[perm]), cam)
    back = np.where(got >= 0, perm[np.maximum(got, 0)], -1)
    same = back == ref
    assert same[margin > 0].all()  # only exact ties may resolve differently when the IDs change
    assert same.mean() > 0.999 and (ref >= 0).mean() > 0.3


@pytest.mark.parametrize("znear", [0.5, 2.0])
def test_near_plane_clipping_removes_only_what_is_behind_the_plane(znear):
    """C5: a long triangle that starts behind the camera and runs away from it.  With the plane at znear its pixels
    are those of the unclipped part in front; moving the plane forward can only remove pixels."""
    verts = np.array([[-0.6, -3.0, 1.0], [0.6, -3.0, 1.0], [0.0, 6.0, -2.5]], dtype=np.float32)
    faces = np.array([[0, 1, 2]], dtype=np.int32)
    # camera at the origin looking along +y, 20 degrees down: camera axes as columns of cam_to_world
    a = np.deg2rad(20.0)
    T = np.eye(4)
    T[:3, 0] = [1, 0, 0]
    T[:3, 1] = [0, -np.sin(a), -np.cos(a)]
    T[:3, 2] = [0, np.cos(a), -np.sin(a)]
    cam = ora.make_camera(T, 80.0, 0, 0, 96, 72, znear=znear)
    X, Y, invz, valid = ora.project(verts, cam)
    assert valid.sum() in (1, 2) and not valid.all()  # the triangle really crosses the plane
    p2f, depth, _ = ora.rasterize(verts, faces, cam, want_depth=True)
    hit = p2f == 0
    assert hit.sum() > 20
    assert (depth[hit] <= 1.0 / znear * (1 + 1e-5)).all()  # nothing nearer than the plane was drawn (depth = 1/z)
    further = ora.rasterize(verts, faces, ora.make_camera(T, 80.0, 0, 0, 96, 72, znear=2 * znear)) == 0
    assert not (further & ~hit).any()
    assert further.sum() < hit.sum()


def test_principal_point_moves_the_picture():
    """cx, cy are offsets from the image centre (cameras.py:75-76): +8 px in cx moves everything 8 columns right."""
    verts, faces = quad(half=1.0)
    a = ora.rasterize(verts, faces, nadir(8.0)) >= 0
    b = ora.rasterize(verts, faces, nadir(8.0, cx=8.0, cy=-5.0)) >= 0
    assert a.sum() > 100 and abs(int(a.sum()) - int(b.sum())) <= 2
    ia, ja = np.nonzero(a)
    ib, jb = np.nonzero(b)
    assert abs((jb.mean() - ja.mean()) - 8.0) < 0.05 and abs((ib.mean() - ia.mean()) + 5.0) < 0.05
