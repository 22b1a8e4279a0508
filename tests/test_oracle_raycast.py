"""Cross-check of the oracle's geometry by a second, independent oracle.

``oracle_raster.c`` (the contract the CUDA kernels are held to) is compared, pixel by pixel, with
``oracle_raycast.c``: float64 ray casting through every pixel centre of the reference's pinhole model
(Moeller-Trumbore in camera space; no snapping, no edge functions, no fill rule, no interpolated depth).  The two
must name the SAME face on every pixel that is

  * edge-safe: no silhouette / shared edge within 2^-7 px of the pixel centre (the ray caster's four corner rays
    agree with its centre ray), and
  * depth-safe: relative 1/z margin between the two nearest faces > 1e-5 in BOTH oracles.

That pins orientation, principal-point sign, row/column conventions, depth order and near-plane behaviour of the
rasterization contract independently of the contract's own arithmetic.  Scenes: the reference's plane fixture
(utils/test_utils.py:69-104) under the 257-px camera of tests/test_derived_cameras.py:339-415, the golden scene,
triangle soups, the synthetic terrain, and the cube / cylinder / cone scene of utils/example_data.py:9-112 seen by
oblique cameras like examples/concept_figure.ipynb.  CPU only.
"""
import numpy as np
import pytest

from geograypher_b200 import synthetic as syn
from geograypher_b200.synthetic import grid_faces
from oracle import oracle as ora
from test_oracle_reference_pins import downward_view, plane_mesh

EPS_EDGE = 2.0**-7
EPS_DEPTH = 1e-5


def _compare(v32, faces, cam, label, min_safe=0.9, min_hit=0.2):
    ids, _, margin = ora.rasterize(v32, faces, cam, want_depth=True, want_margin=True)
    rc_ids, edge_safe, rc_margin = ora.raycast(v32, faces, cam, eps_edge=EPS_EDGE)
    safe = edge_safe & (rc_margin > EPS_DEPTH) & (margin > EPS_DEPTH)
    bad = safe & (ids != rc_ids)
    assert not bad.any(), (f"{label}: {int(bad.sum())} safe pixels differ (first at {np.argwhere(bad)[0]}: "
                           f"raster={ids[bad][0]}, raycast={rc_ids[bad][0]})")
    assert safe.mean() >= min_safe, f"{label}: only {safe.mean():.3f} of the pixels are comparable"
    assert (rc_ids >= 0).mean() >= min_hit, f"{label}: the view barely sees the mesh"
    # the unsafe pixels are what they claim to be: where the two disagree, the other answer is the runner-up or an
    # edge neighbour -- never a face from elsewhere.  Checked loosely: disagreements are a small minority.
    assert (ids != rc_ids).mean() < 0.05
    return safe.mean(), int((ids != rc_ids).sum())


@pytest.mark.parametrize("scale", [1.0, 0.7])
def test_plane_scene(scale):
    verts, faces = plane_mesh()
    sensor = 257
    cam = ora.make_camera(downward_view(4, 100, sensor), 100, 0, 0, sensor, sensor, scale)
    _compare(verts.astype(np.float32), faces, cam, f"plane x{scale}", min_safe=0.85)


def test_plane_scene_with_principal_point_and_tilt():
    """cx, cy != 0 and a camera that is rolled, pitched and yawed: sign and axis conventions of the pinhole model."""
    verts, faces = plane_mesh()
    T = downward_view(4, 100, 200)
    a, b, c = np.deg2rad([17.0, -11.0, 29.0])
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    T[:3, :3] = T[:3, :3] @ Rz @ Ry @ Rx
    cam = ora.make_camera(T, 130.0, 7.25, -4.5, 240, 180)
    _compare(verts.astype(np.float32), faces, cam, "plane tilted", min_safe=0.85)


def test_golden_scene(golden_scene):
    g = golden_scene
    f, cx, cy, W, H = g["intrinsics"]
    v32 = (g["verts"] - g["origin"]).astype(np.float32)
    for k, T in enumerate(g["c2ws"]):
        cam = ora.make_camera(T, f, cx, cy, int(W), int(H), origin=g["origin"])
        _compare(v32, g["faces"], cam, f"golden view {k}", min_safe=0.8)


@pytest.mark.parametrize("name,views", [("tiny", 4), ("c1", 2)])
def test_terrain_surveys(name, views):
    verts, faces, c2ws, cfg = syn.make_survey(name, views)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    W, H = cfg.image_size
    for k, T in enumerate(c2ws):
        cam = ora.make_camera(T, cfg.f, cfg.cx, cfg.cy, W, H, origin=origin)
        _compare(v32, faces, cam, f"{name} view {k}", min_safe=0.8)


@pytest.mark.parametrize("znear", [1e-3, 0.5])
@pytest.mark.parametrize("seed,n_faces,W,H", [(0, 300, 160, 120), (1, 800, 200, 150)])
def test_triangle_soup(seed, n_faces, W, H, znear):
    """The soup of tests/test_gpu_parity.py: slivers, huge faces, faces through the near plane and behind the camera.
    At the default near plane (1e-3) faces tens of metres wide are cut so close to the camera that their clipped
    vertices project millions of pixels away: contract C6 (guard-band clipping) keeps the snapped coordinates exact
    there -- before it existed the two oracles disagreed on a third of this scene's pixels."""
    rng = np.random.default_rng(seed)
    centers = rng.uniform([-30, -30, -5], [30, 30, 60], size=(n_faces, 1, 3))
    size = rng.choice([0.3, 3.0, 40.0], size=(n_faces, 1, 1), p=[0.4, 0.4, 0.2])
    verts = (centers + rng.normal(0, 1, size=(n_faces, 3, 3)) * size).reshape(-1, 3)
    faces = np.arange(3 * n_faces, dtype=np.int32).reshape(-1, 3)
    c2w = np.eye(4)
    c2w[:3, 3] = [0.5, -0.25, -20.0]
    cam = ora.make_camera(c2w, 0.8 * W, 3.0, -2.0, W, H, znear=znear)
    _compare(verts.astype(np.float32), faces, cam, f"soup {seed} znear {znear}", min_safe=0.6)


def test_guard_band_one_huge_face_through_the_camera_plane():
    """One 400 m face that passes a metre under the camera and reaches far behind it, over a small ground patch: its
    near-plane cut projects beyond +-2^20 px, the guard band (C6) clips it, and it covers exactly the pixels the ray
    caster says it covers."""
    ground, gfaces = plane_mesh()
    ground = ground * 10.0  # 40 m x 40 m at z = 0
    big = np.array([[-150.0, -200.0, 7.9], [150.0, -200.0, 7.9], [0.0, 200.0, 8.3]])  # tilted sheet above the ground
    verts = np.concatenate([ground, big])
    faces = np.concatenate([gfaces, [[len(ground), len(ground) + 1, len(ground) + 2]]]).astype(np.int32)
    T = downward_view(4, 100, 200)
    T[:3, 3] = [1.0, -2.0, 9.0]  # just above the sheet, which reaches ~190 m in front of AND behind the camera plane
    a = np.deg2rad(70.0)
    T[:3, :3] = T[:3, :3] @ np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    cam = ora.make_camera(T, 150.0, 0.0, 0.0, 240, 180)
    v32 = verts.astype(np.float32)
    ids = ora.rasterize(v32, faces, cam)
    assert 0.5 < (ids == len(faces) - 1).mean() < 0.98  # the sheet fills the view up to its horizon
    _compare(v32, faces, cam, "huge face", min_safe=0.9, min_hit=0.5)


# ---- the cube / cylinder / cone scene of the reference's concept figure --------------------------------------
def _box(x, y, size):
    """12 triangles (reference: pv.Box(..., quads=False), example_data.py:45-52)."""
    h = size / 2.0
    c = np.array([[x - h, y - h, 0], [x + h, y - h, 0], [x + h, y + h, 0], [x - h, y + h, 0],
                  [x - h, y - h, size], [x + h, y - h, size], [x + h, y + h, size], [x - h, y + h, size]])
    quads = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
    tris = [t for q in quads for t in ((q[0], q[1], q[2]), (q[0], q[2], q[3]))]
    return c, np.array(tris)


def _revolved(x, y, radius, res, apex_up):
    """Cylinder (apex_up None; pv.Cylinder(resolution=10), :57-60) or cone with its apex at z = 0 pointing down
    (pv.Cone(direction=(0,0,-1), resolution=12), :68-74): unit height centred at z = 0.5."""
    ang = 2 * np.pi * np.arange(res) / res
    ring = np.stack([x + radius * np.cos(ang), y + radius * np.sin(ang)], axis=1)
    if apex_up is None:  # cylinder: two rings + two cap centres
        v = np.concatenate([np.c_[ring, np.zeros(res)], np.c_[ring, np.ones(res)], [[x, y, 0.0], [x, y, 1.0]]])
        t = []
        for k in range(res):
            n = (k + 1) % res
            t += [(k, n, res + n), (k, res + n, res + k), (2 * res, n, k), (2 * res + 1, res + k, res + n)]
        return v, np.array(t)
    v = np.concatenate([np.c_[ring, np.ones(res)], [[x, y, 0.0], [x, y, 1.0]]])  # base ring at z = 1, apex at z = 0
    t = []
    for k in range(res):
        n = (k + 1) % res
        t += [(res, n, k), (res + 1, k, n)]
    return v, np.array(t)


def concept_scene():
    """Boxes, cylinders and cones on a ground grid (utils/example_data.py:30-112; the ground is a regular grid
    instead of a Delaunay triangulation, which needs VTK)."""
    parts = [_box(-2.0, -1.0, 1 / np.sqrt(2.0)), _box(2.5, 1.5, 1 / np.sqrt(2.0)),
             _revolved(0.0, 2.0, 0.5, 10, None), _revolved(-3.0, 2.5, 0.5, 10, None),
             _revolved(1.0, -2.0, 0.5, 12, True), _revolved(3.0, -1.0, 0.5, 12, True)]
    n = 40
    ax = np.linspace(-5, 5, n + 1)
    gx, gy = np.meshgrid(ax, ax, indexing="xy")
    parts.append((np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1), grid_faces(n, n)))
    verts, faces, off = [], [], 0
    for v, t in parts:
        verts.append(v)
        faces.append(t + off)
        off += len(v)
    return np.concatenate(verts), np.concatenate(faces).astype(np.int32)


def _look_at(eye, target):
    """cam_to_world of a camera at ``eye`` looking at ``target``: +Z forward, +Y down (cameras/cameras.py:446-477)."""
    eye, target = np.asarray(eye, float), np.asarray(target, float)
    z = (target - eye) / np.linalg.norm(target - eye)
    x = np.cross(z, [0, 0, 1.0])
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, eye
    return T


def test_concept_figure_scene():
    """Occlusion-rich: objects hide the ground and each other from five oblique cameras (the notebook's camera has
    f = 4000 at 3000 x 2200; same field of view at a quarter of the resolution here)."""
    verts, faces = concept_scene()
    v32 = verts.astype(np.float32)
    eyes = [(-7, -7, 6), (7, -6, 5), (0, -9, 4), (6, 7, 7), (-8, 3, 3)]
    hidden_ground = 0
    for k, eye in enumerate(eyes):
        cam = ora.make_camera(_look_at(eye, (0, 0, 0.3)), 1000.0, 12.0, -7.0, 750, 550)
        _compare(v32, faces, cam, f"concept view {k}", min_safe=0.9, min_hit=0.5)
        ids = ora.rasterize(v32, faces, cam)
        hidden_ground += int((ids >= 0).sum() and (ids[ids >= 0] < len(faces) - 3200).sum())
    assert hidden_ground > 5000  # the objects do cover pixels (and therefore hide ground behind them)
