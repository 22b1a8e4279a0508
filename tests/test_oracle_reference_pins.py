"""The reference's own known-answer tests for the path, re-expressed without pyvista (SURVEY.md
Appendix B), run against the oracle.  CPU only.

* tests/test_derived_meshes.py:23-76   (render_flat on a 201x201-vertex plane, nadir camera f=100, 200x200)
* tests/test_derived_cameras.py:339-415 (pix2face dtype / shape / range at scales 0.5, 0.7, 0.9, 1.0)
"""
from itertools import product

import numpy as np
import pytest

from geograypher_b200.synthetic import grid_faces
from oracle import oracle as ora

N = 201


def plane_mesh():
    """Vertex k = r*N + c at x = -2 + 0.02 c, y = -2 + 0.02 r, z = 0 (utils/test_utils.py:69-104)."""
    ax = -2.0 + 0.02 * np.arange(N)
    x, y = np.meshgrid(ax, ax, indexing="xy")
    verts = np.stack([x.ravel(), y.ravel(), np.zeros(N * N)], axis=1)
    return verts, grid_faces(N - 1, N - 1)


def downward_view(scene_width, focal, sensor_width):
    """utils/test_utils.py:42-66."""
    return np.array(
        [[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, scene_width * focal / sensor_width], [0, 0, 0, 1.0]]
    )


def pixel_idx(vector, i, j, stride, color, buffer=0):
    """utils/test_utils.py:132-156."""
    spread = range(-buffer, 2 + buffer)
    for di, dj in product(spread, spread):
        vector[(stride - i - di) * stride + (j + dj)] = color


def test_perspective_camera_known_answer():
    fill_pixels = np.array([[10, 20], [15, 190], [195, 5], [50, 100], [150, 120]])
    empty_pixels = np.array([[30, 40], [160, 180], [120, 40], [100, 150], [180, 100]])
    verts, faces = plane_mesh()
    colors = np.full((N * N, 3), 80, dtype=np.uint8)
    for p in fill_pixels:
        pixel_idx(colors, *p, stride=N, color=[255, 0, 0], buffer=1)
    face_tex = ora.vert_to_face_texture_mean(colors, faces)

    cam = ora.make_camera(downward_view(4, 100, 200), 100, 0, 0, 200, 200)
    p2f = ora.rasterize(verts.astype(np.float32), faces, cam)
    render = ora.render_flat_gather(p2f, face_tex)

    assert render.shape == (200, 200, 3)
    assert np.allclose(render[fill_pixels[:, 0], fill_pixels[:, 1]], [255, 0, 0])
    assert np.allclose(render[empty_pixels[:, 0], empty_pixels[:, 1]], [80, 80, 80])
    # stronger than the reference asks: the plane fills the frame, one cell per pixel
    assert (p2f >= 0).all()
    cell = p2f // 2
    i, j = np.meshgrid(np.arange(200), np.arange(200), indexing="ij")
    np.testing.assert_array_equal(cell, (199 - i) * 200 + j)


@pytest.mark.parametrize("render_img_scale", [0.5, 0.7, 0.9, 1.0])
def test_pix2face_structural_pins(render_img_scale):
    sensor = 2**8 + 1
    verts, faces = plane_mesh()
    cam = ora.make_camera(downward_view(4, 100, sensor), 100, 0, 0, sensor, sensor, render_img_scale)
    ideal = ora.rasterize(verts.astype(np.float32), faces, cam)
    scaled = int(sensor * render_img_scale)
    assert ideal.dtype == np.int64
    assert ideal.shape == (scaled, scaled)
    assert ideal.min() >= -1
    assert ideal.max() < len(faces)
    assert ideal.max() > 0.95 * len(faces)
    for corner in product([slice(None, 10), slice(-10, None)], repeat=2):
        assert len(np.unique(ideal[corner])) > 1
