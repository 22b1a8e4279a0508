"""Deterministic synthetic surveys of the shapes BASELINE.json names.

Everything is seeded with ``numpy.random.default_rng(seed)``; nothing here touches the GPU except
``softmax_predictions_device`` (torch), which bench.py uses to fill device-resident prediction buffers
outside the timed region.  Meshes are produced directly in the camera set's local frame
(``local_to_epsg_4978_transform = I``), like the reference's own fixtures do
(/root/reference/geograypher/utils/test_utils.py:35-38).
"""
from __future__ import annotations

import dataclasses

import numpy as np


# --------------------------------------------------------------------------------------------------
# Named configurations (BASELINE.json "configs")
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class SurveyConfig:
    name: str
    n_cells: int  # grid cells per side -> 2*n^2 faces
    cell_m: float
    mesh_seed: int
    crowns: bool
    cam_grid: tuple  # (lines, per_line)
    cam_spacing: tuple  # (along-line spacing m, between-line spacing m)
    image_size: tuple  # (W, H)
    f: float
    cx: float
    cy: float
    altitude: float
    cam_seed: int
    n_classes: int
    rig: bool = False

    @property
    def n_faces(self):
        return 2 * self.n_cells * self.n_cells

    @property
    def n_cameras(self):
        n = self.cam_grid[0] * self.cam_grid[1]
        return n * 5 if self.rig else n


CONFIGS = {
    # c1: 100 352 faces, 10 pinhole cameras 1024x768, 5 classes
    "c1": SurveyConfig("c1", 224, 1.0, 1, False, (2, 5), (40.0, 50.0), (1024, 768), 800.0, 0.0, 0.0, 60.0, 11, 5),
    # c2/c3/c4: 2 000 000 faces, 500 Metashape-calibrated cameras 5472x3648 (intrinsics from
    # /root/reference/tests/test_derived_cameras.py:31-33), 10 classes
    "c2": SurveyConfig("c2", 1000, 1.0, 2, True, (20, 25), (40.0, 50.0), (5472, 3648), 3705.4729, 11.6739, -27.7498, 120.0, 12, 10),
    # c5: 19 996 488 faces, 400 stations x 5-camera rig, 8192x5460, 10 classes (one-hot votes)
    "c5": SurveyConfig("c5", 3162, 0.5, 5, True, (20, 20), (75.0, 75.0), (8192, 5460), 5500.0, 0.0, 0.0, 150.0, 15, 10, True),
    # tiny scene for smoke / unit tests
    "tiny": SurveyConfig("tiny", 48, 1.0, 3, True, (2, 2), (12.0, 12.0), (160, 120), 110.0, 1.5, -2.25, 40.0, 13, 4),
}


# --------------------------------------------------------------------------------------------------
# Mesh
# --------------------------------------------------------------------------------------------------
def terrain_mesh(n_cells: int, cell_m: float = 1.0, seed: int = 0, crowns: bool = True):
    """Regular-grid terrain: (n+1)^2 vertices over [0, n*cell]^2, two CCW (seen from +Z) triangles per
    cell with a fixed diagonal.  z = 4 octaves of sin*cos relief (about +-15 m) + Gaussian "tree crowns"
    (4e-3 per m^2, height U[5,30] m, sigma U[1.5,4] m).  Returns (verts float64 (V,3), faces int32 (F,3))."""
    rng = np.random.default_rng(seed)
    n = n_cells
    ax = np.arange(n + 1, dtype=np.float64) * cell_m
    x, y = np.meshgrid(ax, ax, indexing="xy")  # row r <-> y, col c <-> x
    z = np.zeros_like(x)
    amp = [8.0, 4.0, 2.0, 1.0]
    for k in range(4):
        fx, fy = rng.uniform(0.5, 1.5, size=2) * (2 * np.pi / (400.0 / 2**k))
        ph = rng.uniform(0, 2 * np.pi, size=2)
        z += amp[k] * np.sin(fx * x + ph[0]) * np.cos(fy * y + ph[1])
    if crowns:
        extent = n * cell_m
        n_crowns = int(round(4e-3 * extent * extent))
        cxy = rng.uniform(0, extent, size=(n_crowns, 2))
        hh = rng.uniform(5.0, 30.0, size=n_crowns)
        sg = rng.uniform(1.5, 4.0, size=n_crowns)
        for (px, py), h, s in zip(cxy, hh, sg):
            r = 4.0 * s
            c0 = max(int(np.floor((px - r) / cell_m)), 0)
            c1 = min(int(np.ceil((px + r) / cell_m)), n)
            r0 = max(int(np.floor((py - r) / cell_m)), 0)
            r1 = min(int(np.ceil((py + r) / cell_m)), n)
            if c1 < c0 or r1 < r0:
                continue
            xs = x[r0 : r1 + 1, c0 : c1 + 1] - px
            ys = y[r0 : r1 + 1, c0 : c1 + 1] - py
            z[r0 : r1 + 1, c0 : c1 + 1] += h * np.exp(-(xs * xs + ys * ys) / (2 * s * s))
    verts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    faces = grid_faces(n, n)
    return verts, faces


def grid_faces(n_rows: int, n_cols: int) -> np.ndarray:
    """Two CCW triangles per cell of an (n_rows+1) x (n_cols+1) vertex grid (vertex k = r*(n_cols+1)+c)."""
    r, c = np.meshgrid(np.arange(n_rows), np.arange(n_cols), indexing="ij")
    v00 = (r * (n_cols + 1) + c).ravel()
    v01 = v00 + 1
    v10 = v00 + (n_cols + 1)
    v11 = v10 + 1
    faces = np.empty((2 * n_rows * n_cols, 3), dtype=np.int32)
    faces[0::2] = np.stack([v00, v01, v11], axis=1)
    faces[1::2] = np.stack([v00, v11, v10], axis=1)
    return faces


# --------------------------------------------------------------------------------------------------
# Cameras
# --------------------------------------------------------------------------------------------------
def _rot(axis: str, deg: float) -> np.ndarray:
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    if axis == "x":
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


NADIR = np.diag([1.0, -1.0, -1.0])  # the reference's downward_view rotation (test_utils.py:59-66)


def lawnmower_cameras(cfg: SurveyConfig, extent_m: float, jitter_deg: float = 2.0):
    """cam_to_world 4x4 list: grid of nadir cameras at altitude ``cfg.altitude`` above z=0, yaw flipping
    0/180 deg on alternate lines, N(0, jitter) roll/pitch/yaw.  With ``cfg.rig`` every station carries a
    5-camera rig (nadir + 4 obliques pitched 30 deg at yaw 0/90/180/270) composed as
    ``c2w_station @ rig_transform`` (rig_cameras.py:73-77 of the reference)."""
    rng = np.random.default_rng(cfg.cam_seed)
    lines, per_line = cfg.cam_grid
    dx, dy = cfg.cam_spacing
    x0 = 0.5 * (extent_m - (per_line - 1) * dx)
    y0 = 0.5 * (extent_m - (lines - 1) * dy)
    rig = [np.eye(3)]
    if cfg.rig:
        rig += [_rot("z", yaw) @ _rot("x", 30.0) for yaw in (0.0, 90.0, 180.0, 270.0)]
    out = []
    for li in range(lines):
        order = range(per_line) if li % 2 == 0 else range(per_line - 1, -1, -1)
        for k in order:
            jr, jp, jy = rng.normal(0.0, jitter_deg, size=3)
            yaw = (0.0 if li % 2 == 0 else 180.0) + jy
            # world <- camera: yaw about world Z, then nadir flip, then small roll/pitch in the camera frame
            R = _rot("z", yaw) @ NADIR @ _rot("x", jp) @ _rot("y", jr)
            for Rr in rig:
                T = np.eye(4)
                T[:3, :3] = R @ Rr
                T[:3, 3] = [x0 + k * dx, y0 + li * dy, cfg.altitude]
                out.append(T)
    return out


def make_survey(name: str, max_cameras: int | None = None):
    """(verts f64, faces i32, [cam_to_world], cfg) for a named configuration."""
    cfg = CONFIGS[name]
    verts, faces = terrain_mesh(cfg.n_cells, cfg.cell_m, cfg.mesh_seed, cfg.crowns)
    cams = lawnmower_cameras(cfg, cfg.n_cells * cfg.cell_m)
    if max_cameras is not None:
        cams = cams[:max_cameras]
    return verts, faces, cams, cfg


# --------------------------------------------------------------------------------------------------
# Predictions
# --------------------------------------------------------------------------------------------------
def class_index_image(cam_index: int, H: int, W: int, n_classes: int, block: int = 32,
                      ignore_frac: float = 0.02, seed: int = 0) -> np.ndarray:
    """uint8 (H, W) class-index image: blocky ``block``-px regions, ``ignore_frac`` of pixels = 255."""
    rng = np.random.default_rng(seed * 1_000_003 + cam_index)
    bh, bw = -(-H // block), -(-W // block)
    blocks = rng.integers(0, n_classes, size=(bh, bw), dtype=np.uint8)
    img = np.repeat(np.repeat(blocks, block, axis=0), block, axis=1)[:H, :W].copy()
    if ignore_frac > 0:
        img[rng.random((H, W)) < ignore_frac] = 255
    return img


def softmax_predictions(cam_index: int, H: int, W: int, n_classes: int, grid=(43, 64)) -> np.ndarray:
    """(H, W, C) float32 softmax of smooth logits (bilinear up-sampling of a small N(0, 1.5) grid per
    class; seed = 1000 + camera index).  NumPy version for small cases and tests."""
    rng = np.random.default_rng(1000 + cam_index)
    g = rng.normal(0.0, 1.5, size=(n_classes,) + tuple(grid)).astype(np.float32)
    yy = np.linspace(0, grid[0] - 1, H, dtype=np.float32)
    xx = np.linspace(0, grid[1] - 1, W, dtype=np.float32)
    y0 = np.clip(np.floor(yy).astype(int), 0, grid[0] - 2)
    x0 = np.clip(np.floor(xx).astype(int), 0, grid[1] - 2)
    fy = (yy - y0)[None, :, None]
    fx = (xx - x0)[None, None, :]
    a = g[:, y0][:, :, x0]
    b = g[:, y0][:, :, x0 + 1]
    c = g[:, y0 + 1][:, :, x0]
    d = g[:, y0 + 1][:, :, x0 + 1]
    logits = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    logits -= logits.max(axis=0, keepdims=True)
    e = np.exp(logits)
    p = e / e.sum(axis=0, keepdims=True)
    return np.ascontiguousarray(np.moveaxis(p, 0, -1).astype(np.float32))


def softmax_predictions_device(cam_index: int, H: int, W: int, n_classes: int, device, out=None, grid=(43, 64)):
    """Same construction on a torch device (values differ from the NumPy version in the last bits; used by
    bench.py, where predictions are inputs and never compared across generators)."""
    import torch

    gen = torch.Generator(device="cpu")
    gen.manual_seed(1000 + cam_index)
    g = (torch.randn((1, n_classes) + tuple(grid), generator=gen) * 1.5).to(device)
    logits = torch.nn.functional.interpolate(g, size=(H, W), mode="bilinear", align_corners=True)[0]
    p = torch.softmax(logits, dim=0).permute(1, 2, 0)
    if out is None:
        return p.contiguous()
    out.copy_(p)
    return out


def voronoi_face_labels(verts, faces, n_sites=200, n_classes=10, nan_frac=0.2, seed=4) -> np.ndarray:
    """(F, 1) float64 per-face label = class of the nearest of ``n_sites`` random sites (by face centroid,
    x/y only); ``nan_frac`` of the sites carry NaN (unlabelled ground)."""
    rng = np.random.default_rng(seed)
    cen = verts[faces].mean(axis=1)[:, :2]
    lo, hi = cen.min(axis=0), cen.max(axis=0)
    sites = rng.uniform(lo, hi, size=(n_sites, 2))
    cls = rng.integers(0, n_classes, size=n_sites).astype(np.float64)
    cls[rng.random(n_sites) < nan_frac] = np.nan
    from scipy.spatial import cKDTree

    _, nearest = cKDTree(sites).query(cen)
    return cls[nearest][:, None]
