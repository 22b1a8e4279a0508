"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of geograypher's multiview-projection hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; ``geograypher_b200`` never does.

Two halves:

* visibility (``pix2face``): C restatement in ``oracle_raster.c`` (see its header for the contract and for
  why sub-pixel parity with VTK / PyTorch3D is UNPINNED), called through ctypes here;
* everything after visibility -- ``project_images``, ``aggregate_projected_images``, the one-hot-vote
  variant, ``render_flat``'s gather, ``save_renders``' uint8 cast, ``inds_to_one_hot`` and
  ``find_argmax_nonzero_value`` -- is in-repo NumPy in the reference and is restated literally below.  These
  restatements are PINNED: ``tests/golden/make_golden.py`` runs the reference's own code (imported from
  /root/reference with its missing third-party imports stubbed) on the same inputs and
  ``tests/test_oracle_golden.py`` checks equality.

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

SUBPIX = 256
DEFAULT_ZNEAR = 1e-3


class OraCamera(ctypes.Structure):
    _fields_ = [
        ("m", ctypes.c_float * 12),
        ("f", ctypes.c_float),
        ("px", ctypes.c_float),
        ("py", ctypes.c_float),
        ("W", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("znear", ctypes.c_float),
    ]


def build(force: bool = False) -> Path:
    """Compile oracle_raster.c with the committed Makefile (gcc, -ffp-contract=off, OpenMP)."""
    so = _HERE / "liboracle_raster.so"
    stale = force
    for name in ("oracle_raster", "oracle_raycast"):
        lib, src = _HERE / f"lib{name}.so", _HERE / f"{name}.c"
        stale = stale or not lib.exists() or lib.stat().st_mtime < src.stat().st_mtime
    if stale:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = build()
        lib = ctypes.CDLL(str(so))
        lib.ora_project.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(OraCamera)] + [
            ctypes.c_void_p
        ] * 4
        lib.ora_project.restype = None
        lib.ora_rasterize.argtypes = [
            ctypes.c_void_p,
            ctypes.c_int64,
            ctypes.c_void_p,
            ctypes.c_int64,
            ctypes.POINTER(OraCamera),
            ctypes.c_void_p,
            ctypes.c_void_p,
            ctypes.c_void_p,
            ctypes.c_int,
        ]
        lib.ora_rasterize.restype = None
        lib.ora_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().ora_num_threads())


# --------------------------------------------------------------------------------------------------
# Camera model (cameras/cameras.py:56-102, 136-152, 179-200; derived_meshes.py:739-794)
# --------------------------------------------------------------------------------------------------
def scaled_image_size(image_height: int, image_width: int, image_scale: float = 1.0):
    """(h, w) = (int(H*s), int(W*s)) -- cameras/cameras.py:197-200."""
    return int(image_height * image_scale), int(image_width * image_scale)


def make_camera(
    cam_to_world,
    f,
    cx,
    cy,
    image_width,
    image_height,
    render_img_scale=1.0,
    origin=None,
    znear=DEFAULT_ZNEAR,
):
    """Float32 camera record of the rasterization contract from geograypher camera parameters.

    world_to_cam = inv(cam_to_world) (cameras.py:84); pinhole with the principal point measured from the
    image centre (cameras.py:75-76, derived_meshes.py:772-780).  ``render_img_scale`` keeps the vertical
    field of view like the pyvista path does (cameras.py:472, meshes.py:1801, 1820-1822):
    f' = f*h'/H, principal point (w'/2 + cx*h'/H, h'/2 + cy*h'/H).  ``origin`` is the float64 point that was
    subtracted from the mesh vertices before they were rounded to float32; it is folded into the
    translation in float64.
    """
    c2w = np.asarray(cam_to_world, dtype=np.float64)
    w2c = np.linalg.inv(c2w)
    m = w2c[:3, :4].copy()
    if origin is not None:
        m[:, 3] = m[:, 3] + m[:, :3] @ np.asarray(origin, dtype=np.float64)
    h, w = scaled_image_size(image_height, image_width, render_img_scale)
    s = h / float(image_height)
    cam = OraCamera()
    m32 = m.astype(np.float32).reshape(-1)
    for k in range(12):
        cam.m[k] = float(m32[k])
    cam.f = np.float32(f * s)
    cam.px = np.float32(w / 2.0 + cx * s)
    cam.py = np.float32(h / 2.0 + cy * s)
    cam.W = w
    cam.H = h
    cam.znear = np.float32(znear)
    return cam


def camera_to_array(cam: OraCamera) -> np.ndarray:
    """(18,) float32 view [m0..m11, f, px, py, W, H, znear] -- handy for tests that feed the same camera
    to the CUDA library."""
    return np.array(
        list(cam.m) + [cam.f, cam.px, cam.py, float(cam.W), float(cam.H), cam.znear],
        dtype=np.float32,
    )


# --------------------------------------------------------------------------------------------------
# Visibility
# --------------------------------------------------------------------------------------------------
def project(verts32: np.ndarray, cam: OraCamera):
    """Stage 1 of the contract: returns (X, Y) int32 fixed point (1/256 px), invz float32, valid bool."""
    verts32 = np.ascontiguousarray(verts32, dtype=np.float32)
    V = verts32.shape[0]
    X = np.empty(V, np.int32)
    Y = np.empty(V, np.int32)
    invz = np.empty(V, np.float32)
    valid = np.empty(V, np.uint8)
    _lib().ora_project(
        verts32.ctypes.data, V, ctypes.byref(cam), X.ctypes.data, Y.ctypes.data, invz.ctypes.data,
        valid.ctypes.data,
    )
    return X, Y, invz, valid.astype(bool)


def rasterize(verts32, faces, cam: OraCamera, want_depth=False, want_margin=False, nthreads=0):
    """pix2face for one view: (H, W) int64 face IDs, -1 where no face (meshes.py:1712-1718).

    Returns ``pix2face`` or ``(pix2face, depth_w, margin)`` (None for the parts not requested).
    """
    verts32 = np.ascontiguousarray(verts32, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    H, W = int(cam.H), int(cam.W)
    out = np.empty((H, W), np.int32)
    depth = np.empty((H, W), np.float64) if want_depth else None
    margin = np.empty((H, W), np.float64) if want_margin else None
    _lib().ora_rasterize(
        verts32.ctypes.data,
        verts32.shape[0],
        faces.ctypes.data,
        faces.shape[0],
        ctypes.byref(cam),
        out.ctypes.data,
        depth.ctypes.data if depth is not None else None,
        margin.ctypes.data if margin is not None else None,
        int(nthreads),
    )
    p2f = out.astype(np.int64)
    if want_depth or want_margin:
        return p2f, depth, margin
    return p2f


def resize_render(image, output_shape, order):
    """save_renders' up-sampling (meshes.py:2312-2321): skimage.transform.resize(image, output_shape, order=order) with
    its defaults is scipy.ndimage.zoom(image, factors, order=order, mode="mirror", grid_mode=True) (skimage >= 0.19;
    no anti-aliasing when up-sampling, preserve_range irrelevant for float input, clip leaves NaN alone)."""
    from scipy import ndimage

    image = np.asarray(image, dtype=float)
    factors = [output_shape[0] / image.shape[0], output_shape[1] / image.shape[1]] + [1] * (image.ndim - 2)
    return ndimage.zoom(image, factors, order=order, mode="mirror", grid_mode=True)


_RAYCAST = None


def raycast(verts32, faces, cam: OraCamera, eps_edge: float = 2.0**-7, nthreads=0):
    """The INDEPENDENT second oracle (oracle_raycast.c): per-pixel-centre ray casting in float64 camera space
    (Moeller-Trumbore), no snapping / edge functions / fill rule.  Returns ``(pix2face int64 (H, W), edge_safe bool,
    depth_margin float64)``: ``edge_safe`` is False where a silhouette or shared edge passes within ``eps_edge`` px
    of the pixel centre (the four corner rays of that square disagree with the centre ray)."""
    global _RAYCAST
    if _RAYCAST is None:
        build()
        lib = ctypes.CDLL(str(_HERE / "liboracle_raycast.so"))
        lib.orc_raycast.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.POINTER(OraCamera), ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int]
        lib.orc_raycast.restype = None
        _RAYCAST = lib
    verts32 = np.ascontiguousarray(verts32, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    H, W = int(cam.H), int(cam.W)
    ids = np.empty((H, W), np.int32)
    safe = np.empty((H, W), np.uint8)
    margin = np.empty((H, W), np.float64)
    _RAYCAST.orc_raycast(verts32.ctypes.data, verts32.shape[0], faces.ctypes.data, faces.shape[0], ctypes.byref(cam),
                         float(eps_edge), ids.ctypes.data, safe.ctypes.data, margin.ctypes.data, int(nthreads))
    return ids.astype(np.int64), safe.astype(bool), margin


# --------------------------------------------------------------------------------------------------
# After visibility: literal NumPy of the reference
# --------------------------------------------------------------------------------------------------
def inds_to_one_hot(inds_image, num_classes, ignore_ind=255):
    """predictors/segmentor.py:37-69 -- (h, w) indices -> (h, w, C) bool; ignore / out-of-range -> all False."""
    del ignore_ind  # the reference computes but never uses it once num_classes is given (:53-57)
    one_hot = np.zeros(inds_image.shape[:2] + (num_classes,), dtype=bool)
    for i in range(num_classes):
        one_hot[..., i] = inds_image == i
    return one_hot


def project_image(pix2face, img, n_faces, compat_negative_index=True):
    """One iteration of project_images, meshes.py:1988-2002.

    NumPy fancy assignment: for a face hit by several pixels the LAST pixel in row-major order wins.
    With ``compat_negative_index`` (the reference's behaviour, TODO at meshes.py:2000) background pixels
    (-1) index face n_faces-1; without it they are skipped.
    """
    n_channels = 1 if img.ndim == 2 else img.shape[-1]
    textured_faces = np.full((n_faces, n_channels), fill_value=np.nan)
    flat_img = np.reshape(img, (img.shape[0] * img.shape[1], -1))
    flat_p2f = pix2face.flatten()
    if compat_negative_index:
        textured_faces[flat_p2f] = flat_img
    else:
        keep = flat_p2f >= 0
        textured_faces[flat_p2f[keep]] = flat_img[keep]
    return textured_faces


def aggregate(pix2faces, images, n_faces, compat_negative_index=True, check_null_image=False):
    """aggregate_projected_images, meshes.py:2046-2084 (sum over views with NaN -> 0, count of views that
    saw a face, mean).  Returns (average, projection_counts, summed_projection)."""
    counts = np.zeros(n_faces)
    summed = None
    for p2f, img in zip(pix2faces, images):
        if check_null_image and not np.any(np.isfinite(img)):
            n_channels = 1 if img.ndim == 2 else img.shape[-1]
            proj = np.full((n_faces, n_channels), np.nan)
        else:
            proj = project_image(p2f, img, n_faces, compat_negative_index)
        if summed is None:
            summed = proj.astype(float)
        else:
            summed = np.nansum([summed, proj], axis=0)
        counts += np.any(np.isfinite(proj), axis=1).astype(int)
    summed[counts == 0] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        average = np.divide(summed, np.expand_dims(counts, 1))
    return average, counts, summed


def aggregate_votes(pix2faces, index_images, n_faces, n_classes, compat_negative_index=True):
    """TexturedPhotogrammetryMeshIndexPredictions.aggregate_projected_images, derived_meshes.py:415-550,
    with dense int arrays instead of scipy CSR: every face whose winning pixel holds a finite class index
    votes once for that class.  Returns (average, counts (F,), summed (F, n_classes))."""
    counts = np.zeros(n_faces, dtype=np.int64)
    summed = np.zeros((n_faces, n_classes), dtype=np.int64)
    for p2f, img in zip(pix2faces, index_images):
        if not np.any(np.isfinite(img)):  # check_null_image=True, meshes.py:1995
            continue
        proj = project_image(p2f, img, n_faces, compat_negative_index)
        idx = np.nonzero(np.isfinite(np.squeeze(proj, axis=1)))[0]
        if len(idx) == 0:
            continue
        counts[idx] += 1
        cls = proj[idx, 0].astype(int)
        summed[idx, cls] += 1
    average = np.zeros((n_faces, n_classes))
    seen = counts > 0
    average[seen] = summed[seen] * np.reciprocal(counts[seen].astype(float))[:, None]
    return average, counts, summed


def render_flat_gather(pix2face, face_texture):
    """Per-view body of render_flat, meshes.py:1921-1937: NaN where pix2face == -1."""
    face_texture = np.asarray(face_texture, dtype=float)
    if face_texture.ndim == 1:
        face_texture = face_texture[:, None]
    img_shape = pix2face.shape[:2]
    flat = pix2face.flatten()
    inds = np.where(flat != -1)[0]
    out = np.full((flat.shape[0], face_texture.shape[1]), fill_value=np.nan)
    out[inds] = face_texture[flat[inds]]
    return out.reshape(img_shape + (face_texture.shape[1],))


def cast_render_to_uint8(rendered, null_value=0):
    """save_renders' cast rule, meshes.py:2323-2334 (NULL_TEXTURE_INT_VALUE = 0, constants.py:27)."""
    rendered = np.array(rendered, dtype=float, copy=True)
    with np.errstate(invalid="ignore"):
        mask = np.logical_or.reduce(
            [rendered < 0, rendered > 255, np.logical_not(np.isfinite(rendered))]
        )
    rendered[mask] = null_value
    return np.squeeze(rendered.astype(np.uint8))


def vert_to_face_texture_mean(vertex_texture, faces):
    """Non-discrete branch of vert_to_face_texture, meshes.py:985-987: mean of the 3 vertex rows."""
    return np.mean(np.asarray(vertex_texture, dtype=float)[faces], axis=1)


def find_argmax_nonzero_value(array, keepdims=False, axis=1):
    """utils/indexing.py:9-32."""
    argmax = np.argmax(array, axis=axis, keepdims=keepdims).astype(float)
    zero_sum_mask = np.sum(array, axis=axis) == 0
    infinite_mask = np.any(~np.isfinite(array), axis=axis)
    argmax[np.logical_or(zero_sum_mask, infinite_mask)] = np.nan
    return argmax


# --------------------------------------------------------------------------------------------------
# Whole path, for bench.py's CPU baseline and for tests
# --------------------------------------------------------------------------------------------------
def pix2face_set(verts32, faces, cams, nthreads=0):
    """(n, H, W) int64 -- meshes.py:1735-1749 (per-camera recursion + np.stack)."""
    return np.stack([rasterize(verts32, faces, c, nthreads=nthreads) for c in cams], axis=0)


# --------------------------------------------------------------------------------------------------
# Lens distortion: literal restatements of the reference, for the warp parity tests
# --------------------------------------------------------------------------------------------------
def metashape_ideal_to_warped(xpix, ypix, f, cx, cy, image_width, image_height, k1=0.0, k2=0.0, k3=0.0, k4=0.0,
                              p1=0.0, p2=0.0, b1=0.0, b2=0.0):
    """MetashapeCameraSet.ideal_to_warped, cameras/derived_cameras.py:163-208."""
    x = (xpix - image_width / 2.0) / f
    y = (ypix - image_height / 2.0) / f
    r = np.sqrt(x**2 + y**2)
    xd = x * (1 + k1 * r**2 + k2 * r**4 + k3 * r**6 + k4 * r**8) + (p1 * (r**2 + 2 * x**2) + 2 * p2 * x * y)
    yd = y * (1 + k1 * r**2 + k2 * r**4 + k3 * r**6 + k4 * r**8) + (p2 * (r**2 + 2 * y**2) + 2 * p1 * x * y)
    return image_width / 2.0 + cx + xd * f + xd * b1 + yd * b2, image_height / 2.0 + cy + yd * f


def ideal_to_warped_map(params, image_scale=1.0):
    """make_distortion_map's forward map, cameras/cameras.py:1027-1053: (2, h, w) [rows, cols] of the position in
    the warped image that every ideal pixel maps to."""
    im_h, im_w = params["image_height"], params["image_width"]
    if np.isclose(image_scale, 1.0):
        h_range, w_range = np.arange(im_h), np.arange(im_w)
    else:
        kw = {"start": 1 / (2 * image_scale), "step": 1 / image_scale}
        h_range = np.arange(stop=im_h, **kw)[: int(im_h * image_scale)]
        w_range = np.arange(stop=im_w, **kw)[: int(im_w * image_scale)]
    rows, cols = np.meshgrid(h_range, w_range, indexing="ij")
    wc, wr = metashape_ideal_to_warped(cols, rows, **params)
    if not np.isclose(image_scale, 1.0):
        wc, wr = wc * image_scale, wr * image_scale
    return np.stack([wr, wc], axis=0)


def inverse_map_griddata(ijmap, downsample=1, fill=-1):
    """inverse_map_interpolation, utils/indexing.py:87-150 (scipy griddata, linear)."""
    from scipy.interpolate import griddata

    H, W = ijmap.shape[1:]
    igrid, jgrid = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    grid = np.stack([igrid.ravel(), jgrid.ravel()], axis=1)
    ds = slice(None, None, downsample)
    sample_y = np.stack([igrid[ds, ds].ravel(), jgrid[ds, ds].ravel()], axis=1)
    sample_x = np.stack([ijmap[0][ds, ds].ravel(), ijmap[1][ds, ds].ravel()], axis=1)
    inv_i = griddata(sample_x, sample_y[:, 0], grid, method="linear", fill_value=fill)
    inv_j = griddata(sample_x, sample_y[:, 1], grid, method="linear", fill_value=fill)
    return np.stack([inv_i.reshape(H, W), inv_j.reshape(H, W)], axis=0)


def nearest_source_index(src_rows, src_cols, h, w):
    """Nearest source pixel of continuous (row, col) coordinates, -1 outside the image: what skimage.transform.warp
    (order=0, mode='constant') does through scipy.ndimage.map_coordinates (utils/image.py:108-117)."""
    rr, cc = np.floor(src_rows + 0.5), np.floor(src_cols + 0.5)
    ok = np.isfinite(rr) & np.isfinite(cc) & (rr >= 0) & (rr < h) & (cc >= 0) & (cc < w)
    idx = np.full(src_rows.shape, -1, dtype=np.int64)
    idx[ok] = (rr[ok] * w + cc[ok]).astype(np.int64)
    return idx


def warp_ids(ids, src_index, fill=-1):
    """Integer-safe nearest warp: out[p] = ids.flat[src_index[p]] or fill."""
    out = np.full(src_index.shape, fill, dtype=ids.dtype)
    ok = src_index >= 0
    out[ok] = ids.ravel()[src_index[ok]]
    return out


def exact_inverse_coordinates(params, image_scale=1.0, iters=40):
    """Exact ideal <- warped source coordinates (rows, cols) by Newton iterations on the forward model in float64:
    the 'corrected reference' the GPU warp is held to (the reference interpolates a down-sampled forward map)."""
    im_h, im_w = params["image_height"], params["image_width"]
    h, w = int(im_h * image_scale), int(im_w * image_scale)
    one = np.isclose(image_scale, 1.0)
    ti, tj = np.meshgrid(np.arange(h, dtype=float), np.arange(w, dtype=float), indexing="ij")
    tx, ty = (tj, ti) if one else (tj / image_scale, ti / image_scale)
    x, y = tx.copy(), ty.copy()
    ok = np.zeros(x.shape, dtype=bool)
    e = 1e-4
    with np.errstate(all="ignore"):
        for _ in range(iters):
            fx, fy = metashape_ideal_to_warped(x, y, **params)
            rx, ry = fx - tx, fy - ty
            ok = (np.abs(rx) < 1e-9) & (np.abs(ry) < 1e-9)
            fxx, fyx = metashape_ideal_to_warped(x + e, y, **params)
            fxy, fyy = metashape_ideal_to_warped(x, y + e, **params)
            a, b, c, d = (fxx - fx) / e, (fxy - fx) / e, (fyx - fy) / e, (fyy - fy) / e
            det = a * d - b * c
            dx, dy = (d * rx - b * ry) / det, (-c * rx + a * ry) / det
            lim = 0.25 * (im_w / 2.0 + im_h / 2.0)
            n = np.maximum(np.abs(dx), np.abs(dy))
            scale = np.where(n > lim, lim / n, 1.0)
            x = np.where(ok, x, x - dx * scale)
            y = np.where(ok, y, y - dy * scale)
    rows, cols = (y, x) if one else (y * image_scale - 0.5, x * image_scale - 0.5)
    rows = np.where(ok, rows, np.nan)
    cols = np.where(ok, cols, np.nan)
    return rows, cols


# --------------------------------------------------------------------------------------------------
# label_polygons, shapely-free restatement of the sjoin path (meshes.py:1141-1306)
# --------------------------------------------------------------------------------------------------
def _point_in_rings(px, py, rings):
    inside = np.zeros(px.shape, dtype=bool)
    for ring in rings:
        ring = np.asarray(ring, dtype=float)
        if len(ring) > 1 and np.array_equal(ring[0], ring[-1]):
            ring = ring[:-1]
        x0, y0 = ring[:, 0], ring[:, 1]
        x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
        for a, b, c, d in zip(x0, y0, x1, y1):
            with np.errstate(divide="ignore", invalid="ignore"):
                hit = ((b > py) != (d > py)) & (px < (c - a) * (py - b) / (d - b) + a)
            inside ^= hit
    return inside


def _segments_cross(ax, ay, bx, by, cx, cy, dx, dy):
    def orient(px, py, qx, qy, rx, ry):
        return (qx - px) * (ry - py) - (qy - py) * (rx - px)

    o1, o2 = orient(ax, ay, bx, by, cx, cy), orient(ax, ay, bx, by, dx, dy)
    o3, o4 = orient(cx, cy, dx, dy, ax, ay), orient(cx, cy, dx, dy, bx, by)
    return ((o1 > 0) != (o2 > 0)) & ((o3 > 0) != (o4 > 0)) & (o1 != 0) & (o2 != 0) & (o3 != 0) & (o4 != 0)


def label_polygons(points, faces, face_labels, polygons_rings, face_weighting=None, xy=None):
    """Per polygon: class with the largest sum of area3D * weight over the labelled faces whose 2-D triangle lies
    within the polygon (all vertices inside by the even-odd rule, no edge crossing, no ring inside the triangle);
    NaN when nothing voted.  Returns (labels list, weights (n_polys, n_classes))."""
    points = np.asarray(points, dtype=float)
    xy = points[:, :2] if xy is None else np.asarray(xy, dtype=float)
    face_labels = np.asarray(face_labels, dtype=float)
    finite = np.isfinite(face_labels)
    n_classes = int(face_labels[finite].max()) + 1 if finite.any() else 1
    tri = xy[faces]  # (F, 3, 2)
    a3 = points[faces]
    area3d = 0.5 * np.linalg.norm(np.cross(a3[:, 1] - a3[:, 0], a3[:, 2] - a3[:, 0]), axis=1)
    w = area3d * (1.0 if face_weighting is None else np.asarray(face_weighting, dtype=float))
    weights = np.zeros((len(polygons_rings), n_classes))
    for p, rings in enumerate(polygons_rings):
        allp = np.concatenate([np.asarray(r, dtype=float) for r in rings])
        lo, hi = allp.min(0), allp.max(0)
        cand = finite & (tri[..., 0].min(1) >= lo[0]) & (tri[..., 1].min(1) >= lo[1]) & (tri[..., 0].max(1) <= hi[0]) & (
            tri[..., 1].max(1) <= hi[1])
        idx = np.nonzero(cand)[0]
        if len(idx) == 0:
            continue
        t = tri[idx]
        ok = np.ones(len(idx), dtype=bool)
        for v in range(3):
            ok &= _point_in_rings(t[:, v, 0], t[:, v, 1], rings)
        for ring in rings:
            ring = np.asarray(ring, dtype=float)
            if len(ring) > 1 and np.array_equal(ring[0], ring[-1]):
                ring = ring[:-1]
            nxt = np.roll(ring, -1, axis=0)
            for (ex0, ey0), (ex1, ey1) in zip(ring, nxt):
                for v in range(3):
                    u = (v + 1) % 3
                    ok &= ~_segments_cross(t[:, v, 0], t[:, v, 1], t[:, u, 0], t[:, u, 1], ex0, ey0, ex1, ey1)
            qx, qy = ring[0]
            s = [(t[:, (v + 1) % 3, 0] - t[:, v, 0]) * (qy - t[:, v, 1]) - (t[:, (v + 1) % 3, 1] - t[:, v, 1]) * (qx - t[:, v, 0])
                 for v in range(3)]
            ok &= ~(((s[0] > 0) & (s[1] > 0) & (s[2] > 0)) | ((s[0] < 0) & (s[1] < 0) & (s[2] < 0)))
        sel = idx[ok]
        np.add.at(weights[p], face_labels[sel].astype(int), w[sel])
    best = weights.max(axis=1)
    labels = np.where(best > 0, weights.argmax(axis=1).astype(float), np.nan)
    return labels.tolist(), weights


def _clip_ring_by_triangle(ring, tri):
    """Sutherland-Hodgman: the (possibly non-convex) ring clipped by the convex, counter-clockwise triangle."""
    out = [tuple(p) for p in ring]
    for e in range(3):
        (x0, y0), (x1, y1) = tri[e], tri[(e + 1) % 3]
        src, out = out, []
        for i in range(len(src)):
            p, q = src[i], src[(i + 1) % len(src)]
            dp = (x1 - x0) * (p[1] - y0) - (y1 - y0) * (p[0] - x0)
            dq = (x1 - x0) * (q[1] - y0) - (y1 - y0) * (q[0] - x0)
            if dp >= 0:
                out.append(p)
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        if len(out) < 3:
            return []
    return out


def _ring_area(ring):
    return 0.5 * abs(sum(ring[i][0] * ring[(i + 1) % len(ring)][1] - ring[(i + 1) % len(ring)][0] * ring[i][1]
                         for i in range(len(ring))))


def label_polygons_overlay(points, faces, face_labels, polygons, face_weighting=None, xy=None):
    """label_polygons with sjoin_overlay=False (meshes.py:1263-1276), pure Python for small cases: every labelled face
    votes in every polygon with  area2D(face n polygon) * (area3D / area2D) * weight.  ``polygons``: list of dicts
    {"exterior": (K, 2), "holes": [(K, 2), ...]}.  The intersection area is computed by clipping the WHOLE exterior
    ring (and each hole, subtracted) with the triangle -- a different decomposition from the CUDA kernel's signed fan
    triangles.  Returns (labels list, weights (n_polys, n_classes))."""
    points = np.asarray(points, dtype=float)
    xy = points[:, :2] if xy is None else np.asarray(xy, dtype=float)
    face_labels = np.asarray(face_labels, dtype=float)
    finite = np.isfinite(face_labels)
    n_classes = int(face_labels[finite].max()) + 1 if finite.any() else 1
    weights = np.zeros((len(polygons), n_classes))
    for f in np.nonzero(finite)[0]:
        tri = [tuple(xy[v]) for v in faces[f]]
        a2 = (tri[1][0] - tri[0][0]) * (tri[2][1] - tri[0][1]) - (tri[1][1] - tri[0][1]) * (tri[2][0] - tri[0][0])
        if a2 == 0:
            continue
        if a2 < 0:
            tri = [tri[0], tri[2], tri[1]]
        p3 = points[faces[f]]
        area3d = 0.5 * np.linalg.norm(np.cross(p3[1] - p3[0], p3[2] - p3[0]))
        ratio = area3d / (0.5 * abs(a2)) * (1.0 if face_weighting is None else float(face_weighting[f]))
        for p, poly in enumerate(polygons):
            inter = _ring_area(_clip_ring_by_triangle(np.asarray(poly["exterior"], dtype=float), tri) or [(0, 0)] * 3)
            for hole in poly.get("holes", []):
                inter -= _ring_area(_clip_ring_by_triangle(np.asarray(hole, dtype=float), tri) or [(0, 0)] * 3)
            if inter > 0:
                weights[p, int(face_labels[f])] += inter * ratio
    best = weights.max(axis=1)
    labels = np.where(best > 0, weights.argmax(axis=1).astype(float), np.nan)
    return labels.tolist(), weights
