"""ctypes binding of libgeograypher_b200.so (the C ABI declared in include/geograypher_b200.h).

The library is the product's only compute path: if it is missing or no CUDA device is present every entry
point raises -- there is no CPU fallback.  Tensors are torch CUDA tensors; only their device pointers cross
the ABI.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgeograypher_b200.so"
MAX_VIEWS_PER_CALL = 32

# enums of include/geograypher_b200.h
PRED_F32, PRED_F64, PRED_U8, PRED_INDEX_U8 = 0, 1, 2, 3
MODE_LAST_PIXEL, MODE_PIXEL_SUM, MODE_VOTE = 0, 1, 2
OUT_F64, OUT_F32, OUT_U8 = 0, 1, 2
FLAG_COMPAT_NEGATIVE_INDEX, FLAG_KEEP_NAN, FLAG_ASSIGN, FLAG_TRUNCATE = 1, 2, 4, 8
ERR_OVERFLOW = -4

EXPORTS = [
    "gg_abi_version", "gg_last_error", "gg_create", "gg_destroy", "gg_sync", "gg_reserve",
    "gg_last_batch_stats", "gg_set_mesh", "gg_project", "gg_rasterize", "gg_aggregate",
    "gg_project_aggregate", "gg_finalize", "gg_render_flat", "gg_stage_count", "gg_stage_name", "gg_profile",
    "gg_profile_read", "gg_drain", "gg_set_pipeline", "gg_build_warp_map", "gg_gather_i32",
    "gg_label_polygons", "gg_get_capacity", "gg_rasterize_render_flat", "gg_project_winners", "gg_accumulate_rows",
    "gg_overflow_info", "gg_resize_render", "gg_label_polygons_overlay", "gg_host_alloc", "gg_host_free",
    "gg_pointer_kind", "gg_gather_rows_host",
]


class GGCamera(ctypes.Structure):
    """gg_camera of the header; identical layout to the oracle's ora_camera."""

    _fields_ = [
        ("m", ctypes.c_float * 12),
        ("f", ctypes.c_float),
        ("px", ctypes.c_float),
        ("py", ctypes.c_float),
        ("W", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("znear", ctypes.c_float),
    ]


class GGDistortion(ctypes.Structure):
    """gg_distortion of the header."""

    _fields_ = [("f", ctypes.c_double), ("cx", ctypes.c_double), ("cy", ctypes.c_double), ("W", ctypes.c_int32),
                ("H", ctypes.c_int32)] + [(k, ctypes.c_double) for k in ("k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")] + [
                    ("image_scale", ctypes.c_double)]


class GeograypherB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libgeograypher_b200 error {code}: {message}")
        self.code = code


def build(verbose: bool = False) -> Path:
    """Compile the CUDA sources in-tree for sm_100a (geograypher_b200/csrc/build.sh)."""
    res = subprocess.run(["bash", str(_PKG / "csrc" / "build.sh")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libgeograypher_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_lib = None


def load():
    """Load the shared library (building it if the .so is absent and nvcc is available) and set prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    camp = ctypes.POINTER(GGCamera)
    lib.gg_abi_version.restype = i32
    lib.gg_last_error.restype = ctypes.c_char_p
    lib.gg_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.gg_destroy.argtypes = [vp]
    lib.gg_destroy.restype = None
    lib.gg_sync.argtypes = [vp, vp]
    lib.gg_reserve.argtypes = [vp, i64, i64]
    lib.gg_last_batch_stats.argtypes = [vp, i32, vp]
    lib.gg_overflow_info.argtypes = [vp, vp]
    lib.gg_host_alloc.argtypes = [i32, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.gg_host_free.argtypes = [vp]
    lib.gg_pointer_kind.argtypes = [vp]
    lib.gg_gather_rows_host.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp, i32]
    lib.gg_get_capacity.argtypes = [vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    lib.gg_set_mesh.argtypes = [vp, vp, i64, vp, i64, vp]
    lib.gg_project.argtypes = [vp, camp, i32, vp, vp, vp, vp, vp]
    lib.gg_rasterize.argtypes = [vp, camp, i32, vp, vp, vp]
    lib.gg_aggregate.argtypes = [vp, vp, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.gg_project_aggregate.argtypes = [vp, camp, i32, ctypes.POINTER(vp), i32, i32, i32, i32, vp, vp, vp, vp]
    lib.gg_project_winners.argtypes = [vp, camp, i32, i32, vp, i64, vp, vp]
    lib.gg_accumulate_rows.argtypes = [vp, vp, i64, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.gg_finalize.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp]
    lib.gg_render_flat.argtypes = [vp, vp, i64, vp, i32, vp, i32, vp]
    lib.gg_rasterize_render_flat.argtypes = [vp, camp, i32, vp, i32, vp, i32, vp, vp]
    lib.gg_stage_name.argtypes = [i32]
    lib.gg_stage_name.restype = ctypes.c_char_p
    lib.gg_drain.argtypes = [vp, vp]
    lib.gg_set_pipeline.argtypes = [vp, i32]
    lib.gg_build_warp_map.argtypes = [i32, ctypes.POINTER(GGDistortion), i32, i32, i32, vp, vp, vp]
    lib.gg_gather_i32.argtypes = [i32, vp, vp, i64, ctypes.c_int32, vp, vp]
    lib.gg_resize_render.argtypes = [i32, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp]
    lib.gg_label_polygons.argtypes = [i32, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i32, i32, vp, vp]
    lib.gg_label_polygons_overlay.argtypes = lib.gg_label_polygons.argtypes
    lib.gg_profile.argtypes = [vp, i32]
    lib.gg_profile_read.argtypes = [vp, vp, vp, i32]
    for name in EXPORTS:
        if name not in ("gg_last_error", "gg_destroy", "gg_stage_name"):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise GeograypherB200Error(rc, load().gg_last_error().decode("utf-8", "replace"))


POINTER_PAGEABLE, POINTER_PINNED, POINTER_DEVICE, POINTER_MANAGED = 0, 1, 2, 3


def pointer_kind(array) -> int:
    """Where a NumPy array's memory lives as far as the GPU is concerned (gg_pointer_kind): pageable host memory
    (the GPU cannot read it), page-locked host memory, device memory or managed memory."""
    return int(load().gg_pointer_kind(ctypes.c_void_p(int(array.ctypes.data))))


def gather_rows_host(images, pairs, offsets, out, n_threads: int = 0, pair_starts=None):
    """out[i] = images[v].reshape(-1, E)[pairs[s, 1]] for every i in offsets[v] .. offsets[v+1], with
    s = pair_starts[v] + i - offsets[v] (``pair_starts`` None: s = i, packed lists) -- gg_gather_rows_host: the host half
    of the pageable-image route, on several cores and without the GIL.  ``images``: C-contiguous arrays of one dtype
    whose last axis has E elements (or (H, W) arrays, E = 1); ``pairs`` (N, 2) int32 (face, pixel); ``out`` (M, E)."""
    n = len(images)
    E = int(out.shape[1]) if out.ndim == 2 else 1
    row_bytes = E * out.dtype.itemsize
    for a in images:
        if not a.flags["C_CONTIGUOUS"] or a.dtype != out.dtype or a.size % E:
            raise ValueError("gather_rows_host: images must be C-contiguous, of the output's dtype, with E-element rows")
    if not (pairs.flags["C_CONTIGUOUS"] and pairs.dtype == np.int32 and out.flags["C_CONTIGUOUS"]):
        raise ValueError("gather_rows_host: pairs must be contiguous int32, out contiguous")
    ptrs = (ctypes.c_void_p * n)(*[int(a.ctypes.data) for a in images])
    npix = np.asarray([a.size // E for a in images], dtype=np.int64)
    offs = np.ascontiguousarray(offsets, dtype=np.int64)
    if offs.shape[0] != n + 1 or int(offs[-1]) > out.shape[0]:
        raise ValueError("gather_rows_host: offsets do not match images / out")
    starts = offs if pair_starts is None else np.ascontiguousarray(pair_starts, dtype=np.int64)
    if starts.shape[0] < n or any(int(starts[v]) + int(offs[v + 1] - offs[v]) > pairs.shape[0] for v in range(n)):
        raise ValueError("gather_rows_host: pair lists run past the end of pairs")
    _check(load().gg_gather_rows_host(ptrs, npix.ctypes.data, pairs.ctypes.data,
                                      None if pair_starts is None else starts.ctypes.data, offs.ctypes.data, n,
                                      row_bytes, out.ctypes.data, int(n_threads)))


def host_array(shape, dtype=np.float32, device: int = 0) -> np.ndarray:
    """A NumPy array in host memory that the GPU can read in place (gg_host_alloc): the container to hand prediction
    images to ``aggregate_projected_images`` in.  The aggregation fetches one row per visible face and view out of it
    over PCIe; unlike page-locked buffers its scattered-read rate does not collapse when many large images are
    resident.  Freed when the array (and every view of it) is garbage-collected."""
    import weakref

    shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
    dtype = np.dtype(dtype)
    n_bytes = max(int(np.prod(shape)) * dtype.itemsize, 1)
    lib = load()
    ptr = ctypes.c_void_p()
    _check(lib.gg_host_alloc(int(device), n_bytes, ctypes.byref(ptr)))
    buf = (ctypes.c_ubyte * n_bytes).from_address(ptr.value)
    weakref.finalize(buf, lib.gg_host_free, ctypes.c_void_p(ptr.value))  # the array's base object keeps buf alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def make_camera(world_to_cam, f, cx, cy, image_width, image_height, render_img_scale=1.0, origin=None,
                znear=1e-3) -> GGCamera:
    """float64 camera parameters -> the float32 record of the rasterization contract.

    Follows PhotogrammetryCamera (reference cameras/cameras.py:56-102): principal point offsets are measured
    from the image centre; the scaled raster is (int(H*s), int(W*s)) (cameras.py:197-200) and keeps the
    vertical field of view (cameras.py:472): f' = f*h'/H.  ``origin`` (float64) is the shift that was
    subtracted from the mesh vertices before rounding them to float32.
    """
    m = np.asarray(world_to_cam, dtype=np.float64)[:3, :4].copy()
    if origin is not None:
        m[:, 3] = m[:, 3] + m[:, :3] @ np.asarray(origin, dtype=np.float64)
    h, w = int(image_height * render_img_scale), int(image_width * render_img_scale)
    s = h / float(image_height)
    cam = GGCamera()
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


def make_distortion(f, cx, cy, image_width, image_height, image_scale=1.0, k1=0.0, k2=0.0, k3=0.0, k4=0.0, p1=0.0,
                    p2=0.0, b1=0.0, b2=0.0) -> GGDistortion:
    return GGDistortion(f=f, cx=cx, cy=cy, W=int(image_width), H=int(image_height), k1=k1, k2=k2, k3=k3, k4=k4, p1=p1,
                        p2=p2, b1=b1, b2=b2, image_scale=image_scale)


def build_warp_map(dist: GGDistortion, h: int, w: int, warped_to_ideal: bool, device: int = 0, want_coords=False):
    """(h, w) int32 CUDA tensor of nearest source indices (-1 = outside); optionally also the (h, w, 2) float32
    continuous (row, col) source coordinates."""
    import torch

    if not torch.cuda.is_available():
        raise GeograypherB200Error(-5, "no CUDA device: geograypher_b200 has no CPU fallback")
    dev = torch.device("cuda", device)
    src = torch.empty((h, w), dtype=torch.int32, device=dev)
    rc = torch.empty((h, w, 2), dtype=torch.float32, device=dev) if want_coords else None
    _check(load().gg_build_warp_map(device, ctypes.byref(dist), h, w, 1 if warped_to_ideal else 0, src.data_ptr(),
                                    rc.data_ptr() if rc is not None else None, _stream_ptr(None)))
    return (src, rc) if want_coords else src


def gather_i32(d_in, d_src_index, fill: int):
    """Nearest-neighbour warp of an int32 raster through a source-index map (both CUDA tensors)."""
    import torch

    assert d_in.dtype == torch.int32 and d_src_index.dtype == torch.int32 and d_in.is_cuda
    out = torch.empty(d_src_index.shape, dtype=torch.int32, device=d_in.device)
    _check(load().gg_gather_i32(d_in.device.index or 0, d_in.contiguous().data_ptr(), d_src_index.data_ptr(),
                                d_src_index.numel(), int(fill), out.data_ptr(), _stream_ptr(None)))
    return out


def resize_render(d_in, h_out: int, w_out: int, order: int, out_dtype=OUT_F64):
    """(h_in, w_in, D) float64 CUDA tensor -> (h_out, w_out, D) tensor: skimage.transform.resize(order = 0 | 1) as
    save_renders uses it (reference meshes.py:2312-2321), optionally with the uint8 cast fused in."""
    import torch

    assert d_in.is_cuda and d_in.dtype == torch.float64 and d_in.dim() == 3
    d_in = d_in.contiguous()
    h_in, w_in, D = (int(x) for x in d_in.shape)
    out = torch.empty((h_out, w_out, D), dtype=torch.float64 if out_dtype == OUT_F64 else torch.uint8, device=d_in.device)
    _check(load().gg_resize_render(d_in.device.index or 0, d_in.data_ptr(), h_in, w_in, D, int(h_out), int(w_out),
                                   int(order), out.data_ptr(), out_dtype, _stream_ptr(None)))
    return out


def _signed_area(ring):
    x, y = ring[:, 0], ring[:, 1]
    return 0.5 * float(np.dot(x, np.roll(y, -1)) - np.dot(np.roll(x, -1), y))


def label_polygons_weights(xyz, xy, faces, labels, face_weight, rings, n_classes, device=0, overlay=False, holes=None):
    """(n_polys, n_classes) float64 NumPy array of summed face weights.  ``rings``: list (one entry per polygon) of
    lists of (K, 2) float arrays (exterior rings and holes alike).  ``overlay``: partially covered faces vote with the
    area of their intersection (sjoin_overlay=False) instead of only faces within a polygon; it needs to know which
    rings are holes (``holes``: per polygon a list of bools; default: only the first ring of a polygon is an
    exterior).  Host arrays in, host array out."""
    import torch

    if not torch.cuda.is_available():
        raise GeograypherB200Error(-5, "no CUDA device: geograypher_b200 has no CPU fallback")
    dev = torch.device("cuda", device)
    flat, ring_off, poly_off, bbox = [], [0], [0], []
    for pi, poly in enumerate(rings):
        pts = []
        for ri, ring in enumerate(poly):
            ring = np.asarray(ring, dtype=np.float64).reshape(-1, 2)
            if len(ring) > 1 and np.array_equal(ring[0], ring[-1]):
                ring = ring[:-1]  # closed rings repeat their first vertex
            if overlay and len(ring) >= 3:  # exteriors counter-clockwise, holes clockwise
                is_hole = holes[pi][ri] if holes is not None else ri > 0
                if (_signed_area(ring) < 0) != bool(is_hole):
                    ring = ring[::-1]
            flat.append(ring)
            pts.append(ring)
            ring_off.append(ring_off[-1] + len(ring))
        poly_off.append(len(ring_off) - 1)
        allp = np.concatenate(pts) if pts else np.zeros((0, 2))
        bbox.append([allp[:, 0].min(), allp[:, 1].min(), allp[:, 0].max(), allp[:, 1].max()] if len(allp) else [1, 1, 0, 0])
    n_polys = len(rings)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_xyz, d_xy = t(xyz, np.float64), t(xy, np.float64)
    d_faces, d_labels = t(faces, np.int32), t(labels, np.float64)
    d_fw = t(face_weight, np.float64) if face_weight is not None else None
    d_pxy = t(np.concatenate(flat) if flat else np.zeros((1, 2)), np.float64)
    d_ro, d_po, d_bb = t(ring_off, np.int32), t(poly_off, np.int32), t(bbox, np.float64)
    d_w = torch.zeros((n_polys, n_classes), dtype=torch.float64, device=dev)
    fn = load().gg_label_polygons_overlay if overlay else load().gg_label_polygons
    _check(fn(device, d_xyz.data_ptr(), d_xy.data_ptr(), d_faces.data_ptr(), d_labels.data_ptr(),
              d_fw.data_ptr() if d_fw is not None else None, int(len(faces)), d_pxy.data_ptr(), d_ro.data_ptr(),
              d_po.data_ptr(), d_bb.data_ptr(), n_polys, int(n_classes), d_w.data_ptr(), _stream_ptr(None)))
    return d_w.cpu().numpy()


_PREFAULT_POOL = None
_PREFAULT_THREADS = 8
_PREFAULT_MIN_BYTES = 16 << 20


def to_fresh_host(tensors):
    """CUDA tensors -> NEW NumPy arrays the caller owns, for results whose size changes from call to call (CSR arrays,
    rasters): the cost of such a delivery is not the PCIe copy but the first touch of the fresh host pages (measured on
    the benchmark host for 1 GiB, ``profiles/r02_d2h_dest.txt``: 2.2 GB/s for ``tensor.cpu()``, 4.9 GB/s into a
    NumPy array -- numpy asks for transparent huge pages --, 2.7 GB/s into a newly page-locked block, 11 GB/s into a
    NumPy array whose pages a few threads have touched first).  So: allocate with NumPy, pre-fault every large array on
    a small thread pool (one write per 4-KiB page; the slice writes release the GIL), and copy array k while the
    pages of the later ones are still being touched.  Accepts one tensor or a list; returns array(s) in the same order."""
    import torch

    global _PREFAULT_POOL
    single = not isinstance(tensors, (list, tuple))
    ts = [tensors] if single else list(tensors)
    outs, futures = [], []
    for t in ts:
        dst = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
        flat = dst.reshape(-1)
        pending = []
        if dst.nbytes >= _PREFAULT_MIN_BYTES:
            if _PREFAULT_POOL is None:
                from concurrent.futures import ThreadPoolExecutor

                _PREFAULT_POOL = ThreadPoolExecutor(_PREFAULT_THREADS, thread_name_prefix="gg-prefault")
            stride = max(1, 4096 // dst.itemsize)
            step = -(-flat.size // _PREFAULT_THREADS)
            pending = [_PREFAULT_POOL.submit(flat[lo:lo + step:stride].fill, 0) for lo in range(0, flat.size, step)]
        outs.append(dst)
        futures.append(pending)
    for t, dst, pending in zip(ts, outs, futures):
        for f in pending:
            f.result()
        if dst.size:
            torch.from_numpy(dst).copy_(t)
    return outs[0] if single else outs


def _stream_ptr(stream):
    if stream is None:
        import torch

        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    return ctypes.c_void_p(int(stream))


class Context:
    """One gg_context: a mesh resident on one device plus scratch.  Not thread-safe."""

    def __init__(self, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise GeograypherB200Error(-5, "no CUDA device: geograypher_b200 has no CPU fallback")
        self.lib = load()
        self.device = int(device)
        self.torch = torch
        h = ctypes.c_void_p()
        _check(self.lib.gg_create(self.device, ctypes.byref(h)))
        self.handle = h
        self.n_faces = 0
        self.n_verts = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.gg_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------------
    def _dev(self):
        return self.torch.device("cuda", self.device)

    @staticmethod
    def _cam_array(cams):
        arr = (GGCamera * len(cams))(*cams)
        return arr

    def sync(self, stream=None):
        _check(self.lib.gg_sync(self.handle, _stream_ptr(stream)))

    def drain(self, stream=None):
        """Make `stream` wait (on the device) for the fused aggregation still in flight on the internal streams."""
        _check(self.lib.gg_drain(self.handle, _stream_ptr(stream)))

    def set_pipeline(self, enable: bool):
        _check(self.lib.gg_set_pipeline(self.handle, 1 if enable else 0))

    def reserve(self, max_faces_per_view=0, max_bin_entries_per_view=0):
        _check(self.lib.gg_reserve(self.handle, int(max_faces_per_view), int(max_bin_entries_per_view)))
        self._reserved_faces = int(max_faces_per_view)

    def get_capacity(self):
        """(face records, (tile, face) pairs) the scratch holds per view; zeros before the first rasterization."""
        recs, bins = ctypes.c_int64(), ctypes.c_int64()
        _check(self.lib.gg_get_capacity(self.handle, ctypes.byref(recs), ctypes.byref(bins)))
        return int(recs.value), int(bins.value)

    def profile(self, enable: bool):
        """Bracket every kernel launch with CUDA events on its stream (per-stage timing for bench.py)."""
        _check(self.lib.gg_profile(self.handle, 1 if enable else 0))

    def profile_read(self, reset: bool = True):
        """{stage: (milliseconds, launches)} accumulated since the last reset; synchronises the device."""
        n = self.lib.gg_stage_count()
        ms = np.zeros(n, dtype=np.float64)
        launches = np.zeros(n, dtype=np.int64)
        _check(self.lib.gg_profile_read(self.handle, ms.ctypes.data, launches.ctypes.data, 1 if reset else 0))
        return {self.lib.gg_stage_name(i).decode(): (float(ms[i]), int(launches[i])) for i in range(n)}

    def last_batch_stats(self, n):
        out = np.zeros((n, 4), dtype=np.int64)
        _check(self.lib.gg_last_batch_stats(self.handle, n, out.ctypes.data))
        return out

    # -- mesh ------------------------------------------------------------------------------------------
    def set_mesh(self, verts32, faces32, stream=None):
        """verts32: (V,3) float32 CUDA tensor (local frame, origin-shifted); faces32: (F,3) int32 CUDA tensor."""
        t = self.torch
        assert verts32.is_cuda and faces32.is_cuda
        assert verts32.dtype == t.float32 and faces32.dtype == t.int32
        verts32, faces32 = verts32.contiguous(), faces32.contiguous()
        _check(self.lib.gg_set_mesh(self.handle, verts32.data_ptr(), verts32.shape[0], faces32.data_ptr(),
                                    faces32.shape[0], _stream_ptr(stream)))
        self.n_verts, self.n_faces = int(verts32.shape[0]), int(faces32.shape[0])

    # -- stage 1 -----------------------------------------------------------------------------------------
    def project(self, cams, stream=None):
        t, n, V = self.torch, len(cams), self.n_verts
        X = t.empty((n, V), dtype=t.int32, device=self._dev())
        Y = t.empty_like(X)
        invz = t.empty((n, V), dtype=t.float32, device=self._dev())
        valid = t.empty((n, V), dtype=t.uint8, device=self._dev())
        _check(self.lib.gg_project(self.handle, self._cam_array(cams), n, X.data_ptr(), Y.data_ptr(),
                                   invz.data_ptr(), valid.data_ptr(), _stream_ptr(stream)))
        return X, Y, invz, valid

    # -- stage 1+2 -----------------------------------------------------------------------------------------
    def rasterize(self, cams, out=None, want_depth=False, stream=None, check=True):
        """pix2face for up to 32 same-size views: (n, H, W) int32 CUDA tensor (-1 = no face)."""
        t, n = self.torch, len(cams)
        H, W = cams[0].H, cams[0].W
        if out is None:
            out = t.empty((n, H, W), dtype=t.int32, device=self._dev())
        depth = t.empty((n, H, W), dtype=t.float32, device=self._dev()) if want_depth else None
        if check:
            self.sync(stream)  # an overflow left behind by earlier unchecked batches belongs to them: raise it now
        for attempt in range(4):
            _check(self.lib.gg_rasterize(self.handle, self._cam_array(cams), n, out.data_ptr(),
                                         depth.data_ptr() if depth is not None else None, _stream_ptr(stream)))
            if not check:
                break
            try:
                self.sync(stream)
                break
            except GeograypherB200Error as e:
                if e.code != ERR_OVERFLOW or attempt == 3:
                    raise
                self._grow_after_overflow(n)
        return (out, depth) if want_depth else out

    def overflow_info(self):
        """(flags, face records wanted, tile entries wanted) found by the last sync() that raised ERR_OVERFLOW, over
        every batch enqueued since the sync before it."""
        out = np.zeros(3, dtype=np.int64)
        _check(self.lib.gg_overflow_info(self.handle, out.ctypes.data))
        return int(out[0]), int(out[1]), int(out[2])

    def _grow_after_overflow(self, n=None):
        """Grow what overflowed -- and only that -- after a sync() raised ERR_OVERFLOW.  The sizes come from the
        high-water marks the kernels keep over ALL batches since the previous sync (the overflowing batch is usually
        not the last one); each grown capacity at least doubles, so a few attempts always suffice.  Records dropped
        by a face-record overflow were never binned, so the tile-entry mark may still be too low: the caller's retry
        loop covers that."""
        del n
        flags, want_recs, want_bins = self.overflow_info()
        recs, bins = self.get_capacity()
        new_recs = max(int(want_recs * 1.25) + 4096, 2 * recs) if flags & 1 else recs
        new_bins = max(int(want_bins * 1.25) + 4096, 2 * bins) if flags & 2 else bins
        if not flags & 3:  # no information (should not happen): fall back to doubling both
            new_recs, new_bins = 2 * recs, 2 * bins
        self.reserve(min(new_recs, max(self.n_faces * 2, 1)), new_bins)

    # -- stage 3 -------------------------------------------------------------------------------------------
    def aggregate(self, pix2face, pred, pred_kind, C, mode, flags, d_sum, d_count, stream=None):
        H, W = int(pix2face.shape[-2]), int(pix2face.shape[-1])
        _check(self.lib.gg_aggregate(self.handle, pix2face.data_ptr(), H, W, pred.data_ptr(), pred_kind, C, mode,
                                     flags, d_sum.data_ptr(), d_count.data_ptr(), _stream_ptr(stream)))

    def project_aggregate(self, cams, preds, pred_kind, C, mode, flags, d_sum, d_count, pix2face_out=None,
                          stream=None, check=True):
        n = len(cams)
        ptrs = (ctypes.c_void_p * n)(*[p.data_ptr() for p in preds])
        if check:
            self.sync(stream)  # see rasterize()
        for attempt in range(4):
            _check(self.lib.gg_project_aggregate(self.handle, self._cam_array(cams), n, ptrs, pred_kind, C, mode,
                                                 flags, d_sum.data_ptr(), d_count.data_ptr(),
                                                 pix2face_out.data_ptr() if pix2face_out is not None else None,
                                                 _stream_ptr(stream)))
            if not check:
                return
            try:
                self.sync(stream)
                return
            except GeograypherB200Error as e:
                # In the fused modes an overflowing batch accumulates nothing (k_resolve_batch bails out), so it can be
                # replayed after growing the scratch.  pixel_sum is not replayable.
                if e.code != ERR_OVERFLOW or attempt == 3 or mode == MODE_PIXEL_SUM:
                    raise
                self._grow_after_overflow(n)

    def project_winners(self, cams, flags=0, stream=None, cap=None):
        """Rasterize up to 32 same-size views and list, per view, every visible face with its last pixel (row-major):
        returns (pairs (n, cap, 2) int32 CUDA tensor of (face, pixel), counts (n,) int32 CUDA tensor).  Asynchronous;
        an overflow of the scratch shows up at the next sync() (grow with _grow_after_overflow and call again).
        ``cap`` (default: what the scratch can produce) bounds the list of a view; a view with more visible faces
        reports its full count and lists only the first ``cap``."""
        t, n = self.torch, len(cams)
        full = self.capacity_hint(cams[0].W, cams[0].H)
        cap = full if cap is None else max(1, min(int(cap), full))
        pairs = t.empty((n, cap, 2), dtype=t.int32, device=self._dev())
        counts = t.empty((n,), dtype=t.int32, device=self._dev())
        _check(self.lib.gg_project_winners(self.handle, self._cam_array(cams), n, flags, pairs.data_ptr(), cap,
                                           counts.data_ptr(), _stream_ptr(stream)))
        return pairs, counts

    def capacity_hint(self, W, H):
        """Upper bound of the (face, pixel) pairs one view can produce with the current scratch: records + 1."""
        faces = max(self.get_capacity()[0], getattr(self, "_reserved_faces", 0))
        if faces <= 0:  # scratch not allocated yet: the library's default for this mesh
            faces = self.n_faces if self.n_faces < (1 << 20) else max(self.n_faces // 8, 1 << 20)
        return int(min(faces, self.n_faces, int(W) * int(H)) + 1)

    def accumulate_rows(self, pairs, n_rows, rows, pred_kind, C, mode, flags, d_sum, d_count, stream=None):
        """Apply ONE view's gathered rows (rows[i] belongs to face pairs[i, 0]); see gg_accumulate_rows."""
        _check(self.lib.gg_accumulate_rows(self.handle, pairs.data_ptr(), int(n_rows), rows.data_ptr(), pred_kind, C, mode,
                                           flags, d_sum.data_ptr(), d_count.data_ptr(), _stream_ptr(stream)))

    def finalize(self, d_sum, d_count, want_avg=True, want_argmax=True, stream=None):
        t = self.torch
        F, C = d_sum.shape
        avg = t.empty_like(d_sum) if want_avg else None
        argmax = t.empty((F,), dtype=t.float64, device=d_sum.device) if want_argmax else None
        _check(self.lib.gg_finalize(self.handle, d_sum.data_ptr(), d_count.data_ptr(), F, C,
                                    avg.data_ptr() if avg is not None else None,
                                    argmax.data_ptr() if argmax is not None else None, _stream_ptr(stream)))
        return avg, argmax

    def rasterize_render_flat(self, cams, face_tex64, out_dtype=OUT_F64, out=None, pix2face_out=None, stream=None,
                              check=True):
        """Fused pix2face + render_flat gather for up to 32 same-size views: (n, H, W, D) CUDA tensor."""
        t, n = self.torch, len(cams)
        H, W, D = cams[0].H, cams[0].W, int(face_tex64.shape[1])
        dt = {OUT_F64: t.float64, OUT_F32: t.float32, OUT_U8: t.uint8}[out_dtype]
        if out is None:
            out = t.empty((n, H, W, D), dtype=dt, device=self._dev())
        if check:
            self.sync(stream)  # see rasterize()
        for attempt in range(4):
            _check(self.lib.gg_rasterize_render_flat(self.handle, self._cam_array(cams), n, face_tex64.data_ptr(), D,
                                                     out.data_ptr(), out_dtype,
                                                     pix2face_out.data_ptr() if pix2face_out is not None else None,
                                                     _stream_ptr(stream)))
            if not check:
                break
            try:
                self.sync(stream)
                break
            except GeograypherB200Error as e:
                if e.code != ERR_OVERFLOW or attempt == 3:
                    raise
                self._grow_after_overflow(n)
        return out

    # -- stage 4 -------------------------------------------------------------------------------------------
    def render_flat(self, pix2face, face_tex64, out_dtype=OUT_F64, out=None, stream=None):
        """pix2face: (..., H, W) int32; face_tex64: (F, D) float64 -> (..., H, W, D)."""
        t = self.torch
        D = int(face_tex64.shape[1])
        dt = {OUT_F64: t.float64, OUT_F32: t.float32, OUT_U8: t.uint8}[out_dtype]
        if out is None:
            out = t.empty(tuple(pix2face.shape) + (D,), dtype=dt, device=pix2face.device)
        _check(self.lib.gg_render_flat(self.handle, pix2face.data_ptr(), pix2face.numel(), face_tex64.data_ptr(), D,
                                       out.data_ptr(), out_dtype, _stream_ptr(stream)))
        return out
