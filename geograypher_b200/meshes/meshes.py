"""TexturedPhotogrammetryMesh: the multiview projection hot path behind the reference's API.

Mirror of the hot-path surface of geograypher/meshes/meshes.py (reference v0.4.0):

    get_mesh_in_cameras_coords   meshes.py:1641-1676
    pix2face                     meshes.py:1678-1856   (VTK render  -> CUDA tiled z-buffer rasterizer)
    render_flat                  meshes.py:1858-1942   (NumPy gather -> CUDA gather)
    project_images               meshes.py:1944-2002   (NumPy fancy assignment -> CUDA last-pixel scatter)
    aggregate_projected_images   meshes.py:2004-2084   (np.nansum loop -> float64 device accumulators)

All arithmetic runs in libgeograypher_b200.so (geograypher_b200/_lib.py); this module only converts between
the reference's host-side objects and device tensors.  There is no CPU fallback: constructing the device
context without a CUDA device raises.  Everything outside the path (CRS handling beyond the local-frame
transform, ROI crops, vector / raster IO, visualisation) is out of scope -- see DESIGN.md.
"""
from __future__ import annotations

import hashlib
import logging
import os
import sys
import typing
from pathlib import Path

import numpy as np

from geograypher_b200 import _lib
from geograypher_b200.cameras import PhotogrammetryCamera, PhotogrammetryCameraSet
from geograypher_b200.constants import (
    CACHE_FOLDER,
    EARTH_CENTERED_EARTH_FIXED_CRS,
    NULL_TEXTURE_INT_VALUE,
    PATH_TYPE,
)
from geograypher_b200.utils import prefetch as _prefetch
from geograypher_b200.utils.indexing import determine_IDs_to_labels


class _HostMapped:
    """A host array the GPU can read in place (page-locked or gg.host_array memory), handed to the kernels by address."""

    def __init__(self, array):
        self.array = array  # keeps the memory alive while batches that read it are in flight

    def data_ptr(self):
        return int(self.array.ctypes.data)


class LocalMesh:
    """The mesh in a camera set's local frame (what the reference passes around as a transformed
    ``pv.PolyData``): float64 ``points`` (V, 3), ``faces`` (F, 3) and the device context that holds the float32
    copy the kernels read."""

    def __init__(self, points, faces, origin, context, key):
        self.points = points
        self.faces = faces
        self.origin = origin
        self.context = context
        self.key = key

    @property
    def n_faces(self):
        return self.faces.shape[0]


def _as_faces(faces) -> np.ndarray:
    faces = np.asarray(faces)
    if faces.ndim == 1:  # pyvista layout [3, i, j, k, 3, ...] (meshes.py:229)
        faces = faces.reshape((-1, 4))[:, 1:4]
    if faces.ndim != 2 or faces.shape[1] != 3:
        raise ValueError("faces must be (F, 3) or pyvista's flat [3, i, j, k, ...] layout")
    return np.ascontiguousarray(faces, dtype=np.int32)


def _is_ecef(crs) -> bool:
    if crs is None:
        return True
    if isinstance(crs, str):
        return crs.upper().replace(" ", "") in ("EPSG:4978", "ECEF")
    to_epsg = getattr(crs, "to_epsg", None)
    return to_epsg is not None and to_epsg() == 4978


class TexturedPhotogrammetryMesh:
    def __init__(
        self,
        mesh,
        input_CRS=EARTH_CENTERED_EARTH_FIXED_CRS,
        downsample_target: float = 1.0,
        texture: typing.Union[PATH_TYPE, np.ndarray, None] = None,
        texture_column_name=None,
        IDs_to_labels: typing.Union[dict, None] = None,
        shift: typing.Union[np.ndarray, None] = None,
        ROI=None,
        ROI_buffer_meters: float = 0,
        log_level: str = "INFO",
        device: int = 0,
        compat_negative_index: bool = False,
        views_per_batch: int = 8,
        use_principal_point: bool = True,
        sparse_host_gather: bool = True,
        prefetch_threads: typing.Union[int, None] = None,
    ):
        """A mesh with per-vertex / per-face textures that can be rendered into, and painted from, posed cameras.

        Args (reference meshes.py:55-104 unless marked *new*):
            mesh: ``(verts (V,3), faces (F,3))`` tuple, a dict / npz path with ``verts`` and ``faces``, or any object
                with ``.points`` and ``.faces`` (pyvista layout accepted).
            input_CRS: CRS of the vertex coordinates.  EPSG:4978 (default) or None need no extra dependency; any
                other CRS is reprojected with pyproj if it is installed (meshes.py:231-286).
            downsample_target, ROI, ROI_buffer_meters: geospatial preprocessing of the reference; only the
                no-op values are supported here.
            texture: ``(V|F, d)`` array or ``.npy`` path.
            IDs_to_labels: mapping from integer IDs to class names for discrete textures.
            shift: (3,) shift applied to the vertices in ``input_CRS``.
            device (*new*): CUDA device index.
            compat_negative_index (*new*): reproduce meshes.py:2000, where background pixels (-1) index the last
                face.  Off by default; results then differ from the reference on face F-1 only.
            views_per_batch (*new*): how many views are rasterized per launch (<= 32).
            use_principal_point (*new*): project with the cameras' principal-point offsets cx, cy, like the
                reference's PyTorch3D renderer and its ``ideal_to_warped`` (derived_meshes.py:772-780).  False
                reproduces the base class, whose pyvista camera has no principal point (cameras.py:446-477).
            sparse_host_gather (*new*): prediction images in ordinary (pageable) NumPy arrays are not uploaded; the
                GPU lists the one pixel per visible face that the aggregation needs and the host gathers those rows.
                False uploads whole images.  Page-locked arrays are read in place by the GPU -- until the distinct
                page-locked images of a call add up to more than ``pinned_direct_limit_bytes`` (an attribute, 4 GiB):
                scattered PCIe reads from page-locked memory slow down with its footprint on some hosts (2.5x on
                the benchmark host at 6 GB), where the host-gather route is then the faster one.
                ``host_array`` memory keeps its rate and is always read in place.
            prefetch_threads (*new*): workers that read the prediction images of the coming views ahead of the
                aggregation.  None (default): on for predictions decoded from files (``LookUpSegmentor``, any segmentor
                with ``io_bound = True``), off otherwise; 0 disables; a number forces it (the segmentor must then be
                thread-safe).
        """
        if downsample_target != 1.0 or ROI is not None:
            raise NotImplementedError(
                "mesh decimation and ROI cropping are geospatial preprocessing outside the projection path; "
                "prepare the mesh with geograypher and pass the arrays"
            )
        self.downsample_target = downsample_target
        self.texture = None
        self.vertex_texture = None
        self.face_texture = None
        self.IDs_to_labels = None
        self.device = int(device)
        self.compat_negative_index = bool(compat_negative_index)
        self.views_per_batch = int(max(1, min(views_per_batch, _lib.MAX_VIEWS_PER_CALL)))
        self.use_principal_point = bool(use_principal_point)
        self.sparse_host_gather = bool(sparse_host_gather)
        self.prefetch_threads = None if prefetch_threads is None else int(prefetch_threads)
        self.pinned_direct_limit_bytes = int(os.environ.get("GG_PINNED_DIRECT_LIMIT", 4 << 30))
        self._context = None
        self._local_cache = None

        self.logger = logging.getLogger(f"mesh_{id(self)}")
        self.logger.setLevel(log_level)
        if not self.logger.hasHandlers():
            self.logger.addHandler(logging.StreamHandler(stream=sys.stdout))

        self.load_mesh(mesh, input_CRS, shift=shift)
        self.load_texture(texture, texture_column_name, IDs_to_labels=IDs_to_labels,
                          background_ID=NULL_TEXTURE_INT_VALUE)

    # ------------------------------------------------------------------------------------------------
    # Set-up
    # ------------------------------------------------------------------------------------------------
    def load_mesh(self, mesh, input_CRS, shift=None, **unused):
        """Vertices are kept in float64 ECEF like the reference does (meshes.py:196-197, 212)."""
        if isinstance(mesh, (str, Path)):
            data = np.load(mesh)
            verts, faces = data["verts"], data["faces"]
        elif isinstance(mesh, dict):
            verts, faces = mesh["verts"], mesh["faces"]
        elif isinstance(mesh, (tuple, list)) and len(mesh) == 2:
            verts, faces = mesh
        elif hasattr(mesh, "points") and hasattr(mesh, "faces"):
            verts, faces = mesh.points, mesh.faces
        else:
            raise TypeError("mesh must be (verts, faces), a dict / .npz with those keys, or have .points/.faces")
        self.points = np.array(verts, dtype=float)  # copy + up-cast
        if self.points.ndim != 2 or self.points.shape[1] != 3:
            raise ValueError("vertices must be (V, 3)")
        self.faces = _as_faces(faces)
        if shift is not None:
            self.points += np.asarray(shift, dtype=float)
        self.CRS = input_CRS
        if not _is_ecef(input_CRS):
            self.reproject_CRS(EARTH_CENTERED_EARTH_FIXED_CRS, inplace=True)

    def reproject_CRS(self, target_CRS, inplace: bool = True):
        """Reference meshes.py:231-286.  Needs pyproj only when the CRS actually changes."""
        if _is_ecef(self.CRS) and _is_ecef(target_CRS):
            points = self.points
        else:
            try:
                import pyproj
            except ImportError as e:
                raise ImportError("re-projecting between CRSs needs pyproj; pass ECEF (EPSG:4978) vertices") from e
            tf = pyproj.Transformer.from_crs(self.CRS, target_CRS, always_xy=True)
            x, y, z = tf.transform(self.points[:, 0], self.points[:, 1], self.points[:, 2])
            points = np.stack([x, y, z], axis=1)
        if inplace:
            self.points = points
            self.CRS = target_CRS
            self._local_cache = None
            return None
        return points

    # -- textures (reference meshes.py:325-531) ---------------------------------------------------------
    def standardize_texture(self, texture_array: np.ndarray):
        if texture_array.ndim == 1:
            texture_array = np.expand_dims(texture_array, axis=1)
        elif texture_array.ndim != 2:
            raise ValueError(f"Input texture should have 1 or 2 dimensions but instead has {texture_array.ndim}")
        return texture_array

    def is_discrete_texture(self):
        return self.IDs_to_labels is not None

    def get_IDs_to_labels(self):
        return self.IDs_to_labels

    def load_texture(self, texture, texture_column_name=None, IDs_to_labels=None, background_ID=None):
        if texture is None:
            if IDs_to_labels is not None:
                self.IDs_to_labels = IDs_to_labels
            return
        if isinstance(texture, (str, Path)):
            texture = np.load(texture, allow_pickle=True)
        texture = self.standardize_texture(np.asarray(texture))
        if texture.shape[1] != 1:
            texture = texture.astype(float)  # multi-column -> real-valued (meshes.py:422-426)
            self.IDs_to_labels = None
        else:
            if IDs_to_labels is None:
                IDs_to_labels = determine_IDs_to_labels(texture, background_ID=background_ID)
            if IDs_to_labels is not None and list(IDs_to_labels.keys()) != list(IDs_to_labels.values()):
                labels_to_IDs = {v: k for k, v in IDs_to_labels.items()}
                if len(labels_to_IDs) != len(IDs_to_labels):
                    raise ValueError("IDs_to_labels is not a one-to-one mapping")
                texture = np.array([labels_to_IDs.get(l, np.nan) for l in texture.squeeze(axis=1).tolist()],
                                   dtype=float)[:, None]
            self.IDs_to_labels = IDs_to_labels
        self.set_texture(texture)

    def set_texture(self, texture_array, is_vertex_texture=None, delete_existing=True):
        texture_array = self.standardize_texture(np.asarray(texture_array))
        if is_vertex_texture is None:
            n_values, n_faces, n_verts = texture_array.shape[0], self.faces.shape[0], self.points.shape[0]
            if n_verts == n_faces:
                raise ValueError("Cannot infer whether texture should be applied to vertices of faces because "
                                 "the number is the same")
            elif n_values == n_verts:
                is_vertex_texture = True
            elif n_values == n_faces:
                is_vertex_texture = False
            else:
                raise ValueError(f"The number of elements in the texture ({n_values}) did not match the number "
                                 f"of faces ({n_faces}) or vertices ({n_verts})")
        if is_vertex_texture:
            self.vertex_texture = texture_array
            if delete_existing:
                self.face_texture = None
        else:
            self.face_texture = texture_array
            if delete_existing:
                self.vertex_texture = None

    def get_texture(self, request_vertex_texture=None, try_verts_faces_conversion=True):
        if self.vertex_texture is None and self.face_texture is None:
            return None
        if request_vertex_texture is None:
            if self.vertex_texture is not None and self.face_texture is not None:
                raise ValueError("Ambigious which texture is requested, set request_vertex_texture appropriately")
            request_vertex_texture = self.vertex_texture is not None
        if request_vertex_texture:
            if self.vertex_texture is not None:
                return self.standardize_texture(self.vertex_texture)
            raise ValueError("Vertex texture not present; face -> vertex conversion is outside the projection path")
        if self.face_texture is not None:
            return self.standardize_texture(self.face_texture)
        if try_verts_faces_conversion:
            face_texture = self.vert_to_face_texture(self.vertex_texture, discrete=self.is_discrete_texture())
            self.set_texture(face_texture, is_vertex_texture=False, delete_existing=False)
            return self.face_texture
        raise ValueError("Face texture not present and conversion was not requested")

    def vert_to_face_texture(self, vert_IDs, discrete=True):
        """Reference meshes.py:947-987: mean of the three vertex rows, or their most common value for discrete
        textures (the reference breaks 3-way ties at random; here the lowest value wins)."""
        if vert_IDs is None:
            raise ValueError("None")
        vert_IDs = np.squeeze(vert_IDs)
        if vert_IDs.ndim != 1 and discrete:
            raise ValueError(f"Can only perform discrete conversion with one dimensional array but instead had "
                             f"{vert_IDs.ndim}")
        per_face = np.asarray(vert_IDs, dtype=float)[self.faces]
        if not discrete:
            return np.mean(per_face, axis=1)
        a, b, c = per_face[:, 0], per_face[:, 1], per_face[:, 2]
        out = np.fmin(np.fmin(a, b), c)  # all different (or NaNs): lowest non-NaN value
        out = np.where(b == c, b, out)
        out = np.where((a == b) | (a == c), a, out)
        return out

    # ------------------------------------------------------------------------------------------------
    # Device context and local frame
    # ------------------------------------------------------------------------------------------------
    def get_mesh_hash(self):
        """sha256 of points + faces (reference meshes.py:1631-1639)."""
        hasher = hashlib.sha256()
        hasher.update(self.points.tobytes())
        hasher.update(self.faces.tobytes())
        return hasher.hexdigest()

    def _get_context(self):
        if self._context is None:
            self._context = _lib.Context(self.device)
        return self._context

    def get_mesh_in_cameras_coords(self, cameras, inplace: bool = False):
        """The mesh in the camera set's local frame, resident on the GPU (reference meshes.py:1641-1676).

        ``v_local = inv(local_to_epsg_4978) @ [v_ecef; 1]`` in float64 on the host, once per camera-set
        transform; the result minus its bounding-box centre is rounded to float32 and uploaded (the origin is
        folded back into every camera's translation in float64, ``_lib.make_camera``).
        """
        import torch

        T = cameras.get_local_to_epsg_4978_transform()
        T = np.eye(4) if T is None else np.asarray(T, dtype=float)
        key = T.tobytes()
        if self._local_cache is not None and self._local_cache.key == key:
            local = self._local_cache
        else:
            Tinv = np.linalg.inv(T)
            points = self.points @ Tinv[:3, :3].T + Tinv[:3, 3]
            origin = 0.5 * (points.min(axis=0) + points.max(axis=0))
            ctx = self._get_context()
            dev = torch.device("cuda", self.device)
            v32 = torch.from_numpy((points - origin).astype(np.float32)).to(dev)
            f32 = torch.from_numpy(self.faces).to(dev)
            ctx.set_mesh(v32, f32)
            local = LocalMesh(points, self.faces, origin, ctx, key)
            self._local_cache = local
        if inplace:
            self.points = local.points
            self.CRS = None
            return None
        return local

    # ------------------------------------------------------------------------------------------------
    # Camera conversion
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _camera_list(cameras):
        if isinstance(cameras, PhotogrammetryCamera) or not hasattr(cameras, "cameras"):
            return [cameras]
        return list(cameras.cameras)

    def _gg_cameras(self, cam_list, local: LocalMesh, scale: float, warp_follows: bool = False):
        """``warp_follows``: the raster will be pushed through the lens model afterwards.  The Metashape model takes an
        ideal raster centred at (W/2, H/2) and adds cx, cy itself (derived_cameras.py:171-208; the reference warps
        only the pyvista render, which has no principal point, cameras.py:449), so the principal point must not be
        applied twice."""
        sizes = {c.get_image_size(scale) for c in cam_list}
        if len(sizes) != 1:
            raise ValueError("Not all cameras have the same image size")  # derived_meshes.py:811-813
        pp = 1.0 if (self.use_principal_point and not warp_follows) else 0.0
        return [
            _lib.make_camera(c.world_to_cam_transform, c.f, pp * c.cx, pp * c.cy, c.image_width, c.image_height,
                             render_img_scale=scale, origin=local.origin)
            for c in cam_list
        ]

    @staticmethod
    def _is_camera_or_set(cameras):
        return isinstance(cameras, (PhotogrammetryCamera, PhotogrammetryCameraSet)) or (
            hasattr(cameras, "cameras") and hasattr(cameras, "get_local_to_epsg_4978_transform")
        ) or hasattr(cameras, "world_to_cam_transform")

    # ------------------------------------------------------------------------------------------------
    # pix2face
    # ------------------------------------------------------------------------------------------------
    def pix2face_device(self, cameras, mesh: LocalMesh = None, render_img_scale: float = 1, out=None,
                        warp_follows: bool = False):
        """pix2face as an (n, h, w) int32 CUDA tensor (-1 = no face); no host round trip.  ``warp_follows``: see
        ``_gg_cameras``."""
        import torch

        if mesh is None:
            mesh = self.get_mesh_in_cameras_coords(cameras)
        cam_list = self._camera_list(cameras)
        gg = self._gg_cameras(cam_list, mesh, render_img_scale, warp_follows=warp_follows)
        n, H, W = len(gg), gg[0].H, gg[0].W
        if out is None:
            out = torch.empty((n, H, W), dtype=torch.int32, device=torch.device("cuda", self.device))
        B = self.views_per_batch
        for s in range(0, n, B):
            mesh.context.rasterize(gg[s : s + B], out=out[s : s + B])
        return out

    def pix2face(
        self,
        cameras,
        mesh: typing.Optional[LocalMesh] = None,
        render_img_scale: float = 1,
        save_to_cache: bool = False,
        cache_folder: typing.Union[None, PATH_TYPE] = CACHE_FOLDER,
        distortion_set=None,
        apply_distortion: bool = True,
    ) -> np.ndarray:
        """For every pixel of every camera, the ID of the mesh face the pixel's ray hits first, -1 if none.

        Same contract as the reference (meshes.py:1678-1718): int64, shape (h, w) for a single camera and
        (n_cameras, h, w) for a camera set (also for a set of length one), with
        ``(h, w) = (int(H * scale), int(W * scale))``.  ``save_to_cache`` / ``cache_folder`` are accepted for
        signature compatibility and unused, like in the reference's PyTorch3D back-end
        (derived_meshes.py:665-668): recomputing on the GPU is cheaper than the disk cache.
        """
        if not self._is_camera_or_set(cameras):
            raise TypeError("cameras must be a PhotogrammetryCamera or a PhotogrammetryCameraSet")
        if mesh is None:
            mesh = self.get_mesh_in_cameras_coords(cameras)
        if distortion_set is None and apply_distortion:
            self.logger.warning("Distortion requested but no distortion parameters provided. Skipping")
            apply_distortion = False

        single = isinstance(cameras, PhotogrammetryCamera) or not hasattr(cameras, "cameras")
        p2f = self.pix2face_device(cameras, mesh=mesh, render_img_scale=render_img_scale, warp_follows=apply_distortion)
        if apply_distortion:
            p2f = self._warp_device(p2f, self._camera_list(cameras), distortion_set, render_img_scale)
        p2f = _lib.to_fresh_host(p2f.to(dtype=__import__("torch").int64))
        return p2f[0] if single else p2f

    @staticmethod
    def _warp_device(p2f, cam_list, distortion_set, scale):
        """Warp (n, h, w) int32 device rasters into the distorted image geometry (reference meshes.py:1842-1854:
        warp_dewarp_image(warped_to_ideal=False, fill_value=-1, interpolation_order=0)).  Camera sets with a GPU lens
        model (MetashapeCameraSet.warp_dewarp_device) never leave the device; any other set goes through its own
        warp_dewarp_image on the host, which raises NotImplementedError when there is no lens model
        (reference cameras.py:1088-1090)."""
        import torch

        if hasattr(distortion_set, "warp_dewarp_device"):
            return torch.stack([
                distortion_set.warp_dewarp_device(cam, p2f[i], warped_to_ideal=False, fill_value=-1, image_scale=scale)
                for i, cam in enumerate(cam_list)])
        host = _lib.to_fresh_host(p2f.to(torch.int64))
        out = np.stack([
            distortion_set.warp_dewarp_image(camera=cam, input_image=host[i], warped_to_ideal=False, fill_value=-1,
                                             interpolation_order=0, image_scale=scale)
            for i, cam in enumerate(cam_list)], axis=0)
        return torch.from_numpy(out.astype(np.int32)).to(p2f.device)

    # ------------------------------------------------------------------------------------------------
    # render_flat
    # ------------------------------------------------------------------------------------------------
    def render_flat_device(self, cameras, batch_size: int = None, render_img_scale: float = 1, out_dtype="float64",
                           check: bool = True):
        """Generator of (n_batch, h, w, d) CUDA tensors: the face texture seen from every camera.  ``check=False``
        leaves the launches asynchronous (the caller synchronises the context once at the end and handles a scratch
        overflow there); the first batch is always checked, which sizes the scratch."""
        import torch

        mesh = self.get_mesh_in_cameras_coords(cameras)
        cam_list = self._camera_list(cameras)
        face_texture = self.get_texture(request_vertex_texture=False, try_verts_faces_conversion=True)
        dev = torch.device("cuda", self.device)
        tex = torch.from_numpy(np.ascontiguousarray(face_texture, dtype=np.float64)).to(dev)
        code = {"float64": _lib.OUT_F64, "float32": _lib.OUT_F32, "uint8": _lib.OUT_U8}[out_dtype]
        gg = self._gg_cameras(cam_list, mesh, render_img_scale)
        B = self.views_per_batch if batch_size is None else max(1, min(batch_size, _lib.MAX_VIEWS_PER_CALL))
        for s in range(0, len(gg), B):  # fused: the face-ID rasters are never written
            yield mesh.context.rasterize_render_flat(gg[s : s + B], tex, out_dtype=code, check=check or s == 0)

    def render_flat(self, cameras, batch_size: int = 1, render_img_scale: float = 1, return_camera: bool = False,
                    **pix2face_kwargs):
        """Render the face texture from the viewpoint of every camera (reference meshes.py:1858-1942).

        Generator of (h, w, d) float64 arrays, NaN where no face is hit, optionally paired with the camera.
        Unlike the reference, trailing cameras are not dropped when ``len(cameras) % batch_size != 0``.
        The device-to-host copy of a batch runs on a copy stream while the next batch is rendered.  Extension
        (keyword ``out_dtype``, undistorted renders only): "float32", or "uint8" with ``save_renders``' cast rule
        (meshes.py:2323-2334) applied on the GPU -- label renders then cross PCIe at one byte per pixel.
        """
        out_dtype = pix2face_kwargs.pop("out_dtype", "float64")
        if isinstance(cameras, PhotogrammetryCamera):
            cameras = PhotogrammetryCameraSet([cameras])
        elif not self._is_camera_or_set(cameras) or not hasattr(cameras, "cameras"):
            raise TypeError("cameras must be a PhotogrammetryCamera or a PhotogrammetryCameraSet")
        apply_distortion = pix2face_kwargs.get("apply_distortion", True)
        distortion_set = pix2face_kwargs.get("distortion_set", None)
        if distortion_set is None and apply_distortion:
            apply_distortion = False
        cam_list = self._camera_list(cameras)
        if not apply_distortion:
            k = 0
            for host in self._to_host_owned(self.render_flat_device(cameras, batch_size, render_img_scale, out_dtype)):
                for img in host:
                    yield (img, cam_list[k]) if return_camera else img
                    k += 1
            return
        # distortion requested: rasterize, warp the face-ID raster into the distorted geometry, then gather
        import torch

        mesh = self.get_mesh_in_cameras_coords(cameras)
        face_texture = self.get_texture(request_vertex_texture=False, try_verts_faces_conversion=True)
        tex = torch.from_numpy(np.ascontiguousarray(face_texture, dtype=np.float64)).to(torch.device("cuda", self.device))
        for k, cam in enumerate(cam_list):
            d_p2f = self._pix2face_for_aggregation(cam, mesh, render_img_scale, pix2face_kwargs)
            img = mesh.context.render_flat(d_p2f[0].contiguous(), tex).cpu().numpy()
            yield (img, cam) if return_camera else img

    # ------------------------------------------------------------------------------------------------
    # project_images / aggregate_projected_images
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _classify_image(img: np.ndarray):
        """numpy prediction image -> (contiguous array, pred_kind, C)."""
        img = np.asarray(img)
        C = 1 if img.ndim == 2 else img.shape[-1]
        if img.dtype == np.bool_:
            return np.ascontiguousarray(img).view(np.uint8), _lib.PRED_U8, C
        if img.dtype == np.uint8:
            return np.ascontiguousarray(img), _lib.PRED_U8, C
        if img.dtype == np.float32:
            return np.ascontiguousarray(img), _lib.PRED_F32, C
        return np.ascontiguousarray(img, dtype=np.float64), _lib.PRED_F64, C

    def _flags(self, extra=0):
        return (_lib.FLAG_COMPAT_NEGATIVE_INDEX if self.compat_negative_index else 0) | extra

    def project_images(self, cameras, batch_size: int = 1, aggregate_img_scale: float = 1,
                       check_null_image: bool = False, **pix2face_kwargs):
        """Per view, the (n_faces, n_channels) float64 array holding, for every face the view sees, the value of
        the face's last pixel in row-major order, NaN elsewhere (reference meshes.py:1944-2002)."""
        import torch

        mesh = self.get_mesh_in_cameras_coords(cameras)
        cam_list = self._camera_list(cameras)
        dev = torch.device("cuda", self.device)
        F = self.faces.shape[0]
        for k, cam in enumerate(cam_list):
            img = cameras.get_image_by_index(k, aggregate_img_scale)
            arr, kind, C = self._classify_image(img)
            d_sum = torch.full((F, C), float("nan"), dtype=torch.float64, device=dev)
            if not check_null_image or np.any(np.isfinite(img)):
                d_count = torch.zeros((F,), dtype=torch.int32, device=dev)
                p2f = self._pix2face_for_aggregation(cam, mesh, aggregate_img_scale, pix2face_kwargs)
                mesh.context.aggregate(p2f[0], torch.from_numpy(arr).to(dev), kind, C, _lib.MODE_LAST_PIXEL,
                                       self._flags(_lib.FLAG_ASSIGN), d_sum, d_count)
            yield _lib.to_fresh_host(d_sum)

    def _pix2face_for_aggregation(self, cameras, mesh, scale, pix2face_kwargs):
        """Device raster for one camera / camera batch, honouring a requested distortion warp."""
        import torch

        apply_distortion = pix2face_kwargs.get("apply_distortion", True)
        distortion_set = pix2face_kwargs.get("distortion_set", None)
        warp = distortion_set is not None and apply_distortion
        p2f = self.pix2face_device(cameras, mesh=mesh, render_img_scale=scale, warp_follows=warp)
        if not warp:
            return p2f
        return self._warp_device(p2f, self._camera_list(cameras), distortion_set, scale)


    @staticmethod
    def _to_device_or_mapped(arr, dev, zero_copy=True):
        """Device tensor for a host prediction image.  The fused last-pixel / vote aggregation reads ONE pixel per
        visible face, so an image that already sits in page-locked host memory is not copied at all: CUDA's unified
        addressing lets the kernel fetch those few rows over PCIe straight from the host buffer.  Pageable arrays
        are uploaded."""
        import torch

        if zero_copy and _lib.pointer_kind(arr) in (_lib.POINTER_PINNED, _lib.POINTER_MANAGED):
            return _HostMapped(arr)
        return torch.from_numpy(arr).to(dev, non_blocking=True)

    def _to_host(self, *tensors):
        """Device tensors -> fresh NumPy arrays owned by the caller.  Every array is backed by its own page-locked
        buffer, so the device-to-host copy runs at the full PCIe rate and is the ONLY copy (no second pass from a
        staging buffer into pageable memory, which used to cost more than the transfer itself).  The buffers come
        from torch's caching host allocator: they go back to its pool when the caller drops the arrays, and a later
        call reuses them."""
        import torch

        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
        for h, t in zip(host, tensors):
            h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [h.numpy() for h in host]  # the array keeps its tensor (and thereby the buffer) alive

    # -- pageable host images: the GPU lists the pixels it needs, the host gathers them -----------------------------
    def _pinned(self, name, shape, dtype):
        """Grow-only page-locked staging tensor kept on the mesh object."""
        import torch

        stage = self.__dict__.setdefault("_sparse_stage", {})
        need = int(np.prod(shape))
        buf = stage.get(name)
        if buf is None or buf.dtype != dtype or buf.numel() < need:
            # page-locking is slow (~1 GB/s): grow geometrically so that slowly growing batches do not re-pin each time
            grown = need if buf is None or buf.dtype != dtype else max(need, buf.numel() * 3 // 2)
            buf = stage[name] = torch.empty((max(grown, 1),), dtype=dtype, pin_memory=True)
        return buf[:need].view(*shape)

    def _sparse_begin(self, ctx, gg, flags, serial, full=False):
        """First half of a batch whose prediction images sit in ordinary (pageable) NumPy arrays.  Uploading a 20-Mpx
        score image costs ~100x the path itself, and the aggregation only needs the row of each visible face's last
        pixel, so: rasterize and list (face, pixel) per view on the GPU (gg_project_winners) and start copying the
        lists to the host on a side stream.  The lists are short and about as long as the previous batch's, so they are
        given that much room (a view that needs more makes ``_sparse_finish`` redo the batch with full-size lists) and
        lists and counts come back in ONE transfer.  Returns the state ``_sparse_finish`` completes; the caller begins
        batch k+1 BEFORE finishing batch k, so the host picks rows while the GPU rasterizes."""
        import torch

        sides = self.__dict__.get("_sparse_streams")
        if sides is None:
            sides = self.__dict__["_sparse_streams"] = [torch.cuda.Stream(device=self.device) for _ in range(2)]
        n = len(gg)
        guess = int(self.__dict__.get("_sparse_guess", 0))
        full = full or guess <= 0
        pairs, counts = ctx.project_winners(gg, flags | (0 if full else _lib.FLAG_TRUNCATE), cap=None if full else guess)
        listed = torch.cuda.Event()
        listed.record()
        slot = serial % 2
        side = sides[slot]  # one copy stream per slot: the lists of batch k must not queue behind "batch k+1 listed"
        # (the host finished reading this slot's list buffers two batches ago, in _sparse_finish)
        h_counts = self._pinned(f"counts{slot}", (n,), torch.int32)
        h_lists = self._pinned(f"lists{slot}", tuple(pairs.shape), torch.int32)
        with torch.cuda.stream(side):
            side.wait_event(listed)
            h_counts.copy_(counts, non_blocking=True)
            h_lists.copy_(pairs, non_blocking=True)
            arrived = torch.cuda.Event()
            arrived.record(side)
        return dict(n=n, gg=gg, pairs=pairs, counts=counts, h_counts=h_counts, h_lists=h_lists, arrived=arrived,
                    slot=slot, serial=serial, full=full)

    def _sparse_finish(self, ctx, st, arrays, kind, C, mode, flags, d_sum, d_count):
        """Second half: pick the listed rows out of the arrays on the host's cores (gg_gather_rows_host), send them
        back and apply them view by view (gg_accumulate_rows) -- the same arithmetic, in the same order, as the fused
        path."""
        import torch

        n, pairs, slot = st["n"], st["pairs"], st["slot"]
        E = 1 if (mode == _lib.MODE_VOTE or kind == _lib.PRED_INDEX_U8) else C  # elements per pixel
        st["arrived"].synchronize()
        m = [int(x) for x in st["h_counts"].tolist()]
        cap = pairs.shape[1]
        # room for the next batches' lists: a quarter more than the longest seen, never shrinking (stable buffer sizes)
        self.__dict__["_sparse_guess"] = max(int(self.__dict__.get("_sparse_guess", 0)),
                                             -(-(int(max(m, default=0) * 1.25) + 1024) // 8192) * 8192)
        if max(m, default=0) > cap:
            if st["full"]:  # the scratch itself overflowed: reported by the next sync, then replayed
                return
            redo = self._sparse_begin(ctx, st["gg"], flags, st["serial"], full=True)
            return self._sparse_finish(ctx, redo, arrays, kind, C, mode, flags, d_sum, d_count)
        offs = np.concatenate([[0], np.cumsum(m)]).astype(np.int64)
        total = int(offs[-1])
        if total == 0:
            return
        np_dtype = arrays[0].dtype
        t_dtype = torch.from_numpy(np.empty(0, dtype=np_dtype)).dtype
        busy = self.__dict__.setdefault("_sparse_busy", [None, None])
        if busy[slot] is not None:  # the upload that last read this slot's row buffer (two batches ago)
            busy[slot].synchronize()
        h_rows = self._pinned(f"rows{slot}_" + str(np_dtype), (total, E), t_dtype)
        # every core picks rows (native threads with software prefetch: np.take holds the GIL)
        _lib.gather_rows_host([np.ascontiguousarray(a) for a in arrays], st["h_lists"].numpy().reshape(-1, 2), offs,
                              h_rows.numpy(), pair_starts=np.arange(n, dtype=np.int64) * cap)
        d_rows = h_rows.to(d_sum.device, non_blocking=True)
        busy[slot] = torch.cuda.Event()
        busy[slot].record()
        for v in range(n):  # view order = the reference's summation order
            if m[v]:
                ctx.accumulate_rows(pairs[v], m[v], d_rows[offs[v] : offs[v + 1]], kind, C, mode, flags, d_sum, d_count)

    def _fetch_prediction(self, cameras, k, scale, image_getter, index_getter):
        """(array, pred_kind, C) of view k: a caller-supplied getter, else the segmentor's class-index image when
        it offers one (expanded on the GPU), else whatever get_image_by_index returns."""
        if image_getter is not None:
            return self._classify_image(image_getter(k))
        if index_getter is not None:
            inds = index_getter(k, scale)
            if inds is not None:
                return np.ascontiguousarray(inds), _lib.PRED_INDEX_U8, cameras.n_image_channels()
        return self._classify_image(cameras.get_image_by_index(k, scale))

    def _accumulate_views(self, cameras, aggregate_img_scale, mode, n_channels=None, pix2face_kwargs=None,
                          image_getter=None, single_view_total=None):
        """Shared driver of the aggregation variants: streams every view's prediction image to the GPU and runs
        rasterize + aggregate there.  Returns (d_sum, d_count, C) still on the device."""
        import torch

        pix2face_kwargs = pix2face_kwargs or {}
        mesh = self.get_mesh_in_cameras_coords(cameras)
        cam_list = self._camera_list(cameras)
        dev = torch.device("cuda", self.device)
        F = self.faces.shape[0]
        apply_distortion = pix2face_kwargs.get("apply_distortion", True) and pix2face_kwargs.get("distortion_set") is not None
        n = len(cam_list)
        single = (n == 1) if single_view_total is None else single_view_total
        flags = self._flags(_lib.FLAG_KEEP_NAN if (single and mode == _lib.MODE_LAST_PIXEL) else 0)
        C = n_channels
        B = 1 if apply_distortion else self.views_per_batch
        index_getter = getattr(cameras, "get_class_index_image_by_index", None) if mode == _lib.MODE_LAST_PIXEL else None
        ctx = mesh.context
        window = 4  # batches in flight between two synchronisations (bounds the memory held by queued predictions)
        for attempt in range(4):
            d_sum = d_count = None
            in_flight = []
            pending = None  # a batch of pageable images whose rows the host has yet to pick
            pinned_seen = {}
            fetch = self._prediction_fetcher(cameras, n, aggregate_img_scale, image_getter, index_getter)
            try:
                for bi, s in enumerate(range(0, n, B)):
                    batch = cam_list[s : s + B]
                    preds, kind = [], None
                    for k in range(s, s + len(batch)):
                        arr, this_kind, this_C = fetch(k)
                        if mode == _lib.MODE_VOTE:
                            this_C = n_channels
                        if C is None:
                            C = this_C
                        if d_sum is None:
                            d_sum = torch.zeros((F, C), dtype=torch.float64, device=dev)
                            d_count = torch.zeros((F,), dtype=torch.int32, device=dev)
                        if kind is None:
                            kind = this_kind
                        elif kind != this_kind:
                            raise ValueError("all prediction images of a batch must share one dtype / layout")
                        preds.append(arr)
                    kinds = [_lib.pointer_kind(a) for a in preds]
                    for a, pk in zip(preds, kinds):  # footprint of the distinct page-locked images seen so far
                        if pk == _lib.POINTER_PINNED and a.ctypes.data not in pinned_seen:
                            pinned_seen[a.ctypes.data] = a.nbytes
                    pinned_large = sum(pinned_seen.values()) > self.pinned_direct_limit_bytes
                    host_gatherable = all(pk == _lib.POINTER_PAGEABLE or (pk == _lib.POINTER_PINNED and pinned_large)
                                          for pk in kinds)
                    sparse = (not apply_distortion and self.sparse_host_gather and mode != _lib.MODE_PIXEL_SUM
                              and len({a.dtype for a in preds}) == 1 and host_gatherable)
                    if sparse:
                        if in_flight:  # the fused calls queued so far use the library's own streams
                            ctx.sync()
                            in_flight.clear()
                        # two-deep: this batch is rasterized while the host picks the rows of the previous one
                        began = self._sparse_begin(ctx, self._gg_cameras(batch, mesh, aggregate_img_scale), flags, bi)
                        if pending is not None:
                            self._sparse_finish(ctx, *pending, mode, flags, d_sum, d_count)
                        pending = (began, preds, kind, C)
                        continue
                    if pending is not None:
                        self._sparse_finish(ctx, *pending, mode, flags, d_sum, d_count)
                        pending = None
                    preds = [self._to_device_or_mapped(a, dev, zero_copy=not apply_distortion) for a in preds]
                    if apply_distortion:
                        p2f = self._pix2face_for_aggregation(batch[0], mesh, aggregate_img_scale, pix2face_kwargs)
                        ctx.aggregate(p2f[0], preds[0], kind, C, mode, flags, d_sum, d_count)
                        continue
                    gg = self._gg_cameras(batch, mesh, aggregate_img_scale)
                    # The calls are asynchronous: batch k+1 is binned while batch k is rasterized and batch k-1 resolved
                    # on the library's internal streams.  Queued predictions are kept alive until the next sync.
                    ctx.project_aggregate(gg, preds, kind, C, mode, flags, d_sum, d_count, check=False)
                    in_flight.append(preds)
                    # Synchronise now and then: it bounds the device memory held by queued (uploaded) predictions and
                    # surfaces a scratch overflow early.  Page-locked host images hold no device memory, so their
                    # queue may run much deeper -- every synchronisation is a bubble in the pipeline.
                    uploaded = sum(1 for b in in_flight for q in b if not isinstance(q, _HostMapped))
                    if uploaded >= window * B or len(in_flight) >= 8 * window:
                        ctx.sync()
                        in_flight.clear()
                if pending is not None:
                    self._sparse_finish(ctx, *pending, mode, flags, d_sum, d_count)
                    pending = None
                ctx.sync()  # raises GG_ERR_OVERFLOW if any batch since the last sync outgrew the scratch
                break
            except _lib.GeograypherB200Error as e:
                # An overflowing batch is skipped as a whole, so the accumulators are consistent but incomplete: grow
                # the scratch and redo the aggregation.
                if e.code != _lib.ERR_OVERFLOW or attempt == 3:
                    self._quiesce(ctx)
                    raise
                ctx._grow_after_overflow()
                C = n_channels
            except BaseException:
                # queued batches still read the prediction tensors and write d_sum / d_count on the library's own
                # streams: they must finish before those tensors go back to torch's allocator
                self._quiesce(ctx)
                raise
            finally:
                close = getattr(fetch, "close", None)
                if close is not None:
                    close()
        return d_sum, d_count, C

    def _prediction_fetcher(self, cameras, n, scale, image_getter, index_getter):
        """``fetch(k) -> (array, pred_kind, C)`` for the views in order.  Predictions that are DECODED per view (a
        segmentor that declares ``io_bound = True`` such as LookUpSegmentor, or a plain camera set reading image files)
        are read ahead on ``self.prefetch_threads`` workers (utils/prefetch.py): a serial 20-Mpx PNG decode takes a
        thousand times longer than the GPU needs for the view.  In-memory predictions (ArraySegmentor, an
        ``image_getter``) and segmentors that say nothing (they may run a network and need not be thread-safe) are
        fetched inline."""
        def one(k):
            return self._fetch_prediction(cameras, k, scale, image_getter, index_getter)

        threads = getattr(self, "prefetch_threads", None)
        if threads is None:  # auto
            segmentor = getattr(cameras, "segmentor", None)
            io_bound = (getattr(segmentor, "io_bound", False) if segmentor is not None
                        else getattr(cameras, "reads_image_files", False))
            threads = _prefetch.default_threads() if (io_bound and image_getter is None and n > 1) else 0
        if threads <= 0:
            return one
        return _prefetch.OrderedPrefetcher(one, n, threads=threads, size_of=lambda item: item[0].nbytes)

    @staticmethod
    def _quiesce(ctx):
        """Wait for everything the context has in flight; a pending overflow report is dropped (the caller is
        already unwinding with another error)."""
        try:
            ctx.sync()
        except _lib.GeograypherB200Error:
            pass

    def aggregate_projected_images(self, cameras, batch_size: int = 1, aggregate_img_scale: float = 1,
                                   return_all: bool = False, return_argmax: bool = False, **kwargs):
        """Aggregate the imagery from multiple cameras into per-face averages (reference meshes.py:2004-2084).

        Per view every visible face takes the value of its last pixel (row-major); values are summed over views
        with NaN counted as 0; a face is counted once per view in which any of its channels is finite; the mean is
        sum / count and faces never seen are NaN.  Sums are accumulated in float64 in view order, so for a single
        GPU the result is bit-identical to the reference's ``np.nansum`` loop.

        Returns ``(average (F, C) float64, {"projection_counts": (F,) float64, "summed_projections": (F, C)})``;
        with ``return_all`` the dict also holds ``all_projections`` (per-view arrays, as slow and as large as in
        the reference); with ``return_argmax`` (*new*) it holds ``argmax`` = find_argmax_nonzero_value(average)
        computed on the GPU.  ``batch_size`` is accepted for compatibility; batching is an internal detail here,
        and no camera is dropped when ``len(cameras) % batch_size != 0`` (the reference does, meshes.py:1976-1979).
        """
        del batch_size
        info = {}
        if return_all:
            info["all_projections"] = list(
                self.project_images(cameras, aggregate_img_scale=aggregate_img_scale, **kwargs)
            )
        pix2face_kwargs = {k: v for k, v in kwargs.items() if k != "check_null_image"}
        d_sum, d_count, _ = self._accumulate_views(cameras, aggregate_img_scale, _lib.MODE_LAST_PIXEL,
                                                   pix2face_kwargs=pix2face_kwargs)
        ctx = self._get_context()
        avg, argmax = ctx.finalize(d_sum, d_count, want_avg=True, want_argmax=return_argmax)
        h_avg, h_sum, h_count = self._to_host(avg, d_sum, d_count.double())
        info["projection_counts"] = h_count
        info["summed_projections"] = h_sum
        if return_argmax:
            info["argmax"] = argmax.cpu().numpy()
        return h_avg, info

    # ------------------------------------------------------------------------------------------------
    # label_polygons
    # ------------------------------------------------------------------------------------------------
    # ------------------------------------------------------------------------------------------------
    # save_renders
    # ------------------------------------------------------------------------------------------------
    def save_IDs_to_labels(self, savepath):
        """Write the ID -> class-name mapping next to the renders (reference meshes.py:2225-2246)."""
        import json

        Path(savepath).parent.mkdir(parents=True, exist_ok=True)
        IDs_to_labels = self.get_IDs_to_labels() or {}
        with open(savepath, "w") as f:
            json.dump({str(int(k)): (v if isinstance(v, str) else float(v)) for k, v in IDs_to_labels.items()}, f,
                      ensure_ascii=False, indent=4)

    def _to_host_pipelined(self, device_batches):
        """Device batches -> host arrays through TWO page-locked staging buffers and a copy stream: the device-to-host
        copy of batch k runs while batch k+1 is rendered, and the consumer works on batch k-1.  Yields
        ``(host_array_view, release)``: the view is only valid until ``release()`` is called, after which its buffer
        receives a later batch."""
        import threading

        import torch

        copy_stream = self.__dict__.get("_copy_stream")
        if copy_stream is None:
            copy_stream = self.__dict__["_copy_stream"] = torch.cuda.Stream(device=self.device)
        slots = [None, None]          # pinned tensors
        events = [None, None]
        free = [threading.Semaphore(1), threading.Semaphore(1)]
        pending = None
        for k, batch in enumerate(device_batches):
            slot = k % 2
            free[slot].acquire()  # the consumer of the batch that used this buffer two rounds ago has let go of it
            if slots[slot] is None or slots[slot].shape != batch.shape or slots[slot].dtype != batch.dtype:
                slots[slot] = torch.empty(batch.shape, dtype=batch.dtype, pin_memory=True)
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                slots[slot].copy_(batch, non_blocking=True)
                events[slot] = torch.cuda.Event()
                events[slot].record(copy_stream)
            batch.record_stream(copy_stream)
            if pending is not None:
                events[pending].synchronize()
                yield slots[pending].numpy(), free[pending].release
            pending = slot
        if pending is not None:
            events[pending].synchronize()
            yield slots[pending].numpy(), free[pending].release

    def _to_host_owned(self, device_batches):
        """Device batches -> host arrays that the CONSUMER owns (``render_flat``'s contract: it may keep every image).
        Each batch gets its own page-locked block from torch's caching host allocator, filled on a copy stream while
        the next batch is rendered; a consumer that drops an image before asking for the next -- the usual loop --
        keeps re-using the same two or three blocks, one that keeps them pays the page-locking of each."""
        import torch

        copy_stream = self.__dict__.get("_copy_stream")
        if copy_stream is None:
            copy_stream = self.__dict__["_copy_stream"] = torch.cuda.Stream(device=self.device)
        pending = None  # (page-locked tensor, event of its copy)
        for batch in device_batches:
            host = torch.empty(batch.shape, dtype=batch.dtype, pin_memory=True)
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                host.copy_(batch, non_blocking=True)
                event = torch.cuda.Event()
                event.record(copy_stream)
            batch.record_stream(copy_stream)
            del batch
            if pending is not None:
                pending[1].synchronize()
                done, pending = pending[0], (host, event)
                yield done.numpy()
                del done
            else:
                pending = (host, event)
            del host
        if pending is not None:
            pending[1].synchronize()
            yield pending[0].numpy()

    def save_renders(self, camera_set, render_image_scale=1.0, output_folder="renders", make_composites: bool = False,
                     save_native_resolution: bool = False, cast_to_uint8: bool = True, save_as_npy: bool = False,
                     uint8_value_for_null_texture=NULL_TEXTURE_INT_VALUE, n_writer_threads: int = 8, **render_kwargs):
        """Render the face texture from every camera and write one file per image (reference meshes.py:2248-2397).

        With ``cast_to_uint8`` the cast rule of the reference (< 0, > 255 or non-finite -> the null value, then
        truncation, meshes.py:2323-2334) is applied on the GPU and only one byte per pixel and channel crosses PCIe.
        ``save_native_resolution`` up-samples scaled renders to the camera's native size on the GPU like the reference
        does on the host (nearest neighbour for discrete textures, bilinear otherwise, meshes.py:2312-2321).  The
        device-to-host copies go through two page-locked buffers on a copy stream (the copy of one batch overlaps the
        rendering of the next) and the files are written by a small thread pool, so neither PCIe nor the disk stalls
        the GPU.  Outputs keep the image's path relative to ``camera_set.image_folder`` (meshes.py:2349-2366) and are
        ``.npy`` arrays (``save_as_npy``) or deflate-compressed TIFFs (needs Pillow).  Composites with the photographs
        are a visualisation feature outside this build.  Returns the number of bytes handed to the writers.
        """
        if make_composites:
            raise NotImplementedError("composites with the photographs are a visualisation step outside this build")
        if uint8_value_for_null_texture != 0 and cast_to_uint8:
            raise NotImplementedError("only NULL_TEXTURE_INT_VALUE = 0 is supported for the fused uint8 cast")
        from concurrent.futures import ThreadPoolExecutor

        output_folder = Path(output_folder)
        output_folder.mkdir(parents=True, exist_ok=True)
        self.logger.info(f"Saving renders to {output_folder}")
        self.save_IDs_to_labels(Path(output_folder, "IDs_to_labels.json"))
        cam_list = self._camera_list(camera_set)

        def write(path, array):
            path.parent.mkdir(parents=True, exist_ok=True)
            array = np.squeeze(array)
            if save_as_npy:
                np.save(str(path.with_suffix(".npy")), array)
                return
            try:
                from PIL import Image
            except ImportError as e:
                raise ImportError("writing TIFFs needs Pillow; use save_as_npy=True") from e
            if not cast_to_uint8:  # uint16 when it fits, else uint32 (reference meshes.py:2381-2388)
                array = array.astype(np.uint16 if np.nanmax(array) <= np.iinfo(np.uint16).max else np.uint32)
            if array.ndim == 3:
                array = array[..., :3]
            Image.fromarray(array).save(str(path.with_suffix(".tif")), compression="tiff_adobe_deflate")

        apply_distortion = render_kwargs.pop("apply_distortion", True)
        if apply_distortion and not hasattr(camera_set, "warp_dewarp_device"):
            apply_distortion = False if render_kwargs.get("distortion_set") is None else apply_distortion
        upsample = bool(save_native_resolution) and render_image_scale != 1
        distorted = apply_distortion and hasattr(camera_set, "warp_dewarp_device")
        render_dtype = "uint8" if (cast_to_uint8 and not upsample) else "float64"

        def device_batches():
            gen = (self._render_flat_distorted_device(camera_set, render_image_scale, render_dtype == "uint8")
                   if distorted else self.render_flat_device(camera_set, None, render_image_scale, render_dtype, check=False))
            if not upsample:
                yield from gen
                return
            import torch

            order = 0 if self.is_discrete_texture() else 1  # meshes.py:2315-2320
            code = _lib.OUT_U8 if cast_to_uint8 else _lib.OUT_F64
            k = 0
            for batch in gen:
                out = []
                for img in batch:
                    h_native, w_native = cam_list[k].get_image_size()
                    out.append(_lib.resize_render(img, h_native, w_native, order, code))
                    k += 1
                yield torch.stack(out)

        ctx = self._get_context()
        for attempt in range(4):
            k, n_bytes = 0, 0
            try:
                with ThreadPoolExecutor(max_workers=max(1, n_writer_threads)) as pool:
                    all_futures = []
                    for host, release in self._to_host_pipelined(device_batches()):
                        futures = []
                        for img in host:
                            cam = cam_list[k]
                            try:
                                rel = Path(cam.get_image_filename()).relative_to(camera_set.image_folder)
                            except (ValueError, TypeError):
                                raise ValueError(
                                    f"Tried to find the relative path of the camera path ({cam.get_image_filename()}) "
                                    f"inside of the camera set image folder ({camera_set.image_folder}), but failed.")
                            futures.append(pool.submit(write, Path(output_folder, rel), img))
                            n_bytes += img.nbytes
                            k += 1
                        # the staging buffer goes back to the pipeline once ITS files are written (the pool is FIFO:
                        # the writes above start before this task does)
                        pool.submit(lambda fs=futures, rel_=release: ([f.exception() for f in fs], rel_()))
                        all_futures += futures
                    pool.shutdown(wait=True)
                    for f in all_futures:
                        f.result()
                ctx.sync()  # unchecked launches: a scratch overflow of any batch shows up here
                return n_bytes
            except _lib.GeograypherB200Error as e:
                if e.code != _lib.ERR_OVERFLOW or attempt == 3:
                    raise
                ctx._grow_after_overflow()  # some renders were skipped: grow the scratch and write everything again

    def _render_flat_distorted_device(self, cameras, scale, cast_to_uint8):
        """render_flat through the lens model, on the device: rasterize, warp the face-ID raster, gather."""
        import torch

        mesh = self.get_mesh_in_cameras_coords(cameras)
        tex = torch.from_numpy(np.ascontiguousarray(self.get_texture(request_vertex_texture=False), dtype=np.float64)).to(
            torch.device("cuda", self.device))
        code = _lib.OUT_U8 if cast_to_uint8 else _lib.OUT_F64
        for cam in self._camera_list(cameras):
            p2f = self._pix2face_for_aggregation(cam, mesh, scale, {"distortion_set": cameras, "apply_distortion": True})
            yield mesh.context.render_flat(p2f.contiguous(), tex, out_dtype=code)

    @staticmethod
    def _polygon_rings(polygons, with_holes=False):
        """Normalise the accepted polygon containers to a list (per polygon) of lists of (K, 2) rings: arrays, dicts
        {"exterior", "holes"}, shapely-like (Multi)Polygons, or anything with a ``.geometry`` column of those.  With
        ``with_holes`` also returns, per polygon, which rings are holes."""
        geoms = getattr(polygons, "geometry", polygons)
        out, holes = [], []
        for g in geoms:
            if hasattr(g, "geoms"):  # MultiPolygon: rings of all parts, even-odd rule
                parts = list(g.geoms)
            else:
                parts = [g]
            rings, is_hole = [], []
            for part in parts:
                if hasattr(part, "exterior"):
                    rings.append(np.asarray(part.exterior.coords)[:, :2])
                    inner = [np.asarray(i.coords)[:, :2] for i in part.interiors]
                elif isinstance(part, dict):
                    rings.append(np.asarray(part["exterior"], dtype=float))
                    inner = [np.asarray(h, dtype=float) for h in part.get("holes", [])]
                else:
                    rings.append(np.asarray(part, dtype=float))
                    inner = []
                rings.extend(inner)
                is_hole += [False] + [True] * len(inner)
            out.append(rings)
            holes.append(is_hole)
        return (out, holes) if with_holes else out

    def label_polygons(self, face_labels, polygons, face_weighting=None, sjoin_overlay=True,
                       return_class_labels=True, unknown_class_label="unknown", buffer_dist_meters=2.0,
                       vertex_xy=None):
        """Assign a class to every polygon from per-face labels (reference meshes.py:1141-1306).

        ``sjoin_overlay=True`` (default, :1259-1261): every face with a finite label whose 2-D triangle lies WITHIN a
        polygon votes for its class with weight ``area3D * face_weighting``.  ``sjoin_overlay=False`` (:1263-1268,
        ``polygons.overlay(faces, how="identity")``): faces are split along the polygon boundaries and every piece
        votes with ``area2D(piece) * (area3D / area2D)(face) * face_weighting``, so partially covered faces count in
        proportion.  The polygon gets the class with the largest total (lowest class ID on ties), NaN /
        ``unknown_class_label`` when nothing voted.  Both run on the GPU (``gg_label_polygons[_overlay]``).

        ``polygons``: sequence of (K, 2) exterior rings, dicts ``{"exterior", "holes"}``, shapely-like polygons, or a
        GeoDataFrame-like object -- in the SAME planar coordinates as ``vertex_xy``.  ``vertex_xy`` (*new*, (V, 2)):
        planar coordinates of the mesh vertices; defaults to the x, y of the stored vertices, which is right for meshes
        kept in a local metric frame (the reference reprojects the mesh to the polygons' CRS with pyproj,
        meshes.py:1194-1209, which is outside this build: pass the reprojected coordinates).  Not reproduced: the
        1e-6 precision snapping of both layers (:1222-1227) and the pre-filter that drops faces reaching more than
        ``buffer_dist_meters`` beyond the dissolved polygons (:1236-1253) -- faces of a survey mesh are far smaller
        than that buffer.
        """
        del buffer_dist_meters
        face_labels = np.squeeze(np.asarray(face_labels, dtype=float))
        if face_labels.ndim != 1:
            raise ValueError(f"Faces labels must be one-dimensional, but is {face_labels.ndim}")
        if face_weighting is not None:
            face_weighting = np.squeeze(np.asarray(face_weighting, dtype=float))
            if face_weighting.ndim != 1:
                raise ValueError(f"Faces labels must be one-dimensional, but is {face_weighting.ndim}")
        rings, holes = self._polygon_rings(polygons, with_holes=True)
        xy = self.points[:, :2] if vertex_xy is None else np.asarray(vertex_xy, dtype=float)
        finite = face_labels[np.isfinite(face_labels)]
        n_classes = int(finite.max()) + 1 if len(finite) else 1
        weights = _lib.label_polygons_weights(self.points, xy, self.faces, face_labels, face_weighting, rings,
                                              n_classes, self.device, overlay=not sjoin_overlay, holes=holes)
        best = weights.max(axis=1)
        predicted = np.where(best > 0, weights.argmax(axis=1).astype(float), np.nan).tolist()
        IDs_to_labels = self.get_IDs_to_labels()
        if return_class_labels and IDs_to_labels is not None:
            predicted = [(IDs_to_labels[int(p)] if np.isfinite(p) else unknown_class_label) for p in predicted]
        return predicted
