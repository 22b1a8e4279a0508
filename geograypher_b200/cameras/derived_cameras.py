"""Calibrated camera sets (reference: geograypher/cameras/derived_cameras.py).

``MetashapeCameraSet`` reads Agisoft Metashape's camera XML (sensors with the Brown-model calibration, the chunk's
local -> ECEF transform, per-image poses) and provides the lens-distortion model that ``pix2face`` applies to its
face-ID rasters (``apply_distortion=True``, reference meshes.py:1842-1854).  The warp itself runs on the GPU.
"""
from __future__ import annotations

import typing
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

from geograypher_b200.cameras.cameras import PhotogrammetryCamera, PhotogrammetryCameraSet
from geograypher_b200.constants import PATH_TYPE
from geograypher_b200.utils.parsing import parse_sensors, parse_transform_metashape

_DISTORTION_KEYS = ("b1", "b2", "k1", "k2", "k3", "k4", "p1", "p2")


def _collect_camera(camera, image_folder, c2ws, filenames, sensor_ids, original_image_folder, active_component_id):
    """One <camera> element -> pose, image path, sensor id; unaligned cameras and cameras of other components are
    skipped (reference derived_cameras.py:15-48)."""
    transform = camera.find("transform")
    if transform is None:
        return
    if active_component_id is not None and camera.get("component_id") != active_component_id:
        return
    c2ws.append(np.array(transform.text.split(), dtype=float).reshape(4, 4))
    name = Path(camera.get("label"))
    if original_image_folder is not None:
        name = name.relative_to(original_image_folder)
    filenames.append(Path(image_folder, name))
    sensor_ids.append(int(camera.get("sensor_id")))


class MetashapeCameraSet(PhotogrammetryCameraSet):
    def __init__(self, camera_file: PATH_TYPE, image_folder: PATH_TYPE,
                 original_image_folder: typing.Optional[PATH_TYPE] = None, validate_images: bool = False,
                 default_sensor_params: dict = {"cx": 0.0, "cy": 0.0}):
        """Parse intrinsics and extrinsics from a Metashape ``.xml`` export (reference derived_cameras.py:52-161).

        ``lon_lats`` are derived from the optimised poses when pyproj is available, otherwise left unset (they are
        not used by the projection path).
        """
        chunk = ET.parse(camera_file).getroot().find("chunk")
        sensors = parse_sensors(chunk.find("sensors"), default_sensor_dict=default_sensor_params)
        chunk_to_epsg4978, active_component_id = parse_transform_metashape(camera_file, return_component_id=True)

        c2ws, filenames, sensor_ids = [], [], []
        for cam_or_group in chunk.find("cameras"):
            members = cam_or_group if cam_or_group.tag == "group" else [cam_or_group]
            for cam in members:
                _collect_camera(cam, image_folder, c2ws, filenames, sensor_ids, original_image_folder,
                                active_component_id)

        lon_lats = None
        if chunk_to_epsg4978 is not None and len(c2ws) > 0:
            try:
                import pyproj

                ecef = np.stack([(chunk_to_epsg4978 @ T[:, 3])[:3] for T in c2ws])
                tf = pyproj.Transformer.from_crs("EPSG:4978", "EPSG:4326")
                lat, lon, _ = tf.transform(xx=ecef[:, 0], yy=ecef[:, 1], zz=ecef[:, 2])
                lon_lats = list(zip(lon, lat))
            except ImportError:
                lon_lats = None

        super().__init__(
            cam_to_world_transforms=c2ws,
            intrinsic_params_per_sensor_type=sensors,
            image_filenames=filenames,
            lon_lats=lon_lats,
            image_folder=image_folder,
            sensor_IDs=sensor_ids,
            validate_images=validate_images,
            local_to_epsg_4978_transform=chunk_to_epsg4978,
        )

    # ------------------------------------------------------------------------------------------------
    # Lens model
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _coefficients(camera):
        params = sorted(camera.distortion_params.keys())
        if not set(params) <= set(_DISTORTION_KEYS):
            raise ValueError(f"Unexpected distortion params found: {params}")
        d = camera.distortion_params
        return {k: (d["k1"] if k == "k1" else d.get(k, 0)) for k in _DISTORTION_KEYS}  # k1 is required

    def ideal_to_warped(self, camera: PhotogrammetryCamera, xpix: np.ndarray, ypix: np.ndarray):
        """Metashape's frame-camera model: pixel coordinates of an ideal pinhole image -> pixel coordinates in the
        distorted image (reference derived_cameras.py:163-208).  The principal-point offsets enter only at the end."""
        c = self._coefficients(camera)
        x = (np.asarray(xpix, dtype=float) - camera.image_width / 2.0) / camera.f
        y = (np.asarray(ypix, dtype=float) - camera.image_height / 2.0) / camera.f
        r2 = x * x + y * y
        radial = 1 + c["k1"] * r2 + c["k2"] * r2**2 + c["k3"] * r2**3 + c["k4"] * r2**4
        xd = x * radial + (c["p1"] * (r2 + 2 * x * x) + 2 * c["p2"] * x * y)
        yd = y * radial + (c["p2"] * (r2 + 2 * y * y) + 2 * c["p1"] * x * y)
        xw = camera.image_width / 2.0 + camera.cx + xd * camera.f + xd * c["b1"] + yd * c["b2"]
        yw = camera.image_height / 2.0 + camera.cy + yd * camera.f
        return xw, yw

    def _gg_distortion(self, camera, image_scale):
        from geograypher_b200 import _lib

        c = self._coefficients(camera)
        return _lib.make_distortion(camera.f, camera.cx, camera.cy, camera.image_width, camera.image_height,
                                    image_scale=image_scale, **c)

    def warp_source_index(self, camera, image_scale: float = 1.0, warped_to_ideal: bool = False, device: int = 0):
        """(h, w) int32 CUDA tensor: for every pixel of the output image the linear index of the nearest source pixel,
        -1 where the source falls outside the image.  Built on the GPU (exact forward model, or its Newton inverse)
        once per (distortion parameters, scale, direction) and cached, like the reference caches its maps
        (cameras.py:1055-1062)."""
        import torch

        from geograypher_b200 import _lib

        key = (self.distortion_key(camera.distortion_params, image_scale), camera.f, camera.cx, camera.cy,
               camera.image_width, camera.image_height, bool(warped_to_ideal), int(device))
        cache = self.__dict__.setdefault("_gpu_warp_maps", {})
        if key not in cache:
            h, w = camera.get_image_size(image_scale)
            cache[key] = _lib.build_warp_map(self._gg_distortion(camera, image_scale), h, w, warped_to_ideal, device)
        return cache[key]

    def warp_dewarp_device(self, camera, d_image, warped_to_ideal: bool = False, fill_value=-1, image_scale=1.0):
        """Nearest-neighbour warp of an (h, w) int32 CUDA tensor (a face-ID raster) without leaving the device.
        Integer-safe: IDs are gathered, never converted to float (the reference's float round trip alters some IDs,
        SURVEY.md headline fact 5)."""
        from geograypher_b200 import _lib

        src = self.warp_source_index(camera, image_scale, warped_to_ideal, d_image.device.index or 0)
        return _lib.gather_i32(d_image, src, int(fill_value))

    def warp_dewarp_image(self, camera, input_image, fill_value=0.0, inversion_downsample: int = 8,
                          interpolation_order: int = 1, warped_to_ideal: bool = True, image_scale: float = 1.0):
        """Apply (``warped_to_ideal=False``) or undo (True) the camera's distortion (reference cameras.py:1092-1156).

        Integer images with ``interpolation_order=0`` -- the pix2face case -- are warped on the GPU.  Other images
        (photographs, bilinear) are outside the projection path and are resampled on the host with SciPy from the same
        exact source coordinates.  ``inversion_downsample`` is accepted for compatibility: the inverse map is computed
        exactly (Newton) instead of by interpolating a down-sampled forward map.
        """
        del inversion_downsample
        input_image = np.asarray(input_image)
        if interpolation_order == 0 and np.issubdtype(input_image.dtype, np.integer) and input_image.ndim == 2:
            import torch

            d_in = torch.from_numpy(input_image.astype(np.int32)).cuda()
            out = self.warp_dewarp_device(camera, d_in, warped_to_ideal, fill_value, image_scale)
            return out.cpu().numpy().astype(input_image.dtype)
        from scipy.ndimage import map_coordinates

        rows, cols = self.warp_source_coordinates(camera, image_scale, warped_to_ideal)
        img = np.atleast_3d(input_image).astype(float)
        out = np.stack(
            [map_coordinates(img[..., c], [rows, cols], order=interpolation_order, mode="grid-constant", cval=float(fill_value))
             for c in range(img.shape[2])], axis=-1)
        return np.squeeze(out).astype(input_image.dtype)

    def warp_source_coordinates(self, camera, image_scale: float = 1.0, warped_to_ideal: bool = True):
        """(rows, cols) float64 source coordinates for every output pixel, in the reference's map convention
        (cameras.py:1027-1062): integer coordinates are pixel positions at scale 1, ``(i + 0.5) / scale`` otherwise.
        Host-side NumPy (used for photographs only)."""
        h, w = camera.get_image_size(image_scale)
        one = np.isclose(image_scale, 1.0)
        rr = np.arange(h, dtype=float) if one else (np.arange(h) + 0.5) / image_scale
        cc = np.arange(w, dtype=float) if one else (np.arange(w) + 0.5) / image_scale
        rows, cols = np.meshgrid(rr, cc, indexing="ij")
        if warped_to_ideal:
            wc, wr = self.ideal_to_warped(camera, cols, rows)
            return (wr, wc) if one else (wr * image_scale, wc * image_scale)
        # ideal -> warped image: invert the forward model with Newton iterations (finite-difference Jacobian)
        ti, tj = np.meshgrid(np.arange(h, dtype=float), np.arange(w, dtype=float), indexing="ij")
        tx, ty = (tj, ti) if one else (tj / image_scale, ti / image_scale)
        x, y = tx.copy(), ty.copy()
        for _ in range(12):
            fx, fy = self.ideal_to_warped(camera, x, y)
            e = 1e-3
            fxx, fyx = self.ideal_to_warped(camera, x + e, y)
            fxy, fyy = self.ideal_to_warped(camera, x, y + e)
            a, b, c_, d = (fxx - fx) / e, (fxy - fx) / e, (fyx - fy) / e, (fyy - fy) / e
            det = a * d - b * c_
            rx, ry = fx - tx, fy - ty
            x = x - (d * rx - b * ry) / det
            y = y - (-c_ * rx + a * ry) / det
        return (y, x) if one else (y * image_scale - 0.5, x * image_scale - 0.5)
