"""Camera containers of the multiview projection path.

Host-side mirror of the reference's ``PhotogrammetryCamera`` / ``PhotogrammetryCameraSet``
(geograypher/cameras/cameras.py:55-200, 661-926): same constructor signatures, attribute names and the
methods the hot path consumes.  Everything geospatial / visual (pyvista frusta, EXIF, triangulation,
ROI sub-setting) is outside the path and not provided.
"""
from __future__ import annotations

import hashlib
import json
import os
from copy import copy, deepcopy
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from geograypher_b200.constants import EXAMPLE_INTRINSICS, PATH_TYPE


def _imread(path) -> np.ndarray:
    path = Path(path)
    if path.suffix == ".npy":
        return np.load(path)
    try:
        from PIL import Image
    except ImportError as e:
        raise ImportError("Reading image files needs Pillow") from e
    return np.asarray(Image.open(path))


def _resize_bilinear(image: np.ndarray, shape) -> np.ndarray:
    """Bilinear resize sampled at output pixel centres (what skimage.transform.resize's default order=1 does,
    without its anti-aliasing filter)."""
    h, w = image.shape[:2]
    ys = np.clip((np.arange(shape[0]) + 0.5) * h / shape[0] - 0.5, 0, h - 1)
    xs = np.clip((np.arange(shape[1]) + 0.5) * w / shape[1] - 0.5, 0, w - 1)
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    fy = (ys - y0).reshape(-1, 1, *([1] * (image.ndim - 2)))
    fx = (xs - x0).reshape(1, -1, *([1] * (image.ndim - 2)))
    img = image.astype(float)
    top = img[y0][:, x0] * (1 - fx) + img[y0][:, x1] * fx
    bot = img[y1][:, x0] * (1 - fx) + img[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


class PhotogrammetryCamera:
    def __init__(
        self,
        image_filename: PATH_TYPE,
        cam_to_world_transform: np.ndarray,
        f: float,
        cx: float,
        cy: float,
        image_width: int,
        image_height: int,
        distortion_params: Dict[str, float] = {},
        lon_lat: Union[None, Tuple[float, float]] = None,
        local_to_epsg_4978_transform: Union[np.ndarray, None] = None,
    ):
        """One posed pinhole camera (reference cameras.py:56-102).

        Args:
            image_filename: the image this camera took (may be None for synthetic cameras)
            cam_to_world_transform: 4x4 camera-to-world transform in the camera set's local frame; camera axes
                are +X right, +Y down, +Z forward
            f: focal length in pixels
            cx, cy: principal point offset from the image centre, pixels
            image_width, image_height: image size in pixels
            distortion_params: lens-distortion coefficients (carried, see DESIGN.md section "next rows")
            lon_lat: optional location
            local_to_epsg_4978_transform: 4x4 local frame -> EPSG:4978
        """
        self.image_filename = image_filename
        self.cam_to_world_transform = np.asarray(cam_to_world_transform, dtype=float)
        self.world_to_cam_transform = np.linalg.inv(self.cam_to_world_transform)
        self.f = f
        self.cx = cx
        self.cy = cy
        self.image_width = image_width
        self.image_height = image_height
        self.distortion_params = distortion_params
        self._local_to_epsg_4978_transform = local_to_epsg_4978_transform
        self.lon_lat = (None, None) if lon_lat is None else lon_lat
        self.image_size = (image_height, image_width)
        self.image = None
        self.cache_image = False

    def get_camera_hash(self, include_image_hash: bool = False):
        """sha256 of the camera geometry (reference cameras.py:104-134)."""
        settings = {
            "transform": self.cam_to_world_transform.tolist(),
            "f": self.f,
            "cx": self.cx,
            "cy": self.cy,
            "image_width": self.image_width,
            "image_height": self.image_height,
            "distortion_params": self.distortion_params,
            "lon_lat": self.lon_lat,
        }
        if include_image_hash:
            settings["image_filename"] = str(self.image_filename)
        return hashlib.sha256(json.dumps(settings, sort_keys=True).encode("utf-8")).hexdigest()

    def get_camera_properties(self):
        """Reference cameras.py:136-152."""
        return {
            "focal_length": self.f,
            "principal_point_x": self.cx,
            "principal_point_y": self.cy,
            "image_height": self.image_height,
            "image_width": self.image_width,
            "distortion_params": self.distortion_params,
            "world_to_cam_transform": self.world_to_cam_transform,
        }

    def get_image(self, image_scale: float = 1.0) -> np.ndarray:
        """uint8 images are returned as float in [0, 1]; optional resize (reference cameras.py:154-174)."""
        if self.image is None:
            image = _imread(self.image_filename)
            if image.dtype == np.uint8:
                image = image / 255.0
            if self.cache_image:
                self.image = image
        else:
            image = self.image
        if image_scale != 1.0:
            image = _resize_bilinear(
                image, (int(image.shape[0] * image_scale), int(image.shape[1] * image_scale))
            )
        return image

    def get_image_filename(self):
        return self.image_filename

    def get_image_size(self, image_scale=1.0):
        """(h, w) = (int(H*s), int(W*s)) (reference cameras.py:179-200)."""
        if self.image_size is None:
            image = self.image if self.image is not None else self.get_image()
            self.image_size = image.shape[:2]
        return (int(self.image_size[0] * image_scale), int(self.image_size[1] * image_scale))

    def get_local_to_epsg_4978_transform(self):
        return self._local_to_epsg_4978_transform


class PhotogrammetryCameraSet:
    reads_image_files = True  # get_image_by_index decodes a file per view: the aggregation may read views ahead

    def __init__(
        self,
        cameras: Union[None, PhotogrammetryCamera, List[PhotogrammetryCamera]] = None,
        cam_to_world_transforms: Optional[List[np.ndarray]] = None,
        intrinsic_params_per_sensor_type: Dict[int, Dict[str, float]] = {0: EXAMPLE_INTRINSICS},
        image_filenames: Optional[List[PATH_TYPE]] = None,
        lon_lats: Optional[List[Union[None, Tuple[float, float]]]] = None,
        image_folder: Optional[PATH_TYPE] = None,
        sensor_IDs: Optional[List[int]] = None,
        validate_images: bool = False,
        local_to_epsg_4978_transform: np.ndarray = np.eye(4),
    ):
        """A set of cameras in one local frame (reference cameras.py:662-781)."""
        self._local_to_epsg_4978_transform = local_to_epsg_4978_transform
        self._maps_ideal_to_warped = {}
        self._maps_warped_to_ideal = {}

        if cameras is not None:
            if isinstance(cameras, PhotogrammetryCamera):
                cameras = [cameras]
            names = [str(c.image_filename) for c in cameras if c.image_filename is not None]
            if len(names) == len(cameras) and len(names) > 0:
                self.image_folder = (
                    Path(names[0]).parent if len(names) == 1 else Path(os.path.commonpath(names))
                )
            else:
                self.image_folder = image_folder
            self.cameras = list(cameras)
            return

        n_transforms = len(cam_to_world_transforms)
        if image_filenames is None:
            image_filenames = [None] * n_transforms
        if sensor_IDs is None and len(intrinsic_params_per_sensor_type) == 1:
            sensor_IDs = [list(intrinsic_params_per_sensor_type.keys())[0]] * n_transforms
        elif sensor_IDs is None or len(sensor_IDs) != n_transforms:
            raise ValueError(
                f"Number of sensor_IDs ({None if sensor_IDs is None else len(sensor_IDs)}) is different than "
                f"the number of transforms ({n_transforms})"
            )
        if lon_lats is None:
            lon_lats = [None] * n_transforms

        self.cam_to_world_transforms = cam_to_world_transforms
        self.intrinsic_params_per_sensor_type = intrinsic_params_per_sensor_type
        self.image_filenames = image_filenames
        self.lon_lats = lon_lats
        self.sensor_IDs = sensor_IDs
        self.image_folder = image_folder

        if validate_images:
            keep = [i for i, fn in enumerate(image_filenames) if fn is not None and Path(fn).is_file()]
            self.image_filenames = [self.image_filenames[i] for i in keep]
            self.cam_to_world_transforms = [self.cam_to_world_transforms[i] for i in keep]
            self.sensor_IDs = [self.sensor_IDs[i] for i in keep]
            self.lon_lats = [self.lon_lats[i] for i in keep]

        self.cameras = []
        for fn, c2w, sensor_ID, lon_lat in zip(
            self.image_filenames, self.cam_to_world_transforms, self.sensor_IDs, self.lon_lats
        ):
            params = self.intrinsic_params_per_sensor_type[sensor_ID]
            if params is None:  # sensor without a full calibration (cameras.py:766-768)
                continue
            self.cameras.append(
                PhotogrammetryCamera(
                    fn, c2w, lon_lat=lon_lat, local_to_epsg_4978_transform=local_to_epsg_4978_transform, **params
                )
            )

    def __len__(self):
        return self.n_cameras()

    def __getitem__(self, slice):
        subset = self.cameras[slice]
        if isinstance(subset, PhotogrammetryCamera):
            return subset
        return PhotogrammetryCameraSet(
            subset,
            image_folder=getattr(self, "image_folder", None),
            local_to_epsg_4978_transform=self._local_to_epsg_4978_transform,
        )

    def n_cameras(self) -> int:
        return len(self.cameras)

    def n_image_channels(self) -> int:
        return 3

    def get_image_folder(self):
        return self.image_folder

    def get_subset_cameras(self, inds: List[int], deep: bool = True):
        """The cameras at ``inds`` as a new set (reference cameras.py: a deep copy).  ``deep=False`` (*new*) shares the
        camera objects with this set instead of copying them."""
        subset = deepcopy(self) if deep else copy(self)
        subset.cameras = [subset.cameras[i] for i in inds]
        return subset

    def get_image_by_index(self, index: int, image_scale: float = 1.0) -> np.ndarray:
        return self[index].get_image(image_scale=image_scale)

    def get_image_filename(self, index: Union[int, None], absolute=True):
        if index is None:
            return [self.get_image_filename(i, absolute=absolute) for i in range(len(self.cameras))]
        filename = self.cameras[index].get_image_filename()
        if filename is None:
            return None
        return Path(filename) if absolute else Path(filename).relative_to(self.get_image_folder())

    def get_local_to_epsg_4978_transform(self):
        """4x4 local frame -> EPSG:4978 (reference cameras.py:911-926)."""
        return self._local_to_epsg_4978_transform

    # ---- lens distortion hooks (reference cameras.py:968-1156) ------------------------------------------
    def distortion_key(self, parameters: Dict[str, float], image_scale: float = 1.0) -> str:
        """Repeatable cache key of a distortion-parameter dict, 8 decimals (reference cameras.py:968-993)."""
        keys = sorted(parameters.keys())
        return "|".join([f"{k}:{parameters[k]:.8f}" for k in keys] + [f"image_scale:{image_scale:.8f}"])

    def ideal_to_warped(self, camera, xpix, ypix):
        """Only calibrated (Metashape) camera sets define a distortion model (reference cameras.py:1088-1090)."""
        raise NotImplementedError("Distortion is only defined for camera sets with a lens model")

    def warp_dewarp_image(self, camera, input_image, warped_to_ideal=True, fill_value=0.0,
                          interpolation_order=1, image_scale=1.0, **kwargs):
        self.ideal_to_warped(camera, np.zeros(1), np.zeros(1))
