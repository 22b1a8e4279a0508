"""Camera set whose "images" are segmentation results (reference: geograypher/cameras/segmentor.py)."""
import inspect
import typing
from copy import copy, deepcopy

import numpy as np

from geograypher_b200.cameras.cameras import PhotogrammetryCameraSet
from geograypher_b200.predictors.segmentor import Segmentor


class SegmentorPhotogrammetryCameraSet(PhotogrammetryCameraSet):
    def __init__(self, base_camera_set: PhotogrammetryCameraSet, segmentor: Segmentor,
                 dont_load_base_image: bool = True):
        """Wraps a camera set so that get_image_by_index returns the segmentor's output
        (reference cameras/segmentor.py:10-31; like the reference it does not call super().__init__)."""
        self.base_camera_set = base_camera_set
        self.segmentor = segmentor
        self.dont_load_base_image = dont_load_base_image
        self.cameras = self.base_camera_set.cameras
        self._local_to_epsg_4978_transform = self.base_camera_set._local_to_epsg_4978_transform
        self.image_folder = getattr(base_camera_set, "image_folder", None)
        self._indices = list(range(len(self.cameras)))  # position of each camera in the original set

    def _segment(self, index: int, image_scale: float, method: str):
        raw = None if self.dont_load_base_image else self.base_camera_set.get_image_by_index(index, image_scale)
        filename = self.base_camera_set.get_image_filename(index, absolute=True)
        fn = getattr(self.segmentor, method)
        kwargs = {"filename": filename, "image_scale": image_scale}
        if self._takes_index(method, fn):
            kwargs["index"] = self._indices[index]  # reference-style segmentors only take filename / image_scale
        return fn(raw, **kwargs)

    def _takes_index(self, method, fn):
        cache = self.__dict__.setdefault("_takes_index_cache", {})
        if method not in cache:
            params = inspect.signature(fn).parameters
            cache[method] = "index" in params or any(
                p.kind == inspect.Parameter.VAR_KEYWORD for p in params.values()
            )
        return cache[method]

    def get_image_by_index(self, index: int, image_scale: float = 1) -> np.ndarray:
        return self._segment(index, image_scale, "segment_image")

    def get_class_index_image_by_index(self, index: int, image_scale: float = 1):
        """(h, w) uint8 class indices when the segmentor can provide them, else None.  The GPU aggregation expands
        them on the fly, which moves C times fewer bytes than the (h, w, C) one-hot array."""
        if not hasattr(self.segmentor, "segment_image_indices"):
            return None
        try:
            inds = self._segment(index, image_scale, "segment_image_indices")
        except NotImplementedError:
            return None
        inds = np.asarray(inds)
        return inds if inds.dtype == np.uint8 and inds.ndim == 2 else None

    def get_raw_image_by_index(self, index: int, image_scale: float = 1) -> np.ndarray:
        return self.base_camera_set.get_image_by_index(index=index, image_scale=image_scale)

    def get_subset_cameras(self, inds: typing.List[int], deep: bool = True):
        subset = deepcopy(self) if deep else copy(self)
        subset.cameras = [subset.cameras[i] for i in inds]
        subset._indices = [subset._indices[i] for i in inds]
        subset.base_camera_set = subset.base_camera_set.get_subset_cameras(inds, deep=deep)
        return subset

    def __getitem__(self, slice):
        if isinstance(slice, (int, np.integer)):
            return self.cameras[slice]
        return self.get_subset_cameras(list(range(len(self.cameras)))[slice])

    def n_image_channels(self) -> int:
        return self.segmentor.num_classes
