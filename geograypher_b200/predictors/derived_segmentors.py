"""Segmentors that feed the aggregation path (reference: geograypher/predictors/derived_segmentors.py)."""
from pathlib import Path

import numpy as np

from geograypher_b200.constants import PATH_TYPE
from geograypher_b200.predictors.segmentor import Segmentor


def _read_index_image(path) -> np.ndarray:
    path = Path(path)
    if path.suffix == ".npy":
        return np.load(path)
    try:
        from PIL import Image

        return np.asarray(Image.open(path))
    except ImportError as e:  # same pattern as the reference's optional back-ends (derived_meshes.py:581-584)
        raise ImportError("Reading image files needs Pillow; .npy index images work without it") from e


def _nearest_resize(image: np.ndarray, shape) -> np.ndarray:
    """Nearest-neighbour resize to ``shape`` (rows, cols), sampling at output pixel centres like
    skimage.transform.resize(order=0) does."""
    rows = np.minimum(((np.arange(shape[0]) + 0.5) * image.shape[0] / shape[0]).astype(int), image.shape[0] - 1)
    cols = np.minimum(((np.arange(shape[1]) + 0.5) * image.shape[1] / shape[1]).astype(int), image.shape[1] - 1)
    return image[rows][:, cols]


class LookUpSegmentor(Segmentor):
    """Class-index PNGs stored next to the images (derived_segmentors.py:32-51)."""

    io_bound = True  # every call decodes a file: the aggregation reads the coming views ahead on a thread pool

    def __init__(self, base_folder, lookup_folder, num_classes=10):
        super().__init__(num_classes=num_classes)
        self.base_folder = Path(base_folder)
        self.lookup_folder = lookup_folder

    def segment_image_indices(self, image, filename: PATH_TYPE, image_scale: float = 1.0) -> np.ndarray:
        """(h, w) uint8 class indices -- what the GPU aggregation consumes directly."""
        relative_path = Path(filename).relative_to(self.base_folder)
        lookup_path = Path(self.lookup_folder, relative_path)
        suffix = ".png" if lookup_path.with_suffix(".png").exists() or not lookup_path.with_suffix(".npy").exists() else ".npy"
        inds = _read_index_image(lookup_path.with_suffix(suffix))
        if image_scale != 1:
            inds = _nearest_resize(inds, (int(inds.shape[0] * image_scale), int(inds.shape[1] * image_scale)))
        return inds

    def segment_image(self, image, filename: PATH_TYPE, image_scale: float = 1.0):
        inds = self.segment_image_indices(image, filename=filename, image_scale=image_scale)
        return self.inds_to_one_hot(inds, num_classes=self.num_classes)


class ArraySegmentor(Segmentor):
    """Predictions held in memory, keyed by the camera's position in the base camera set.

    Not in the reference: it stands in for segmentors that read model outputs from disk when the predictions
    already live in arrays (tests, benchmarks, pipelines that run the network in-process).  ``images[i]`` is
    either an (h, w) class-index image (``one_hot=True`` -> expanded like LookUpSegmentor does) or any
    (h, w[, C]) array handed through unchanged.
    """

    def __init__(self, images, num_classes=None, one_hot=False):
        super().__init__(num_classes=num_classes)
        self.images = images
        self.one_hot = one_hot

    def __deepcopy__(self, memo):
        # Camera sets deep-copy themselves when they are sub-set (get_subset_cameras, as in the reference).  The
        # images are inputs, never written: share them instead of duplicating gigabytes -- a copy would also lose the
        # page-locking that lets the GPU read them in place.
        clone = type(self).__new__(type(self))
        memo[id(self)] = clone
        clone.__dict__.update(self.__dict__)
        clone.images = list(self.images) if isinstance(self.images, list) else self.images
        return clone

    def segment_image_indices(self, image, index: int, **kwargs):
        if not self.one_hot:
            raise NotImplementedError("index images are only available when one_hot=True")
        return np.asarray(self.images[index])

    def segment_image(self, image, index: int = None, **kwargs):
        out = self.images[index]
        if self.one_hot:
            return self.inds_to_one_hot(np.asarray(out), num_classes=self.num_classes)
        return out
