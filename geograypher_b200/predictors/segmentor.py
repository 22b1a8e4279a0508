"""Prediction-image contract of the aggregation path (reference: geograypher/predictors/segmentor.py)."""
import typing

import numpy as np


class Segmentor:
    """Base class with the reference's interface (predictors/segmentor.py:6-35)."""

    def __init__(self, num_classes=None):
        self.num_classes = num_classes

    def setup(self, **kwargs) -> None:
        pass

    def segment_image(self, image: np.ndarray, **kwargs):
        raise NotImplementedError("Abstract base class")

    def segment_image_batch(self, images: typing.List[np.ndarray], **kwargs):
        return [self.segment_image(image, **kwargs) for image in images]

    @staticmethod
    def inds_to_one_hot(inds_image: np.ndarray, num_classes: typing.Union[int, None] = None,
                        ignore_ind: int = 255) -> np.ndarray:
        """(m, n) class indices -> (m, n, num_classes) bool (predictors/segmentor.py:37-69).

        A pixel whose index is ``ignore_ind`` or >= num_classes becomes an all-False row, which still counts
        as an observation of the face it lands on (meshes.py:2064-2067).  The GPU path does not need this
        expansion: pass the index image itself (GG_PRED_INDEX_U8) and the kernel expands it on the fly.
        """
        inds_image = np.asarray(inds_image)
        if num_classes is None:
            num_classes = int(np.max(inds_image)) + 1  # the reference does not mask ignore_ind here (:53-57)
        classes = np.arange(num_classes, dtype=np.int64).reshape((1,) * inds_image.ndim + (-1,))
        return inds_image[..., None].astype(np.int64) == classes
