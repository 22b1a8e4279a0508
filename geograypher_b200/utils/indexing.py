"""Host-side helpers of the hot path (reference: geograypher/utils/indexing.py)."""
import typing

import numpy as np


def find_argmax_nonzero_value(array: np.ndarray, keepdims: bool = False, axis: int = 1) -> np.ndarray:
    """Argmax along ``axis`` as float, NaN for rows whose sum is zero or that hold a non-finite value.

    Same contract as the reference's utils/indexing.py:9-32.  This is the NumPy form for arrays that already
    live on the host; ``TexturedPhotogrammetryMesh.aggregate_projected_images(..., return_argmax=True)``
    computes the same thing on the GPU (``gg_finalize``) without a second pass over host memory.
    """
    array = np.asarray(array)
    argmax = np.argmax(array, axis=axis, keepdims=keepdims).astype(float)
    invalid = np.logical_or(np.sum(array, axis=axis) == 0, np.any(~np.isfinite(array), axis=axis))
    argmax[invalid] = np.nan
    return argmax


def determine_IDs_to_labels(texture_array: np.ndarray, all_discrete_texture_values=None,
                            background_ID: typing.Optional[int] = None):
    """{ID: label} for a one-column texture, or None when the values are genuinely continuous
    (reference utils/indexing.py:35-84)."""
    if texture_array.dtype == float:
        finite = texture_array[np.isfinite(texture_array)]
        if not np.allclose(finite, finite.astype(int)):
            return None
    source = texture_array if all_discrete_texture_values is None else all_discrete_texture_values
    IDs_to_labels, i = {}, 0
    for value in np.unique(source):
        if i == background_ID:
            i += 1
        IDs_to_labels[i] = value
        i += 1
    return IDs_to_labels
