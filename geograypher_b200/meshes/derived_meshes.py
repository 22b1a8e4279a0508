"""Aggregation variants of the reference (geograypher/meshes/derived_meshes.py) on the CUDA path."""
from __future__ import annotations

import numpy as np

from geograypher_b200 import _lib
from geograypher_b200.constants import CHUNKED_MESH_BUFFER_DIST_METERS
from geograypher_b200.meshes.meshes import TexturedPhotogrammetryMesh


class TexturedPhotogrammetryMeshIndexPredictions(TexturedPhotogrammetryMesh):
    def aggregate_projected_images(self, cameras, n_classes: int, batch_size: int = 1,
                                   aggregate_img_scale: float = 1, return_all: bool = False,
                                   as_sparse: bool = True, **kwargs):
        """One-hot voting (reference derived_meshes.py:415-550).

        The images are (h, w) arrays of class indices with NaN for "no prediction".  Per view, every face whose
        last pixel (row-major) holds a finite index votes once for that class: ``counts[f] += 1`` and
        ``summed[f, class] += 1``; ``average = summed / counts``.  Votes are integers, so the result does not
        depend on the number of GPUs or on the order of the views.

        Returns ``(average, {"projection_counts", "summed_projections"})`` as ``scipy.sparse.csr_array`` objects of
        shape (F, n_classes) / (F, 1) like the reference; ``as_sparse=False`` (*new*) returns dense arrays.
        The accumulators are dense on the device (F x n_classes float64).
        """
        del batch_size
        info = {}
        if return_all:
            info["all_projections"] = list(
                self.project_images(cameras, aggregate_img_scale=aggregate_img_scale, check_null_image=True, **kwargs)
            )
        pix2face_kwargs = {k: v for k, v in kwargs.items() if k != "check_null_image"}
        d_sum, d_count, _ = self._accumulate_views(cameras, aggregate_img_scale, _lib.MODE_VOTE,
                                                   n_channels=int(n_classes), pix2face_kwargs=pix2face_kwargs)
        counts = d_count.cpu().numpy().astype(np.int64)
        summed = d_sum.cpu().numpy().astype(np.int64)
        average = np.zeros(summed.shape, dtype=float)
        seen = counts > 0
        average[seen] = summed[seen] * np.reciprocal(counts[seen].astype(float))[:, None]
        if as_sparse:
            from scipy.sparse import csr_array

            info["projection_counts"] = csr_array(counts[:, None])
            info["summed_projections"] = csr_array(summed)
            return csr_array(average), info
        info["projection_counts"] = counts
        info["summed_projections"] = summed
        return average, info


class TexturedPhotogrammetryMeshChunked(TexturedPhotogrammetryMesh):
    """The reference chunks the mesh by camera clusters because its rasterizer cannot hold a large survey
    (derived_meshes.py:23-317).  On a B200 the whole mesh stays resident and every view is frustum-culled on the
    GPU, so the chunked methods give the same results as the base class; the chunking arguments are accepted and
    ignored."""

    def aggregate_projected_images(self, cameras, batch_size: int = 1, aggregate_img_scale: float = 1,
                                   n_clusters: int = 8, buffer_dist_meters: float = CHUNKED_MESH_BUFFER_DIST_METERS,
                                   vis_clusters: bool = False, **kwargs):
        del n_clusters, buffer_dist_meters, vis_clusters
        average, info = super().aggregate_projected_images(cameras, batch_size=batch_size,
                                                           aggregate_img_scale=aggregate_img_scale, **kwargs)
        info["projection_counts"] = info["projection_counts"].astype(int)  # int in the reference (:264)
        return average, info

    def render_flat(self, cameras, batch_size: int = 1, render_img_scale: float = 1, n_clusters: int = 8,
                    buffer_dist_meters: float = 50, vis_clusters: bool = False, **pix2face_kwargs):
        del n_clusters, buffer_dist_meters, vis_clusters
        return super().render_flat(cameras, batch_size=batch_size, render_img_scale=render_img_scale,
                                   **pix2face_kwargs)
