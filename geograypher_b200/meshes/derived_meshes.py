"""Aggregation variants of the reference (geograypher/meshes/derived_meshes.py) on the CUDA path."""
from __future__ import annotations

import numpy as np

from geograypher_b200 import _lib
from geograypher_b200.cameras import PhotogrammetryCamera, PhotogrammetryCameraSet
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
        if as_sparse:
            average, info["projection_counts"], info["summed_projections"] = self._votes_to_csr(d_sum, d_count)
            return average, info
        import torch

        counts, summed = _lib.to_fresh_host([d_count.to(torch.int64), d_sum.to(torch.int64)])
        average = np.zeros(summed.shape, dtype=float)
        seen = counts > 0
        average[seen] = summed[seen] * np.reciprocal(counts[seen].astype(float))[:, None]
        info["projection_counts"] = counts
        info["summed_projections"] = summed
        return average, info

    @staticmethod
    def _votes_to_csr(d_sum, d_count):
        """(average, counts, summed) as ``scipy.sparse.csr_array`` objects of shape (F, C), (F, 1), (F, C) -- the
        reference's return types (derived_meshes.py:527-550) -- built from the dense device accumulators WITHOUT a
        dense pass on the host: the non-zero pattern, the row pointers and the per-entry means are computed where the
        accumulators live, and only the CSR arrays (a few bytes per observed face, not 8 * C per face of the mesh)
        cross PCIe.  At 20 M faces and 10 classes the dense host route (copy, two casts, a masked divide, three
        dense -> CSR conversions) takes seconds, more than a few hundred views of GPU work.  Same values, index
        order and dtypes as ``csr_array(dense)`` of the dense route."""
        import torch
        from scipy.sparse import csr_array

        F, C = d_sum.shape
        nz = d_sum != 0
        rows, cols = nz.nonzero(as_tuple=True)  # row-major order = CSR order, column indices sorted within a row
        index_dtype = torch.int32 if max(int(rows.numel()), F, C) < 2**31 - 1 else torch.int64
        indptr = torch.zeros(F + 1, dtype=torch.int64, device=d_sum.device)
        torch.cumsum(nz.sum(dim=1), dim=0, out=indptr[1:])
        sums = d_sum[rows, cols]
        # the dense route's arithmetic: summed * reciprocal(counts), both float64
        means = sums * torch.reciprocal(d_count[rows].to(torch.float64))
        seen = d_count > 0
        count_indptr = torch.zeros(F + 1, dtype=torch.int64, device=d_sum.device)
        torch.cumsum(seen, dim=0, out=count_indptr[1:])
        count_data = d_count[seen].to(torch.int64)
        indptr, cols = indptr.to(index_dtype), cols.to(index_dtype)
        # delivered into fresh, pre-faulted NumPy arrays (_lib.to_fresh_host: the first touch of the host pages, not
        # PCIe, is what such a delivery costs); the column indices and row pointers cross twice so that the two
        # (F, C) results share no array
        (h_indptr, h_cols, h_sums, h_means, h_count_indptr, h_count_data, h_cols2, h_indptr2) = _lib.to_fresh_host(
            [indptr, cols, sums.to(torch.int64), means, count_indptr.to(index_dtype), count_data, cols, indptr])
        average = csr_array((h_means, h_cols, h_indptr), shape=(F, C))
        summed = csr_array((h_sums, h_cols2, h_indptr2), shape=(F, C))
        counts = csr_array((h_count_data, np.zeros(len(h_count_data), dtype=h_cols.dtype), h_count_indptr), shape=(F, 1))
        return average, counts, summed


def _planar_xy(points):
    """(N, 2) metric planar coordinates of (N, 3) points.  Earth-centred coordinates (|p| of the order of the
    Earth's radius) are projected onto the local tangent plane (east, north) at their centroid -- the reference
    reprojects to a projected CRS with pyproj (derived_meshes.py:62-75), which agrees with the tangent plane to well
    under a metre over a survey; coordinates that already are a local metric frame are used as they are."""
    points = np.asarray(points, dtype=float)
    if len(points) == 0 or np.median(np.linalg.norm(points, axis=1)) < 1e6:
        return points[:, :2].copy(), None
    c = points.mean(axis=0)
    up = c / np.linalg.norm(c)
    east = np.cross([0.0, 0.0, 1.0], up)
    east /= np.linalg.norm(east)
    north = np.cross(up, east)
    frame = (c, east, north)
    return np.stack([(points - c) @ east, (points - c) @ north], axis=1), frame


def _kmeans_labels(xy, n_clusters, seed=0):
    """Cluster ID per point: scikit-learn's KMeans like the reference (derived_meshes.py:77-80) when it is installed,
    else Lloyd's algorithm with k-means++ seeding."""
    if n_clusters > len(xy):
        raise ValueError(f"n_samples={len(xy)} should be >= n_clusters={n_clusters}.")
    try:
        from sklearn.cluster import KMeans

        return KMeans(n_clusters=n_clusters, n_init=10, random_state=seed).fit_predict(xy)
    except ImportError:
        pass
    rng = np.random.default_rng(seed)
    centers = [xy[rng.integers(len(xy))]]
    for _ in range(1, n_clusters):
        d2 = np.min([((xy - c) ** 2).sum(axis=1) for c in centers], axis=0)
        centers.append(xy[rng.choice(len(xy), p=d2 / d2.sum())] if d2.sum() > 0 else xy[rng.integers(len(xy))])
    centers = np.array(centers)
    for _ in range(100):
        labels = np.argmin(((xy[:, None, :] - centers[None]) ** 2).sum(axis=2), axis=1)
        new = np.array([xy[labels == k].mean(axis=0) if np.any(labels == k) else centers[k] for k in range(n_clusters)])
        if np.allclose(new, centers):
            break
        centers = new
    return labels


class TexturedPhotogrammetryMeshChunked(TexturedPhotogrammetryMesh):
    """Chunked operations (reference derived_meshes.py:23-317): the cameras are clustered, every cluster works on the
    sub-mesh within ``buffer_dist_meters`` of its cameras, and the per-chunk results are merged by original face ID.

    The reference chunks because its rasterizer cannot hold a large survey; here the whole mesh fits one GPU and every
    view is frustum-culled, so chunking buys nothing -- but it CHANGES RESULTS (a cluster's cameras neither see nor
    are occluded by faces farther than the buffer from every camera of the cluster), so it is reproduced faithfully.
    ``chunked=False`` (*new*) runs the base-class path on the whole mesh instead."""

    def camera_clusters(self, cameras, n_clusters: int = 8):
        """(cluster ID per camera, (n, 2) planar camera positions, tangent frame or None)."""
        cam_list = self._camera_list(cameras)
        T = cameras.get_local_to_epsg_4978_transform() if hasattr(cameras, "get_local_to_epsg_4978_transform") else None
        T = np.eye(4) if T is None else np.asarray(T, dtype=float)
        local = np.array([np.asarray(c.cam_to_world_transform, dtype=float)[:3, 3] for c in cam_list])
        world = local @ T[:3, :3].T + T[:3, 3]
        pts = np.concatenate([world, self.points], axis=0)
        xy, frame = _planar_xy(pts)
        return _kmeans_labels(xy[: len(cam_list)], n_clusters), xy[: len(cam_list)], xy[len(cam_list):]

    def get_mesh_chunks_for_cameras(self, cameras, n_clusters: int = 8,
                                    buffer_dist_meters: float = CHUNKED_MESH_BUFFER_DIST_METERS,
                                    vis_clusters: bool = False, include_texture: bool = False):
        """Generator of ``(sub_mesh, sub_camera_set, face_IDs)`` (reference derived_meshes.py:26-151).

        A chunk keeps every face with at least one vertex within ``buffer_dist_meters`` (planar distance) of a camera
        of the cluster -- the reference buffers the camera points into a polygon, selects the vertices inside it and
        extracts the cells adjacent to them (meshes.py:693-714).  ``face_IDs`` index the faces of the full mesh."""
        del vis_clusters
        from scipy.spatial import cKDTree

        if isinstance(cameras, PhotogrammetryCamera):
            cameras = PhotogrammetryCameraSet([cameras])
        labels, cam_xy, vert_xy = self.camera_clusters(cameras, n_clusters)
        texture = self.get_texture(request_vertex_texture=False) if include_texture else None
        for cluster in range(n_clusters):
            cam_inds = np.where(labels == cluster)[0]
            sub_cameras = cameras.get_subset_cameras([int(i) for i in cam_inds])
            dist, _ = cKDTree(cam_xy[cam_inds]).query(vert_xy) if len(cam_inds) else (np.full(len(vert_xy), np.inf), None)
            vert_in = dist <= buffer_dist_meters
            face_IDs = np.where(vert_in[self.faces].any(axis=1))[0]
            if len(face_IDs) == 0:
                yield None, sub_cameras, face_IDs
                continue
            sub_faces = self.faces[face_IDs]
            used, inverse = np.unique(sub_faces.reshape(-1), return_inverse=True)
            sub_texture = texture[face_IDs] if texture is not None else None
            IDs_to_labels = None
            if sub_texture is not None and self.is_discrete_texture():
                vals = np.unique(sub_texture)
                IDs_to_labels = {u: u for u in vals[np.isfinite(vals)]}  # identity: no second remapping (:118-131)
            sub_mesh = TexturedPhotogrammetryMesh(
                (self.points[used], inverse.reshape(-1, 3).astype(np.int32)), input_CRS=self.CRS, texture=sub_texture,
                IDs_to_labels=IDs_to_labels, device=self.device, compat_negative_index=self.compat_negative_index,
                views_per_batch=self.views_per_batch, use_principal_point=self.use_principal_point,
                sparse_host_gather=self.sparse_host_gather, log_level=self.logger.level)
            yield sub_mesh, sub_cameras, face_IDs

    def aggregate_projected_images(self, cameras, batch_size: int = 1, aggregate_img_scale: float = 1,
                                   n_clusters: int = 8, buffer_dist_meters: float = CHUNKED_MESH_BUFFER_DIST_METERS,
                                   vis_clusters: bool = False, chunked: bool = True, **kwargs):
        """Reference derived_meshes.py:222-317: per-chunk sums and counts merged into the full arrays by face ID
        (NaN sums counted as 0), then the same epilogue as the parent class.  Counts are ``int`` like there (:264)."""
        if not chunked:
            average, info = super().aggregate_projected_images(cameras, batch_size=batch_size,
                                                               aggregate_img_scale=aggregate_img_scale, **kwargs)
            info["projection_counts"] = info["projection_counts"].astype(int)
            return average, info
        kwargs.pop("return_all", None)
        summed = np.zeros((self.faces.shape[0], cameras.n_image_channels()), dtype=float)
        counts = np.zeros(self.faces.shape[0], dtype=int)
        for sub_mesh, sub_cameras, face_IDs in self.get_mesh_chunks_for_cameras(
                cameras, n_clusters=n_clusters, buffer_dist_meters=buffer_dist_meters, vis_clusters=vis_clusters):
            if len(face_IDs) == 0 or len(sub_cameras) == 0:
                continue
            _, sub_info = sub_mesh.aggregate_projected_images(sub_cameras, batch_size=batch_size,
                                                              aggregate_img_scale=aggregate_img_scale,
                                                              return_all=False, **kwargs)
            summed[face_IDs] = np.nansum([summed[face_IDs], sub_info["summed_projections"]], axis=0)
            counts[face_IDs] = counts[face_IDs] + sub_info["projection_counts"].astype(int)
        summed[counts == 0] = np.nan
        with np.errstate(invalid="ignore", divide="ignore"):
            average = np.divide(summed, np.expand_dims(counts, 1))
        return average, {"projection_counts": counts, "summed_projections": summed}

    def render_flat(self, cameras, batch_size: int = 1, render_img_scale: float = 1, n_clusters: int = 8,
                    buffer_dist_meters: float = CHUNKED_MESH_BUFFER_DIST_METERS, vis_clusters: bool = False,
                    chunked: bool = True, **pix2face_kwargs):
        """Reference derived_meshes.py:153-220: renders come cluster by cluster (NOT in the order of ``cameras``), each
        from its cluster's sub-mesh."""
        if not chunked:
            yield from super().render_flat(cameras, batch_size=batch_size, render_img_scale=render_img_scale,
                                           **pix2face_kwargs)
            return
        for sub_mesh, sub_cameras, face_IDs in self.get_mesh_chunks_for_cameras(
                cameras, n_clusters=n_clusters, buffer_dist_meters=buffer_dist_meters, vis_clusters=vis_clusters,
                include_texture=True):
            if sub_mesh is None:
                continue  # the reference cannot render a cluster without mesh either
            yield from sub_mesh.render_flat(sub_cameras, batch_size=batch_size, render_img_scale=render_img_scale,
                                            **pix2face_kwargs)
