/*
 * geograypher_b200.h -- C ABI of the B200-native multiview-projection library (libgeograypher_b200.so).
 *
 * The reference (open-forest-observatory/geograypher v0.4.0, pure Python) has NO native interface; its
 * extension point for this path is method override on TexturedPhotogrammetryMesh
 * (geograypher/meshes/derived_meshes.py:642 overrides pix2face; :222 and :415 override
 * aggregate_projected_images).  Each entry point below names the reference method(s) whose arithmetic it
 * replaces; geograypher_b200/meshes/meshes.py binds them with ctypes behind the reference's signatures.
 * INTEGRATION.md shows the stub a geograypher maintainer would add.
 *
 * Conventions
 *   - every d_* pointer is a DEVICE pointer owned by the caller (e.g. torch tensor storage); h_* pointers
 *     are host memory read before the call returns;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is enqueued on it
 *     and the call returns without synchronising unless stated;
 *   - return value: 0 on success, negative gg_status on failure; gg_last_error() gives the message of the
 *     last failure on the calling thread; nothing throws across the ABI;
 *   - a gg_context is bound to one device, owns the mesh copy + scratch, and is not thread-safe (one
 *     context per mesh per device, like the reference's one pix2face_plotter per mesh, meshes.py:111-114).
 *   - there is no CPU fallback: without a CUDA device gg_create fails.
 */
#ifndef GEOGRAYPHER_B200_H
#define GEOGRAYPHER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_ABI_VERSION 1
#define GG_MAX_VIEWS_PER_CALL 32

typedef struct gg_context gg_context;

typedef enum {
    GG_OK = 0,
    GG_ERR_INVALID = -1,  /* bad argument */
    GG_ERR_CUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
    GG_ERR_NO_MESH = -3,  /* gg_set_mesh has not been called */
    GG_ERR_OVERFLOW = -4, /* scratch capacity exceeded in the last batch: call gg_reserve and retry */
    GG_ERR_NO_DEVICE = -5
} gg_status;

/*
 * One pinhole view of the rasterization contract.  Built on the host in float64 from
 * PhotogrammetryCamera (cameras/cameras.py:56-102, get_camera_properties :136-152) and rounded once:
 *   m  = rows 0..2 of world_to_cam (= inv(cam_to_world), cameras.py:84), with the mesh origin shift folded in
 *   f  = focal length in pixels, px = W/2 + cx, py = H/2 + cy  (derived_meshes.py:772-780)
 *   W,H = raster size (get_image_size(scale), cameras.py:179-200)
 */
typedef struct {
    float m[12];
    float f, px, py;
    int32_t W, H;
    float znear;
} gg_camera;

/* prediction image layouts accepted by the aggregation entry points */
typedef enum {
    GG_PRED_F32 = 0,       /* (H,W,C) float32, NaN = null                                  */
    GG_PRED_F64 = 1,       /* (H,W,C) float64 (what PhotogrammetryCamera.get_image yields) */
    GG_PRED_U8 = 2,        /* (H,W,C) uint8 / bool (LookUpSegmentor one-hot)               */
    GG_PRED_INDEX_U8 = 3   /* (H,W) uint8 class index; expanded on the fly exactly like
                              Segmentor.inds_to_one_hot (predictors/segmentor.py:37-69): value c<C -> one-hot,
                              anything else (255 = ignore) -> all-zero row                  */
} gg_pred_kind;

typedef enum {
    GG_MODE_LAST_PIXEL = 0, /* reference parity: per view the LAST pixel (row-major) of each face sets the
                               face's row (meshes.py:2001), rows are summed over views with NaN->0 and a face
                               counts once per view in which any channel is finite (meshes.py:2057-2067) */
    GG_MODE_PIXEL_SUM = 1,  /* non-reference: every pixel adds its scores, count = number of pixels */
    GG_MODE_VOTE = 2        /* TexturedPhotogrammetryMeshIndexPredictions (derived_meshes.py:480-520): the
                               face's last pixel holds a class index (C == 1 image, NaN = null); counts[f]+=1,
                               sum[f, class] += 1 */
} gg_agg_mode;

typedef enum { GG_OUT_F64 = 0, GG_OUT_F32 = 1, GG_OUT_U8 = 2 } gg_out_dtype;

/* ---- lifetime ------------------------------------------------------------------------------------- */
int gg_abi_version(void);
const char *gg_last_error(void);
int gg_create(int device, gg_context **out);
/* Host memory for prediction images that the fused aggregation reads IN PLACE over PCIe (one row per visible face and
   view): managed memory whose preferred location is the host and that is mapped into `device`.  On the benchmark host
   scattered reads from such a buffer keep their rate as the buffers grow (0.23 G 40-byte rows/s at 10 GB) where
   cudaHostAlloc'ed memory drops to 0.10 G rows/s (profiles/r02_pcie_rows_*.txt).  No counterpart in the reference
   (its images are NumPy arrays); page-locked (cudaHostAlloc / torch pin_memory) buffers are accepted as well.
   gg_pointer_kind: 0 pageable or unknown, 1 page-locked host, 2 device, 3 managed. */
int gg_host_alloc(int device, size_t bytes, void **out);
int gg_host_free(void *p);
int gg_pointer_kind(const void *p);
/* Host-side helper of the pageable-image route (gg_project_winners -> this -> gg_accumulate_rows): copy, for every
   listed (face, pixel) pair of every view, the `row_bytes` bytes of that pixel out of the view's image into one packed
   table, on `n_threads` persistent host threads (0 = up to 16) with software prefetch.  View v owns output rows
   h_offsets[v] .. h_offsets[v+1]; its pairs start at pair index h_pair_starts[v] (NULL: packed like the output, i.e.
   h_offsets).  Pixel indices are clamped to the image like np.take(mode="clip").  Replaces the reference's NumPy fancy
   indexing `img[pix2face-ordered rows]` (meshes.py:1991-2001), which holds the GIL.  No GPU work; no context needed. */
int gg_gather_rows_host(const void *const *h_images, const int64_t *h_pixels_per_image, const int32_t *h_pairs,
                        const int64_t *h_pair_starts, const int64_t *h_offsets, int n_views, int64_t row_bytes,
                        void *h_out, int n_threads);
void gg_destroy(gg_context *ctx);
/* Synchronise `stream` and the internal streams, and report a deferred failure of the work enqueued since the last
   gg_sync (GG_ERR_OVERFLOW: the batches that overflowed the scratch were skipped as a whole; CUDA errors). */
int gg_sync(gg_context *ctx, void *stream);
/* gg_project_aggregate runs its fused modes as a two-stage software pipeline on two internal streams (binning of
   batch k+1 overlaps the rasterization of batch k).  The accumulators it writes are complete, with respect to
   `stream`, after gg_drain (no host synchronisation), gg_finalize or gg_sync; every other entry point drains
   implicitly.  gg_set_pipeline(ctx, 0) turns the overlap off (everything is then enqueued on the caller's stream). */
int gg_drain(gg_context *ctx, void *stream);
int gg_set_pipeline(gg_context *ctx, int enable);
/* Scratch sizing: max face records per view (0 = number of faces) and max (tile, face) pairs per view. */
int gg_reserve(gg_context *ctx, int64_t max_faces_per_view, int64_t max_bin_entries_per_view);
/* Current scratch capacities per view (0 before the first rasterization). */
int gg_get_capacity(gg_context *ctx, int64_t *h_faces_per_view, int64_t *h_bin_entries_per_view);
/* Counters of the most recent rasterization batch, valid after gg_sync: per view [n_visible_blocks,
   n_face_records, n_bin_entries, overflow_flag].  out must hold 4*n int64. */
int gg_last_batch_stats(gg_context *ctx, int n, int64_t *h_out);
/* What the most recent gg_sync that returned GG_ERR_OVERFLOW found, over ALL batches enqueued since the sync before
   it (the batch that overflowed is usually not the last one): h_out3 = [flags (1 = face records, 2 = tile entries),
   most face records any view wanted, most tile entries any view wanted].  Grow only what overflowed. */
int gg_overflow_info(gg_context *ctx, int64_t *h_out3);

/* ---- instrumentation: every kernel launch is counted per stage; with gg_profile(ctx, 1) each launch is also
        bracketed by CUDA events on its stream.  gg_profile_read synchronises the device and returns the
        accumulated milliseconds and launch counts per stage (arrays of gg_stage_count() entries). ---------- */
int gg_stage_count(void);
const char *gg_stage_name(int stage);
int gg_profile(gg_context *ctx, int enable);
int gg_profile_read(gg_context *ctx, double *h_ms, int64_t *h_launches, int reset);

/* ---- mesh (replaces the per-call pv.PolyData / pytorch3d Meshes construction, meshes.py:1806-1817,
        derived_meshes.py:592-640).  d_verts: V x 3 float32 in the camera set's local frame (after
        get_mesh_in_cameras_coords, meshes.py:1641-1676, minus the origin shift); d_faces: F x 3 int32.
        Copies the mesh, builds the per-block bounds used for frustum culling.  Synchronous. -------- */
int gg_set_mesh(gg_context *ctx, const float *d_verts, int64_t V, const int32_t *d_faces, int64_t F,
                void *stream);

/* ---- stage 1: batched camera projection of all vertices (contract C1+C2).  Outputs are n x V. ------- */
int gg_project(gg_context *ctx, const gg_camera *h_cams, int n, int32_t *d_X, int32_t *d_Y, float *d_invz,
               uint8_t *d_valid, void *stream);

/* ---- stage 1+2: pix2face (meshes.py:1678-1856 / derived_meshes.py:642-737).  n <= GG_MAX_VIEWS_PER_CALL
        views of identical W x H.  d_pix2face: n x H x W int32, -1 = no face.  d_depth (optional, may be
        NULL): n x H x W float32 1/z_cam of the winning face (0 where none). ---------------------------- */
int gg_rasterize(gg_context *ctx, const gg_camera *h_cams, int n, int32_t *d_pix2face, float *d_depth,
                 void *stream);

/* ---- stage 3: per-face aggregation of ONE view's prediction image given its pix2face raster
        (project_images + the accumulation of aggregate_projected_images, meshes.py:1988-2001, 2056-2067;
        vote mode: derived_meshes.py:480-520).  d_sum: F x C float64 (vote mode: F x n_classes),
        d_count: F int32; both are accumulated into, so zero them before the first view.
        `flags` is a bit mask: GG_FLAG_COMPAT_NEG reproduces meshes.py:2000 (background
        pixels index face F-1); GG_FLAG_KEEP_NAN keeps NaN scores instead of turning them into 0 (what the
        reference does when the whole aggregation has a single view, meshes.py:2056-2057); GG_FLAG_ASSIGN
        writes the face rows instead of accumulating (project_images' per-view output; count[f]=1 marks
        written rows). ------------------------------------------------------------------------------------ */
#define GG_FLAG_COMPAT_NEG 1
#define GG_FLAG_KEEP_NAN 2
#define GG_FLAG_ASSIGN 4
#define GG_FLAG_TRUNCATE 8 /* gg_project_winners: a view with more than cap_pairs_per_view visible faces lists the first
                              cap and reports its full count in d_counts -- not an overflow; the caller calls again
                              with room for it (lists are usually ~1 % of what the scratch could produce) */
int gg_aggregate(gg_context *ctx, const int32_t *d_pix2face, int H, int W, const void *d_pred, int pred_kind,
                 int C, int mode, int flags, double *d_sum, int32_t *d_count, void *stream);

/* ---- stage 1+2+3 fused over n views: rasterize and aggregate without a round trip of the rasters through
        the host.  h_pred[i] is the device pointer of view i's prediction image.  d_pix2face may be NULL. -- */
int gg_project_aggregate(gg_context *ctx, const gg_camera *h_cams, int n, const void *const *h_pred,
                         int pred_kind, int C, int mode, int flags, double *d_sum, int32_t *d_count,
                         int32_t *d_pix2face, void *stream);

/* ---- the fused path split in two, for prediction images in PAGEABLE host memory (which the GPU cannot read in
        place and which cost ~100x the path to upload whole).  gg_project_winners rasterizes n views and lists, per
        view, every visible face with its last pixel in row-major order (the pixel whose value project_images keeps,
        meshes.py:2001): d_pairs[(view*cap + i)*2 + {0,1}] = {face, pixel}, i < d_counts[view] (order unspecified).
        With GG_FLAG_COMPAT_NEG the last background pixel is listed for face F-1 (meshes.py:2000).  The host gathers
        pred[pixel, :] for those pairs and gg_accumulate_rows applies ONE view's rows (row i belongs to the face of
        pair i; C elements per row, one element for GG_PRED_INDEX_U8 and GG_MODE_VOTE) with the arithmetic of
        gg_aggregate.  Call it view by view, in view order, to get the reference's float64 sums bit for bit. ------- */
int gg_project_winners(gg_context *ctx, const gg_camera *h_cams, int n, int flags, int32_t *d_pairs,
                       int64_t cap_pairs_per_view, int32_t *d_counts, void *stream);
int gg_accumulate_rows(gg_context *ctx, const int32_t *d_pairs, int64_t n_rows, const void *d_rows, int pred_kind, int C,
                       int mode, int flags, double *d_sum, int32_t *d_count, void *stream);

/* ---- epilogue of aggregate_projected_images (meshes.py:2069-2082) + find_argmax_nonzero_value
        (utils/indexing.py:9-32): avg = sum / count (NaN rows where count == 0; d_sum rows with count == 0 are
        set to NaN in place like meshes.py:2070), argmax as float64 with NaN for all-zero / non-finite rows.
        d_avg and d_argmax may each be NULL. ----------------------------------------------------------------- */
int gg_finalize(gg_context *ctx, double *d_sum, const int32_t *d_count, int64_t F, int C, double *d_avg,
                double *d_argmax, void *stream);

/* ---- stage 4: render_flat gather (meshes.py:1921-1937): out[p,:] = tex[pix2face[p],:] or NaN.
        d_face_tex: F x D float64.  out_dtype GG_OUT_U8 applies save_renders' cast rule
        (meshes.py:2323-2334: <0, >255, non-finite -> 0, then truncate). ------------------------------------ */
int gg_render_flat(gg_context *ctx, const int32_t *d_pix2face, int64_t n_pixels, const double *d_face_tex,
                   int D, void *d_out, int out_dtype, void *stream);

/* ---- stage 1+2+4 fused: rasterize n views and write the rendered face texture directly (render_flat,
        meshes.py:1858-1942); d_out: n x H x W x D of out_dtype; d_pix2face (n x H x W int32) may be NULL. ---------- */
int gg_rasterize_render_flat(gg_context *ctx, const gg_camera *h_cams, int n, const double *d_face_tex, int D,
                             void *d_out, int out_dtype, int32_t *d_pix2face, void *stream);

/* ---- lens distortion: Metashape frame-camera model of MetashapeCameraSet.ideal_to_warped
        (cameras/derived_cameras.py:163-208).  f, cx, cy, W, H are the FULL-resolution intrinsics; image_scale the
        render scale of the rasters being warped (cameras.py:1027-1053). ---------------------------------------- */
typedef struct {
    double f, cx, cy;
    int32_t W, H;
    double k1, k2, k3, k4, p1, p2, b1, b2;
    double image_scale;
} gg_distortion;

/* For every pixel of an h x w output image: linear index of the nearest source pixel, -1 if the source falls outside
   the image (replaces make_distortion_map + inverse_map_interpolation, cameras.py:995-1062, utils/indexing.py:87-150).
   warped_to_ideal = 0: the output is the distorted image and samples an ideal (pinhole) raster -- the pix2face case,
   solved exactly with Newton iterations; 1: the output is the ideal image and samples the distorted one.
   d_src_rc (optional, 2*h*w float32): the continuous (row, col) source coordinates. */
int gg_build_warp_map(int device, const gg_distortion *h_dist, int h, int w, int warped_to_ideal,
                      int32_t *d_src_index, float *d_src_rc, void *stream);
/* out[p] = src_index[p] >= 0 ? in[src_index[p]] : fill -- the nearest-neighbour warp of a face-ID raster
   (utils/image.py:72-126 with interpolation_order = 0), integer-safe. */
int gg_gather_i32(int device, const int32_t *d_in, const int32_t *d_src_index, int64_t n_out, int32_t fill,
                  int32_t *d_out, void *stream);
/* save_renders' up-sampling of a render to the camera's native size (meshes.py:2312-2321:
   skimage.transform.resize(rendered, native_size, order = 0 for discrete textures, 1 otherwise), i.e.
   scipy.ndimage.zoom(grid_mode=True, mode="mirror")), optionally with the uint8 rule of meshes.py:2323-2334 fused in.
   d_in: h_in x w_in x D float64 (NaN = no face); d_out: h_out x w_out x D of out_dtype (GG_OUT_F64 | GG_OUT_U8). */
int gg_resize_render(int device, const double *d_in, int h_in, int w_in, int D, int h_out, int w_out, int order,
                     void *d_out, int out_dtype, void *stream);

/* ---- label_polygons (meshes.py:1141-1306 with sjoin_overlay=True): every face with a finite label whose
        2-D triangle lies within polygon p adds  area3D(face) * face_weight  to d_weights[p][class].
        d_xyz: V x 3 float64 (3-D area), d_xy: V x 2 float64 (planar coordinates shared with the polygons),
        d_labels: F float64 (NaN = unlabelled), d_face_weight: F float64 or NULL.  Polygons are sets of rings
        (exteriors and holes alike, even-odd rule): vertices d_poly_xy, ring r spans [ring_offsets[r], ring_offsets[r+1]),
        polygon p owns rings [poly_ring_offsets[p], poly_ring_offsets[p+1]), d_poly_bbox: n_polys x 4 (xmin ymin xmax ymax).
        d_weights (n_polys x n_classes float64) is accumulated into. -------------------------------------------------- */
int gg_label_polygons(int device, const double *d_xyz, const double *d_xy, const int32_t *d_faces,
                      const double *d_labels, const double *d_face_weight, int64_t F, const double *d_poly_xy,
                      const int32_t *d_ring_offsets, const int32_t *d_poly_ring_offsets, const double *d_poly_bbox,
                      int n_polys, int n_classes, double *d_weights, void *stream);
/* The same with sjoin_overlay=False (meshes.py:1263-1276: polygons.overlay(faces, how="identity")): every labelled face
   adds  area2D(face n polygon) * (area3D / area2D)(face) * face_weight  to each polygon it overlaps.  Rings must be
   oriented: exteriors counter-clockwise, holes clockwise (the areas of the pieces are signed sums over ring edges). */
int gg_label_polygons_overlay(int device, const double *d_xyz, const double *d_xy, const int32_t *d_faces,
                              const double *d_labels, const double *d_face_weight, int64_t F, const double *d_poly_xy,
                              const int32_t *d_ring_offsets, const int32_t *d_poly_ring_offsets,
                              const double *d_poly_bbox, int n_polys, int n_classes, double *d_weights, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOGRAYPHER_B200_H */
