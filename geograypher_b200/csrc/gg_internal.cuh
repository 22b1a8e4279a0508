// Internal declarations shared by the translation units of libgeograypher_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges are no-ops unless a profiler injects itself
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/geograypher_b200.h"

// ---- rasterization contract constants (DESIGN.md "Rasterization contract") -------------------------
#define GG_SUBPIX 256
#define GG_HALF 128
#define GG_SUBPIX_LOG2 8
#define GG_COORD_CLAMP 536870912.0f  // 2^29 sub-pixel units
#define GG_GUARD_PX 1048576.0f       // contract C6: guard band of +-2^20 px (half of what the clamp can hold)

// ---- tiling ------------------------------------------------------------------------------------------
#define GG_TILE_W 32           // one warp rasterizes one 32 x 8 px tile; lane = 8 consecutive px of one row
#define GG_TILE_H 8
#ifndef GG_RASTER_WARPS
#define GG_RASTER_WARPS 1      // independent warps per CTA (measured: 1 warp x 4 tiles beats 4 warps x 2 tiles by 3 %:
                               // a CTA's slot is not held by its slowest warp)
#endif
#define GG_RASTER_THREADS (32 * GG_RASTER_WARPS)
#ifndef GG_DENSE_TILES_PER_WARP
#define GG_DENSE_TILES_PER_WARP 1  // the same for the dense (pixel_sum) mode
#endif
#ifndef GG_TILES_PER_WARP
#define GG_TILES_PER_WARP 4    // consecutive tiles of a row per warp (not in the dense mode); see k_raster_tiles
#endif
#ifndef GG_RASTER_MIN_BLOCKS
#define GG_RASTER_MIN_BLOCKS (32 / GG_RASTER_WARPS)  // 64 registers per thread: 32 warps per SM
#endif
#ifndef GG_NSETS
#define GG_NSETS 3             // scratch-slot sets the software pipeline rotates through
#endif
#define GG_CHUNK 32            // faces staged per warp per pass (one per lane)
#define GG_BLOCK_FACES 128     // faces per cull block

struct __align__(16) GGFaceRec {  // one surviving face of one view, orientation-normalised (area2 > 0); 128 B
    int32_t A[3], B[3];    // edge gradients in sub-pixel units: E_k(P) = A_k*Px + B_k*Py + const, A = -dy, B = dx
    long long C[3];        // E_k at the centre of pixel (0,0), top-left bias folded in (>= 0 <=> inside)
    double w00, gx, gy;    // 1/z plane per pixel: w(j, i) = w00 + gx*j + gy*i
    float w0, w1, w2;      // 1/z_cam at the vertices (exact-depth path)
    int32_t face;          // face ID
    uint16_t jmin, jmax, imin, imax;  // pixel-centre index range, clamped to the raster
    unsigned long long tmask;  // which tiles of the bounding box the triangle can touch (bit = row-major index in the
                               // box), or ~0 when the box has more than 64 tiles (the fill pass then re-tests)
    int dup;                   // 1: an earlier record of this view carries the same face (second triangle of a clipped
                               // face): per-face passes over the records skip it
    int pad[5];                // one 128-byte line per record: written and read as 8 x 16 B
};
static_assert(sizeof(GGFaceRec) == 128, "GGFaceRec layout");

struct __align__(16) GGTileFace {  // a face record re-expressed relative to one 32 x 8 tile; 64 B
    int e[3], sx[3], sy[3];        // fast path: biased edge functions at the tile-origin pixel centre + per-pixel steps
    float w_org, gx, gy;           // fast path: 1/z plane relative to the tile origin
    unsigned lanemask;             // lanes (8-px strips) whose pixels intersect the face's pixel range
    int face;
    int rec;                       // record index (winner slot; source of the exact path)
    unsigned fast;                 // 1: 32-bit edges and the float plane are safe for this tile
};
static_assert(sizeof(GGTileFace) == 64, "GGTileFace layout");

struct GGCamBatch {  // passed by value as a __grid_constant__ kernel parameter (<= 4 KB)
    gg_camera cam[GG_MAX_VIEWS_PER_CALL];
};

struct GGViewScratch {  // device pointers of one batch slot
    int32_t *vis_blocks;   // [n_blocks]
    GGFaceRec *recs;       // [cap_recs]
    int32_t *tile_count;   // [n_tiles]   faces per tile after setup; fill cursor (list start .. list end) afterwards
    int32_t *tile_offset;  // [n_tiles]   start of the tile's list in bins (lists are not in tile order)
    GGTileFace *bins;      // [cap_bins]  per-(tile, face) setups grouped by tile
    int32_t *winner;       // [F]  last (row-major) pixel won by each FACE in this view, -1 = none (fused aggregation)
    int32_t *counters;     // [8]: 0 n_vis_blocks, 1 n_recs, 2 n_bin_entries, 3 overflow flag, 4 bg winner,
                           //      5 rec index of face F-1 (or -1)
};

struct GGViewBatch {
    GGViewScratch v[GG_MAX_VIEWS_PER_CALL];
};

struct GGPredBatch {  // device pointers of the views' prediction images
    const void *p[GG_MAX_VIEWS_PER_CALL];
};

// ---- per-stage launch counting and (optional) CUDA-event timing ------------------------------------------
enum GGStage {
    GG_ST_MESH = 0, GG_ST_PROJECT, GG_ST_CULL, GG_ST_SETUP, GG_ST_SCAN, GG_ST_FILL, GG_ST_RASTER, GG_ST_LAST_PIXEL,
    GG_ST_RESOLVE, GG_ST_PIXEL_SUM, GG_ST_FINALIZE, GG_ST_RENDER_FLAT, GG_ST_MISC, GG_ST_STAGE, GG_ST_COUNT
};

struct GGProfPending {
    int stage;
    cudaEvent_t a, b;
};

struct GGProfiler {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    std::vector<GGProfPending> pending;
    double ms[GG_ST_COUNT] = {0};
    int64_t launches[GG_ST_COUNT] = {0};
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

struct gg_context {
    int device = 0;
    GGProfiler prof;
    // mesh
    int64_t V = 0, F = 0;
    float4 *d_verts = nullptr;   // [V] xyz + pad
    int4 *d_faces = nullptr;     // [F] i0,i1,i2,face_id
    int dense_prefetch = 1;      // GG_MODE_PIXEL_SUM: L2 prefetch of the score tiles (GG_DENSE_PREFETCH=0 turns it off)
    // Rasterizer variant whose lanes walk their own face lists (k_raster_tiles<..., LL = true>).  It pays off when tile
    // lists are long (c5, 10.5 faces per tile: +9 %) and costs on short ones (c2, 4.3 per tile: -6 %), so by default
    // gg_sync picks it from the longest view of the batches it has just waited for (tile entries per tile >=
    // lane_lists_min_entries).  GG_LANE_LISTS=0 / 1 pins it off / on.
    int lane_lists = 0;
    int lane_lists_auto = 1;
    float lane_lists_min_entries = 7.0f;
    int64_t last_n_tiles = 0;    // tiles per view of the most recent rasterization
    float *d_block_lo = nullptr; // [n_blocks*3]
    float *d_block_hi = nullptr; // [n_blocks*3]
    int64_t n_blocks = 0;
    // scratch
    int64_t cap_recs = 0, cap_bins = 0;
    int64_t req_recs = 0, req_bins = 0;  // user reservation (0 = automatic)
    int n_slots = 0;
    int64_t slot_tiles = 0;
    char *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    GGViewBatch vset[GG_NSETS];   // sets of batch slots: batch k+1 is binned while batch k is rasterized and batch
                                  // k-1 resolved (the third set keeps a slow resolve -- rows fetched over PCIe -- from
                                  // holding back the binning of batch k+2)
    int cur = 0;                  // set used by the work being enqueued / most recently enqueued
    // software pipeline of the fused aggregation (gg_project_aggregate): binning on sA, raster + resolve on sB
    cudaStream_t sA = nullptr, sB = nullptr, sC = nullptr;  // binning / rasterizer / resolve
    cudaEvent_t ev_mid[GG_NSETS] = {};            // rasterizer of a set done (resolve may start)
    cudaEvent_t ev_user = nullptr, ev_bin[GG_NSETS] = {}, ev_ras[GG_NSETS] = {};
    bool ras_pending[GG_NSETS] = {};
    int parity = 0;
    bool pipeline = true;
    int32_t *d_winner = nullptr;  // [F] last-pixel winner per face (dense, unfused aggregation)
    int64_t winner_cap = 0;
    int32_t *d_wdense = nullptr;  // [n_slots * F] per-view per-face winners of the fused aggregation
    int64_t wdense_cap = 0;
    int setup_ctas = 0, fill_ctas = 0;  // > 0: CTAs per SM, summed over the views of a batch, of k_setup_faces / k_fill_bins
    char *d_stage = nullptr;      // rows fetched from prediction images that live in host memory
    size_t stage_bytes = 0;
    int stage_host_rows = 1;      // GG_STAGE_HOST_ROWS=0: resolve reads the host images directly
    int stage_inflight = 2560;    // PCIe read requests (32-byte sectors) the host-row fetch keeps in flight (GG_STAGE_INFLIGHT)
    int stage_grid = 0;           // > 0: absolute grid size of the host-row fetch kernel (GG_STAGE_GRID, experiments)
    int stage_ctas = 0;           // CTAs per SM of the host-row fetch kernel (GG_STAGE_CTAS)
    int32_t *d_raster = nullptr;  // internal n x H x W raster when the caller does not want pix2face back
    int64_t raster_cap = 0;
    int32_t *d_sticky = nullptr;  // [4] since the last gg_sync: OR of the batches' overflow flags, most face records any
                                  // view wanted, most tile entries any view wanted, unused
    int64_t last_overflow[3] = {0, 0, 0};  // what the last failing gg_sync read from d_sticky (gg_overflow_info)
    int last_batch_n = 0;
    int sm_count = 148;
};

void gg_set_error(const std::string &msg);
const char *gg_stage_label(int stage);  // gg_api.cu: the names gg_stage_name() returns
int gg_cuda_fail(cudaError_t e, const char *what);

// Launch a kernel (or any stream operation) under a stage label: counts it, and brackets it with events when
// profiling is enabled.  Usage: GG_LAUNCH(ctx, GG_ST_RASTER, st, k_raster<<<g, b, 0, st>>>(...));
#define GG_LAUNCH(ctx, stage, st, ...)                                      \
    do {                                                                    \
        GGProfPending _p{stage, nullptr, nullptr};                          \
        if ((ctx)->prof.on) {                                               \
            nvtxRangePushA(gg_stage_label(stage)); /* NVTX range per stage */ \
            _p.a = (ctx)->prof.get();                                       \
            _p.b = (ctx)->prof.get();                                       \
            cudaEventRecord(_p.a, st);                                      \
        }                                                                   \
        __VA_ARGS__;                                                        \
        (ctx)->prof.launches[stage] += 1;                                   \
        if ((ctx)->prof.on) {                                               \
            cudaEventRecord(_p.b, st);                                      \
            (ctx)->prof.pending.push_back(_p);                              \
            nvtxRangePop();                                                 \
        }                                                                   \
        GG_CUDA(cudaGetLastError());                                        \
    } while (0)

#define GG_CUDA(call)                                            \
    do {                                                         \
        cudaError_t _e = (call);                                 \
        if (_e != cudaSuccess) return gg_cuda_fail(_e, #call);   \
    } while (0)

// gg_raster.cu
int gg_ensure_scratch(gg_context *ctx, int n_views, int W, int H);
// make `st` wait for everything the internal pipeline streams still have in flight
int gg_pipeline_drain(gg_context *ctx, cudaStream_t st);
int gg_launch_mesh_blocks(gg_context *ctx, cudaStream_t st);
int gg_launch_project(gg_context *ctx, const gg_camera *cams, int n, int32_t *dX, int32_t *dY, float *dinvz,
                      uint8_t *dvalid, cudaStream_t st);
// h_pred != nullptr selects the fused dense per-pixel-sum epilogue (GG_MODE_PIXEL_SUM, C <= 32)
// st_bin == st_ras: everything on one stream (set 0).  Otherwise binning goes to st_bin, the rasterizer to st_ras,
// linked by ev_bin[ctx->cur].
int gg_launch_rasterize(gg_context *ctx, const gg_camera *cams, int n, int32_t *d_pix2face, float *d_depth,
                        int want_winners, int compat_bg, cudaStream_t st_bin, cudaStream_t st_ras,
                        const void *const *h_pred = nullptr, int pred_kind = 0, int C = 0, double *d_sum = nullptr,
                        int32_t *d_count = nullptr, const double *d_tex = nullptr, int D = 0, void *d_out = nullptr,
                        int out_dtype = 0);
// gg_aggregate.cu: consume the per-face winners of the last rasterization batch (all views, in view order)
int gg_launch_resolve_batch(gg_context *ctx, int n, const void *const *h_pred, int pred_kind, int C, int mode, int flags,
                            double *d_sum, int32_t *d_count, cudaStream_t st);
// gg_aggregate.cu
int gg_launch_aggregate(gg_context *ctx, const int32_t *d_pix2face, int H, int W, const void *d_pred,
                        int pred_kind, int C, int mode, int compat, double *d_sum, int32_t *d_count,
                        cudaStream_t st);
int gg_launch_compact_winners(gg_context *ctx, int n, int flags, int32_t *d_pairs, int64_t cap, int32_t *d_counts,
                              cudaStream_t st);
int gg_launch_accumulate_rows(gg_context *ctx, const int32_t *d_pairs, int64_t n_rows, const void *d_rows, int pred_kind,
                              int C, int mode, int flags, double *d_sum, int32_t *d_count, cudaStream_t st);
int gg_launch_finalize(gg_context *ctx, double *d_sum, const int32_t *d_count, int64_t F, int C, double *d_avg,
                       double *d_argmax, cudaStream_t st);
int gg_launch_render_flat(gg_context *ctx, const int32_t *d_pix2face, int64_t P, const double *d_tex, int D,
                          void *d_out, int out_dtype, cudaStream_t st);
