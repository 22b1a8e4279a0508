// Stage 3 (per-face aggregation of prediction images), its epilogue (mean + argmax) and stage 4 (render_flat
// gather).
//
// Reference semantics restated (file:line under /root/reference/geograypher):
//   meshes/meshes.py:1988-2001   textured_faces[flat_pix2face] = flat_img   -> LAST pixel (row-major) per face
//   meshes/meshes.py:2056-2067   sum over views with NaN -> 0; a face counts once per view with a finite value
//   meshes/meshes.py:2069-2082   rows never seen -> NaN; mean = sum / count
//   meshes/derived_meshes.py:480-520   one-hot votes from the face's last pixel
//   meshes/meshes.py:1921-1937   render_flat gather;  :2323-2334 uint8 cast rule
//   utils/indexing.py:9-32       find_argmax_nonzero_value
//   predictors/segmentor.py:59-69  inds_to_one_hot (GG_PRED_INDEX_U8 expands it on the fly)
#include <algorithm>
#include "gg_internal.cuh"

namespace {

// GG_FLAG_COMPAT_NEG (1): meshes.py:2000, background pixels (-1) index face F-1
// GG_FLAG_KEEP_NAN (2): single-view corner of meshes.py:2056-2057, the first projection keeps its NaNs
// GG_FLAG_ASSIGN (4): project_images (meshes.py:1991-2002), write the face's row instead of accumulating;
//                     count[f] = 1 marks a row that was written

// ---- pass A: last pixel (row-major) of every face in this raster -------------------------------------
// One thread handles 4 consecutive pixels; it issues an atomicMax only for the last pixel of each run of equal
// face IDs, and skips its final run when the next lane starts with the same face (that lane, or a later one,
// holds a larger pixel index for it).
__global__ void __launch_bounds__(256) k_last_pixel(const int32_t *__restrict__ p2f, int64_t P, int32_t last_face,
                                                    int compat, int aligned16, int32_t *__restrict__ winner) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p0 = t * 4;
    int f[4];
    if (aligned16 && p0 + 3 < P) {
        const int4 v = *reinterpret_cast<const int4 *>(p2f + p0);
        f[0] = v.x;
        f[1] = v.y;
        f[2] = v.z;
        f[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = (p0 + i < P) ? p2f[p0 + i] : -2;
    }
    if (compat) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (f[i] == -1) f[i] = last_face;
    }
    const int next_first = __shfl_down_sync(0xffffffffu, f[0], 1);
    const bool has_next = (threadIdx.x & 31) != 31;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int nxt = (i < 3) ? f[i + 1] : (has_next ? next_first : -3);
        if (f[i] >= 0 && f[i] != nxt) atomicMax(&winner[f[i]], (int32_t)(p0 + i));
    }
}

template <typename T>
__device__ __forceinline__ double load_score(const T *pred, int64_t idx) {
    return (double)pred[idx];
}

// ---- pass B: one thread per face; consumes and resets the winner ---------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_resolve_dense(int32_t *__restrict__ winner, int64_t F,
                                                       const T *__restrict__ pred, int C, int flags,
                                                       double *__restrict__ sum, int32_t *__restrict__ count) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int32_t p = winner[f];
    if (p < 0) return;
    winner[f] = -1;
    bool any_finite = false;
    if (flags & GG_FLAG_ASSIGN) {
        for (int c = 0; c < C; ++c) sum[f * C + c] = load_score(pred, (int64_t)p * C + c);
        count[f] = 1;
        return;
    }
    for (int c = 0; c < C; ++c) {
        const double v = load_score(pred, (int64_t)p * C + c);
        any_finite = any_finite || isfinite(v);
        if (!isnan(v) || (flags & GG_FLAG_KEEP_NAN)) sum[f * C + c] += v;
    }
    if (any_finite) count[f] += 1;
}

__global__ void __launch_bounds__(256) k_resolve_index(int32_t *__restrict__ winner, int64_t F,
                                                       const uint8_t *__restrict__ pred, int C, int flags,
                                                       double *__restrict__ sum, int32_t *__restrict__ count) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int32_t p = winner[f];
    if (p < 0) return;
    winner[f] = -1;
    const int cls = pred[p];
    if (flags & GG_FLAG_ASSIGN) {
        for (int c = 0; c < C; ++c) sum[f * C + c] = (c == cls) ? 1.0 : 0.0;
        count[f] = 1;
        return;
    }
    if (cls < C) sum[f * C + cls] += 1.0;
    count[f] += 1;  // a one-hot row is always finite, even the all-zero row of an ignored pixel
}

template <typename T>
__global__ void __launch_bounds__(256) k_resolve_vote(int32_t *__restrict__ winner, int64_t F,
                                                      const T *__restrict__ pred, int C, double *__restrict__ sum,
                                                      int32_t *__restrict__ count) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int32_t p = winner[f];
    if (p < 0) return;
    winner[f] = -1;
    const double v = (double)pred[p];
    if (!isfinite(v)) return;
    count[f] += 1;
    const long long cls = (long long)v;
    if (cls >= 0 && cls < C) sum[f * C + cls] += 1.0;
}

// ---- fused path: consume the per-face winners that k_raster_tiles<true> left for a batch of views -------------
// blockIdx.y = view, grid-stride over that view's face records (the record count lives on the device).  A face seen
// by several views of the batch is handled by the thread of the EARLIEST such view, which then walks the later
// views in order: every face's float64 sum is accumulated in view order (bit-identical to the reference's loop,
// meshes.py:2056-2062) with no atomics.  The extra index r == n_recs covers the meshes.py:2000 quirk when face F-1
// has no record of its own but background pixels were written to its slot.
template <typename T>
__global__ void __launch_bounds__(256) k_resolve_batch(const __grid_constant__ GGViewBatch views, int n_views, int64_t F,
                                                       const __grid_constant__ GGPredBatch preds, int C, int pred_kind,
                                                       int mode, int flags, double *__restrict__ sum,
                                                       int32_t *__restrict__ count) {
    // A scratch overflow anywhere in the batch voids the whole batch: nothing is accumulated, so the host can grow
    // the scratch and replay the same views in the same order.
    for (int v = 0; v < n_views; ++v)
        if (views.v[v].counters[3] != 0) return;
    const int view = blockIdx.y;
    const GGViewScratch &vs = views.v[view];
    const int n_recs = vs.counters[1];
    const bool compat = (flags & GG_FLAG_COMPAT_NEG) != 0;
    const int n = n_recs + ((compat && vs.counters[5] < 0) ? 1 : 0);
    const bool keep_nan = (flags & GG_FLAG_KEEP_NAN) != 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (r < n_recs && vs.recs[r].dup) continue;  // second triangle of a clipped face: the first one speaks for it
        const int64_t f = r < n_recs ? (int64_t)vs.recs[r].face : F - 1;
        if (vs.winner[f] < 0) continue;
        bool earlier = false;
        for (int u = 0; u < view; ++u) earlier = earlier || (views.v[u].winner[f] >= 0);
        if (earlier) continue;
        if (mode == GG_MODE_VOTE) {
            int cnt = 0;
            for (int u = view; u < n_views; ++u) {
                const int p = views.v[u].winner[f];
                if (p < 0) continue;
                const double v = (double)((const T *)preds.p[u])[p];
                if (!isfinite(v)) continue;
                cnt += 1;
                const long long cls = (long long)v;
                if (cls >= 0 && cls < C) sum[f * C + cls] += 1.0;
            }
            count[f] += cnt;
        } else if (pred_kind == GG_PRED_INDEX_U8) {
            int cnt = 0;
            for (int u = view; u < n_views; ++u) {
                const int p = views.v[u].winner[f];
                if (p < 0) continue;
                const int cls = (int)((const T *)preds.p[u])[p];
                if (cls < C) sum[f * C + cls] += 1.0;
                cnt += 1;  // a one-hot row is always finite, even the all-zero row of an ignored pixel
            }
            count[f] += cnt;
        } else {
            unsigned fin = 0;  // bit u: view u contributed a finite value
            for (int c0 = 0; c0 < C; c0 += 8) {
                double acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = (c0 + j < C) ? sum[f * C + c0 + j] : 0.0;
                for (int u = view; u < n_views; ++u) {
                    const int p = views.v[u].winner[f];
                    if (p < 0) continue;
                    const T *row = (const T *)preds.p[u] + (int64_t)p * C + c0;
                    double vals[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) vals[j] = (c0 + j < C) ? (double)row[j] : 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (c0 + j < C) {
                            if (isfinite(vals[j])) fin |= 1u << u;
                            if (!isnan(vals[j]) || keep_nan) acc[j] += vals[j];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c0 + j < C) sum[f * C + c0 + j] = acc[j];
            }
            count[f] += __popc(fin);
        }
    }
}

// ---- non-reference dense mode, unfused: every pixel adds its scores -------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_pixel_sum(const int32_t *__restrict__ p2f, int64_t P,
                                                   const T *__restrict__ pred, int C, int index_kind,
                                                   double *__restrict__ sum, int32_t *__restrict__ count) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int f = p2f[p];
    if (f < 0) return;
    if (index_kind) {
        const int cls = (int)pred[p];
        if (cls < C) atomicAdd(&sum[(int64_t)f * C + cls], 1.0);
    } else {
        for (int c = 0; c < C; ++c) {
            const double v = (double)pred[p * C + c];
            if (!isnan(v)) atomicAdd(&sum[(int64_t)f * C + c], v);
        }
    }
    atomicAdd(&count[f], 1);
}

// ---- epilogue: mean + argmax -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_finalize(double *__restrict__ sum, const int32_t *__restrict__ count,
                                                  int64_t F, int C, double *__restrict__ avg,
                                                  double *__restrict__ argmax) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int32_t n = count[f];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (n == 0) {
        for (int c = 0; c < C; ++c) {
            sum[f * C + c] = nan;
            if (avg) avg[f * C + c] = nan;
        }
        if (argmax) argmax[f] = nan;
        return;
    }
    double best = 0.0, total = 0.0;
    int best_c = 0;
    bool bad = false;
    const double dn = (double)n;
    for (int c = 0; c < C; ++c) {
        const double a = sum[f * C + c] / dn;
        if (avg) avg[f * C + c] = a;
        bad = bad || !isfinite(a);
        total += a;
        if (c == 0 || a > best) {
            best = a;
            best_c = c;
        }
    }
    if (argmax) argmax[f] = (bad || total == 0.0) ? nan : (double)best_c;
}

// ---- stage 4: gather ---------------------------------------------------------------------------------------
template <typename OUT>
__device__ __forceinline__ OUT convert_out(double v);
template <>
__device__ __forceinline__ double convert_out<double>(double v) {
    return v;
}
template <>
__device__ __forceinline__ float convert_out<float>(double v) {
    return (float)v;
}
template <>
__device__ __forceinline__ uint8_t convert_out<uint8_t>(double v) {
    // save_renders: < 0, > 255 or non-finite -> NULL_TEXTURE_INT_VALUE (0), then astype(uint8) truncates
    if (!(v >= 0.0) || v > 255.0 || !isfinite(v)) return 0;
    return (uint8_t)v;
}

template <typename OUT>
__global__ void __launch_bounds__(256) k_render_flat(const int32_t *__restrict__ p2f, int64_t P,
                                                     const double *__restrict__ tex, int D, OUT *__restrict__ out) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const int f = p2f[p];
        for (int d = 0; d < D; ++d) out[p * D + d] = convert_out<OUT>(f >= 0 ? tex[(int64_t)f * D + d] : nan);
    }
}

__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

static int ensure_winner(gg_context *ctx, cudaStream_t st) {
    if (ctx->winner_cap >= ctx->F && ctx->d_winner) return GG_OK;
    if (ctx->d_winner) {
        GG_CUDA(cudaDeviceSynchronize());
        GG_CUDA(cudaFree(ctx->d_winner));
        ctx->d_winner = nullptr;
    }
    GG_CUDA(cudaMalloc(&ctx->d_winner, (size_t)ctx->F * 4));
    ctx->winner_cap = ctx->F;
    GG_LAUNCH(ctx, GG_ST_MISC, st, k_fill_i32<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->d_winner, ctx->F, -1));
    return GG_OK;
}

int gg_launch_aggregate(gg_context *ctx, const int32_t *d_pix2face, int H, int W, const void *d_pred,
                        int pred_kind, int C, int mode, int flags, double *d_sum, int32_t *d_count,
                        cudaStream_t st) {
    const int64_t P = (int64_t)H * W, F = ctx->F;
    if (P >= (1LL << 31)) {
        gg_set_error("gg_aggregate: raster larger than 2^31 pixels");
        return GG_ERR_INVALID;
    }
    if (mode == GG_MODE_PIXEL_SUM) {
        const unsigned g = (unsigned)((P + 255) / 256);
        switch (pred_kind) {
            case GG_PRED_F32: GG_LAUNCH(ctx, GG_ST_PIXEL_SUM, st, k_pixel_sum<float><<<g, 256, 0, st>>>(d_pix2face, P, (const float *)d_pred, C, 0, d_sum, d_count)); break;
            case GG_PRED_F64: GG_LAUNCH(ctx, GG_ST_PIXEL_SUM, st, k_pixel_sum<double><<<g, 256, 0, st>>>(d_pix2face, P, (const double *)d_pred, C, 0, d_sum, d_count)); break;
            case GG_PRED_U8: GG_LAUNCH(ctx, GG_ST_PIXEL_SUM, st, k_pixel_sum<uint8_t><<<g, 256, 0, st>>>(d_pix2face, P, (const uint8_t *)d_pred, C, 0, d_sum, d_count)); break;
            case GG_PRED_INDEX_U8: GG_LAUNCH(ctx, GG_ST_PIXEL_SUM, st, k_pixel_sum<uint8_t><<<g, 256, 0, st>>>(d_pix2face, P, (const uint8_t *)d_pred, C, 1, d_sum, d_count)); break;
            default: gg_set_error("gg_aggregate: bad pred_kind"); return GG_ERR_INVALID;
        }
        return GG_OK;
    }
    int rc = ensure_winner(ctx, st);
    if (rc != GG_OK) return rc;
    const int64_t threads = (P + 3) / 4;
    GG_LAUNCH(ctx, GG_ST_LAST_PIXEL, st,
              k_last_pixel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_pix2face, P, (int32_t)(F - 1),
                                                                              flags & GG_FLAG_COMPAT_NEG,
                                                                              (((uintptr_t)d_pix2face) & 15) == 0,
                                                                              ctx->d_winner));
    const unsigned gf = (unsigned)((F + 255) / 256);
    if (mode == GG_MODE_LAST_PIXEL) {
        switch (pred_kind) {
            case GG_PRED_F32: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_dense<float><<<gf, 256, 0, st>>>(ctx->d_winner, F, (const float *)d_pred, C, flags, d_sum, d_count)); break;
            case GG_PRED_F64: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_dense<double><<<gf, 256, 0, st>>>(ctx->d_winner, F, (const double *)d_pred, C, flags, d_sum, d_count)); break;
            case GG_PRED_U8: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_dense<uint8_t><<<gf, 256, 0, st>>>(ctx->d_winner, F, (const uint8_t *)d_pred, C, flags, d_sum, d_count)); break;
            case GG_PRED_INDEX_U8: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_index<<<gf, 256, 0, st>>>(ctx->d_winner, F, (const uint8_t *)d_pred, C, flags, d_sum, d_count)); break;
            default: gg_set_error("gg_aggregate: bad pred_kind"); return GG_ERR_INVALID;
        }
    } else if (mode == GG_MODE_VOTE) {
        switch (pred_kind) {
            case GG_PRED_F32: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_vote<float><<<gf, 256, 0, st>>>(ctx->d_winner, F, (const float *)d_pred, C, d_sum, d_count)); break;
            case GG_PRED_F64: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_vote<double><<<gf, 256, 0, st>>>(ctx->d_winner, F, (const double *)d_pred, C, d_sum, d_count)); break;
            case GG_PRED_U8:
            case GG_PRED_INDEX_U8: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_vote<uint8_t><<<gf, 256, 0, st>>>(ctx->d_winner, F, (const uint8_t *)d_pred, C, d_sum, d_count)); break;
            default: gg_set_error("gg_aggregate: bad pred_kind"); return GG_ERR_INVALID;
        }
    } else {
        gg_set_error("gg_aggregate: bad mode");
        return GG_ERR_INVALID;
    }
    return GG_OK;
}

// The winner arrays are dense per face and only ~1.5 % of their entries are written per view: instead of clearing them
// before every batch, whoever consumed a batch's winners puts the touched entries back to -1.  Runs also for a batch
// that was voided by a scratch overflow (the views that did not overflow have written their winners).
__global__ void __launch_bounds__(256) k_reset_winners(const __grid_constant__ GGViewBatch views, int64_t F, int flags) {
    const GGViewScratch &vs = views.v[blockIdx.y];
    const int n_recs = vs.counters[1];
    const bool compat = (flags & GG_FLAG_COMPAT_NEG) != 0;
    const int n = n_recs + (compat ? 1 : 0);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
        vs.winner[r < n_recs ? (int64_t)vs.recs[r].face : F - 1] = -1;
}

static int launch_reset_winners(gg_context *ctx, int n, int flags, cudaStream_t st) {
    GG_LAUNCH(ctx, GG_ST_RESOLVE, st,
              k_reset_winners<<<dim3((unsigned)(ctx->sm_count * 2), n), 256, 0, st>>>(ctx->vset[ctx->cur], ctx->F, flags));
    return GG_OK;
}

// ---- prediction images in (pinned) HOST memory ------------------------------------------------------------------
// The fused last-pixel / vote aggregation needs one row per visible face.  When the images were left in page-locked
// host memory the rows are fetched over PCIe: first ALL of them, in parallel, into a device staging table (one thread
// per element, so a warp's loads of one row share their sectors), then k_resolve_batch walks the views in order on
// the staged copy.  The winners are re-pointed from pixel indices to staging rows in between.
// Two steps.  (1) k_compact_stage walks the view's face records (a chain of dependent L2 reads: record -> face ->
// winner) and lists the winning pixel of every visible face densely, re-pointing the face's winner from the pixel to
// its row in the staging table.  (2) k_fetch_rows is then nothing but the gather itself: one thread per ELEMENT of
// the listed rows, flat indexing, so that consecutive lanes read consecutive bytes of a row and a warp covers 32 / E
// rows -- the shape that reaches the link's scattered-read rate in scripts/pcie_rows_bench.cu (~0.26 G 40-byte
// rows/s; the one-kernel version that did the record walk and the fetch together ran 3x below it).  The link needs
// only a few hundred reads in flight, so the fetch is a SMALL grid (GG_STAGE_CTAS CTAs per SM): launched on the
// high-priority resolve stream it slips into the first slots the rasterizer of the NEXT batch frees and runs beside
// it instead of time-slicing the machine with it.
__global__ void __launch_bounds__(256) k_compact_stage(const __grid_constant__ GGViewBatch views, int n_views, int64_t F,
                                                       int flags, int32_t *__restrict__ row_pix, int64_t rows_per_view,
                                                       int32_t *__restrict__ row_count) {
    for (int v = 0; v < n_views; ++v)
        if (views.v[v].counters[3] != 0) return;
    const int view = blockIdx.y;
    const GGViewScratch &vs = views.v[view];
    const int n_recs = vs.counters[1];
    const bool compat = (flags & GG_FLAG_COMPAT_NEG) != 0;
    const int n = n_recs + ((compat && vs.counters[5] < 0) ? 1 : 0);
    int32_t *__restrict__ pix = row_pix + (int64_t)view * rows_per_view;
    for (int r0 = blockIdx.x * blockDim.x; r0 < n; r0 += gridDim.x * blockDim.x) {
        const int r = r0 + threadIdx.x;
        int64_t f = -1;
        int p = -1;
        if (r < n && !(r < n_recs && vs.recs[r].dup)) {
            f = r < n_recs ? (int64_t)vs.recs[r].face : F - 1;
            p = vs.winner[f];
        }
        const bool keep = p >= 0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        int base = 0;
        if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(&row_count[view], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const int row = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1));  // < rows_per_view: one row per record
            pix[row] = p;
            vs.winner[f] = row;  // nobody else reads this view's winner of this face before k_resolve_batch
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 8) k_fetch_rows(int n_views, const __grid_constant__ GGPredBatch preds, int E,
                                                       unsigned div_magic, const int32_t *__restrict__ row_pix,
                                                       int64_t rows_per_view, const int32_t *__restrict__ row_count,
                                                       T *__restrict__ stage) {
    for (int view = 0; view < n_views; ++view) {
        const T *__restrict__ pred = (const T *)preds.p[view];
        const int32_t *__restrict__ pix = row_pix + (int64_t)view * rows_per_view;
        T *__restrict__ out = stage + (int64_t)view * rows_per_view * E;
        const unsigned total = (unsigned)row_count[view] * (unsigned)E;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const unsigned row = E == 1 ? i : __umulhi(i, div_magic);  // i / E (exact: i < 2^27, see the launch)
            const unsigned e = i - row * (unsigned)E;
            out[i] = pred[(int64_t)pix[row] * E + e];
        }
    }
}

template <typename T>
static int stage_host_rows(gg_context *ctx, int n, GGPredBatch &pb, int E, int flags, cudaStream_t st) {
    const int64_t rows_per_view = ctx->cap_recs + 1;
    if (rows_per_view * E >= (1LL << 27)) {
        gg_set_error("gg_project_aggregate: too many face records x channels for the host-row staging table");
        return GG_ERR_INVALID;
    }
    // [n][rows_per_view][E] staged rows, then [n][rows_per_view] winning pixels, then [32] row counts
    const size_t b_rows = (size_t)n * rows_per_view * E * sizeof(T), b_rows_al = (b_rows + 255) / 256 * 256;
    const size_t b_pix = (size_t)n * rows_per_view * sizeof(int32_t), b_pix_al = (b_pix + 255) / 256 * 256;
    const size_t need = b_rows_al + b_pix_al + 256;
    if (ctx->stage_bytes < need) {
        if (ctx->d_stage) {
            GG_CUDA(cudaDeviceSynchronize());
            GG_CUDA(cudaFree(ctx->d_stage));
            ctx->d_stage = nullptr;
        }
        GG_CUDA(cudaMalloc(&ctx->d_stage, need));
        ctx->stage_bytes = need;
    }
    T *stage = (T *)ctx->d_stage;
    int32_t *row_pix = (int32_t *)(ctx->d_stage + b_rows_al);
    int32_t *row_count = (int32_t *)(ctx->d_stage + b_rows_al + b_pix_al);
    GG_CUDA(cudaMemsetAsync(row_count, 0, GG_MAX_VIEWS_PER_CALL * sizeof(int32_t), st));
    GG_LAUNCH(ctx, GG_ST_RESOLVE, st,
              k_compact_stage<<<dim3((unsigned)(ctx->sm_count * 2), n), 256, 0, st>>>(ctx->vset[ctx->cur], n, ctx->F, flags,
                                                                                    row_pix, rows_per_view, row_count));
    // floor(i / E) == umulhi(i, ceil(2^32 / E)) for i * E < 2^32 -- far beyond the 2^27 elements checked above
    const unsigned magic = E > 1 ? (unsigned)((0x100000000ULL + (unsigned)E - 1) / (unsigned)E) : 0u;
    // Grid of the fetch: what matters is how many PCIe read requests are in flight.  Every thread has one load
    // outstanding and a row of E elements is ceil(E * sizeof(T) / 32) sectors, so the grid is sized for
    // ctx->stage_inflight sectors (2 560): measured on the benchmark host, 500 c2 views end to end take 66 ms at that
    // depth against 81 ms with 2 CTAs per SM (~15 k sectors in flight) and 106 ms with 8 -- requests beyond what the
    // link can hold queue up in front of it and slow it down (profiles/r02_e2e_stage_grid.txt).
    const int sectors_per_row = (int)((E * sizeof(T) + 31) / 32);
    const int64_t threads = (int64_t)ctx->stage_inflight / sectors_per_row * E;
    int grid = (int)std::min<int64_t>(std::max<int64_t>((threads + 255) / 256, 4), (int64_t)ctx->sm_count * 2);
    if (ctx->stage_ctas > 0) grid = ctx->sm_count * ctx->stage_ctas;
    if (ctx->stage_grid > 0) grid = ctx->stage_grid;
    GG_LAUNCH(ctx, GG_ST_STAGE, st,
              k_fetch_rows<T><<<(unsigned)grid, 256, 0, st>>>(n, pb, E, magic, row_pix, rows_per_view, row_count, stage));
    for (int i = 0; i < n; ++i) pb.p[i] = stage + (int64_t)i * rows_per_view * E;
    return GG_OK;
}

// ---- the path split in two for prediction images that sit in PAGEABLE host memory ------------------------------
// The GPU cannot read pageable memory and uploading a whole image costs 100x more than the path itself, so the host
// gathers the few rows that matter: k_compact_winners lists, per view, (face, last pixel) of every visible face; the
// host picks those rows out of its array and sends them back; k_accumulate_rows applies one view's rows.
__global__ void __launch_bounds__(256) k_compact_winners(const __grid_constant__ GGViewBatch views, int n_views, int64_t F,
                                                         int flags, int32_t *__restrict__ pairs, int64_t cap,
                                                         int32_t *__restrict__ counts, int32_t *__restrict__ sticky) {
    for (int v = 0; v < n_views; ++v)
        if (views.v[v].counters[3] != 0) return;  // overflowed batch: counts stay 0, the sticky flag reports it
    const int view = blockIdx.y;
    const GGViewScratch &vs = views.v[view];
    const int n_recs = vs.counters[1];
    const bool compat = (flags & GG_FLAG_COMPAT_NEG) != 0;
    const int n = n_recs + ((compat && vs.counters[5] < 0) ? 1 : 0);
    for (int r0 = blockIdx.x * blockDim.x; r0 < n; r0 += gridDim.x * blockDim.x) {
        const int r = r0 + threadIdx.x;
        int f = -1, p = -1;
        if (r < n && !(r < n_recs && vs.recs[r].dup)) {
            f = r < n_recs ? vs.recs[r].face : (int)(F - 1);
            p = vs.winner[f];
        }
        const bool keep = p >= 0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        int base = 0;
        if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(&counts[view], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const int64_t idx = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
            if (idx < cap) {
                pairs[2 * ((int64_t)view * cap + idx)] = f;
                pairs[2 * ((int64_t)view * cap + idx) + 1] = p;
            } else if (!(flags & GG_FLAG_TRUNCATE)) {
                atomicOr(sticky, 1);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_accumulate_rows(const int32_t *__restrict__ pairs, int64_t n_rows,
                                                         const T *__restrict__ rows, int C, int pred_kind, int mode,
                                                         int flags, double *__restrict__ sum, int32_t *__restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int64_t f = pairs[2 * i];
    const bool keep_nan = (flags & GG_FLAG_KEEP_NAN) != 0;
    if (mode == GG_MODE_VOTE) {
        const double v = (double)rows[i];
        if (!isfinite(v)) return;
        const long long cls = (long long)v;
        if (cls >= 0 && cls < C) sum[f * C + cls] += 1.0;
        count[f] += 1;
    } else if (pred_kind == GG_PRED_INDEX_U8) {
        const int cls = (int)rows[i];
        if (cls < C) sum[f * C + cls] += 1.0;
        count[f] += 1;
    } else {
        bool fin = false;
        for (int c = 0; c < C; ++c) {
            const double v = (double)rows[i * C + c];
            fin = fin || isfinite(v);
            if (!isnan(v) || keep_nan) sum[f * C + c] += v;
        }
        count[f] += fin ? 1 : 0;
    }
}

int gg_launch_compact_winners(gg_context *ctx, int n, int flags, int32_t *d_pairs, int64_t cap, int32_t *d_counts,
                              cudaStream_t st) {
    GG_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)n * sizeof(int32_t), st));
    GG_LAUNCH(ctx, GG_ST_RESOLVE, st,
              k_compact_winners<<<dim3((unsigned)(ctx->sm_count * 2), n), 256, 0, st>>>(ctx->vset[ctx->cur], n, ctx->F, flags,
                                                                                      d_pairs, cap, d_counts, ctx->d_sticky));
    return launch_reset_winners(ctx, n, flags, st);
}

int gg_launch_accumulate_rows(gg_context *ctx, const int32_t *d_pairs, int64_t n_rows, const void *d_rows, int pred_kind,
                              int C, int mode, int flags, double *d_sum, int32_t *d_count, cudaStream_t st) {
    if (n_rows == 0) return GG_OK;
    const unsigned g = (unsigned)((n_rows + 255) / 256);
    switch (pred_kind) {
        case GG_PRED_F32: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_accumulate_rows<float><<<g, 256, 0, st>>>(d_pairs, n_rows, (const float *)d_rows, C, pred_kind, mode, flags, d_sum, d_count)); break;
        case GG_PRED_F64: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_accumulate_rows<double><<<g, 256, 0, st>>>(d_pairs, n_rows, (const double *)d_rows, C, pred_kind, mode, flags, d_sum, d_count)); break;
        case GG_PRED_U8:
        case GG_PRED_INDEX_U8: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_accumulate_rows<uint8_t><<<g, 256, 0, st>>>(d_pairs, n_rows, (const uint8_t *)d_rows, C, pred_kind, mode, flags, d_sum, d_count)); break;
        default: gg_set_error("gg_accumulate_rows: bad pred_kind"); return GG_ERR_INVALID;
    }
    return GG_OK;
}

static int resolve_batch_body(gg_context *ctx, int n, const void *const *h_pred, int pred_kind, int C, int mode, int flags,
                              double *d_sum, int32_t *d_count, cudaStream_t st) {
    GGPredBatch pb;
    for (int i = 0; i < GG_MAX_VIEWS_PER_CALL; ++i) pb.p[i] = i < n ? h_pred[i] : nullptr;
    bool on_host = ctx->stage_host_rows != 0;
    for (int i = 0; i < n && on_host; ++i) {
        cudaPointerAttributes attr;
        on_host = cudaPointerGetAttributes(&attr, h_pred[i]) == cudaSuccess &&
                  (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);  // gg_host_alloc: host-resident
    }
    (void)cudaGetLastError();
    if (on_host) {
        const int E = (mode == GG_MODE_VOTE || pred_kind == GG_PRED_INDEX_U8) ? 1 : C;  // elements per pixel
        int rc = GG_OK;
        switch (pred_kind) {
            case GG_PRED_F32: rc = stage_host_rows<float>(ctx, n, pb, E, flags, st); break;
            case GG_PRED_F64: rc = stage_host_rows<double>(ctx, n, pb, E, flags, st); break;
            case GG_PRED_U8:
            case GG_PRED_INDEX_U8: rc = stage_host_rows<uint8_t>(ctx, n, pb, E, flags, st); break;
            default: gg_set_error("gg_project_aggregate: bad pred_kind"); return GG_ERR_INVALID;
        }
        if (rc != GG_OK) return rc;
    }
    const dim3 g((unsigned)(ctx->sm_count * 2), n);
    const int64_t F = ctx->F;
    switch (pred_kind) {
        case GG_PRED_F32: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_batch<float><<<g, 256, 0, st>>>(ctx->vset[ctx->cur], n, F, pb, C, pred_kind, mode, flags, d_sum, d_count)); break;
        case GG_PRED_F64: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_batch<double><<<g, 256, 0, st>>>(ctx->vset[ctx->cur], n, F, pb, C, pred_kind, mode, flags, d_sum, d_count)); break;
        case GG_PRED_U8:
        case GG_PRED_INDEX_U8: GG_LAUNCH(ctx, GG_ST_RESOLVE, st, k_resolve_batch<uint8_t><<<g, 256, 0, st>>>(ctx->vset[ctx->cur], n, F, pb, C, pred_kind, mode, flags, d_sum, d_count)); break;
        default: gg_set_error("gg_project_aggregate: bad pred_kind"); return GG_ERR_INVALID;
    }
    return GG_OK;
}

int gg_launch_resolve_batch(gg_context *ctx, int n, const void *const *h_pred, int pred_kind, int C, int mode, int flags,
                            double *d_sum, int32_t *d_count, cudaStream_t st) {
    const int rc = resolve_batch_body(ctx, n, h_pred, pred_kind, C, mode, flags, d_sum, d_count, st);
    const int rc_reset = launch_reset_winners(ctx, n, flags, st);  // also when the body failed: the arrays must end clean
    return rc != GG_OK ? rc : rc_reset;
}

int gg_launch_finalize(gg_context *ctx, double *d_sum, const int32_t *d_count, int64_t F, int C, double *d_avg,
                       double *d_argmax, cudaStream_t st) {
    GG_LAUNCH(ctx, GG_ST_FINALIZE, st, k_finalize<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(d_sum, d_count, F, C, d_avg, d_argmax));
    return GG_OK;
}

int gg_launch_render_flat(gg_context *ctx, const int32_t *d_pix2face, int64_t P, const double *d_tex, int D,
                          void *d_out, int out_dtype, cudaStream_t st) {
    const int64_t want = (P + 255) / 256;
    const unsigned g = (unsigned)(want < (int64_t)ctx->sm_count * 32 ? (want > 0 ? want : 1) : ctx->sm_count * 32);
    switch (out_dtype) {
        case GG_OUT_F64: GG_LAUNCH(ctx, GG_ST_RENDER_FLAT, st, k_render_flat<double><<<g, 256, 0, st>>>(d_pix2face, P, d_tex, D, (double *)d_out)); break;
        case GG_OUT_F32: GG_LAUNCH(ctx, GG_ST_RENDER_FLAT, st, k_render_flat<float><<<g, 256, 0, st>>>(d_pix2face, P, d_tex, D, (float *)d_out)); break;
        case GG_OUT_U8: GG_LAUNCH(ctx, GG_ST_RENDER_FLAT, st, k_render_flat<uint8_t><<<g, 256, 0, st>>>(d_pix2face, P, d_tex, D, (uint8_t *)d_out)); break;
        default: gg_set_error("gg_render_flat: bad out_dtype"); return GG_ERR_INVALID;
    }
    return GG_OK;
}
