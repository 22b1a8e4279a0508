// Stage 1 (camera projection) and stage 2 (tiled z-buffer rasterizer) of the multiview projection path.
//
// Replaces the VTK/OpenGL render of TexturedPhotogrammetryMesh.pix2face
// (/root/reference/geograypher/meshes/meshes.py:1776-1836) and the PyTorch3D MeshRasterizer call of
// TexturedPhotogrammetryMeshPyTorch3dRendering.pix2face (derived_meshes.py:691-737).
//
// Per batch of views (all kernels are batched over the views with blockIdx.y / blockIdx.z):
//   k_cull_blocks   frustum-cull 128-face blocks by their bounding boxes          -> visible block list
//   k_setup_faces   gather + project the 3 vertices of every face of a visible block (contract C1/C2),
//                   reject faces that cover no pixel centre, emit a 48-byte record, count tiles
//   k_scan_tiles    exclusive scan of the per-tile counts
//   k_fill_bins     write record indices into per-tile lists
//   k_raster_tiles  one CTA per 64x32-pixel tile: stage the tile's face records in shared memory, every
//                   thread owns 8 consecutive pixels of one row and keeps (depth, face) in registers;
//                   exact integer edge functions with top-left rule (C3), nearest 1/z wins, ties -> lowest
//                   face ID (C4); writes int32 face IDs coalesced.
// The result does not depend on the order in which faces land in a tile list, so the atomics used for
// compaction do not make it non-deterministic.
#include "gg_internal.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------
// Contract C1 + C2: float32 projection with a fixed operation order and no FMA contraction, then snap to
// 1/256 px.  __fmul_rn / __fadd_rn / __fdiv_rn are never fused or reordered by nvcc.
// ------------------------------------------------------------------------------------------------------
struct Proj {
    int X, Y;
    float invz;
    bool ok;
};

__device__ __forceinline__ Proj project_vertex(float x, float y, float z, const gg_camera &c) {
    Proj p;
    float t;
    t = __fmul_rn(c.m[0], x);
    t = __fadd_rn(t, __fmul_rn(c.m[1], y));
    t = __fadd_rn(t, __fmul_rn(c.m[2], z));
    const float xc = __fadd_rn(t, c.m[3]);
    t = __fmul_rn(c.m[4], x);
    t = __fadd_rn(t, __fmul_rn(c.m[5], y));
    t = __fadd_rn(t, __fmul_rn(c.m[6], z));
    const float yc = __fadd_rn(t, c.m[7]);
    t = __fmul_rn(c.m[8], x);
    t = __fadd_rn(t, __fmul_rn(c.m[9], y));
    t = __fadd_rn(t, __fmul_rn(c.m[10], z));
    const float zc = __fadd_rn(t, c.m[11]);
    p.ok = (zc >= c.znear) && isfinite(xc) && isfinite(yc) && isfinite(zc);
    p.X = 0;
    p.Y = 0;
    p.invz = 0.f;
    if (p.ok) {
        const float sx = __fadd_rn(__fdiv_rn(__fmul_rn(c.f, xc), zc), c.px);
        const float sy = __fadd_rn(__fdiv_rn(__fmul_rn(c.f, yc), zc), c.py);
        float fx = rintf(__fmul_rn(sx, (float)GG_SUBPIX));
        float fy = rintf(__fmul_rn(sy, (float)GG_SUBPIX));
        if (!isfinite(fx) || !isfinite(fy)) {
            p.ok = false;
        } else {
            fx = fminf(fmaxf(fx, -GG_COORD_CLAMP), GG_COORD_CLAMP);
            fy = fminf(fmaxf(fy, -GG_COORD_CLAMP), GG_COORD_CLAMP);
            p.X = __float2int_rn(fx);
            p.Y = __float2int_rn(fy);
            p.invz = __fdiv_rn(1.0f, zc);
        }
    }
    return p;
}

__device__ __forceinline__ int warp_append(int *counter, bool keep) {
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m == 0) return -1;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return keep ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// ------------------------------------------------------------------------------------------------------
// Mesh preprocessing: bounding box of every block of GG_BLOCK_FACES consecutive faces.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GG_BLOCK_FACES) k_mesh_blocks(const float4 *__restrict__ verts,
                                                               const int4 *__restrict__ faces, int64_t F,
                                                               float *__restrict__ lo, float *__restrict__ hi) {
    const int64_t b = blockIdx.x;
    const int64_t fi = b * GG_BLOCK_FACES + threadIdx.x;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (fi < F) {
        const int4 f = faces[fi];
        const int idx[3] = {f.x, f.y, f.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 v = verts[idx[k]];
            mn[0] = fminf(mn[0], v.x);
            mx[0] = fmaxf(mx[0], v.x);
            mn[1] = fminf(mn[1], v.y);
            mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z);
            mx[2] = fmaxf(mx[2], v.z);
        }
    }
    __shared__ float s_mn[GG_BLOCK_FACES / 32][3], s_mx[GG_BLOCK_FACES / 32][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; ++k) {
            s_mn[w][k] = mn[k];
            s_mx[w][k] = mx[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = s_mn[0][threadIdx.x], c = s_mx[0][threadIdx.x];
        for (int i = 1; i < GG_BLOCK_FACES / 32; ++i) {
            a = fminf(a, s_mn[i][threadIdx.x]);
            c = fmaxf(c, s_mx[i][threadIdx.x]);
        }
        lo[b * 3 + threadIdx.x] = a;
        hi[b * 3 + threadIdx.x] = c;
    }
}

// ------------------------------------------------------------------------------------------------------
// Stage 1 on its own: project all vertices for n views (gg_project).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_project(const float4 *__restrict__ verts, int64_t V,
                                                 const __grid_constant__ GGCamBatch cams, int32_t *__restrict__ X,
                                                 int32_t *__restrict__ Y, float *__restrict__ invz,
                                                 uint8_t *__restrict__ valid) {
    const int view = blockIdx.y;
    const gg_camera &c = cams.cam[view];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = verts[i];
        const Proj p = project_vertex(v.x, v.y, v.z, c);
        const int64_t o = (int64_t)view * V + i;
        X[o] = p.X;
        Y[o] = p.Y;
        invz[o] = p.invz;
        valid[o] = p.ok ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------------
// Per-view frustum culling of face blocks (conservative; never changes the result).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cull_blocks(const float *__restrict__ lo, const float *__restrict__ hi,
                                                     int n_blocks, const __grid_constant__ GGCamBatch cams,
                                                     const __grid_constant__ GGViewBatch views) {
    const int view = blockIdx.y;
    const gg_camera &c = cams.cam[view];
    const GGViewScratch &vs = views.v[view];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (b < n_blocks) {
        const float l[3] = {lo[b * 3], lo[b * 3 + 1], lo[b * 3 + 2]};
        const float h[3] = {hi[b * 3], hi[b * 3 + 1], hi[b * 3 + 2]};
        float zmin = INFINITY, zmax = -INFINITY;
        float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        bool finite = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float x = (k & 1) ? h[0] : l[0], y = (k & 2) ? h[1] : l[1], z = (k & 4) ? h[2] : l[2];
            const float xc = c.m[0] * x + c.m[1] * y + c.m[2] * z + c.m[3];
            const float yc = c.m[4] * x + c.m[5] * y + c.m[6] * z + c.m[7];
            const float zc = c.m[8] * x + c.m[9] * y + c.m[10] * z + c.m[11];
            finite = finite && isfinite(xc) && isfinite(yc) && isfinite(zc);
            zmin = fminf(zmin, zc);
            zmax = fmaxf(zmax, zc);
            if (zc > 0.f) {
                const float sx = c.f * xc / zc + c.px, sy = c.f * yc / zc + c.py;
                xmin = fminf(xmin, sx);
                xmax = fmaxf(xmax, sx);
                ymin = fminf(ymin, sy);
                ymax = fmaxf(ymax, sy);
            }
        }
        const float zslack = 1e-4f * fabsf(zmax) + 1e-6f;
        if (!finite) {
            keep = true;  // let the per-face test decide
        } else if (zmax + zslack < c.znear) {
            keep = false;  // entirely behind the near plane: every face has an invalid vertex (C5)
        } else if (zmin - zslack < c.znear || zmin <= 0.f) {
            keep = true;  // straddles the near plane: the screen box is unbounded
        } else {
            const float s = 2.0f + 1e-4f * (fabsf(xmin) + fabsf(xmax) + fabsf(ymin) + fabsf(ymax));
            keep = (xmax + s >= 0.f) && (xmin - s <= (float)c.W) && (ymax + s >= 0.f) && (ymin - s <= (float)c.H);
        }
    }
    const int pos = warp_append(&vs.counters[0], keep);
    if (pos >= 0) vs.vis_blocks[pos] = b;
}

// ------------------------------------------------------------------------------------------------------
// Face setup: one thread per face of a visible block.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GG_BLOCK_FACES) k_setup_faces(const float4 *__restrict__ verts,
                                                               const int4 *__restrict__ faces, int64_t F,
                                                               int64_t cap_recs,
                                                               const __grid_constant__ GGCamBatch cams,
                                                               const __grid_constant__ GGViewBatch views) {
    const int view = blockIdx.y;
    const gg_camera &c = cams.cam[view];
    const GGViewScratch &vs = views.v[view];
    const int n_vis = vs.counters[0];
    const int tiles_x = (c.W + GG_TILE_W - 1) / GG_TILE_W;
    for (int vb = blockIdx.x; vb < n_vis; vb += gridDim.x) {
        const int64_t fi = (int64_t)vs.vis_blocks[vb] * GG_BLOCK_FACES + threadIdx.x;
        bool keep = false;
        GGFaceRec r;
        if (fi < F) {
            const int4 f = faces[fi];
            const float4 a = verts[f.x], b = verts[f.y], d = verts[f.z];
            const Proj p0 = project_vertex(a.x, a.y, a.z, c);
            Proj p1 = project_vertex(b.x, b.y, b.z, c);
            Proj p2 = project_vertex(d.x, d.y, d.z, c);
            if (p0.ok && p1.ok && p2.ok) {
                long long area2 = (long long)(p1.X - p0.X) * (long long)(p2.Y - p0.Y) -
                                  (long long)(p2.X - p0.X) * (long long)(p1.Y - p0.Y);
                if (area2 != 0) {
                    if (area2 < 0) {
                        const Proj t = p1;
                        p1 = p2;
                        p2 = t;
                    }
                    const int xmin = min(p0.X, min(p1.X, p2.X)), xmax = max(p0.X, max(p1.X, p2.X));
                    const int ymin = min(p0.Y, min(p1.Y, p2.Y)), ymax = max(p0.Y, max(p1.Y, p2.Y));
                    // pixel centres 256*j+128 inside [xmin, xmax]; arithmetic shift == floor division
                    int jmin = (xmin + (GG_SUBPIX - GG_HALF - 1)) >> GG_SUBPIX_LOG2;
                    int jmax = (xmax - GG_HALF) >> GG_SUBPIX_LOG2;
                    int imin = (ymin + (GG_SUBPIX - GG_HALF - 1)) >> GG_SUBPIX_LOG2;
                    int imax = (ymax - GG_HALF) >> GG_SUBPIX_LOG2;
                    jmin = max(jmin, 0);
                    imin = max(imin, 0);
                    jmax = min(jmax, c.W - 1);
                    imax = min(imax, c.H - 1);
                    if (jmin <= jmax && imin <= imax) {
                        keep = true;
                        r.x0 = p0.X;
                        r.y0 = p0.Y;
                        r.x1 = p1.X;
                        r.y1 = p1.Y;
                        r.x2 = p2.X;
                        r.y2 = p2.Y;
                        r.w0 = p0.invz;
                        r.w1 = p1.invz;
                        r.w2 = p2.invz;
                        r.face = f.w;
                        r.jmin = (uint16_t)jmin;
                        r.jmax = (uint16_t)jmax;
                        r.imin = (uint16_t)imin;
                        r.imax = (uint16_t)imax;
                    }
                }
            }
        }
        const int idx = warp_append(&vs.counters[1], keep);
        if (keep) {
            if (idx < cap_recs) {
                vs.recs[idx] = r;
                if (r.face == (int32_t)(F - 1)) vs.counters[5] = idx;
                const int tx0 = r.jmin / GG_TILE_W, tx1 = r.jmax / GG_TILE_W;
                const int ty0 = r.imin / GG_TILE_H, ty1 = r.imax / GG_TILE_H;
                for (int ty = ty0; ty <= ty1; ++ty)
                    for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&vs.tile_count[ty * tiles_x + tx], 1);
            } else {
                atomicOr(&vs.counters[3], 1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Exclusive scan of tile counts (one CTA per view); leaves tile_count zeroed for use as the fill cursor.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_tiles(int n_tiles, int64_t cap_recs, int64_t cap_bins,
                                                     const __grid_constant__ GGViewBatch views) {
    const GGViewScratch &vs = views.v[blockIdx.x];
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) {
        s_carry = 0;
        if (vs.counters[1] > cap_recs) vs.counters[1] = (int)cap_recs;  // records beyond capacity were dropped
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < n_tiles) ? vs.tile_count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[w] = x;
        __syncthreads();
        if (w == 0) {
            int s = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            s_warp[lane] = s;
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (w > 0 ? s_warp[w - 1] : 0) + x - v;
        if (i < n_tiles) {
            vs.tile_offset[i] = excl;
            vs.tile_count[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        vs.tile_offset[n_tiles] = s_carry;
        vs.counters[2] = s_carry;
        if ((int64_t)s_carry > cap_bins) atomicOr(&vs.counters[3], 2);
    }
}

__global__ void __launch_bounds__(256) k_fill_bins(const __grid_constant__ GGCamBatch cams,
                                                   const __grid_constant__ GGViewBatch views) {
    const int view = blockIdx.y;
    const GGViewScratch &vs = views.v[view];
    if (vs.counters[3] != 0) return;
    const int n_recs = vs.counters[1];
    const int tiles_x = (cams.cam[view].W + GG_TILE_W - 1) / GG_TILE_W;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_recs; r += gridDim.x * blockDim.x) {
        const GGFaceRec &rec = vs.recs[r];
        const int tx0 = rec.jmin / GG_TILE_W, tx1 = rec.jmax / GG_TILE_W;
        const int ty0 = rec.imin / GG_TILE_H, ty1 = rec.imax / GG_TILE_H;
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                const int t = ty * tiles_x + tx;
                const int pos = vs.tile_offset[t] + atomicAdd(&vs.tile_count[t], 1);
                vs.bins[pos] = r;
            }
    }
}

// ------------------------------------------------------------------------------------------------------
// Tile rasterizer.
// ------------------------------------------------------------------------------------------------------
struct TileFace {
    long long e[3];   // edge functions at the tile-origin pixel centre, top-left bias folded in (>= 0 inside)
    long long sx[3];  // step per pixel in x
    long long sy[3];  // step per pixel in y
    double inv_area;  // 1 / area2 (exact-depth path)
    float w_org, gx, gy;  // 1/z plane relative to the tile origin (fast path)
    float w0, w1, w2;
    int face;
    int flags;  // bit0: 32-bit edges + plane depth are safe; bits 1..3: bias of edge k
    short bx0, bx1, by0, by1;  // pixel range inside the tile (inclusive)
};

#define TF_FAST 1

__device__ __forceinline__ void setup_tile_face(TileFace &tf, const GGFaceRec &r, int tile_x0, int tile_y0) {
    const long long Px = (long long)GG_SUBPIX * tile_x0 + GG_HALF, Py = (long long)GG_SUBPIX * tile_y0 + GG_HALF;
    const long long X[3] = {r.x0, r.x1, r.x2}, Y[3] = {r.y0, r.y1, r.y2};
    long long A[3], B[3], E[3];
    int flags = 0;
    long long ebound = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int k1 = (k + 1) % 3;
        A[k] = -(Y[k1] - Y[k]);
        B[k] = (X[k1] - X[k]);
        E[k] = B[k] * (Py - Y[k]) + A[k] * (Px - X[k]);
        const bool inclusive = (A[k] > 0) || (A[k] == 0 && B[k] > 0);  // contract C3 (top-left rule)
        const long long bias = inclusive ? 0 : 1;
        flags |= (int)bias << (1 + k);
        tf.e[k] = E[k] - bias;
        tf.sx[k] = A[k] * GG_SUBPIX;
        tf.sy[k] = B[k] * GG_SUBPIX;
        const long long bound = llabs(tf.e[k]) + (GG_TILE_W - 1) * llabs(tf.sx[k]) + (GG_TILE_H - 1) * llabs(tf.sy[k]);
        ebound = max(ebound, bound);
    }
    const long long area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
    const double inv_area = 1.0 / (double)area2;
    const double w0 = r.w0, w1 = r.w1, w2 = r.w2;
    // barycentric weights: E1 -> v0, E2 -> v1, E0 -> v2
    const double w_org = ((double)E[1] * w0 + (double)E[2] * w1 + (double)E[0] * w2) * inv_area;
    const double gx = (double)GG_SUBPIX * ((double)A[1] * w0 + (double)A[2] * w1 + (double)A[0] * w2) * inv_area;
    const double gy = (double)GG_SUBPIX * ((double)B[1] * w0 + (double)B[2] * w1 + (double)B[0] * w2) * inv_area;
    tf.inv_area = inv_area;
    tf.w_org = (float)w_org;
    tf.gx = (float)gx;
    tf.gy = (float)gy;
    tf.w0 = r.w0;
    tf.w1 = r.w1;
    tf.w2 = r.w2;
    tf.face = r.face;
    const double wmin = fmin(w0, fmin(w1, w2));
    const double spread = fabs(w_org) + (GG_TILE_W - 1) * fabs(gx) + (GG_TILE_H - 1) * fabs(gy);
    if (ebound < 2147483647LL && spread <= 16.0 * wmin) flags |= TF_FAST;
    tf.flags = flags;
    tf.bx0 = (short)max((int)r.jmin - tile_x0, 0);
    tf.bx1 = (short)min((int)r.jmax - tile_x0, GG_TILE_W - 1);
    tf.by0 = (short)max((int)r.imin - tile_y0, 0);
    tf.by1 = (short)min((int)r.imax - tile_y0, GG_TILE_H - 1);
}

__global__ void __launch_bounds__(GG_RASTER_THREADS) k_raster_tiles(const __grid_constant__ GGCamBatch cams,
                                                                    const __grid_constant__ GGViewBatch views,
                                                                    int32_t *__restrict__ pix2face,
                                                                    float *__restrict__ depth) {
    const int view = blockIdx.z;
    const gg_camera &c = cams.cam[view];
    const GGViewScratch &vs = views.v[view];
    const int W = c.W, H = c.H;
    const int tiles_x = (W + GG_TILE_W - 1) / GG_TILE_W;
    const int tile = blockIdx.y * tiles_x + blockIdx.x;
    const int tile_x0 = blockIdx.x * GG_TILE_W, tile_y0 = blockIdx.y * GG_TILE_H;

    __shared__ TileFace s_faces[GG_CHUNK];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = (warp & 1) * 32, wy0 = (warp >> 1) * 8;  // warp region: 32 x 8 px
    const int tx0 = wx0 + (lane & 3) * 8;                     // this thread: 8 px of row ty
    const int ty = wy0 + (lane >> 2);

    float bw[8];
    int bf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bw[i] = 0.f;
        bf[i] = -1;
    }

    const bool overflow = vs.counters[3] != 0;
    const int beg = overflow ? 0 : vs.tile_offset[tile];
    const int end = overflow ? 0 : vs.tile_offset[tile + 1];

    for (int base = beg; base < end; base += GG_CHUNK) {
        const int n = min(GG_CHUNK, end - base);
        if (threadIdx.x < n) setup_tile_face(s_faces[threadIdx.x], vs.recs[vs.bins[base + threadIdx.x]], tile_x0, tile_y0);
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            const TileFace &tf = s_faces[k];
            // warp-uniform reject, then per-thread reject
            if (tf.bx1 < wx0 || tf.bx0 > wx0 + 31 || tf.by1 < wy0 || tf.by0 > wy0 + 7) continue;
            if (ty < tf.by0 || ty > tf.by1 || tf.bx1 < tx0 || tf.bx0 > tx0 + 7) continue;
            const int face = tf.face;
            if (tf.flags & TF_FAST) {
                const int s0 = (int)tf.sx[0], s1 = (int)tf.sx[1], s2 = (int)tf.sx[2];
                int e0 = (int)tf.e[0] + s0 * tx0 + (int)tf.sy[0] * ty;
                int e1 = (int)tf.e[1] + s1 * tx0 + (int)tf.sy[1] * ty;
                int e2 = (int)tf.e[2] + s2 * tx0 + (int)tf.sy[2] * ty;
                const float wrow = fmaf(tf.gy, (float)ty, tf.w_org);
                const float gx = tf.gx;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if ((e0 | e1 | e2) >= 0) {
                        const float w = fmaf(gx, (float)(tx0 + i), wrow);
                        if (w > bw[i] || (w == bw[i] && face < bf[i])) {
                            bw[i] = w;
                            bf[i] = face;
                        }
                    }
                    e0 += s0;
                    e1 += s1;
                    e2 += s2;
                }
            } else {
                long long e0 = tf.e[0] + tf.sx[0] * tx0 + tf.sy[0] * ty;
                long long e1 = tf.e[1] + tf.sx[1] * tx0 + tf.sy[1] * ty;
                long long e2 = tf.e[2] + tf.sx[2] * tx0 + tf.sy[2] * ty;
                const long long b0 = (tf.flags >> 1) & 1, b1 = (tf.flags >> 2) & 1, b2 = (tf.flags >> 3) & 1;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if ((e0 | e1 | e2) >= 0) {
                        const double wd = ((double)(e1 + b1) * (double)tf.w0 + (double)(e2 + b2) * (double)tf.w1 +
                                           (double)(e0 + b0) * (double)tf.w2) * tf.inv_area;
                        const float w = (float)wd;
                        if (w > bw[i] || (w == bw[i] && face < bf[i])) {
                            bw[i] = w;
                            bf[i] = face;
                        }
                    }
                    e0 += tf.sx[0];
                    e1 += tf.sx[1];
                    e2 += tf.sx[2];
                }
            }
        }
        __syncthreads();
    }

    // ---- write the 8 pixels of this thread ----
    const int row = tile_y0 + ty, col = tile_x0 + tx0;
    if (row < H && col < W) {
        const int64_t o = ((int64_t)view * H + row) * W + col;
        if (col + 7 < W && ((o & 3) == 0)) {
            if (pix2face) {
                int4 *p = reinterpret_cast<int4 *>(pix2face + o);
                p[0] = make_int4(bf[0], bf[1], bf[2], bf[3]);
                p[1] = make_int4(bf[4], bf[5], bf[6], bf[7]);
            }
            if (depth) {
                float4 *p = reinterpret_cast<float4 *>(depth + o);
                p[0] = make_float4(bw[0], bw[1], bw[2], bw[3]);
                p[1] = make_float4(bw[4], bw[5], bw[6], bw[7]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (col + i < W) {
                    if (pix2face) pix2face[o + i] = bf[i];
                    if (depth) depth[o + i] = bw[i];
                }
            }
        }
    }
}

}  // namespace

// ======================================================================================================
// Host side
// ======================================================================================================
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int gg_ensure_scratch(gg_context *ctx, int n_views, int W, int H) {
    const int64_t tiles = (int64_t)((W + GG_TILE_W - 1) / GG_TILE_W) * ((H + GG_TILE_H - 1) / GG_TILE_H);
    const int64_t cap_recs = ctx->req_recs > 0 ? ctx->req_recs : ctx->F;
    const int64_t cap_bins = ctx->req_bins > 0 ? ctx->req_bins : (4 * cap_recs > (1 << 20) ? 4 * cap_recs : (1 << 20));
    if (n_views <= ctx->n_slots && tiles <= ctx->slot_tiles && cap_recs == ctx->cap_recs && cap_bins == ctx->cap_bins)
        return GG_OK;
    if (ctx->d_scratch) {
        GG_CUDA(cudaDeviceSynchronize());
        GG_CUDA(cudaFree(ctx->d_scratch));
        ctx->d_scratch = nullptr;
    }
    const int slots = n_views > ctx->n_slots ? n_views : ctx->n_slots;
    const int64_t slot_tiles = tiles > ctx->slot_tiles ? tiles : ctx->slot_tiles;
    const size_t b_vis = align_up((size_t)ctx->n_blocks * 4, 256);
    const size_t b_rec = align_up((size_t)cap_recs * sizeof(GGFaceRec), 256);
    const size_t b_cnt = align_up((size_t)slot_tiles * 4, 256);
    const size_t b_off = align_up((size_t)(slot_tiles + 1) * 4, 256);
    const size_t b_bin = align_up((size_t)cap_bins * 4, 256);
    const size_t b_ctr = 256;
    const size_t per_slot = b_vis + b_rec + b_cnt + b_off + b_bin + b_ctr;
    GG_CUDA(cudaMalloc(&ctx->d_scratch, per_slot * slots));
    ctx->scratch_bytes = per_slot * slots;
    for (int s = 0; s < slots; ++s) {
        char *p = ctx->d_scratch + per_slot * s;
        GGViewScratch &v = ctx->views.v[s];
        v.vis_blocks = (int32_t *)p;
        p += b_vis;
        v.recs = (GGFaceRec *)p;
        p += b_rec;
        v.tile_count = (int32_t *)p;
        p += b_cnt;
        v.tile_offset = (int32_t *)p;
        p += b_off;
        v.bins = (int32_t *)p;
        p += b_bin;
        v.counters = (int32_t *)p;
    }
    ctx->n_slots = slots;
    ctx->slot_tiles = slot_tiles;
    ctx->cap_recs = cap_recs;
    ctx->cap_bins = cap_bins;
    return GG_OK;
}

int gg_launch_mesh_blocks(gg_context *ctx, cudaStream_t st) {
    GG_LAUNCH(ctx, GG_ST_MESH, st,
              k_mesh_blocks<<<(unsigned)ctx->n_blocks, GG_BLOCK_FACES, 0, st>>>(ctx->d_verts, ctx->d_faces, ctx->F,
                                                                                 ctx->d_block_lo, ctx->d_block_hi));
    return GG_OK;
}

int gg_launch_project(gg_context *ctx, const gg_camera *cams, int n, int32_t *dX, int32_t *dY, float *dinvz,
                      uint8_t *dvalid, cudaStream_t st) {
    GGCamBatch cb;
    for (int i = 0; i < n; ++i) cb.cam[i] = cams[i];
    const int64_t want = (ctx->V + 255) / 256;
    const unsigned gx = (unsigned)(want < (int64_t)ctx->sm_count * 16 ? (want > 0 ? want : 1) : ctx->sm_count * 16);
    GG_LAUNCH(ctx, GG_ST_PROJECT, st, k_project<<<dim3(gx, n), 256, 0, st>>>(ctx->d_verts, ctx->V, cb, dX, dY, dinvz, dvalid));
    return GG_OK;
}

int gg_launch_rasterize(gg_context *ctx, const gg_camera *cams, int n, int32_t *d_pix2face, float *d_depth,
                        cudaStream_t st) {
    const int W = cams[0].W, H = cams[0].H;
    int rc = gg_ensure_scratch(ctx, n, W, H);
    if (rc != GG_OK) return rc;
    GGCamBatch cb;
    for (int i = 0; i < n; ++i) cb.cam[i] = cams[i];
    const int tiles_x = (W + GG_TILE_W - 1) / GG_TILE_W, tiles_y = (H + GG_TILE_H - 1) / GG_TILE_H;
    const int n_tiles = tiles_x * tiles_y;
    for (int i = 0; i < n; ++i) {
        GG_CUDA(cudaMemsetAsync(ctx->views.v[i].tile_count, 0, (size_t)n_tiles * 4, st));
        GG_CUDA(cudaMemsetAsync(ctx->views.v[i].counters, 0, 16, st));
        GG_CUDA(cudaMemsetAsync(ctx->views.v[i].counters + 4, 0xFF, 8, st));
    }
    ctx->last_batch_n = n;
    const int nb = (int)ctx->n_blocks;
    GG_LAUNCH(ctx, GG_ST_CULL, st,
              k_cull_blocks<<<dim3((nb + 255) / 256, n), 256, 0, st>>>(ctx->d_block_lo, ctx->d_block_hi, nb, cb, ctx->views));
    const int gsetup = nb < ctx->sm_count * 8 ? nb : ctx->sm_count * 8;
    GG_LAUNCH(ctx, GG_ST_SETUP, st,
              k_setup_faces<<<dim3(gsetup, n), GG_BLOCK_FACES, 0, st>>>(ctx->d_verts, ctx->d_faces, ctx->F, ctx->cap_recs,
                                                                        cb, ctx->views));
    GG_LAUNCH(ctx, GG_ST_SCAN, st, k_scan_tiles<<<n, 1024, 0, st>>>(n_tiles, ctx->cap_recs, ctx->cap_bins, ctx->views));
    GG_LAUNCH(ctx, GG_ST_FILL, st, k_fill_bins<<<dim3(ctx->sm_count * 4, n), 256, 0, st>>>(cb, ctx->views));
    GG_LAUNCH(ctx, GG_ST_RASTER, st,
              k_raster_tiles<<<dim3(tiles_x, tiles_y, n), GG_RASTER_THREADS, 0, st>>>(cb, ctx->views, d_pix2face, d_depth));
    return GG_OK;
}
