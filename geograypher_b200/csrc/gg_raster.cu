// Stage 1 (camera projection) and stage 2 (tiled z-buffer rasterizer) of the multiview projection path.
//
// Replaces the VTK/OpenGL render of TexturedPhotogrammetryMesh.pix2face
// (/root/reference/geograypher/meshes/meshes.py:1776-1836) and the PyTorch3D MeshRasterizer call of
// TexturedPhotogrammetryMeshPyTorch3dRendering.pix2face (derived_meshes.py:691-737).
//
// Per batch of up to 32 views (every kernel is batched over the views with blockIdx.y / blockIdx.z; the cameras travel
// as a __grid_constant__ kernel parameter):
//   k_reset_views   zero the per-tile counts and the counters of the batch's scratch slots
//   k_cull_blocks   frustum-cull blocks of 128 consecutive (Z-ordered) faces by their bounding boxes -> visible blocks
//   k_setup_faces   gather + project the 3 vertices of every face of a visible block (contract C1/C2), clip against the
//                   near plane (C5: up to two triangles per face), reject faces that cover no pixel centre, emit a
//                   128-byte record (edge equations, float64 1/z plane, tile mask), count the tiles it touches
//   k_reserve_tiles warp-scan of the tile counts + one atomicAdd per warp reserves each tile's list (no global scan)
//   k_fill_bins     one 64-byte tile-relative setup per (tile, face) pair, 8 lanes per record
//   k_raster_tiles  a warp rasterizes 32x8-pixel tiles (four consecutive ones, the next tile's setups prefetched into
//                   shared memory while the current one is worked on): every lane owns 8 consecutive pixels of one
//                   row and keeps (1/z, face) -- the dense mode also the list position -- in registers; exact integer
//                   edge functions with top-left rule (C3), nearest 1/z wins, ties -> lowest face ID (C4).  Epilogues
//                   by mode: face-ID / depth rasters, last pixel per face (fused last-pixel / vote aggregation),
//                   dense per-pixel score sums, fused render_flat gather.
// The result does not depend on the order in which faces land in a tile list, so the atomics used for
// compaction do not make it non-deterministic.
#include <cstring>

#include <algorithm>
#include "gg_internal.cuh"

#ifndef GG_SETUP_MIN_BLOCKS
#define GG_SETUP_MIN_BLOCKS 8  // 64 registers: the binning kernels are latency bound, occupancy matters more than spills
#endif
#ifndef GG_FILL_MIN_BLOCKS
#define GG_FILL_MIN_BLOCKS 4
#endif
#ifndef GG_FILL_LANES
#define GG_FILL_LANES 8  // lanes that share one face record in k_fill_bins and split its tiles (power of two)
#endif

namespace {

// ------------------------------------------------------------------------------------------------------
// Contract C1 + C2: float32 projection with a fixed operation order and no FMA contraction, then snap to
// 1/256 px.  __fmul_rn / __fadd_rn / __fdiv_rn are never fused or reordered by nvcc.
// ------------------------------------------------------------------------------------------------------
struct Proj {
    int X, Y;
    float invz;
    bool ok;
    bool far;  // the snapped coordinates would not even be finite (contract C6 clips such a triangle first)
};

struct Cam3 {  // camera-space point
    float x, y, z;
};

// C1, first half: camera-space coordinates
__device__ __forceinline__ Cam3 cam_space(float x, float y, float z, const gg_camera &c) {
    Cam3 p;
    float t;
    t = __fmul_rn(c.m[0], x);
    t = __fadd_rn(t, __fmul_rn(c.m[1], y));
    t = __fadd_rn(t, __fmul_rn(c.m[2], z));
    p.x = __fadd_rn(t, c.m[3]);
    t = __fmul_rn(c.m[4], x);
    t = __fadd_rn(t, __fmul_rn(c.m[5], y));
    t = __fadd_rn(t, __fmul_rn(c.m[6], z));
    p.y = __fadd_rn(t, c.m[7]);
    t = __fmul_rn(c.m[8], x);
    t = __fadd_rn(t, __fmul_rn(c.m[9], y));
    t = __fadd_rn(t, __fmul_rn(c.m[10], z));
    p.z = __fadd_rn(t, c.m[11]);
    return p;
}

// C1, second half + C2: perspective division and snapping of a camera-space point in front of the near plane
__device__ __forceinline__ Proj to_screen(const Cam3 &q, const gg_camera &c) {
    Proj p;
    p.ok = (q.z >= c.znear) && isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
    p.X = 0;
    p.Y = 0;
    p.invz = 0.f;
    p.far = false;
    if (p.ok) {
        const float sx = __fadd_rn(__fdiv_rn(__fmul_rn(c.f, q.x), q.z), c.px);
        const float sy = __fadd_rn(__fdiv_rn(__fmul_rn(c.f, q.y), q.z), c.py);
        float fx = rintf(__fmul_rn(sx, (float)GG_SUBPIX));
        float fy = rintf(__fmul_rn(sy, (float)GG_SUBPIX));
        if (!isfinite(fx) || !isfinite(fy)) {
            p.ok = false;
            p.far = true;
        } else {
            fx = fminf(fmaxf(fx, -GG_COORD_CLAMP), GG_COORD_CLAMP);
            fy = fminf(fmaxf(fy, -GG_COORD_CLAMP), GG_COORD_CLAMP);
            p.X = __float2int_rn(fx);
            p.Y = __float2int_rn(fy);
            p.invz = __fdiv_rn(1.0f, q.z);
        }
    }
    return p;
}

__device__ __forceinline__ Proj project_vertex(float x, float y, float z, const gg_camera &c) {
    return to_screen(cam_space(x, y, z, c), c);
}

// C5: intersection of the edge from P (in front of the near plane) to Q (behind it) with the plane z = znear
__device__ __forceinline__ Cam3 clip_edge(const Cam3 &P, const Cam3 &Q, float znear) {
    const float t = __fdiv_rn(__fsub_rn(P.z, znear), __fsub_rn(P.z, Q.z));
    Cam3 r;
    r.x = __fadd_rn(P.x, __fmul_rn(t, __fsub_rn(Q.x, P.x)));
    r.y = __fadd_rn(P.y, __fmul_rn(t, __fsub_rn(Q.y, P.y)));
    r.z = znear;
    return r;
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

// Whole-struct stores / loads as 16-byte vectors (full 32-byte sectors, no partial-sector read-modify-write in L2).
template <typename S>
__device__ __forceinline__ void store_vec16(S *dst, const S &src) {
    static_assert(sizeof(S) % 16 == 0, "16-byte multiple");
    const int4 *s4 = reinterpret_cast<const int4 *>(&src);
    int4 *d4 = reinterpret_cast<int4 *>(dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(S) / 16); ++i) d4[i] = s4[i];
}

template <typename S>
__device__ __forceinline__ S load_vec16(const S *src) {
    S out;
    const int4 *s4 = reinterpret_cast<const int4 *>(src);
    int4 *d4 = reinterpret_cast<int4 *>(&out);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(S) / 16); ++i) d4[i] = __ldg(s4 + i);
    return out;
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

// One launch resets the per-view counters and tile counts of a whole batch.
__global__ void __launch_bounds__(256) k_reset_views(int n_tiles, const __grid_constant__ GGViewBatch views) {
    const GGViewScratch &vs = views.v[blockIdx.y];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) vs.tile_count[t] = 0;
    if (blockIdx.x == 0 && threadIdx.x < 8) vs.counters[threadIdx.x] = (threadIdx.x == 4 || threadIdx.x == 5) ? -1 : 0;
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

// Conservative triangle / tile test: the tile is skipped when all its pixel centres lie on the outer side of one
// edge (the edge function is linear, so its maximum over the tile is at the corner chosen by the gradient signs).
__device__ __forceinline__ bool tile_may_touch(const GGFaceRec &r, int tx, int ty) {
    const int x0 = tx * GG_TILE_W, y0 = ty * GG_TILE_H;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int xc = r.A[k] > 0 ? x0 + GG_TILE_W - 1 : x0;
        const int yc = r.B[k] > 0 ? y0 + GG_TILE_H - 1 : y0;
        const long long e = r.C[k] + (long long)r.A[k] * GG_SUBPIX * xc + (long long)r.B[k] * GG_SUBPIX * yc;
        if (e < 0) return false;
    }
    return true;
}

// Screen-space record of one (sub-)triangle given in camera space.  Returns 1 and the record, 0 when the triangle
// cannot cover a pixel centre, 2 (only with check_guard) when a vertex projects beyond the guard band of contract C6:
// the caller then clips the triangle against the band first.
__device__ __forceinline__ int build_record(const Cam3 &qa, const Cam3 &qb, const Cam3 &qc, int face_id,
                                            const gg_camera &c, GGFaceRec &r, bool check_guard) {
    const Proj p0 = to_screen(qa, c);
    Proj p1 = to_screen(qb, c);
    Proj p2 = to_screen(qc, c);
    if (!(p0.ok && p1.ok && p2.ok)) return (check_guard && (p0.far || p1.far || p2.far)) ? 2 : 0;
    const int xmin = min(p0.X, min(p1.X, p2.X)), xmax = max(p0.X, max(p1.X, p2.X));
    const int ymin = min(p0.Y, min(p1.Y, p2.Y)), ymax = max(p0.Y, max(p1.Y, p2.Y));
    // contract C6: a vertex beyond the guard band of +-2^20 px.  Tested on the snapped integers: beyond 2^16 px the
    // float32 coordinate times 256 is an integer already, so |X| > 2^28 <=> |sx| > 2^20 exactly (the oracle's test).
    constexpr int kGuard = (int)GG_GUARD_PX * GG_SUBPIX;
    if (check_guard && (xmin < -kGuard || xmax > kGuard || ymin < -kGuard || ymax > kGuard)) return 2;
    const long long area2 = (long long)(p1.X - p0.X) * (long long)(p2.Y - p0.Y) -
                            (long long)(p2.X - p0.X) * (long long)(p1.Y - p0.Y);
    if (area2 == 0) return 0;
    if (area2 < 0) {  // make the interior the positive side
        const Proj t = p1;
        p1 = p2;
        p2 = t;
    }
    // pixel centres 256*j+128 inside [xmin, xmax]; arithmetic shift == floor division
    int jmin = (xmin + (GG_SUBPIX - GG_HALF - 1)) >> GG_SUBPIX_LOG2;
    int jmax = (xmax - GG_HALF) >> GG_SUBPIX_LOG2;
    int imin = (ymin + (GG_SUBPIX - GG_HALF - 1)) >> GG_SUBPIX_LOG2;
    int imax = (ymax - GG_HALF) >> GG_SUBPIX_LOG2;
    jmin = max(jmin, 0);
    imin = max(imin, 0);
    jmax = min(jmax, c.W - 1);
    imax = min(imax, c.H - 1);
    if (jmin > jmax || imin > imax) return 0;
    const long long X[3] = {p0.X, p1.X, p2.X}, Y[3] = {p0.Y, p1.Y, p2.Y};
    long long Ak[3], Bk[3], E0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int k1 = (k + 1) % 3;
        Ak[k] = -(Y[k1] - Y[k]);
        Bk[k] = (X[k1] - X[k]);
        E0[k] = Bk[k] * (GG_HALF - Y[k]) + Ak[k] * (GG_HALF - X[k]);  // at pixel (0,0) centre
        // contract C3 (top-left rule): E == 0 is inside only on left / top edges
        const bool inclusive = (Ak[k] > 0) || (Ak[k] == 0 && Bk[k] > 0);
        r.A[k] = (int)Ak[k];
        r.B[k] = (int)Bk[k];
        r.C[k] = E0[k] - (inclusive ? 0 : 1);
    }
    const long long a2 = area2 < 0 ? -area2 : area2;
    const double inv_area = 1.0 / (double)a2;
    const double w0 = p0.invz, w1 = p1.invz, w2 = p2.invz;
    // barycentric weights: E1 -> v0, E2 -> v1, E0 -> v2
    r.w00 = ((double)E0[1] * w0 + (double)E0[2] * w1 + (double)E0[0] * w2) * inv_area;
    r.gx = (double)GG_SUBPIX * ((double)Ak[1] * w0 + (double)Ak[2] * w1 + (double)Ak[0] * w2) * inv_area;
    r.gy = (double)GG_SUBPIX * ((double)Bk[1] * w0 + (double)Bk[2] * w1 + (double)Bk[0] * w2) * inv_area;
    r.w0 = p0.invz;
    r.w1 = p1.invz;
    r.w2 = p2.invz;
    r.face = face_id;
    r.jmin = (uint16_t)jmin;
    r.jmax = (uint16_t)jmax;
    r.imin = (uint16_t)imin;
    r.imax = (uint16_t)imax;
    return 1;
}

// Count the tiles a record touches (one atomicAdd each), fill its tile mask and store it at recs[idx]; beyond the
// capacity only the overflow flags are raised.
__device__ __forceinline__ void emit_record(GGFaceRec &r, int idx, const GGViewScratch &vs, int64_t F, int64_t cap_recs,
                                            int32_t *__restrict__ sticky, int tiles_x) {
    if (idx < cap_recs) {
        const int tx0 = r.jmin / GG_TILE_W, tx1 = r.jmax / GG_TILE_W;
        const int ty0 = r.imin / GG_TILE_H, ty1 = r.imax / GG_TILE_H;
        const int ntx = tx1 - tx0 + 1, nty = ty1 - ty0 + 1;
        unsigned long long tmask = 0;
        const bool small = ntx * nty <= 64;
        // tile_may_touch for every tile of the box, evaluated incrementally: the three edge functions at
        // each tile's innermost corner differ from tile to tile by constants (same int64 values)
        long long e_row[3], dex[3], dey[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int xc = tx0 * GG_TILE_W + (r.A[k] > 0 ? GG_TILE_W - 1 : 0);
            const int yc = ty0 * GG_TILE_H + (r.B[k] > 0 ? GG_TILE_H - 1 : 0);
            e_row[k] = r.C[k] + (long long)r.A[k] * GG_SUBPIX * xc + (long long)r.B[k] * GG_SUBPIX * yc;
            dex[k] = (long long)r.A[k] * (GG_SUBPIX * GG_TILE_W);
            dey[k] = (long long)r.B[k] * (GG_SUBPIX * GG_TILE_H);
        }
        int bit = 0;
        for (int ty = ty0; ty <= ty1; ++ty) {
            long long e0 = e_row[0], e1 = e_row[1], e2 = e_row[2];
            int32_t *row_count = vs.tile_count + ty * tiles_x;
            for (int tx = tx0; tx <= tx1; ++tx, ++bit) {
                if ((e0 | e1 | e2) >= 0) {
                    atomicAdd(&row_count[tx], 1);
                    if (small) tmask |= 1ull << bit;
                }
                e0 += dex[0];
                e1 += dex[1];
                e2 += dex[2];
            }
            e_row[0] += dey[0];
            e_row[1] += dey[1];
            e_row[2] += dey[2];
        }
        r.tmask = small ? tmask : ~0ull;
        store_vec16(&vs.recs[idx], r);
        if (r.face == (int32_t)(F - 1)) vs.counters[5] = idx;
    } else {
        atomicOr(&vs.counters[3], 1);
        atomicOr(sticky, 1);
    }
}

// ---- contract C6: guard-band clipping (rare: faces cut by the near plane close to the camera, huge faces) ----------
// Signed distance (>= 0: inside) of a camera-space point to guard plane k: 0 right, 1 left, 2 bottom, 3 top.
__device__ __forceinline__ float guard_dist(const Cam3 &p, const gg_camera &c, int k) {
    switch (k) {
        case 0: return __fsub_rn(__fmul_rn(__fsub_rn(GG_GUARD_PX, c.px), p.z), __fmul_rn(c.f, p.x));
        case 1: return __fadd_rn(__fmul_rn(__fadd_rn(GG_GUARD_PX, c.px), p.z), __fmul_rn(c.f, p.x));
        case 2: return __fsub_rn(__fmul_rn(__fsub_rn(GG_GUARD_PX, c.py), p.z), __fmul_rn(c.f, p.y));
        default: return __fadd_rn(__fmul_rn(__fadd_rn(GG_GUARD_PX, c.py), p.z), __fmul_rn(c.f, p.y));
    }
}

// The (up to two) camera-space triangles of face fi after clipping against the near plane (contract C5).
__device__ __forceinline__ int face_triangles(const float4 *__restrict__ verts, const int4 *__restrict__ faces, int64_t fi,
                                              const gg_camera &c, Cam3 (&tri)[2][3], int &face_id) {
    const int4 f = faces[fi];
    face_id = f.w;
    const float4 a = verts[f.x], b = verts[f.y], d = verts[f.z];
    const Cam3 p[3] = {cam_space(a.x, a.y, a.z, c), cam_space(b.x, b.y, b.z, c), cam_space(d.x, d.y, d.z, c)};
    bool finite = true;
    int nfront = 0;
    bool front[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        finite = finite && isfinite(p[k].x) && isfinite(p[k].y) && isfinite(p[k].z);
        front[k] = p[k].z >= c.znear;
        nfront += front[k] ? 1 : 0;
    }
    if (finite && nfront == 3) {
        tri[0][0] = p[0];
        tri[0][1] = p[1];
        tri[0][2] = p[2];
        return 1;
    }
    if (finite && nfront == 1) {  // rotate so that the vertex in front comes first: (A, B, C)
        const int ia = front[0] ? 0 : (front[1] ? 1 : 2);
        const Cam3 A = p[ia], B = p[(ia + 1) % 3], C = p[(ia + 2) % 3];
        tri[0][0] = A;
        tri[0][1] = clip_edge(A, B, c.znear);
        tri[0][2] = clip_edge(A, C, c.znear);
        return 1;
    }
    if (finite && nfront == 2) {  // rotate so that the vertex behind comes last: (A, B, C)
        const int ic = !front[0] ? 0 : (!front[1] ? 1 : 2);
        const Cam3 A = p[(ic + 1) % 3], B = p[(ic + 2) % 3], C = p[ic];
        const Cam3 rbc = clip_edge(B, C, c.znear), rac = clip_edge(A, C, c.znear);
        tri[0][0] = A;
        tri[0][1] = B;
        tri[0][2] = rbc;
        tri[1][0] = A;
        tri[1][1] = rbc;
        tri[1][2] = rac;
        return 2;
    }
    return 0;
}

// Sutherland-Hodgman against the four guard planes in float32 with a fixed operation order (the oracle does the
// same arithmetic), fan triangulation, one record per fan triangle appended with its own atomicAdd.
// The triangle is re-derived from the face, so that nothing but (fi, t) stays live across this call in the caller.
__device__ __noinline__ bool setup_guarded(const float4 *__restrict__ verts, const int4 *__restrict__ faces, int64_t fi, int t,
                                           const gg_camera &c, const GGViewScratch &vs, int64_t F, int64_t cap_recs,
                                           int32_t *__restrict__ sticky, int tiles_x, bool kept_before) {
    Cam3 tri[2][3];
    int face_id;
    if (face_triangles(verts, faces, fi, c, tri, face_id) <= t) return kept_before;
    Cam3 buf[2][8];
    int n = 3, cur = 0;
    buf[0][0] = tri[t][0];
    buf[0][1] = tri[t][1];
    buf[0][2] = tri[t][2];
    for (int k = 0; k < 4 && n >= 3; ++k) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const Cam3 P = buf[cur][i], Q = buf[cur][(i + 1) % n];
            const float dP = guard_dist(P, c, k), dQ = guard_dist(Q, c, k);
            const bool inP = dP >= 0.0f, inQ = dQ >= 0.0f;
            if (inP && m < 8) buf[1 - cur][m++] = P;
            if (inP != inQ && m < 8) {  // from the inside vertex towards the outside one
                const Cam3 I = inP ? P : Q, O = inP ? Q : P;
                const float dI = inP ? dP : dQ, dO = inP ? dQ : dP;
                const float t = __fdiv_rn(dI, __fsub_rn(dI, dO));
                Cam3 R;
                R.x = __fadd_rn(I.x, __fmul_rn(t, __fsub_rn(O.x, I.x)));
                R.y = __fadd_rn(I.y, __fmul_rn(t, __fsub_rn(O.y, I.y)));
                R.z = __fadd_rn(I.z, __fmul_rn(t, __fsub_rn(O.z, I.z)));
                buf[1 - cur][m++] = R;
            }
        }
        n = m;
        cur = 1 - cur;
    }
    for (int i = 1; i + 1 < n; ++i) {
        GGFaceRec r;
        if (build_record(buf[cur][0], buf[cur][i], buf[cur][i + 1], face_id, c, r, false) == 1) {
            r.dup = kept_before ? 1 : 0;
            kept_before = true;
            emit_record(r, atomicAdd(&vs.counters[1], 1), vs, F, cap_recs, sticky, tiles_x);
        }
    }
    return kept_before;
}

// ------------------------------------------------------------------------------------------------------
// Face setup: one thread per face of a visible block.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GG_BLOCK_FACES, GG_SETUP_MIN_BLOCKS) k_setup_faces(const float4 *__restrict__ verts,
                                                               const int4 *__restrict__ faces, int64_t F,
                                                               int64_t cap_recs, int32_t *__restrict__ sticky,
                                                               const __grid_constant__ GGCamBatch cams,
                                                               const __grid_constant__ GGViewBatch views) {
    const int view = blockIdx.y;
    const gg_camera &c = cams.cam[view];
    const GGViewScratch &vs = views.v[view];
    const int n_vis = vs.counters[0];
    const int tiles_x = (c.W + GG_TILE_W - 1) / GG_TILE_W;
    for (int vb = blockIdx.x; vb < n_vis; vb += gridDim.x) {
        const int64_t fi = (int64_t)vs.vis_blocks[vb] * GG_BLOCK_FACES + threadIdx.x;
        // up to two triangles per face: faces crossing the near plane are clipped in camera space (contract C5)
        Cam3 tri[2][3];
        int n_tri = 0, face_id = -1;
        if (fi < F) n_tri = face_triangles(verts, faces, fi, c, tri, face_id);
        bool kept_before = false;
        unsigned guarded = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {  // every lane takes part in both rounds (warp-aggregated append)
            int code = 0;
            GGFaceRec r;
            if (t < n_tri) code = build_record(tri[t][0], tri[t][1], tri[t][2], face_id, c, r, true);
            const bool keep = code == 1;
            guarded |= code == 2 ? 1u << t : 0u;
            r.dup = kept_before ? 1 : 0;
            kept_before = kept_before || keep;
            const int idx = warp_append(&vs.counters[1], keep);
            if (keep) emit_record(r, idx, vs, F, cap_recs, sticky, tiles_x);
        }
        if (guarded) {  // contract C6 (rare, divergent): clip against the guard band, one record per fan triangle
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (guarded & (1u << t))
                    kept_before = setup_guarded(verts, faces, fi, t, c, vs, F, cap_recs, sticky, tiles_x, kept_before);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Reserve list space per tile.  Lists need not be stored in tile order, so instead of a scan every warp sums its
// 32 counts and claims a range with one atomicAdd.  tile_count becomes the fill pass's cursor (start of the list; the
// fill pass leaves it at the end of the list).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reserve_tiles(int n_tiles, int64_t cap_recs, int32_t *__restrict__ sticky,
                                                       const __grid_constant__ GGViewBatch views) {
    const GGViewScratch &vs = views.v[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (t == 0) {
        vs.counters[6] = vs.counters[1];  // records wanted (reported by gg_last_batch_stats)
        atomicMax(&sticky[1], vs.counters[1]);  // high-water mark since the last gg_sync (gg_overflow_info)
        if (vs.counters[1] > cap_recs) vs.counters[1] = (int)cap_recs;  // records beyond capacity were dropped
    }
    const int v = (t < n_tiles) ? vs.tile_count[t] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    const int total = __shfl_sync(0xffffffffu, x, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&vs.counters[2], total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (t < n_tiles) {
        vs.tile_offset[t] = base + x - v;
        vs.tile_count[t] = base + x - v;  // the fill pass's cursor: absolute, so that one atomicAdd yields the slot
    }
}

// Tile-relative setup of one face: done once per (tile, face) pair by the fill pass so that the rasterizer only
// streams ready-made 64-byte records.
__device__ __forceinline__ GGTileFace setup_tile_face(const GGFaceRec &r, int rec, int tile_x0, int tile_y0) {
    GGTileFace tf;
    bool fits = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const long long sx = (long long)r.A[k] * GG_SUBPIX, sy = (long long)r.B[k] * GG_SUBPIX;
        const long long e = r.C[k] + sx * tile_x0 + sy * tile_y0;
        // conservative float bound of |E| over the tile (2x head-room under 2^31)
        const float bound = fabsf((float)e) + (float)(GG_TILE_W - 1) * fabsf((float)sx) + (float)(GG_TILE_H - 1) * fabsf((float)sy);
        fits = fits && (bound < 1073741824.0f);
        tf.e[k] = (int)e;
        tf.sx[k] = (int)sx;
        tf.sy[k] = (int)sy;
    }
    const double w_org = r.w00 + r.gx * (double)tile_x0 + r.gy * (double)tile_y0;
    tf.w_org = (float)w_org;
    tf.gx = (float)r.gx;
    tf.gy = (float)r.gy;
    const float wmin = fminf(r.w0, fminf(r.w1, r.w2));
    const float spread = fabsf(tf.w_org) + (float)(GG_TILE_W - 1) * fabsf(tf.gx) + (float)(GG_TILE_H - 1) * fabsf(tf.gy);
    fits = fits && (spread <= 16.0f * wmin);
    const int bx0 = max((int)r.jmin - tile_x0, 0), bx1 = min((int)r.jmax - tile_x0, GG_TILE_W - 1);
    const int by0 = max((int)r.imin - tile_y0, 0), by1 = min((int)r.imax - tile_y0, GG_TILE_H - 1);
    // lane = row * 4 + strip: rows by0..by1, strips (bx0 >> 3)..(bx1 >> 3)
    const unsigned strips = (0xfu >> (3 - (bx1 >> 3))) & (0xfu << (bx0 >> 3)) & 0xfu;
    unsigned rows = (0xffu >> (7 - by1)) & (0xffu << by0) & 0xffu;  // bit r = row r
    rows = (rows | (rows << 12)) & 0x000f000fu;                      // spread the 8 row bits to one bit per nibble
    rows = (rows | (rows << 6)) & 0x03030303u;
    rows = (rows | (rows << 3)) & 0x11111111u;
    tf.lanemask = rows * strips;
    tf.face = r.face;
    tf.rec = rec;
    tf.fast = fits ? 1u : 0u;
    return tf;
}

__global__ void __launch_bounds__(256, GG_FILL_MIN_BLOCKS) k_fill_bins(int64_t cap_bins, int32_t *__restrict__ sticky,
                                                   const __grid_constant__ GGCamBatch cams,
                                                   const __grid_constant__ GGViewBatch views) {
    const int view = blockIdx.y;
    const GGViewScratch &vs = views.v[view];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&sticky[2], vs.counters[2]);
    if ((int64_t)vs.counters[2] > cap_bins) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicOr(&vs.counters[3], 2);
            atomicOr(sticky, 2);
        }
        return;
    }
    if (vs.counters[3] != 0) return;
    const int n_recs = vs.counters[1];
    const int tiles_x = (cams.cam[view].W + GG_TILE_W - 1) / GG_TILE_W;
    // GG_FILL_LANES lanes share one record and split its tiles, so that the atomicAdd -> store chains of a face run in parallel
    const int sub = threadIdx.x & (GG_FILL_LANES - 1);
    const int groups = (gridDim.x * blockDim.x) / GG_FILL_LANES;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) / GG_FILL_LANES; r < n_recs; r += groups) {
        const GGFaceRec rec = load_vec16(&vs.recs[r]);
        const int tx0 = rec.jmin / GG_TILE_W, tx1 = rec.jmax / GG_TILE_W;
        const int ty0 = rec.imin / GG_TILE_H, ty1 = rec.imax / GG_TILE_H;
        const int ntx = tx1 - tx0 + 1;
        const unsigned long long tmask = rec.tmask;
        if (tmask != ~0ull) {
            // lane `sub` looks at bits sub, sub + 8, ... of the mask (a box of up to 64 tiles): no bit counting, and the
            // 8 lanes of a record stay in step
            const float inv_ntx = 1.0f / (float)ntx;
            for (int b = sub; b < 64 && (tmask >> b) != 0; b += GG_FILL_LANES) {
                if (!((tmask >> b) & 1ull)) continue;
                const int by = (int)(((float)b + 0.5f) * inv_ntx);  // b / ntx, exact for these small integers
                const int tx = tx0 + (b - by * ntx), ty = ty0 + by;
                const int t = ty * tiles_x + tx;
                store_vec16(&vs.bins[atomicAdd(&vs.tile_count[t], 1)],
                            setup_tile_face(rec, r, tx * GG_TILE_W, ty * GG_TILE_H));
            }
        } else {
            const int total = ntx * (ty1 - ty0 + 1);
            for (int i = sub; i < total; i += GG_FILL_LANES) {
                const int tx = tx0 + i % ntx, ty = ty0 + i / ntx;
                if (!tile_may_touch(rec, tx, ty)) continue;
                const int t = ty * tiles_x + tx;
                store_vec16(&vs.bins[atomicAdd(&vs.tile_count[t], 1)],
                            setup_tile_face(rec, r, tx * GG_TILE_W, ty * GG_TILE_H));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Tile rasterizer: a WARP per 32 x 8 px tile at a time, no block-level synchronisation.
// ------------------------------------------------------------------------------------------------------
// Exact evaluation of one face at one pixel (slow path: long edges or steep depth planes).
__device__ __noinline__ bool exact_cover(const GGFaceRec &r, int j, int i, float &w) {
    long long E[3];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        E[k] = r.C[k] + (long long)r.A[k] * GG_SUBPIX * j + (long long)r.B[k] * GG_SUBPIX * i;
        inside = inside && (E[k] >= 0);
        const bool inclusive = (r.A[k] > 0) || (r.A[k] == 0 && r.B[k] > 0);
        E[k] += inclusive ? 0 : 1;  // undo the fill-rule bias for the depth interpolation
    }
    if (!inside) return false;
    const double area = (double)(E[0] + E[1] + E[2]);
    w = (float)(((double)E[1] * (double)r.w0 + (double)E[2] * (double)r.w1 + (double)E[0] * (double)r.w2) / area);
    return true;
}

__device__ __forceinline__ int4 lds128(unsigned addr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// MODE 0: rasters only.  MODE 1: + last pixel of every face (fused last-pixel / vote aggregation).
// MODE 2: + dense per-pixel score sums (GG_MODE_PIXEL_SUM), T = element type of the score images.
#ifndef GG_DENSE_MIN_BLOCKS
#define GG_DENSE_MIN_BLOCKS GG_RASTER_MIN_BLOCKS  // 6 (80 registers, no spills) measured 8 % slower: occupancy wins
#endif
#define GG_RM_PLAIN 0
#define GG_RM_WINNERS 1
#define GG_RM_DENSE 2
#define GG_DENSE_ACC_FLOATS (GG_CHUNK * 32 + 64)  // per warp: slots of the list positions + one scratch row
#define GG_RM_GATHER 3  // fused render_flat: out[p, :] = tex[face, :] (T = output element type)
#define GG_RM_WINNERS_ONLY 4  // GG_RM_WINNERS without face-ID / depth rasters (the fused aggregation never asks for them)

struct GGDenseArgs {
    GGPredBatch preds;
    double *sum;     // [F][C]
    int32_t *count;  // [F]
    int C;
    int index_kind;  // 1: (H,W) uint8 class index expanded on the fly
    int vec_ok;      // every image base and row start is 16-byte aligned: two channels per load
    int l2_prefetch; // ... and so is every tile row: bulk L2 prefetch of the tile's scores ahead of the epilogue
    // GG_RM_GATHER
    const double *tex;  // [F][D]
    void *out;          // [n][H][W][D]
    int D;
};

template <typename OUT>
__device__ __forceinline__ OUT gather_convert(double v);
template <>
__device__ __forceinline__ double gather_convert<double>(double v) {
    return v;
}
template <>
__device__ __forceinline__ float gather_convert<float>(double v) {
    return (float)v;
}
template <>
__device__ __forceinline__ uint8_t gather_convert<uint8_t>(double v) {
    // save_renders' rule (meshes.py:2323-2334): < 0, > 255 or non-finite -> 0, then astype(uint8) truncates
    if (!(v >= 0.0) || v > 255.0 || !isfinite(v)) return 0;
    return (uint8_t)v;
}

template <typename T>
__device__ __forceinline__ float dense_load(const T *__restrict__ pred, int64_t pix, int C, int ch, int index_kind) {
    if (index_kind) return ((int)pred[pix] == ch) ? 1.f : 0.f;
    const float v = (float)pred[pix * C + ch];
    return v == v ? v : 0.f;  // NaN = null -> contributes nothing
}

// LL ("lane lists", not in the dense mode): instead of the whole warp stepping through the tile's list face by face --
// every face costs the full per-pixel block however few lanes it touches -- every lane first collects the faces whose
// lane mask holds ITS strip (one bit per face of the chunk) and then walks its own list: the warp runs
// max-over-lanes iterations, lanes work on different faces in the same iteration (lane-private shared-memory reads).
template <int MODE, typename T, int CT, bool LL = false>
__global__ void __launch_bounds__(GG_RASTER_THREADS, MODE == GG_RM_DENSE ? GG_DENSE_MIN_BLOCKS : GG_RASTER_MIN_BLOCKS) k_raster_tiles(const __grid_constant__ GGCamBatch cams,
                                                                       const __grid_constant__ GGViewBatch views,
                                                                       int n_tiles, int32_t *__restrict__ pix2face,
                                                                       float *__restrict__ depth, int compat_bg,
                                                                       const __grid_constant__ GGDenseArgs dense) {
    constexpr bool WINNERS = (MODE == GG_RM_WINNERS || MODE == GG_RM_WINNERS_ONLY);
    constexpr bool RASTER_OUT = (MODE != GG_RM_WINNERS_ONLY);  // pix2face / depth may be requested
    constexpr bool NEED_POS = (MODE == GG_RM_DENSE);  // only the dense epilogue indexes by list position
    // grid: (groups of GG_RASTER_WARPS * TPW tiles along x, tile rows, views); a warp rasterizes TPW consecutive
    // tiles of a row, one after the other.  TPW > 1: the list bounds of all its tiles come back in ONE memory round
    // trip (lane j loads tile j's), and the setups of tile j+1 travel into the second shared-memory buffer (cp.async)
    // while tile j is rasterized, so that only a warp's first tile waits for memory.
    constexpr int TPW = (MODE == GG_RM_DENSE) ? GG_DENSE_TILES_PER_WARP : GG_TILES_PER_WARP;
    const int view = blockIdx.z;
    const gg_camera &c = cams.cam[view];
    const GGViewScratch &vs = views.v[view];
    const int W = c.W, H = c.H;
    const int tiles_x = (W + GG_TILE_W - 1) / GG_TILE_W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_xb = (blockIdx.x * GG_RASTER_WARPS + warp) * TPW;
    if (tile_xb >= tiles_x) return;  // whole warp leaves; no block-level barriers below
    const int tile_y0 = blockIdx.y * GG_TILE_H;

    __shared__ GGTileFace s_all[GG_RASTER_WARPS][TPW > 1 ? 2 : 1][GG_CHUNK];
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(&s_all[warp][0][0]);

    const int tx0 = (lane & 3) * 8;  // this lane: pixels tx0..tx0+7 of row ty
    const int ty = lane >> 2;
    const float fty = (float)ty, ftx0 = (float)tx0;

    const bool overflow = vs.counters[3] != 0;
    int my_beg = 0, my_end = 0;  // lane j < TPW: bounds of the warp's j-th tile (the fill cursor ends at the list's end)
    if (lane < TPW && tile_xb + lane < tiles_x) {
        const int t = blockIdx.y * tiles_x + tile_xb + lane;
        my_beg = vs.tile_offset[t];
        // loaded unconditionally, beside the overflow flag (the compiler would predicate it on the flag and chain two
        // memory round trips)
        asm volatile("ld.global.s32 %0, [%1];" : "=r"(my_end) : "l"(vs.tile_count + t));
    }
    if (overflow) my_end = my_beg;
    auto stage_async = [&](int buf, int beg, int n) {  // n <= GG_CHUNK setups -> buffer buf, one commit group
        const char *src = reinterpret_cast<const char *>(vs.bins + beg);
        const unsigned dst = s_base + (unsigned)buf * (unsigned)(GG_CHUNK * sizeof(GGTileFace));
        for (int idx = lane; idx < n * 4; idx += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + idx * 16), "l"(src + idx * 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (TPW > 1) {
        const int b0 = __shfl_sync(0xffffffffu, my_beg, 0), e0 = __shfl_sync(0xffffffffu, my_end, 0);
        stage_async(0, b0, min(GG_CHUNK, e0 - b0));
    }

#pragma unroll 1
  for (int j = 0; j < TPW; ++j) {
    const int tile_x = tile_xb + j;
    if (tile_x >= tiles_x) break;
    const int tile_x0 = tile_x * GG_TILE_W;
    const int beg = __shfl_sync(0xffffffffu, my_beg, j);
    const int len = __shfl_sync(0xffffffffu, my_end, j) - beg;
    const unsigned s_faces_addr = s_base + (TPW > 1 ? (unsigned)(j & 1) * (unsigned)(GG_CHUNK * sizeof(GGTileFace)) : 0u);
    GGTileFace *s_faces = &s_all[warp][TPW > 1 ? (j & 1) : 0][0];
    if (TPW > 1) {
        if (j + 1 < TPW && tile_x + 1 < tiles_x) {  // next tile's setups into the other buffer, then wait for this tile's
            const int nb = __shfl_sync(0xffffffffu, my_beg, j + 1), ne = __shfl_sync(0xffffffffu, my_end, j + 1);
            stage_async((j + 1) & 1, nb, min(GG_CHUNK, ne - nb));
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    }

    // Per pixel: the winner's key (bits of its 1/z, which is positive -> ordered like the float; then ~face so that the
    // lower ID wins a tie) compared as ONE 64-bit quantity, and its position in the tile's list (-1 = none).
    // The key lives in a register PAIR read as one float64: for these bit patterns (sign clear, exponent field below
    // all-ones) the float64 order equals the unsigned 64-bit order, and DSETP -- one instruction on the FP64 pipe --
    // replaces the two ISETPs of a 64-bit integer compare on the ALU pipe, which is the pipe that limits this kernel.
    double key[8];
    int bp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        key[i] = 0.0;  // 1/z = 0, ~face = 0 (face -1)
        bp[i] = -1;
    }

    if (MODE == GG_RM_DENSE) {
        // The epilogue will stream this tile's scores: ask for them now (one bulk L2 prefetch per tile row, issued by
        // one lane each) so that they travel from HBM while the faces are rasterized.
        const int C = CT > 0 ? CT : dense.C;
        if (dense.l2_prefetch && len > 0 && W - tile_x0 >= GG_TILE_W && lane < min(GG_TILE_H, H - tile_y0)) {
            const T *src = (const T *)dense.preds.p[view] + ((int64_t)(tile_y0 + lane) * W + tile_x0) * C;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(GG_TILE_W * C * sizeof(T)))
                         : "memory");
        }
    }

    for (int base = 0; base < len; base += GG_CHUNK) {
        const int n = min(GG_CHUNK, len - base);
        if (TPW == 1 || base > 0) {  // (TPW > 1: the first chunk was prefetched)
           // stream n ready-made 64-byte setups into shared memory, 16 B per lane and pass: the usual list of <= 8
           // faces takes ONE pass (a four-way unrolled, predicated copy cost ~90 instructions per tile)
            if (TPW > 1) __syncwarp();  // every lane is done with the previous chunk in this buffer
            const int4 *src = reinterpret_cast<const int4 *>(vs.bins + beg + base);
#pragma unroll 1
            for (int idx = lane; idx < n * 4; idx += 64) {
                const bool two = idx + 32 < n * 4;
                const int4 v0 = __ldg(src + idx);
                int4 v1 = v0;
                if (two) v1 = __ldg(src + idx + 32);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_faces_addr + idx * 16), "r"(v0.x), "r"(v0.y),
                             "r"(v0.z), "r"(v0.w)
                             : "memory");
                if (two)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_faces_addr + idx * 16 + 512), "r"(v1.x),
                                 "r"(v1.y), "r"(v1.z), "r"(v1.w)
                                 : "memory");
            }
        }
        __syncwarp();
        if (LL) {
            static_assert(!LL || !NEED_POS, "lane lists carry no list position");
            unsigned mine = 0;  // bit k: face k of this chunk may touch this lane's strip
            {
                unsigned ma = s_faces_addr + 48;
                for (int k = 0; k < n; ++k, ma += 64) {
                    unsigned lm;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lm) : "r"(ma));
                    mine |= ((lm >> lane) & 1u) << k;
                }
            }
            while (__any_sync(0xffffffffu, mine != 0u)) {
                const bool act = mine != 0u;
                const unsigned kk = act ? (unsigned)(__ffs((int)mine) - 1) : 0u;
                mine &= mine - 1u;  // (0 stays 0)
                const unsigned fa = s_faces_addr + kk * 64u;
                const int4 q3 = lds128(fa + 48);  // lanemask face rec fast
                const unsigned nface = ~(unsigned)q3.y;
                if (q3.w) {
                    const int4 q0 = lds128(fa);       // e0 e1 e2 sx0
                    const int4 q1 = lds128(fa + 16);  // sx1 sx2 sy0 sy1
                    const int4 q2 = lds128(fa + 32);  // sy2 w_org gx gy
                    const int s0 = q0.w, s1 = q1.x, s2 = q1.y;
                    int e0 = q0.x + s0 * tx0 + q1.z * ty;
                    int e1 = q0.y + s1 * tx0 + q1.w * ty;
                    int e2 = q0.z + s2 * tx0 + q2.x * ty;
                    e0 = act ? e0 : -1;  // a lane whose list has run out covers nothing
                    const int s0m = act ? s0 : 0;
                    const float gx = __int_as_float(q2.z);
                    const float wrow = fmaf(gx, ftx0, fmaf(__int_as_float(q2.w), fty, __int_as_float(q2.y)));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float w = fmaf(gx, (float)i, wrow);
                        const double cand = __hiloint2double(__float_as_int(w), (int)nface);
                        asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %0;\n\tsetp.ge.and.s32 p, %2, 0, p;\n\t"
                            "selp.f64 %0, %1, %0, p;\n\t}"
                            : "+d"(key[i])
                            : "d"(cand), "r"(e0 | e1 | e2));
                        e0 += s0m;
                        e1 += s1;
                        e2 += s2;
                    }
                } else if (act) {
                    const GGFaceRec &r = vs.recs[q3.z];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float w;
                        if (exact_cover(r, tile_x0 + tx0 + i, tile_y0 + ty, w) && w > 0.f) {
                            const double cand = __hiloint2double(__float_as_int(w), (int)nface);
                            if (cand > key[i]) key[i] = cand;
                        }
                    }
                }
            }
        } else {
        unsigned fa = s_faces_addr;  // shared-memory address of setup k: one register, bumped by 64 per face
        for (int k = 0; k < n; ++k, fa += 64) {
            const int4 q3 = lds128(fa + 48);  // lanemask face rec fast
            if (!((((unsigned)q3.x) >> lane) & 1u)) continue;
            const unsigned nface = ~(unsigned)q3.y;
            const int pos = base + k;
            if (q3.w) {
                const int4 q0 = lds128(fa);       // e0 e1 e2 sx0
                const int4 q1 = lds128(fa + 16);  // sx1 sx2 sy0 sy1
                const int4 q2 = lds128(fa + 32);  // sy2 w_org gx gy
                const int s0 = q0.w, s1 = q1.x, s2 = q1.y;
                int e0 = q0.x + s0 * tx0 + q1.z * ty;
                int e1 = q0.y + s1 * tx0 + q1.w * ty;
                int e2 = q0.z + s2 * tx0 + q2.x * ty;
                const float gx = __int_as_float(q2.z);
                const float wrow = fmaf(gx, ftx0, fmaf(__int_as_float(q2.w), fty, __int_as_float(q2.y)));
#pragma unroll
                for (int i = 0; i < 8; ++i) {  // branch-free: selects cost less than the divergence bookkeeping
                    const float w = fmaf(gx, (float)i, wrow);
                    const double cand = __hiloint2double(__float_as_int(w), (int)nface);
                    if (NEED_POS) {
                        const bool upd = ((e0 | e1 | e2) >= 0) & (cand > key[i]);  // w <= 0 (never inside a face) loses
                        key[i] = upd ? cand : key[i];
                        bp[i] = upd ? pos : bp[i];
                    } else {
                        // the same update with the two conditions folded into ONE predicate by hand (left to itself the
                        // compiler nests two selects per key half when the predicate has a single consumer)
                        asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %0;\n\tsetp.ge.and.s32 p, %2, 0, p;\n\t"
                            "selp.f64 %0, %1, %0, p;\n\t}"
                            : "+d"(key[i])
                            : "d"(cand), "r"(e0 | e1 | e2));
                    }
                    e0 += s0;
                    e1 += s1;
                    e2 += s2;
                }
            } else {
                const GGFaceRec &r = vs.recs[q3.z];
#pragma unroll
                for (int i = 0; i < 8; ++i) {  // unrolled so that bw / bf / br stay in registers
                    float w;
                    if (exact_cover(r, tile_x0 + tx0 + i, tile_y0 + ty, w) && w > 0.f) {
                        const double cand = __hiloint2double(__float_as_int(w), (int)nface);
                        if (cand > key[i]) {
                            key[i] = cand;
                            if (NEED_POS) bp[i] = pos;
                        }
                    }
                }
            }
        }
        }  // !LL
        __syncwarp();
    }

    float bw[8];
    int bf[8];  // 1/z and face ID of the winners, -1 = none
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bw[i] = __int_as_float(__double2hiint(key[i]));
        bf[i] = ~__double2loint(key[i]);
    }

    // ---- write the 8 pixels of this lane ----
    const int row = tile_y0 + ty, col = tile_x0 + tx0;
    const bool row_ok = row < H;
    if (RASTER_OUT && row_ok && col < W) {
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

    if (WINNERS) {
        // Last pixel (row-major) won by every face.  Only the end of a run of equal winners inside this lane's 8
        // pixels can be the face's last pixel of the row; a run continued by the next strip is left to that strip.
        // Every run-end is ONE predicated max-reduction straight into the view's winner array (fire and forget, L2
        // resident): the per-pixel state carries no list position, and no shared-memory pre-reduction or flush.
        // nf = the key's low word = ~face (0 = no winner; 1 = a face ID no mesh has, the "nothing follows" mark).
        unsigned nf[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) nf[i] = (unsigned)__double2loint(key[i]);
        const unsigned next_first = __shfl_down_sync(0xffffffffu, nf[0], 1);
        const bool has_next = (lane & 3) != 3;
        int bgmax = -1;
        const int pix0 = row * W + col;
        int *const winner = vs.winner;
        const char *const wbase = reinterpret_cast<const char *>(winner) - 4;  // &winner[~nf] = wbase - 4 * (int)nf
        if (tile_x0 + GG_TILE_W <= W && tile_y0 + GG_TILE_H <= H) {  // tile entirely inside the image (warp-uniform)
            const unsigned after = has_next ? next_first : 1u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned nxt = (i < 7) ? nf[i < 7 ? i + 1 : 7] : after;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\tsetp.ne.and.u32 p, %0, 0, p;\n\t"
                             "@p red.global.max.s32 [%2], %3;\n\t}" ::"r"(nf[i]), "r"(nxt), "l"(wbase - (int64_t)(int)nf[i] * 4),
                             "r"(pix0 + i)
                             : "memory");
            }
            if (compat_bg) {
#pragma unroll
                for (int i = 0; i < 8; ++i) bgmax = nf[i] == 0u ? pix0 + i : bgmax;  // pixel index grows with i
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool in_img = row_ok && (col + i < W);
                const int pix = pix0 + i;
                const unsigned nxt =
                    (i < 7) ? ((col + i + 1 < W) ? nf[i < 7 ? i + 1 : 7] : 1u) : ((has_next && col + 8 < W) ? next_first : 1u);
                if (in_img) {
                    if (nf[i] == 0u) bgmax = pix;  // pixel index grows with i
                    else if (nf[i] != nxt) atomicMax(&winner[(int)~nf[i]], pix);
                }
            }
        }
        if (compat_bg) {  // meshes.py:2000: background pixels index the last face
            bgmax = __reduce_max_sync(0xffffffffu, bgmax);
            if (lane == 0 && bgmax >= 0) atomicMax(&winner[compat_bg - 1], bgmax);
        }
    }

    if (MODE == GG_RM_GATHER) {
        // Fused render_flat (meshes.py:1921-1937): every pixel takes the texture row of its face, NaN where no face
        // was hit; the face-ID raster itself need not be written.
        const int D = dense.D;
        T *out = reinterpret_cast<T *>(dense.out);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        if (row_ok && col < W) {
            const int64_t o = ((int64_t)view * H + row) * W + col;
            if (D == 1 && sizeof(T) == 1 && col + 7 < W && ((o & 7) == 0)) {
                unsigned lo = 0, hi = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned b = (unsigned)gather_convert<uint8_t>(bf[i] >= 0 ? dense.tex[bf[i]] : nan);
                    if (i < 4) lo |= b << (8 * i);
                    else hi |= b << (8 * (i - 4));
                }
                *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(dense.out) + o) = make_uint2(lo, hi);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (col + i < W) {
                        for (int d = 0; d < D; ++d)
                            out[(o + i) * D + d] = gather_convert<T>(bf[i] >= 0 ? dense.tex[(int64_t)bf[i] * D + d] : nan);
                    }
                }
            }
        }
    }

    if (MODE == GG_RM_DENSE) {
        // Per-pixel scatter-add of the view's (H, W, C) scores into per-face sums; the raster never leaves the SM.
        // The lanes are re-mapped from "8 pixels each" to (pixel group, channel block): a lane owns V consecutive
        // channels (V = 2 when C is even and the rows are aligned, else 1) of every g-th pixel, g = 32 / (C / V), so
        // every load instruction covers g*C contiguous elements of the tile row and a lane's channels never change.
        // A lane sums in registers while the winning face stays the same and spills into its PRIVATE shared-memory
        // slots s_acc[list position][lane*V + j] when it changes (no shared-memory float atomics: they are CAS
        // loops).  The fast path adds the raw values; a null (NaN) or infinite score shows up in the tile's totals
        // and sends the (rare) tile through the filtering loop instead.  The tile then issues one float64 atomicAdd
        // per (face, channel).
        extern __shared__ float s_dyn[];
        constexpr int V = (CT > 0 && CT % 2 == 0 && sizeof(T) * 2 <= 16) ? 2 : 1;  // channels per lane and load
        constexpr int kSlots = 32 * V;              // accumulator slots per list position
        constexpr int kPosCap = GG_CHUNK / V;       // list positions that have slots; later ones use direct atomics
        __shared__ unsigned char s_pos_all[GG_RASTER_WARPS][GG_TILE_W * GG_TILE_H];
        __shared__ int s_cnt_all[GG_RASTER_WARPS][GG_CHUNK];
        const int C = CT > 0 ? CT : dense.C;  // CT > 0: channel count known at compile time (fully unrolled rows)
        float *s_acc = s_dyn + warp * GG_DENSE_ACC_FLOATS;
        unsigned char *s_pos = s_pos_all[warp];
        int *s_cnt = s_cnt_all[warp];
        const T *__restrict__ pred = (const T *)dense.preds.p[view];
        {
            unsigned lo = 0, hi = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned b = (bp[i] >= 0 && bp[i] < kPosCap) ? (unsigned)bp[i] : (unsigned)kPosCap;  // no slot
                if (i < 4) lo |= b << (8 * i);
                else hi |= b << (8 * (i - 4));
            }
            *reinterpret_cast<uint2 *>(&s_pos[ty * GG_TILE_W + tx0]) = make_uint2(lo, hi);
        }
        const int nk = min(len, kPosCap);
        for (int i = lane; i < nk * kSlots; i += 32) s_acc[i] = 0.f;
        s_cnt[lane] = 0;
        __syncwarp();
        {  // pixel counts per list position: integer shared-memory atomics are native, one per run of equal winners
            int run = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool in_img = row_ok && (col + i < W);
                if (in_img && bp[i] >= 0 && bp[i] < kPosCap) {
                    run += 1;
                    const bool last = (i == 7) || (bp[i + 1 < 8 ? i + 1 : 7] != bp[i]) || !(col + i + 1 < W);
                    if (last) {
                        atomicAdd(&s_cnt[bp[i]], run);
                        run = 0;
                    }
                }
            }
        }
        // list positions beyond the shared-memory table (tiles with more than kPosCap faces): direct atomics
        // (done first: the winners in bf / bp are not needed after this and their registers are free for the loads)
        if (len > kPosCap) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (bp[i] >= kPosCap && row_ok && col + i < W) {
                    const int64_t pix = (int64_t)row * W + col + i;
                    for (int cch = 0; cch < C; ++cch)
                        atomicAdd(&dense.sum[(int64_t)bf[i] * C + cch], (double)dense_load<T>(pred, pix, C, cch, dense.index_kind));
                    atomicAdd(&dense.count[bf[i]], 1);
                }
            }
        }
        const int cols = min(GG_TILE_W, W - tile_x0), rows = min(GG_TILE_H, H - tile_y0);
        const bool index_kind = dense.index_kind != 0;

        // -- general loop: any channel count <= 32, any tile width, class-index images, nulls filtered -----------
        auto filtered_rows = [&]() {
            const int g = 32 / C;   // pixels per step
            const int gC = g * C;   // elements per step
            const int grp = lane / C, ch = lane - grp * C;
            const int steps = (cols + g - 1) / g;
            if (lane >= gC) return;
            float acc = 0.f;
            unsigned cur = (unsigned)kPosCap;
            const int stride = index_kind ? g : gC;     // elements per step
            const int first = index_kind ? grp : lane;  // this lane's element in step 0
            const int last = (index_kind ? cols : cols * C) - 1;  // last element of the tile row
            for (int r = 0; r < rows; ++r) {
                const int64_t pix0 = (int64_t)(tile_y0 + r) * W + tile_x0;
                const T *__restrict__ rp = pred + (index_kind ? pix0 : pix0 * C);
                const unsigned char *prow = s_pos + r * GG_TILE_W;
                for (int j0 = 0; j0 < steps; j0 += 4) {
                    T raw[4];
                    unsigned ps[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)  // four independent, unconditional loads (index clamped to the row)
                        raw[u] = rp[min((j0 + u) * stride + first, last)];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int x = (j0 + u) * g + grp;
                        const unsigned pp = prow[min(x, GG_TILE_W - 1)];
                        ps[u] = x < cols ? pp : (unsigned)kPosCap;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float v;
                        if (index_kind) v = ((int)raw[u] == ch) ? 1.f : 0.f;
                        else {
                            v = (float)raw[u];
                            v = v == v ? v : 0.f;  // NaN = null -> contributes nothing
                        }
                        if (ps[u] != cur) {
                            if (cur < (unsigned)kPosCap) s_acc[cur * kSlots + lane] += acc;
                            acc = 0.f;
                            cur = ps[u];
                        }
                        acc += v;
                    }
                }
            }
            if (cur < (unsigned)kPosCap) s_acc[cur * kSlots + lane] += acc;
        };

        bool fast = false;
        if (CT > 0) fast = cols == GG_TILE_W && nk > 0 && (V == 1 || dense.vec_ok);
        if (CT > 0 && fast) {
            // Compile-time channel count and a full-width tile: the whole row (kSteps loads per lane) is issued from
            // one base pointer with immediate offsets, then every value is added, unfiltered, straight into the
            // lane's private slot of its pixel's list position (pixels without a slot go to a scratch row): no state
            // is carried from pixel to pixel, so lanes never diverge.
            constexpr int kC = CT > 0 ? CT : 1;
            constexpr int kLP = kC / V;                 // lanes per pixel
            constexpr int kG = 32 / kLP;                // pixels per step
            constexpr int kSteps = (GG_TILE_W + kG - 1) / kG;
            struct alignas(sizeof(T) * V) Vec {
                T v[V];
            };
            struct alignas(sizeof(float) * V) Acc {
                float v[V];
            };
            const int grp = lane / kLP;
            if (lane < kG * kLP) {
                const Vec *__restrict__ lp = reinterpret_cast<const Vec *>(pred + ((int64_t)tile_y0 * W + tile_x0) * kC) + lane;
                const int64_t row_stride = (int64_t)W * kLP;  // in Vec units
                const unsigned char *q = s_pos + grp;
                Acc *slot = reinterpret_cast<Acc *>(s_acc) + lane;
                auto load_row = [&](Vec(&raw)[kSteps], int r) {
                    const Vec *__restrict__ rp = lp + r * row_stride;
#pragma unroll
                    for (int u = 0; u < kSteps; ++u) {
                        const bool in_row = (u + 1) * kG <= GG_TILE_W || u * kG + grp < GG_TILE_W;
                        if (in_row) raw[u] = rp[u * (kG * kLP)];
                    }
                };
                auto add_row = [&](const Vec(&raw)[kSteps], int r) {
                    const unsigned char *qr = q + r * GG_TILE_W;
#pragma unroll
                    for (int u = 0; u < kSteps; ++u) {
                        const bool in_row = (u + 1) * kG <= GG_TILE_W || u * kG + grp < GG_TILE_W;
                        if (in_row) {
                            Acc *a = slot + (unsigned)qr[u * kG] * 32;  // kSlots floats per list position
                            Acc t = *a;
#pragma unroll
                            for (int j = 0; j < V; ++j) t.v[j] += (float)raw[u].v[j];
                            *a = t;
                        }
                    }
                };
                // three rows in flight: the loads of the rows after next are issued before the current one is added
                Vec rawA[kSteps], rawB[kSteps], rawC[kSteps];
                load_row(rawA, 0);
                if (1 < rows) load_row(rawB, 1);
                for (int r = 0; r < rows; r += 3) {
                    if (r + 2 < rows) load_row(rawC, r + 2);
                    add_row(rawA, r);
                    if (r + 1 < rows) {
                        if (r + 3 < rows) load_row(rawA, r + 3);
                        add_row(rawB, r + 1);
                    }
                    if (r + 2 < rows) {
                        if (r + 4 < rows) load_row(rawB, r + 4);
                        add_row(rawC, r + 2);
                    }
                }
            }
        } else if (nk > 0) {
            filtered_rows();
        }
        __syncwarp();
        // slot of (pixel group q, channel c) is q*C + c on both paths; groups beyond a path's own are zero
        const int g_all = kSlots / C;
        // One float64 atomicAdd per (face, channel).  On the unfiltered path a total that is not finite means a null or
        // infinite score somewhere under that face in this tile: those entries are held back, the tile is summed again
        // by the filtering loop, and only they are added from the second pass.
        unsigned redo = 0;
        {
            int it = 0;
            for (int idx = lane; idx < nk * C; idx += 32, ++it) {
                const int k = idx / C, cch = idx - k * C;
                const int n_px = s_cnt[k];
                float total = 0.f;
                for (int q = 0; q < g_all; ++q) total += s_acc[k * kSlots + q * C + cch];
                if (n_px > 0) {
                    const int64_t face = len <= GG_CHUNK ? s_faces[k].face : vs.bins[beg + k].face;
                    if (cch == 0) atomicAdd(&dense.count[face], n_px);
                    if (CT > 0 && fast && !(fabsf(total) <= 3.0e38f)) redo |= 1u << it;
                    else atomicAdd(&dense.sum[face * C + cch], (double)total);
                }
            }
        }
        if (CT > 0 && fast && __any_sync(0xffffffffu, redo != 0)) {
            for (int i = lane; i < nk * kSlots; i += 32) s_acc[i] = 0.f;
            __syncwarp();
            filtered_rows();
            __syncwarp();
            int it = 0;
            for (int idx = lane; idx < nk * C; idx += 32, ++it) {
                if (!((redo >> it) & 1u)) continue;
                const int k = idx / C, cch = idx - k * C;
                float total = 0.f;
                for (int q = 0; q < g_all; ++q) total += s_acc[k * kSlots + q * C + cch];
                const int64_t face = len <= GG_CHUNK ? s_faces[k].face : vs.bins[beg + k].face;
                atomicAdd(&dense.sum[face * C + cch], (double)total);
            }
        }
    }
  }  // the warp's next tile
}

}  // namespace

// ======================================================================================================
// Host side
// ======================================================================================================
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int gg_ensure_scratch(gg_context *ctx, int n_views, int W, int H) {
    const int64_t tiles = (int64_t)((W + GG_TILE_W - 1) / GG_TILE_W) * ((H + GG_TILE_H - 1) / GG_TILE_H);
    // defaults: a view rarely sees more than 1/8 of a large mesh, and a tile list is a few faces long; both are
    // grown on demand (GG_ERR_OVERFLOW -> gg_reserve -> replay)
    const int64_t auto_recs = ctx->F < (1 << 20) ? ctx->F : ((ctx->F / 8 > (1 << 20)) ? ctx->F / 8 : (1 << 20));
    const int64_t cap_recs = ctx->req_recs > 0 ? ctx->req_recs : auto_recs;
    const int64_t cap_bins = ctx->req_bins > 0 ? ctx->req_bins : (16 * tiles > (1 << 20) ? 16 * tiles : (1 << 20));
    if (n_views <= ctx->n_slots && tiles <= ctx->slot_tiles && cap_recs == ctx->cap_recs && cap_bins == ctx->cap_bins)
        return GG_OK;
    if (ctx->d_scratch) {
        GG_CUDA(cudaDeviceSynchronize());
        GG_CUDA(cudaFree(ctx->d_scratch));
        ctx->d_scratch = nullptr;
    }
    const int slots = n_views > ctx->n_slots ? n_views : ctx->n_slots;  // per set; GG_NSETS sets are allocated
    const int64_t slot_tiles = tiles > ctx->slot_tiles ? tiles : ctx->slot_tiles;
    const size_t b_vis = align_up((size_t)ctx->n_blocks * 4, 256);
    const size_t b_rec = align_up((size_t)cap_recs * sizeof(GGFaceRec), 256);
    const size_t b_cnt = align_up((size_t)slot_tiles * 4, 256);
    const size_t b_off = align_up((size_t)slot_tiles * 4, 256);
    const size_t b_bin = align_up((size_t)cap_bins * sizeof(GGTileFace), 256);
    const size_t b_ctr = 256;
    const size_t per_slot = b_vis + b_rec + b_cnt + b_off + b_bin + b_ctr;
    GG_CUDA(cudaMalloc(&ctx->d_scratch, per_slot * slots * GG_NSETS));
    ctx->scratch_bytes = per_slot * slots * GG_NSETS;
    for (int s = 0; s < GG_NSETS * slots; ++s) {
        char *p = ctx->d_scratch + per_slot * s;
        GGViewScratch &v = ctx->vset[s / slots].v[s % slots];
        v.vis_blocks = (int32_t *)p;
        p += b_vis;
        v.recs = (GGFaceRec *)p;
        p += b_rec;
        v.tile_count = (int32_t *)p;
        p += b_cnt;
        v.tile_offset = (int32_t *)p;
        p += b_off;
        v.bins = (GGTileFace *)p;
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

template <typename T>
static int launch_dense(gg_context *ctx, const GGCamBatch &cb, dim3 rgrid, int n_tiles, int32_t *d_pix2face,
                        const GGDenseArgs &da, cudaStream_t st) {
    const size_t dyn = (size_t)GG_RASTER_WARPS * GG_DENSE_ACC_FLOATS * sizeof(float);
#define GG_DENSE_CASE(CT)                                                                                             \
    case CT:                                                                                                          \
        GG_LAUNCH(ctx, GG_ST_RASTER, st,                                                                              \
                  (k_raster_tiles<GG_RM_DENSE, T, CT><<<rgrid, GG_RASTER_THREADS, dyn, st>>>(cb, ctx->vset[ctx->cur], n_tiles,   \
                                                                                            d_pix2face, nullptr, 0, da))); \
        break;
    switch (da.index_kind ? 0 : da.C) {  // compile-time channel counts for the common cases, generic otherwise
        GG_DENSE_CASE(3)
        GG_DENSE_CASE(4)
        GG_DENSE_CASE(5)
        GG_DENSE_CASE(8)
        GG_DENSE_CASE(10)
        GG_DENSE_CASE(16)
        default:
            GG_LAUNCH(ctx, GG_ST_RASTER, st,
                      (k_raster_tiles<GG_RM_DENSE, T, 0><<<rgrid, GG_RASTER_THREADS, dyn, st>>>(cb, ctx->vset[ctx->cur], n_tiles,
                                                                                               d_pix2face, nullptr, 0, da)));
    }
#undef GG_DENSE_CASE
    return GG_OK;
}

int gg_pipeline_drain(gg_context *ctx, cudaStream_t st) {
    for (int s = 0; s < GG_NSETS; ++s) {
        if (ctx->ras_pending[s]) {
            GG_CUDA(cudaStreamWaitEvent(st, ctx->ev_ras[s], 0));
            ctx->ras_pending[s] = false;
        }
    }
    return GG_OK;
}

int gg_launch_rasterize(gg_context *ctx, const gg_camera *cams, int n, int32_t *d_pix2face, float *d_depth,
                        int want_winners, int compat_bg, cudaStream_t st_bin, cudaStream_t st_ras,
                        const void *const *h_pred, int pred_kind, int C, double *d_sum, int32_t *d_count,
                        const double *d_tex, int D, void *d_out, int out_dtype) {
    const int W = cams[0].W, H = cams[0].H;
    int rc = gg_ensure_scratch(ctx, n, W, H);
    if (rc != GG_OK) return rc;
    const bool piped = st_bin != st_ras;
    if (!piped) ctx->cur = 0;
    GGCamBatch cb;
    for (int i = 0; i < n; ++i) cb.cam[i] = cams[i];
    const int tiles_x = (W + GG_TILE_W - 1) / GG_TILE_W, tiles_y = (H + GG_TILE_H - 1) / GG_TILE_H;
    const int n_tiles = tiles_x * tiles_y;
    ctx->last_n_tiles = n_tiles;
    ctx->last_batch_n = n;
    if (want_winners) {
        if (ctx->wdense_cap < (int64_t)n * ctx->F) {
            if (ctx->d_wdense) {
                GG_CUDA(cudaDeviceSynchronize());
                GG_CUDA(cudaFree(ctx->d_wdense));
                ctx->d_wdense = nullptr;
            }
            GG_CUDA(cudaMalloc(&ctx->d_wdense, (size_t)GG_NSETS * n * ctx->F * 4));
            // -1 = "not seen".  Set once: every consumer of a batch's winners (gg_launch_resolve_batch,
            // gg_launch_compact_winners) puts back the entries the batch touched, which is ~1 % of a full clear.
            GG_CUDA(cudaMemset(ctx->d_wdense, 0xFF, (size_t)GG_NSETS * n * ctx->F * 4));
            ctx->wdense_cap = (int64_t)n * ctx->F;
        }
        for (int i = 0; i < n; ++i)
            ctx->vset[ctx->cur].v[i].winner = ctx->d_wdense + ((int64_t)ctx->cur * (ctx->wdense_cap / ctx->F) + i) * ctx->F;
    }
    cudaStream_t st = st_bin;
    GG_LAUNCH(ctx, GG_ST_MISC, st,
              k_reset_views<<<dim3((n_tiles + 1023) / 1024, n), 256, 0, st>>>(n_tiles, ctx->vset[ctx->cur]));
    const int nb = (int)ctx->n_blocks;
    GG_LAUNCH(ctx, GG_ST_CULL, st,
              k_cull_blocks<<<dim3((nb + 255) / 256, n), 256, 0, st>>>(ctx->d_block_lo, ctx->d_block_hi, nb, cb, ctx->vset[ctx->cur]));
    // Grid sizes of the two binning kernels: 8 / 4 CTAs per SM and view.  GG_SETUP_CTAS / GG_FILL_CTAS (CTAs per SM
    // over the whole batch) shrink them to a small resident set that runs beside the previous batch's rasterizer
    // instead of displacing its CTAs: +2.6 % on c2 at 2 / 2, but -5 % on c5, whose binning then becomes the critical
    // path -- hence not the default.
    const int setup_ctas = ctx->setup_ctas, fill_ctas = ctx->fill_ctas;
    int gsetup = nb < ctx->sm_count * 8 ? nb : ctx->sm_count * 8;
    if (setup_ctas > 0) gsetup = std::max(1, std::min(gsetup, ctx->sm_count * setup_ctas / n));
    const int gfill = fill_ctas > 0 ? std::max(1, ctx->sm_count * fill_ctas / n) : ctx->sm_count * 4;
    GG_LAUNCH(ctx, GG_ST_SETUP, st,
              k_setup_faces<<<dim3(gsetup, n), GG_BLOCK_FACES, 0, st>>>(ctx->d_verts, ctx->d_faces, ctx->F, ctx->cap_recs,
                                                                        ctx->d_sticky, cb, ctx->vset[ctx->cur]));
    GG_LAUNCH(ctx, GG_ST_SCAN, st,
              k_reserve_tiles<<<dim3((n_tiles + 255) / 256, n), 256, 0, st>>>(n_tiles, ctx->cap_recs, ctx->d_sticky, ctx->vset[ctx->cur]));
    GG_LAUNCH(ctx, GG_ST_FILL, st, k_fill_bins<<<dim3(gfill, n), 256, 0, st>>>(ctx->cap_bins, ctx->d_sticky, cb, ctx->vset[ctx->cur]));
    if (piped) {
        GG_CUDA(cudaEventRecord(ctx->ev_bin[ctx->cur], st_bin));
        GG_CUDA(cudaStreamWaitEvent(st_ras, ctx->ev_bin[ctx->cur], 0));
    }
    st = st_ras;
    const int per_cta_d = GG_RASTER_WARPS * GG_DENSE_TILES_PER_WARP;
    const dim3 rgrid((tiles_x + per_cta_d - 1) / per_cta_d, tiles_y, n);  // dense mode
    const int per_cta = GG_RASTER_WARPS * GG_TILES_PER_WARP;
    const dim3 rgrid_t((tiles_x + per_cta - 1) / per_cta, tiles_y, n);
    GGDenseArgs da;
    memset(&da, 0, sizeof(da));
    if (h_pred) {  // fused dense per-pixel sums
        for (int i = 0; i < n; ++i) da.preds.p[i] = h_pred[i];
        da.sum = d_sum;
        da.count = d_count;
        da.C = C;
        da.index_kind = pred_kind == GG_PRED_INDEX_U8;
        da.vec_ok = 1;
        for (int i = 0; i < n; ++i) da.vec_ok &= ((uintptr_t)h_pred[i] % 16) == 0;
        const size_t elem = pred_kind == GG_PRED_F64 ? 8 : (pred_kind == GG_PRED_F32 ? 4 : 1);
        da.l2_prefetch = ctx->dense_prefetch && da.vec_ok && !da.index_kind && ((size_t)W * C * elem) % 16 == 0;
        for (int i = 0; i < n && da.l2_prefetch; ++i) {  // only device memory is prefetched into L2
            cudaPointerAttributes attr;
            da.l2_prefetch = cudaPointerGetAttributes(&attr, h_pred[i]) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
        }
        (void)cudaGetLastError();
        switch (pred_kind) {
            case GG_PRED_F32: return launch_dense<float>(ctx, cb, rgrid, n_tiles, d_pix2face, da, st);
            case GG_PRED_F64: return launch_dense<double>(ctx, cb, rgrid, n_tiles, d_pix2face, da, st);
            case GG_PRED_U8:
            case GG_PRED_INDEX_U8: return launch_dense<uint8_t>(ctx, cb, rgrid, n_tiles, d_pix2face, da, st);
            default: gg_set_error("bad pred_kind"); return GG_ERR_INVALID;
        }
    }
    if (d_tex) {  // fused render_flat
        da.tex = d_tex;
        da.out = d_out;
        da.D = D;
        // (GG_LANE_LISTS picks the kernel variant whose lanes walk their own face lists; see k_raster_tiles)
#define GG_RASTER_LAUNCH(MODE_, T_, ...)                                                                               \
    do {                                                                                                                \
        if (ctx->lane_lists)                                                                                            \
            GG_LAUNCH(ctx, GG_ST_RASTER, st,                                                                            \
                      (k_raster_tiles<MODE_, T_, 0, true><<<rgrid_t, GG_RASTER_THREADS, 0, st>>>(__VA_ARGS__)));          \
        else                                                                                                            \
            GG_LAUNCH(ctx, GG_ST_RASTER, st,                                                                            \
                      (k_raster_tiles<MODE_, T_, 0, false><<<rgrid_t, GG_RASTER_THREADS, 0, st>>>(__VA_ARGS__)));         \
    } while (0)
        switch (out_dtype) {
            case GG_OUT_F64:
                GG_RASTER_LAUNCH(GG_RM_GATHER, double, cb, ctx->vset[ctx->cur], n_tiles, d_pix2face, nullptr, 0, da);
                break;
            case GG_OUT_F32:
                GG_RASTER_LAUNCH(GG_RM_GATHER, float, cb, ctx->vset[ctx->cur], n_tiles, d_pix2face, nullptr, 0, da);
                break;
            case GG_OUT_U8:
                GG_RASTER_LAUNCH(GG_RM_GATHER, uint8_t, cb, ctx->vset[ctx->cur], n_tiles, d_pix2face, nullptr, 0, da);
                break;
            default: gg_set_error("bad out_dtype"); return GG_ERR_INVALID;
        }
        return GG_OK;
    }
    if (want_winners && !d_pix2face && !d_depth)  // the fused aggregation: no raster leaves the SM
        GG_RASTER_LAUNCH(GG_RM_WINNERS_ONLY, float, cb, ctx->vset[ctx->cur], n_tiles, nullptr, nullptr,
                         compat_bg ? (int)ctx->F : 0, da);
    else if (want_winners)
        GG_RASTER_LAUNCH(GG_RM_WINNERS, float, cb, ctx->vset[ctx->cur], n_tiles, d_pix2face, d_depth,
                         compat_bg ? (int)ctx->F : 0, da);
    else
        GG_RASTER_LAUNCH(GG_RM_PLAIN, float, cb, ctx->vset[ctx->cur], n_tiles, d_pix2face, d_depth, 0, da);
#undef GG_RASTER_LAUNCH
    return GG_OK;
}
