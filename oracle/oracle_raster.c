/*
 * oracle/oracle_raster.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the visibility ("pix2face") step of geograypher's multiview
 * projection path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path never does.
 *
 * What it restates (file:line under /root/reference):
 *   - TexturedPhotogrammetryMesh.pix2face            geograypher/meshes/meshes.py:1678-1856
 *     "for each pixel, the ID of the nearest mesh face along the ray, -1 if none".
 *   - the pinhole model of the PyTorch3D back-end    geograypher/meshes/derived_meshes.py:739-794
 *     x = f*Xc/Zc + W/2 + cx ,  y = f*Yc/Zc + H/2 + cy   (camera frame +X right, +Y down, +Z fwd;
 *     geograypher/cameras/cameras.py:446-477 uses the same frame: up = -Y, look = +Z).
 *   - float32 vertex / camera data inside the rasterizer (derived_meshes.py:631, 761-764).
 *
 * The arithmetic of the rasterization itself lives in un-vendored third-party code
 * (VTK 9.2.6 OpenGL through pyvista 0.42.2, poetry.lock; or PyTorch3D MeshRasterizer), which is
 * absent from /root/reference and cannot be run in the build container.  PARITY UNPINNED at the
 * sub-pixel level: this file restates the published algorithm both share (sample at pixel
 * centres, nearest positive depth wins) and fixes the unspecified parts in an explicit
 * contract (DESIGN.md "Rasterization contract"):
 *   C1  projection in IEEE float32, round-to-nearest, fixed operation order, no FMA contraction
 *   C2  screen coordinates snapped to 1/256 px fixed point (GPU-style 8 sub-pixel bits)
 *   C3  coverage = exact integer edge functions, pixel centre (j+0.5, i+0.5), top-left fill rule
 *   C4  depth = 1/z_cam interpolated linearly in screen space (perspective-correct), nearest wins,
 *       exact ties -> lowest face ID
 *   C5  faces are clipped against the plane z_cam = znear in camera space (float32, fixed operation order); a face
 *       entirely behind it, or with a non-finite coordinate, is dropped
 *   C6  guard band: a (sub-)triangle with a vertex that projects beyond +-2^20 px is clipped, in camera space and in
 *       float32 with a fixed operation order, against the four planes  sx = +-2^20, sy = +-2^20  (Sutherland-Hodgman,
 *       intersections always computed from the inside vertex towards the outside one) and fan-triangulated, so that
 *       the snap of C2 never has to clamp a coordinate (clamping x and y independently would turn the edge)
 * It is pinned against the reference's own known-answer tests for this path
 * (tests/test_derived_meshes.py:23-76, tests/test_derived_cameras.py:339-415) in
 * tests/test_oracle_reference_pins.py.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off is REQUIRED: contract C1 forbids fused multiply-add in the projection.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_SUBPIX 256
#define ORA_HALF 128
#define ORA_CLAMP 536870912.0f /* 2^29 sub-pixel units */

typedef struct {
    float m[12];  /* world_to_cam rows 0..2, row-major 3x4 */
    float f;      /* focal length, pixels (already scaled for render_img_scale) */
    float px, py; /* principal point in pixels: W/2 + cx, H/2 + cy */
    int32_t W, H; /* raster size */
    float znear;  /* contract C5 */
} ora_camera;

/* Contract C1, first half: camera-space coordinates.  Relies on -ffp-contract=off and SSE float math. */
static inline void cam_space(const float *v, const ora_camera *c, float *out) {
    const float x = v[0], y = v[1], z = v[2];
    const float *m = c->m;
    float t;
    t = m[0] * x;
    t = t + m[1] * y;
    t = t + m[2] * z;
    out[0] = t + m[3];
    t = m[4] * x;
    t = t + m[5] * y;
    t = t + m[6] * z;
    out[1] = t + m[7];
    t = m[8] * x;
    t = t + m[9] * y;
    t = t + m[10] * z;
    out[2] = t + m[11];
}

/* Contract C1 second half + C2: perspective division and snapping of a camera-space point with zc >= znear. */
static inline int to_screen(const float *pc, const ora_camera *c, int32_t *X, int32_t *Y, float *invz) {
    const float xc = pc[0], yc = pc[1], zc = pc[2];
    *X = 0;
    *Y = 0;
    *invz = 0.0f;
    if (!(zc >= c->znear) || !isfinite(xc) || !isfinite(yc) || !isfinite(zc)) return 0;
    float sx = (c->f * xc) / zc + c->px;
    float sy = (c->f * yc) / zc + c->py;
    float fx = nearbyintf(sx * (float)ORA_SUBPIX); /* round-half-even */
    float fy = nearbyintf(sy * (float)ORA_SUBPIX);
    if (!isfinite(fx) || !isfinite(fy)) return 0;
    if (fx > ORA_CLAMP) fx = ORA_CLAMP;
    if (fx < -ORA_CLAMP) fx = -ORA_CLAMP;
    if (fy > ORA_CLAMP) fy = ORA_CLAMP;
    if (fy < -ORA_CLAMP) fy = -ORA_CLAMP;
    *X = (int32_t)fx;
    *Y = (int32_t)fy;
    *invz = 1.0f / zc;
    return 1;
}

static inline int project_vertex(const float *v, const ora_camera *c, int32_t *X, int32_t *Y, float *invz) {
    float pc[3];
    cam_space(v, c, pc);
    return to_screen(pc, c, X, Y, invz);
}

/* Contract C5: intersection of the edge from P (in front, zp >= znear) to Q (behind) with the plane z = znear. */
static inline void clip_edge(const float *P, const float *Q, float znear, float *R) {
    const float t = (P[2] - znear) / (P[2] - Q[2]);
    float d;
    d = Q[0] - P[0];
    R[0] = P[0] + t * d;
    d = Q[1] - P[1];
    R[1] = P[1] + t * d;
    R[2] = znear;
}

/* Stage 1 on its own: project every vertex (used to check the CUDA projection bit for bit). */
void ora_project(const float *verts, int64_t V, const ora_camera *cam, int32_t *X, int32_t *Y,
                 float *invz, uint8_t *valid) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < V; ++i) {
        valid[i] = (uint8_t)project_vertex(verts + 3 * i, cam, X + i, Y + i, invz + i);
    }
}

static inline int64_t floordiv(int64_t a, int64_t b) {
    int64_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

/* Contract C3 tie rule: a sample exactly on an edge (E == 0) belongs to the triangle iff the edge is a
 * "left" edge (A > 0: interior lies towards +x) or a "top" edge (A == 0 and B > 0: horizontal edge with
 * the interior towards +y, i.e. below it on a y-down screen).  (A,B) = gradient of E. */
static inline int edge_inclusive(int64_t A, int64_t B) { return (A > 0) || (A == 0 && B > 0); }

/* Rasterize one (sub-)triangle given in camera space (all three points in front of the near plane) into the
 * row band [r0, r1).  Strictly nearer replaces: with faces visited in increasing ID this is the lowest-ID tie-break. */
static void raster_tri(const float *pa, const float *pb, const float *pc, int32_t face, const ora_camera *cam, int r0,
                       int r1, int32_t *pix2face, double *wbest, double *wsecond) {
    const int W = cam->W;
    int32_t X[3], Y[3];
    float IZ[3];
    if (!to_screen(pa, cam, &X[0], &Y[0], &IZ[0]) || !to_screen(pb, cam, &X[1], &Y[1], &IZ[1]) ||
        !to_screen(pc, cam, &X[2], &Y[2], &IZ[2]))
        return;
    int64_t x0 = X[0], y0 = Y[0], x1 = X[1], y1 = Y[1], x2 = X[2], y2 = Y[2];
    double w0 = IZ[0], w1 = IZ[1], w2 = IZ[2];
    /* bounding box in fixed point -> pixel-centre index range */
    int64_t xmin = x0 < x1 ? x0 : x1;
    if (x2 < xmin) xmin = x2;
    int64_t xmax = x0 > x1 ? x0 : x1;
    if (x2 > xmax) xmax = x2;
    int64_t ymin = y0 < y1 ? y0 : y1;
    if (y2 < ymin) ymin = y2;
    int64_t ymax = y0 > y1 ? y0 : y1;
    if (y2 > ymax) ymax = y2;
    /* centres c = 256*j + 128 with xmin <= c <= xmax */
    int64_t jmin = floordiv(xmin - ORA_HALF + ORA_SUBPIX - 1, ORA_SUBPIX); /* ceil */
    int64_t jmax = floordiv(xmax - ORA_HALF, ORA_SUBPIX);
    int64_t imin = floordiv(ymin - ORA_HALF + ORA_SUBPIX - 1, ORA_SUBPIX);
    int64_t imax = floordiv(ymax - ORA_HALF, ORA_SUBPIX);
    if (jmin < 0) jmin = 0;
    if (jmax > W - 1) jmax = W - 1;
    if (imin < r0) imin = r0;
    if (imax > r1 - 1) imax = r1 - 1;
    if (jmin > jmax || imin > imax) return;
    int64_t area2 = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    if (area2 == 0) return;
    if (area2 < 0) { /* make the interior the positive side */
        int64_t tx = x1, ty = y1;
        double tw = w1;
        x1 = x2;
        y1 = y2;
        w1 = w2;
        x2 = tx;
        y2 = ty;
        w2 = tw;
        area2 = -area2;
    }
    /* E_k(P) = A_k*Px + B_k*Py + C_k ; edge k runs v_k -> v_{k+1} */
    const int64_t A0 = -(y1 - y0), B0 = (x1 - x0);
    const int64_t A1 = -(y2 - y1), B1 = (x2 - x1);
    const int64_t A2 = -(y0 - y2), B2 = (x0 - x2);
    const int inc0 = edge_inclusive(A0, B0), inc1 = edge_inclusive(A1, B1), inc2 = edge_inclusive(A2, B2);
    const double inv_area = 1.0 / (double)area2;
    const int64_t dE0 = A0 * ORA_SUBPIX, dE1 = A1 * ORA_SUBPIX, dE2 = A2 * ORA_SUBPIX; /* step per pixel along a row */
    for (int64_t i = imin; i <= imax; ++i) {
        const int64_t Py = ORA_SUBPIX * i + ORA_HALF;
        const int64_t Px0 = ORA_SUBPIX * jmin + ORA_HALF;
        int64_t E0 = B0 * (Py - y0) + A0 * (Px0 - x0); /* exact integers: stepping them is exact too */
        int64_t E1 = B1 * (Py - y1) + A1 * (Px0 - x1);
        int64_t E2 = B2 * (Py - y2) + A2 * (Px0 - x2);
        for (int64_t j = jmin; j <= jmax; ++j, E0 += dE0, E1 += dE1, E2 += dE2) {
            if ((E0 | E1 | E2) < 0) continue;
            if ((E0 == 0 && !inc0) || (E1 == 0 && !inc1) || (E2 == 0 && !inc2)) continue;
            /* barycentric weights: E1 -> v0, E2 -> v1, E0 -> v2 */
            const double w = ((double)E1 * w0 + (double)E2 * w1 + (double)E0 * w2) * inv_area;
            const size_t p = (size_t)i * (size_t)W + (size_t)j;
            if (w > wbest[p]) {
                if (wsecond && pix2face[p] != face) wsecond[p] = wbest[p];
                wbest[p] = w;
                pix2face[p] = face;
            } else if (wsecond && pix2face[p] != face && w > wsecond[p]) {
                wsecond[p] = w;
            }
        }
    }
}


/* ---- contract C6: guard-band clipping ------------------------------------------------------------------- */
#define ORA_GUARD 1048576.0f /* 2^20 px: half of what the snapped coordinates can hold */

/* float32 screen coordinates of C1 (before the snap) beyond the guard band? */
static inline int beyond_guard(const float *pc, const ora_camera *c) {
    const float sx = (c->f * pc[0]) / pc[2] + c->px;
    const float sy = (c->f * pc[1]) / pc[2] + c->py;
    return !(fabsf(sx) <= ORA_GUARD) || !(fabsf(sy) <= ORA_GUARD);
}

/* signed distance (>= 0: inside) of a camera-space point to guard plane k: 0 right, 1 left, 2 bottom, 3 top */
static inline float guard_dist(const float *p, const ora_camera *c, int k) {
    float a, b;
    switch (k) {
        case 0: a = (ORA_GUARD - c->px) * p[2]; b = c->f * p[0]; return a - b;
        case 1: a = (ORA_GUARD + c->px) * p[2]; b = c->f * p[0]; return a + b;
        case 2: a = (ORA_GUARD - c->py) * p[2]; b = c->f * p[1]; return a - b;
        default: a = (ORA_GUARD + c->py) * p[2]; b = c->f * p[1]; return a + b;
    }
}

/* Clip the triangle (a, b, c) against the four guard planes; out receives at most 8 points; returns their number. */
static int guard_clip(const float *pa, const float *pb, const float *pc, const ora_camera *cam, float out[8][3]) {
    float buf[2][8][3];
    int n = 3, cur = 0;
    memcpy(buf[0][0], pa, 12);
    memcpy(buf[0][1], pb, 12);
    memcpy(buf[0][2], pc, 12);
    for (int k = 0; k < 4 && n >= 3; ++k) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const float *P = buf[cur][i], *Q = buf[cur][(i + 1) % n];
            const float dP = guard_dist(P, cam, k), dQ = guard_dist(Q, cam, k);
            const int inP = dP >= 0.0f, inQ = dQ >= 0.0f;
            if (inP && m < 8) memcpy(buf[1 - cur][m++], P, 12);
            if (inP != inQ && m < 8) { /* from the inside vertex towards the outside one */
                const float *I = inP ? P : Q, *O = inP ? Q : P;
                const float dI = inP ? dP : dQ, dO = inP ? dQ : dP;
                const float t = dI / (dI - dO);
                float *R = buf[1 - cur][m++];
                for (int d = 0; d < 3; ++d) {
                    const float delta = O[d] - I[d];
                    R[d] = I[d] + t * delta;
                }
            }
        }
        n = m;
        cur = 1 - cur;
    }
    if (n < 3) return 0;
    memcpy(out, buf[cur], sizeof(float) * 3 * (size_t)n);
    return n;
}

/* raster_tri behind the guard band (C6) */
static void raster_tri_guarded(const float *pa, const float *pb, const float *pc, int32_t face, const ora_camera *cam, int r0,
                               int r1, int32_t *pix2face, double *wbest, double *wsecond) {
    if (!beyond_guard(pa, cam) && !beyond_guard(pb, cam) && !beyond_guard(pc, cam)) {
        raster_tri(pa, pb, pc, face, cam, r0, r1, pix2face, wbest, wsecond);
        return;
    }
    float poly[8][3];
    const int n = guard_clip(pa, pb, pc, cam, poly);
    for (int i = 1; i + 1 < n; ++i) raster_tri(poly[0], poly[i], poly[i + 1], face, cam, r0, r1, pix2face, wbest, wsecond);
}

/* Up to two camera-space sub-triangles of face fi after clipping against z = znear (contract C5).  Returns their
 * number (0: dropped).  tri[t][k] points to 3 floats; clipped points live in buf. */
static int face_subtris(const float *PC, const int32_t *faces, int64_t V, int64_t fi, float znear,
                        const float *tri[2][3], float buf[4][3]) {
    const int32_t idx[3] = {faces[3 * fi], faces[3 * fi + 1], faces[3 * fi + 2]};
    if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0 || idx[0] >= V || idx[1] >= V || idx[2] >= V) return 0;
    const float *p[3] = {PC + 3 * (size_t)idx[0], PC + 3 * (size_t)idx[1], PC + 3 * (size_t)idx[2]};
    int front[3], nfront = 0, finite = 1;
    for (int k = 0; k < 3; ++k) {
        finite = finite && isfinite(p[k][0]) && isfinite(p[k][1]) && isfinite(p[k][2]);
        front[k] = p[k][2] >= znear;
        nfront += front[k];
    }
    if (!finite || nfront == 0) return 0;
    if (nfront == 3) {
        tri[0][0] = p[0];
        tri[0][1] = p[1];
        tri[0][2] = p[2];
        return 1;
    }
    if (nfront == 1) { /* C5: rotate so that the vertex in front comes first: (A, B, C) */
        const int a = front[0] ? 0 : (front[1] ? 1 : 2);
        const float *A = p[a], *B = p[(a + 1) % 3], *C = p[(a + 2) % 3];
        clip_edge(A, B, znear, buf[0]);
        clip_edge(A, C, znear, buf[1]);
        tri[0][0] = A;
        tri[0][1] = buf[0];
        tri[0][2] = buf[1];
        return 1;
    }
    /* two in front: rotate so that the vertex behind comes last: (A, B, C) */
    const int c = !front[0] ? 0 : (!front[1] ? 1 : 2);
    const float *A = p[(c + 1) % 3], *B = p[(c + 2) % 3], *C = p[c];
    clip_edge(B, C, znear, buf[2]);
    clip_edge(A, C, znear, buf[3]);
    tri[0][0] = A;
    tri[0][1] = B;
    tri[0][2] = buf[2];
    tri[1][0] = A;
    tri[1][1] = buf[2];
    tri[1][2] = buf[3];
    return 2;
}

/* Rows [*i0, *i1] whose pixel centres a sub-triangle can cover (the same arithmetic as raster_tri's bounding box);
 * returns 0 when it cannot cover any pixel centre of the W x H raster. */
static int subtri_rows(const float *pa, const float *pb, const float *pc, const ora_camera *cam, int *i0, int *i1) {
    int32_t X[3], Y[3];
    float IZ[3];
    if (!to_screen(pa, cam, &X[0], &Y[0], &IZ[0]) || !to_screen(pb, cam, &X[1], &Y[1], &IZ[1]) ||
        !to_screen(pc, cam, &X[2], &Y[2], &IZ[2]))
        return 0;
    int64_t xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
    for (int k = 1; k < 3; ++k) {
        if (X[k] < xmin) xmin = X[k];
        if (X[k] > xmax) xmax = X[k];
        if (Y[k] < ymin) ymin = Y[k];
        if (Y[k] > ymax) ymax = Y[k];
    }
    int64_t jmin = floordiv(xmin - ORA_HALF + ORA_SUBPIX - 1, ORA_SUBPIX), jmax = floordiv(xmax - ORA_HALF, ORA_SUBPIX);
    int64_t imin = floordiv(ymin - ORA_HALF + ORA_SUBPIX - 1, ORA_SUBPIX), imax = floordiv(ymax - ORA_HALF, ORA_SUBPIX);
    if (jmin < 0) jmin = 0;
    if (jmax > cam->W - 1) jmax = cam->W - 1;
    if (imin < 0) imin = 0;
    if (imax > cam->H - 1) imax = cam->H - 1;
    if (jmin > jmax || imin > imax) return 0;
    *i0 = (int)imin;
    *i1 = (int)imax;
    return 1;
}

/*
 * Rasterize one view.
 *   pix2face : H*W int32, -1 = no face                                  (required)
 *   depth_w  : H*W double, best 1/z_cam (0 where no face)               (optional)
 *   margin   : H*W double, (w_best - w_second)/w_best, 1 if no runner-up (optional; contract: a pixel is
 *              "depth-safe" iff margin > eps_depth)
 * The raster is cut into row bands (one per task); every band visits, in increasing face ID, the faces whose rows
 * reach into it (a per-view culling pass lists them), so the total work does not grow with the number of threads.
 */
void ora_rasterize(const float *verts, int64_t V, const int32_t *faces, int64_t F, const ora_camera *cam,
                   int32_t *pix2face, double *depth_w, double *margin, int nthreads) {
    const int W = cam->W, H = cam->H;
    float *PC = (float *)malloc(sizeof(float) * 3 * (size_t)(V > 0 ? V : 1)); /* camera-space vertices */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < V; ++i) cam_space(verts + 3 * i, cam, PC + 3 * i);

    const size_t P = (size_t)W * (size_t)H;
    double *wbest = (double *)malloc(sizeof(double) * P);
    double *wsecond = margin ? (double *)malloc(sizeof(double) * P) : NULL;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)P; ++p) {
        pix2face[p] = -1;
        wbest[p] = 0.0;
        if (wsecond) wsecond[p] = 0.0;
    }

#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    int nbands = nthreads * 8;
    if (nbands > H) nbands = H;
    if (nbands < 1) nbands = 1;
    const float znear = cam->znear;

    /* culling pass: first / last band of every face (-1: the face covers no pixel centre) */
    int32_t *b0 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F > 0 ? F : 1));
    int32_t *b1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F > 0 ? F : 1));
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t fi = 0; fi < F; ++fi) {
        const float *tri[2][3];
        float buf[4][3];
        const int nt = face_subtris(PC, faces, V, fi, znear, tri, buf);
        int lo = H, hi = -1;
        for (int t = 0; t < nt; ++t) {
            int i0, i1;
            if (beyond_guard(tri[t][0], cam) || beyond_guard(tri[t][1], cam) || beyond_guard(tri[t][2], cam)) {
                lo = 0; /* clipped against the guard band while rasterizing: any row (rare, conservative) */
                hi = H - 1;
            } else if (subtri_rows(tri[t][0], tri[t][1], tri[t][2], cam, &i0, &i1)) {
                if (i0 < lo) lo = i0;
                if (i1 > hi) hi = i1;
            }
        }
        if (hi < 0) {
            b0[fi] = b1[fi] = -1;
        } else { /* band of row r = the b with H*b/nbands <= r < H*(b+1)/nbands */
            int ba = (int)(((int64_t)lo * nbands) / H), bb = (int)(((int64_t)hi * nbands) / H);
            while (ba > 0 && (int64_t)H * ba / nbands > lo) --ba;
            while (ba + 1 < nbands && (int64_t)H * (ba + 1) / nbands <= lo) ++ba;
            while (bb > 0 && (int64_t)H * bb / nbands > hi) --bb;
            while (bb + 1 < nbands && (int64_t)H * (bb + 1) / nbands <= hi) ++bb;
            b0[fi] = ba;
            b1[fi] = bb;
        }
    }
    /* per-band face lists in increasing face ID (counting sort; the fill is a single ordered sweep) */
    int64_t *start = (int64_t *)calloc((size_t)nbands + 1, sizeof(int64_t));
    for (int64_t fi = 0; fi < F; ++fi)
        if (b0[fi] >= 0)
            for (int b = b0[fi]; b <= b1[fi]; ++b) start[b + 1] += 1;
    for (int b = 0; b < nbands; ++b) start[b + 1] += start[b];
    int32_t *list = (int32_t *)malloc(sizeof(int32_t) * (size_t)(start[nbands] > 0 ? start[nbands] : 1));
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)nbands);
    for (int b = 0; b < nbands; ++b) cursor[b] = start[b];
    for (int64_t fi = 0; fi < F; ++fi)
        if (b0[fi] >= 0)
            for (int b = b0[fi]; b <= b1[fi]; ++b) list[cursor[b]++] = (int32_t)fi;

#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int band = 0; band < nbands; ++band) {
        const int r0 = (int)((int64_t)H * band / nbands);
        const int r1 = (int)((int64_t)H * (band + 1) / nbands);
        for (int64_t q = start[band]; q < start[band + 1]; ++q) {
            const int64_t fi = list[q];
            const float *tri[2][3];
            float buf[4][3];
            const int nt = face_subtris(PC, faces, V, fi, znear, tri, buf);
            for (int t = 0; t < nt; ++t)
                raster_tri_guarded(tri[t][0], tri[t][1], tri[t][2], (int32_t)fi, cam, r0, r1, pix2face, wbest, wsecond);
        }
    }
    if (depth_w) memcpy(depth_w, wbest, sizeof(double) * P);
    if (margin) {
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < (int64_t)P; ++p)
            margin[p] = (pix2face[p] >= 0) ? (wbest[p] - wsecond[p]) / wbest[p] : 1.0;
    }
    free(cursor);
    free(list);
    free(start);
    free(b0);
    free(b1);
    free(wbest);
    if (wsecond) free(wsecond);
    free(PC);
}

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
