/*
 * oracle/oracle_raycast.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A SECOND, INDEPENDENT statement of what TexturedPhotogrammetryMesh.pix2face computes
 * (/root/reference/geograypher/meshes/meshes.py:1678-1718: "for each pixel, the ID of the nearest mesh face
 * along the pixel's ray, -1 if none"), used only to cross-check oracle_raster.c (tests/test_oracle_raycast.py).
 * It shares NOTHING with the rasterization contract of oracle_raster.c: no screen-space snapping, no edge
 * functions, no fill rule, no interpolated depth.  Per pixel it casts the ray of the reference's pinhole model
 *
 *     d = ((j + 0.5 - px) / f, (i + 0.5 - py) / f, 1)        px = W/2 + cx, py = H/2 + cy
 *
 * (cameras/cameras.py:588-592, derived_meshes.py:772-780; pixel (row i, col j), centre at +0.5; camera frame +X
 * right, +Y down, +Z forward) from the camera centre and intersects it with every candidate triangle in CAMERA
 * SPACE with the Moeller-Trumbore test in float64.  The nearest hit with z_cam >= znear wins.
 *
 * The two oracles can only be compared where the answer does not hinge on sub-pixel conventions, so this file
 * also says where that is: besides the centre ray it casts four rays through the corners of a square of half-size
 * `eps_edge` pixels around the centre; a pixel is EDGE-SAFE iff all five rays see the same face (no silhouette or
 * shared edge passes within eps_edge of the sample), and its DEPTH MARGIN is (w1 - w2) / w1 for the two nearest
 * hits of the centre ray, w = 1 / z_cam (1 when there is no second hit).
 *
 * Inputs are the float32 vertices and the float32 camera record the rasterizers receive, widened to float64; all
 * arithmetic is float64.  Build: oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float m[12];
    float f;
    float px, py;
    int32_t W, H;
    float znear;
} orc_camera; /* same layout as ora_camera / gg_camera */

typedef struct {
    double p0[3], e1[3], e2[3]; /* camera-space vertex 0 and the two edges from it */
    int32_t j0, j1, i0, i1;     /* candidate pixel box (inclusive), conservative */
    int32_t face;
} orc_tri;

static inline void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* Moeller-Trumbore, ray origin at the camera centre (0,0,0), direction d (d_z = 1, so t IS z_cam).
 * Returns 1 and *t when the ray hits the triangle (boundary included). */
static inline int ray_tri(const orc_tri *T, const double *d, double *t) {
    double h[3], q[3], s[3];
    cross3(d, T->e2, h);
    const double a = dot3(T->e1, h);
    if (a == 0.0 || !isfinite(a)) return 0; /* ray parallel to the triangle's plane, or degenerate triangle */
    const double inv = 1.0 / a;
    s[0] = -T->p0[0];
    s[1] = -T->p0[1];
    s[2] = -T->p0[2];
    const double u = inv * dot3(s, h);
    if (!(u >= 0.0 && u <= 1.0)) return 0;
    cross3(s, T->e1, q);
    const double v = inv * dot3(d, q);
    if (!(v >= 0.0 && u + v <= 1.0)) return 0;
    *t = inv * dot3(T->e2, q);
    return isfinite(*t);
}

/*
 * pix2face  : H*W int32   face of the centre ray, -1 = none
 * edge_safe : H*W uint8   1 iff the centre ray and the four corner rays (+-eps_edge px) see the same face
 * margin    : H*W double  (w1 - w2) / w1 of the centre ray's two nearest hits on DIFFERENT faces
 */
void orc_raycast(const float *verts, int64_t V, const int32_t *faces, int64_t F, const orc_camera *cam, double eps_edge,
                 int32_t *pix2face, uint8_t *edge_safe, double *margin, int nthreads) {
    const int W = cam->W, H = cam->H;
    const double f = cam->f, px = cam->px, py = cam->py, znear = cam->znear;
    double *PC = (double *)malloc(sizeof(double) * 3 * (size_t)(V > 0 ? V : 1));
    for (int64_t i = 0; i < V; ++i) {
        const double x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
        for (int r = 0; r < 3; ++r)
            PC[3 * i + r] = (double)cam->m[4 * r] * x + (double)cam->m[4 * r + 1] * y + (double)cam->m[4 * r + 2] * z +
                            (double)cam->m[4 * r + 3];
    }
    orc_tri *tris = (orc_tri *)malloc(sizeof(orc_tri) * (size_t)(F > 0 ? F : 1));
    int64_t nt = 0;
    for (int64_t fi = 0; fi < F; ++fi) {
        const int32_t a = faces[3 * fi], b = faces[3 * fi + 1], c = faces[3 * fi + 2];
        if (a < 0 || b < 0 || c < 0 || a >= V || b >= V || c >= V) continue;
        const double *p[3] = {PC + 3 * (size_t)a, PC + 3 * (size_t)b, PC + 3 * (size_t)c};
        int finite = 1, behind = 0, front = 0;
        for (int k = 0; k < 3; ++k) {
            finite = finite && isfinite(p[k][0]) && isfinite(p[k][1]) && isfinite(p[k][2]);
            if (p[k][2] < znear) behind++;
            else front++;
        }
        if (!finite || front == 0) continue;
        orc_tri *T = &tris[nt];
        for (int k = 0; k < 3; ++k) {
            T->p0[k] = p[0][k];
            T->e1[k] = p[1][k] - p[0][k];
            T->e2[k] = p[2][k] - p[0][k];
        }
        T->face = (int32_t)fi;
        if (behind > 0) { /* crosses the near plane: its projection is unbounded, look at every pixel */
            T->j0 = 0;
            T->j1 = W - 1;
            T->i0 = 0;
            T->i1 = H - 1;
        } else {
            double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
            for (int k = 0; k < 3; ++k) {
                const double sx = f * p[k][0] / p[k][2] + px, sy = f * p[k][1] / p[k][2] + py;
                if (sx < xmin) xmin = sx;
                if (sx > xmax) xmax = sx;
                if (sy < ymin) ymin = sy;
                if (sy > ymax) ymax = sy;
            }
            const double pad = 1.5 + eps_edge; /* pixel centres within the projected box, generously */
            double j0 = floor(xmin - pad), j1 = ceil(xmax + pad), i0 = floor(ymin - pad), i1 = ceil(ymax + pad);
            if (j0 < 0) j0 = 0;
            if (i0 < 0) i0 = 0;
            if (j1 > W - 1) j1 = W - 1;
            if (i1 > H - 1) i1 = H - 1;
            if (!(j0 <= j1 && i0 <= i1)) continue;
            T->j0 = (int32_t)j0;
            T->j1 = (int32_t)j1;
            T->i0 = (int32_t)i0;
            T->i1 = (int32_t)i1;
        }
        nt++;
    }

    const size_t P = (size_t)W * (size_t)H;
    /* per pixel: five rays (0 = centre, 1..4 = corners): nearest depth and face; second-nearest depth of the centre */
    double *zbest = (double *)malloc(sizeof(double) * P * 5);
    int32_t *fbest = (int32_t *)malloc(sizeof(int32_t) * P * 5);
    double *zsecond = (double *)malloc(sizeof(double) * P);
    for (size_t p = 0; p < P * 5; ++p) {
        zbest[p] = INFINITY;
        fbest[p] = -1;
    }
    for (size_t p = 0; p < P; ++p) zsecond[p] = INFINITY;
    static const double ox[5] = {0, -1, 1, -1, 1}, oy[5] = {0, -1, -1, 1, 1};

#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    int nbands = nthreads * 4;
    if (nbands > H) nbands = H;
    if (nbands < 1) nbands = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int band = 0; band < nbands; ++band) {
        const int r0 = (int)((int64_t)H * band / nbands), r1 = (int)((int64_t)H * (band + 1) / nbands);
        for (int64_t q = 0; q < nt; ++q) { /* increasing face ID: strictly nearer replaces = lowest ID on exact ties */
            const orc_tri *T = &tris[q];
            const int i0 = T->i0 > r0 ? T->i0 : r0, i1 = T->i1 < r1 - 1 ? T->i1 : r1 - 1;
            for (int i = i0; i <= i1; ++i) {
                for (int j = T->j0; j <= T->j1; ++j) {
                    const size_t p = (size_t)i * (size_t)W + (size_t)j;
                    for (int r = 0; r < 5; ++r) {
                        const double d[3] = {(j + 0.5 + ox[r] * eps_edge - px) / f, (i + 0.5 + oy[r] * eps_edge - py) / f, 1.0};
                        double t;
                        if (!ray_tri(T, d, &t) || !(t >= znear)) continue;
                        if (t < zbest[5 * p + r]) {
                            if (r == 0 && fbest[5 * p] != T->face) zsecond[p] = zbest[5 * p];
                            zbest[5 * p + r] = t;
                            fbest[5 * p + r] = T->face;
                        } else if (r == 0 && fbest[5 * p] != T->face && t < zsecond[p]) {
                            zsecond[p] = t;
                        }
                    }
                }
            }
        }
    }
    for (size_t p = 0; p < P; ++p) {
        pix2face[p] = fbest[5 * p];
        uint8_t same = 1;
        for (int r = 1; r < 5; ++r) same = same && (fbest[5 * p + r] == fbest[5 * p]);
        edge_safe[p] = same;
        if (fbest[5 * p] >= 0 && isfinite(zsecond[p])) {
            const double w1 = 1.0 / zbest[5 * p], w2 = 1.0 / zsecond[p];
            margin[p] = (w1 - w2) / w1;
        } else {
            margin[p] = 1.0;
        }
    }
    free(zsecond);
    free(fbest);
    free(zbest);
    free(tris);
    free(PC);
}
