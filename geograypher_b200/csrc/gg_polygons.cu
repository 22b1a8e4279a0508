// label_polygons on the GPU.
//
// Reference: TexturedPhotogrammetryMesh.label_polygons, geograypher/meshes/meshes.py:1141-1306 with the default
// sjoin_overlay=True: every labelled face whose 2-D triangle lies WITHIN a polygon (gpd.sjoin predicate "within",
// :1259-1261) votes for its class with weight  area2D * (area3D / area2D) * face_weighting = area3D * face_weighting
// (:1211-1219, :1273; utils/numeric.py:305-327); the polygon takes the class with the largest summed weight.
//
// One thread per face, brute force over the polygons with a bounding-box reject, then an exact test against the
// polygon's rings in float64: all three vertices inside (even-odd rule over all rings, so holes and multi-part
// polygons work), no proper crossing between a triangle edge and a ring edge, and no ring swallowed by the triangle.
// Contacts of measure zero (a vertex exactly on a polygon edge) are not resolved the way shapely's snapped-precision
// predicates do; the reference has no test that pins them.
#include "gg_internal.cuh"

namespace {

__device__ __forceinline__ double orient(double ax, double ay, double bx, double by, double cx, double cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}

__device__ __forceinline__ bool proper_cross(double ax, double ay, double bx, double by, double cx, double cy, double dx,
                                             double dy) {
    const double o1 = orient(ax, ay, bx, by, cx, cy), o2 = orient(ax, ay, bx, by, dx, dy);
    const double o3 = orient(cx, cy, dx, dy, ax, ay), o4 = orient(cx, cy, dx, dy, bx, by);
    return ((o1 > 0) != (o2 > 0)) && ((o3 > 0) != (o4 > 0)) && o1 != 0 && o2 != 0 && o3 != 0 && o4 != 0;
}

__global__ void __launch_bounds__(128) k_label_polygons(const double *__restrict__ xyz, const double *__restrict__ xy,
                                                        const int32_t *__restrict__ faces, const double *__restrict__ labels,
                                                        const double *__restrict__ face_w, int64_t F,
                                                        const double *__restrict__ pxy, const int32_t *__restrict__ ring_off,
                                                        const int32_t *__restrict__ poly_ring_off,
                                                        const double *__restrict__ poly_bbox, int n_polys, int n_classes,
                                                        double *__restrict__ weights) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double lab = labels[f];
    if (!isfinite(lab)) return;  // meshes.py:1195: only faces with a label take part
    const long long cls = (long long)lab;
    if (cls < 0 || cls >= n_classes) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const double x0 = xy[2 * i0], y0 = xy[2 * i0 + 1], x1 = xy[2 * i1], y1 = xy[2 * i1 + 1], x2 = xy[2 * i2],
                 y2 = xy[2 * i2 + 1];
    const double bxmin = fmin(x0, fmin(x1, x2)), bxmax = fmax(x0, fmax(x1, x2));
    const double bymin = fmin(y0, fmin(y1, y2)), bymax = fmax(y0, fmax(y1, y2));
    double weight = -1.0;  // computed on first use
    for (int p = 0; p < n_polys; ++p) {
        const double *bb = poly_bbox + 4 * p;
        if (bxmin < bb[0] || bymin < bb[1] || bxmax > bb[2] || bymax > bb[3]) continue;
        bool in0 = false, in1 = false, in2 = false, bad = false;
        for (int r = poly_ring_off[p]; r < poly_ring_off[p + 1] && !bad; ++r) {
            const int a = ring_off[r], b = ring_off[r + 1];
            for (int k = a; k < b; ++k) {
                const int kn = (k + 1 < b) ? k + 1 : a;
                const double ex0 = pxy[2 * k], ey0 = pxy[2 * k + 1], ex1 = pxy[2 * kn], ey1 = pxy[2 * kn + 1];
                // crossing number of a horizontal ray towards +x from each vertex
                if ((ey0 > y0) != (ey1 > y0) && x0 < (ex1 - ex0) * (y0 - ey0) / (ey1 - ey0) + ex0) in0 = !in0;
                if ((ey0 > y1) != (ey1 > y1) && x1 < (ex1 - ex0) * (y1 - ey0) / (ey1 - ey0) + ex0) in1 = !in1;
                if ((ey0 > y2) != (ey1 > y2) && x2 < (ex1 - ex0) * (y2 - ey0) / (ey1 - ey0) + ex0) in2 = !in2;
                if (proper_cross(x0, y0, x1, y1, ex0, ey0, ex1, ey1) || proper_cross(x1, y1, x2, y2, ex0, ey0, ex1, ey1) ||
                    proper_cross(x2, y2, x0, y0, ex0, ey0, ex1, ey1)) {
                    bad = true;
                    break;
                }
            }
            if (!bad && b > a) {  // a ring that lies entirely inside the triangle (a hole smaller than the face)
                const double qx = pxy[2 * a], qy = pxy[2 * a + 1];
                const double s0 = orient(x0, y0, x1, y1, qx, qy), s1 = orient(x1, y1, x2, y2, qx, qy),
                             s2 = orient(x2, y2, x0, y0, qx, qy);
                if ((s0 > 0 && s1 > 0 && s2 > 0) || (s0 < 0 && s1 < 0 && s2 < 0)) bad = true;
            }
        }
        if (bad || !(in0 && in1 && in2)) continue;
        if (weight < 0.0) {  // 3-D triangle area (utils/numeric.py:305-327) times the optional face weighting
            const double ax = xyz[3 * i1] - xyz[3 * i0], ay = xyz[3 * i1 + 1] - xyz[3 * i0 + 1], az = xyz[3 * i1 + 2] - xyz[3 * i0 + 2];
            const double bx = xyz[3 * i2] - xyz[3 * i0], by = xyz[3 * i2 + 1] - xyz[3 * i0 + 1], bz = xyz[3 * i2 + 2] - xyz[3 * i0 + 2];
            const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
            weight = 0.5 * sqrt(cx * cx + cy * cy + cz * cz) * (face_w ? face_w[f] : 1.0);
            if (!(weight >= 0.0)) weight = 0.0;
        }
        atomicAdd(&weights[(int64_t)p * n_classes + cls], weight);
    }
}


// ---- sjoin_overlay = False (meshes.py:1263-1276): polygons.overlay(faces, how="identity") splits every face along
// the polygon boundaries and each piece votes with  area2D(face n polygon) * (area3D / area2D)(face) * face_weighting.
// area2D(T n P) is computed without building the pieces: the indicator of a polygon is the signed sum of the
// indicators of the fan triangles (O, a_k, a_k+1) over all its ring edges (exterior rings counter-clockwise, holes
// clockwise -- the binding orients them), so  area(T n P) = sum_k sign_k * area(T n fan_k), and T n fan_k is a
// triangle-triangle intersection: Sutherland-Hodgman with at most 6 output vertices.  O = centroid of T keeps the fan
// triangles that matter well conditioned.  float64 throughout.
__device__ __forceinline__ double clip_tri_tri_area(const double tx[3], const double ty[3], double ax, double ay, double bx,
                                                    double by, double cx, double cy) {
    // subject = (a, b, c), clip = T (counter-clockwise)
    double px[8], py[8], qx[8], qy[8];
    int n = 3;
    px[0] = ax; py[0] = ay; px[1] = bx; py[1] = by; px[2] = cx; py[2] = cy;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const double ex0 = tx[e], ey0 = ty[e], ex1 = tx[(e + 1) % 3], ey1 = ty[(e + 1) % 3];
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const double di = (ex1 - ex0) * (py[i] - ey0) - (ey1 - ey0) * (px[i] - ex0);  // >= 0: inside (left of the edge)
            const double dj = (ex1 - ex0) * (py[j] - ey0) - (ey1 - ey0) * (px[j] - ex0);
            if (di >= 0.0) {
                qx[m] = px[i]; qy[m] = py[i]; ++m;
            }
            if ((di >= 0.0) != (dj >= 0.0)) {
                const double t = di / (di - dj);
                qx[m] = px[i] + t * (px[j] - px[i]); qy[m] = py[i] + t * (py[j] - py[i]); ++m;
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
        if (n < 3) return 0.0;
    }
    double a2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        a2 += px[i] * py[j] - px[j] * py[i];
    }
    return 0.5 * fabs(a2);
}

__global__ void __launch_bounds__(128) k_label_polygons_overlay(
    const double *__restrict__ xyz, const double *__restrict__ xy, const int32_t *__restrict__ faces,
    const double *__restrict__ labels, const double *__restrict__ face_w, int64_t F, const double *__restrict__ pxy,
    const int32_t *__restrict__ ring_off, const int32_t *__restrict__ poly_ring_off, const double *__restrict__ poly_bbox,
    int n_polys, int n_classes, double *__restrict__ weights) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double lab = labels[f];
    if (!isfinite(lab)) return;
    const long long cls = (long long)lab;
    if (cls < 0 || cls >= n_classes) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    double tx[3] = {xy[2 * i0], xy[2 * i1], xy[2 * i2]}, ty[3] = {xy[2 * i0 + 1], xy[2 * i1 + 1], xy[2 * i2 + 1]};
    const double a2 = orient(tx[0], ty[0], tx[1], ty[1], tx[2], ty[2]);
    if (!(a2 != 0.0) || !isfinite(a2)) return;  // no 2-D area: nothing to overlay (the 3D/2D ratio is undefined)
    if (a2 < 0.0) {  // make T counter-clockwise
        const double sx = tx[1], sy = ty[1];
        tx[1] = tx[2]; ty[1] = ty[2]; tx[2] = sx; ty[2] = sy;
    }
    const double area2d = 0.5 * fabs(a2);
    const double bxmin = fmin(tx[0], fmin(tx[1], tx[2])), bxmax = fmax(tx[0], fmax(tx[1], tx[2]));
    const double bymin = fmin(ty[0], fmin(ty[1], ty[2])), bymax = fmax(ty[0], fmax(ty[1], ty[2]));
    const double ox = (tx[0] + tx[1] + tx[2]) / 3.0, oy = (ty[0] + ty[1] + ty[2]) / 3.0;
    double ratio = -1.0;
    for (int p = 0; p < n_polys; ++p) {
        const double *bb = poly_bbox + 4 * p;
        if (bxmax < bb[0] || bymax < bb[1] || bxmin > bb[2] || bymin > bb[3]) continue;
        double inter = 0.0;
        for (int r = poly_ring_off[p]; r < poly_ring_off[p + 1]; ++r) {
            const int a = ring_off[r], b = ring_off[r + 1];
            for (int k = a; k < b; ++k) {
                const int kn = (k + 1 < b) ? k + 1 : a;
                const double ex0 = pxy[2 * k], ey0 = pxy[2 * k + 1], ex1 = pxy[2 * kn], ey1 = pxy[2 * kn + 1];
                const double s = orient(ox, oy, ex0, ey0, ex1, ey1);
                if (s == 0.0) continue;
                // both orientations of the fan triangle are clipped as counter-clockwise subjects
                const double ar = s > 0.0 ? clip_tri_tri_area(tx, ty, ox, oy, ex0, ey0, ex1, ey1)
                                          : clip_tri_tri_area(tx, ty, ox, oy, ex1, ey1, ex0, ey0);
                inter += s > 0.0 ? ar : -ar;
            }
        }
        if (!(inter > 0.0)) continue;
        if (inter > area2d) inter = area2d;
        if (ratio < 0.0) {  // (area3D / area2D) * face weighting, meshes.py:1211-1219
            const double ax = xyz[3 * i1] - xyz[3 * i0], ay = xyz[3 * i1 + 1] - xyz[3 * i0 + 1], az = xyz[3 * i1 + 2] - xyz[3 * i0 + 2];
            const double bx = xyz[3 * i2] - xyz[3 * i0], by = xyz[3 * i2 + 1] - xyz[3 * i0 + 1], bz = xyz[3 * i2 + 2] - xyz[3 * i0 + 2];
            const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
            ratio = 0.5 * sqrt(cx * cx + cy * cy + cz * cz) / area2d * (face_w ? face_w[f] : 1.0);
            if (!(ratio >= 0.0) || !isfinite(ratio)) ratio = 0.0;
        }
        atomicAdd(&weights[(int64_t)p * n_classes + cls], inter * ratio);
    }
}

}  // namespace

extern "C" int gg_label_polygons(int device, const double *d_xyz, const double *d_xy, const int32_t *d_faces,
                                 const double *d_labels, const double *d_face_weight, int64_t F, const double *d_poly_xy,
                                 const int32_t *d_ring_offsets, const int32_t *d_poly_ring_offsets,
                                 const double *d_poly_bbox, int n_polys, int n_classes, double *d_weights, void *stream) {
    if (!d_xyz || !d_xy || !d_faces || !d_labels || !d_poly_xy || !d_ring_offsets || !d_poly_ring_offsets || !d_poly_bbox ||
        !d_weights || F < 1 || n_polys < 1 || n_classes < 1) {
        gg_set_error("gg_label_polygons: bad arguments");
        return GG_ERR_INVALID;
    }
    GG_CUDA(cudaSetDevice(device));
    k_label_polygons<<<(unsigned)((F + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d_xyz, d_xy, d_faces, d_labels, d_face_weight, F, d_poly_xy, d_ring_offsets, d_poly_ring_offsets, d_poly_bbox, n_polys,
        n_classes, d_weights);
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}

extern "C" int gg_label_polygons_overlay(int device, const double *d_xyz, const double *d_xy, const int32_t *d_faces,
                                         const double *d_labels, const double *d_face_weight, int64_t F,
                                         const double *d_poly_xy, const int32_t *d_ring_offsets,
                                         const int32_t *d_poly_ring_offsets, const double *d_poly_bbox, int n_polys,
                                         int n_classes, double *d_weights, void *stream) {
    if (!d_xyz || !d_xy || !d_faces || !d_labels || !d_poly_xy || !d_ring_offsets || !d_poly_ring_offsets || !d_poly_bbox ||
        !d_weights || F < 1 || n_polys < 1 || n_classes < 1) {
        gg_set_error("gg_label_polygons_overlay: bad arguments");
        return GG_ERR_INVALID;
    }
    GG_CUDA(cudaSetDevice(device));
    k_label_polygons_overlay<<<(unsigned)((F + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d_xyz, d_xy, d_faces, d_labels, d_face_weight, F, d_poly_xy, d_ring_offsets, d_poly_ring_offsets, d_poly_bbox, n_polys,
        n_classes, d_weights);
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}
