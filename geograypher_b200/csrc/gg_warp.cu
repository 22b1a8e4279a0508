// Lens-distortion warp of face-ID rasters.
//
// Replaces the reference's post-warp of pix2face (geograypher/meshes/meshes.py:1842-1854 ->
// cameras/cameras.py:1092-1156 warp_dewarp_image -> utils/image.py:72-126 skimage warp) for Metashape's frame
// camera model (cameras/derived_cameras.py:163-208).  Differences, both deliberate:
//   * the ideal<-warped map is the exact inverse of the forward model (Newton, float64) instead of a scipy
//     griddata interpolation of an 8x down-sampled forward map (utils/indexing.py:87-150);
//   * IDs are gathered as integers; the reference normalises them to [0,1] floats and back, which alters some IDs.
// Map conventions follow cameras.py:1027-1053: at scale 1 integer coordinates are pixel positions; at another
// scale s output pixel i stands for the full-resolution coordinate (i + 0.5) / s and warped coordinates are
// multiplied by s.
#include "gg_internal.cuh"

namespace {

struct Dist {
    double f, cx, cy, halfW, halfH, k1, k2, k3, k4, p1, p2, b1, b2, s;
    int one;  // image_scale is (numerically) 1
};

// forward model + Jacobian with respect to the ideal pixel coordinates
__device__ __forceinline__ void ideal_to_warped(const Dist &d, double x, double y, double &xw, double &yw, double *J) {
    const double u = (x - d.halfW) / d.f, v = (y - d.halfH) / d.f;
    const double r2 = u * u + v * v;
    const double R = 1.0 + r2 * (d.k1 + r2 * (d.k2 + r2 * (d.k3 + r2 * d.k4)));
    const double xd = u * R + (d.p1 * (r2 + 2.0 * u * u) + 2.0 * d.p2 * u * v);
    const double yd = v * R + (d.p2 * (r2 + 2.0 * v * v) + 2.0 * d.p1 * u * v);
    xw = d.halfW + d.cx + xd * d.f + xd * d.b1 + yd * d.b2;
    yw = d.halfH + d.cy + yd * d.f;
    if (J) {
        const double Rp = d.k1 + r2 * (2.0 * d.k2 + r2 * (3.0 * d.k3 + r2 * 4.0 * d.k4));
        const double xdu = R + 2.0 * u * u * Rp + 6.0 * d.p1 * u + 2.0 * d.p2 * v;
        const double xdv = 2.0 * u * v * Rp + 2.0 * d.p1 * v + 2.0 * d.p2 * u;
        const double ydu = 2.0 * u * v * Rp + 2.0 * d.p2 * u + 2.0 * d.p1 * v;
        const double ydv = R + 2.0 * v * v * Rp + 6.0 * d.p2 * v + 2.0 * d.p1 * u;
        J[0] = ((d.f + d.b1) * xdu + d.b2 * ydu) / d.f;  // d xw / d x
        J[1] = ((d.f + d.b1) * xdv + d.b2 * ydv) / d.f;  // d xw / d y
        J[2] = ydu;                                       // d yw / d x   (f * ydu / f)
        J[3] = ydv;                                       // d yw / d y
    }
}

__global__ void __launch_bounds__(256) k_build_warp_map(Dist d, int h, int w, int warped_to_ideal,
                                                        int32_t *__restrict__ src_index, float *__restrict__ src_rc) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)h * w) return;
    const int i = (int)(p / w), j = (int)(p - (int64_t)i * w);
    double sr, sc;  // source (row, col) in the reference's map convention
    bool ok = true;
    if (warped_to_ideal) {
        const double x = d.one ? (double)j : ((double)j + 0.5) / d.s, y = d.one ? (double)i : ((double)i + 0.5) / d.s;
        double xw, yw;
        ideal_to_warped(d, x, y, xw, yw, nullptr);
        sc = d.one ? xw : xw * d.s;
        sr = d.one ? yw : yw * d.s;
    } else {
        const double tx = d.one ? (double)j : (double)j / d.s, ty = d.one ? (double)i : (double)i / d.s;
        double x = tx, y = ty;
        ok = false;
        for (int it = 0; it < 30; ++it) {
            double xw, yw, J[4];
            ideal_to_warped(d, x, y, xw, yw, J);
            const double rx = xw - tx, ry = yw - ty;
            if (!(fabs(rx) < 1e300) || !(fabs(ry) < 1e300)) break;
            if (fabs(rx) < 1e-9 && fabs(ry) < 1e-9) {
                ok = true;
                break;
            }
            const double det = J[0] * J[3] - J[1] * J[2];
            if (!(fabs(det) > 1e-12)) break;
            double dx = (J[3] * rx - J[1] * ry) / det, dy = (-J[2] * rx + J[0] * ry) / det;
            const double lim = 0.25 * (d.halfW + d.halfH);  // damp wild steps far outside the calibrated range
            const double n = fmax(fabs(dx), fabs(dy));
            if (n > lim) {
                dx *= lim / n;
                dy *= lim / n;
            }
            x -= dx;
            y -= dy;
        }
        sc = d.one ? x : x * d.s - 0.5;
        sr = d.one ? y : y * d.s - 0.5;
    }
    int32_t idx = -1;
    if (ok && isfinite(sr) && isfinite(sc)) {
        const double rr = floor(sr + 0.5), rc = floor(sc + 0.5);  // nearest pixel
        if (rr >= 0.0 && rr < (double)h && rc >= 0.0 && rc < (double)w) idx = (int32_t)((int64_t)rr * w + (int64_t)rc);
    }
    src_index[p] = idx;
    if (src_rc) {
        src_rc[2 * p] = ok ? (float)sr : -1.f;
        src_rc[2 * p + 1] = ok ? (float)sc : -1.f;
    }
}

__global__ void __launch_bounds__(256) k_gather_i32(const int32_t *__restrict__ in, const int32_t *__restrict__ src_index,
                                                    int64_t n_out, int32_t fill, int32_t *__restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_out; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = src_index[p];
        out[p] = s >= 0 ? in[s] : fill;
    }
}


// ---- save_renders' up-sampling to the native resolution (meshes.py:2312-2321) ---------------------------------------
// skimage.transform.resize(rendered, native_size, order = 0 | 1) is scipy.ndimage.zoom(..., grid_mode=True,
// mode="mirror") (skimage >= 0.19; numpy-style "reflect" is ndimage's "mirror"): output pixel i samples the input at
// x = (i + 0.5) * n_in / n_out - 0.5, nearest = floor(x + 0.5), linear between floor(x) and floor(x) + 1, coordinates
// outside [0, n_in - 1] mirrored about the end pixels.  NaN (no face) propagates through the linear weights exactly
// like it does there.  The uint8 rule of save_renders (meshes.py:2323-2334) can be applied on the way out.
__device__ __forceinline__ double mirror_coord(double x, int n) {
    if (n == 1) return 0.0;
    if (x < 0.0) x = -x;
    const double hi = (double)(n - 1);
    if (x > hi) x = 2.0 * hi - x;
    return x;
}

template <typename OUT>
__device__ __forceinline__ OUT resize_out(double v);
template <>
__device__ __forceinline__ double resize_out<double>(double v) {
    return v;
}
template <>
__device__ __forceinline__ uint8_t resize_out<uint8_t>(double v) {
    if (!(v >= 0.0) || v > 255.0 || !isfinite(v)) return 0;
    return (uint8_t)v;
}

template <typename OUT>
__global__ void __launch_bounds__(256) k_resize(const double *__restrict__ in, int h_in, int w_in, int D, int h_out,
                                                int w_out, int order, OUT *__restrict__ out) {
    const double sy = (double)h_in / (double)h_out, sx = (double)w_in / (double)w_out;
    const int64_t n = (int64_t)h_out * w_out;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / w_out), j = (int)(p - (int64_t)i * w_out);
        // no FMA contraction here: a coordinate that is an exact integer in ndimage's arithmetic must be one here
        // too, or floor() picks another pair of neighbours (visible through NaN propagation)
        const double y = __dadd_rn(__dmul_rn((double)i + 0.5, sy), -0.5), x = __dadd_rn(__dmul_rn((double)j + 0.5, sx), -0.5);
        if (order == 0) {
            const int yi = min(max((int)floor(y + 0.5), 0), h_in - 1), xi = min(max((int)floor(x + 0.5), 0), w_in - 1);
            for (int d = 0; d < D; ++d) out[p * D + d] = resize_out<OUT>(in[((int64_t)yi * w_in + xi) * D + d]);
        } else {
            const double ym = mirror_coord(y, h_in), xm = mirror_coord(x, w_in);
            const int y0 = min((int)floor(ym), h_in - 1), x0 = min((int)floor(xm), w_in - 1);
            // the upper neighbour of the last pixel is its mirror image (index n - 2), with weight 0: it only matters
            // for NaN propagation, which ndimage does through zero weights too
            const int y1 = y0 + 1 < h_in ? y0 + 1 : max(h_in - 2, 0), x1 = x0 + 1 < w_in ? x0 + 1 : max(w_in - 2, 0);
            const double ty = ym - (double)y0, tx = xm - (double)x0;
            for (int d = 0; d < D; ++d) {
                const double a = in[((int64_t)y0 * w_in + x0) * D + d], b = in[((int64_t)y0 * w_in + x1) * D + d];
                const double c = in[((int64_t)y1 * w_in + x0) * D + d], e = in[((int64_t)y1 * w_in + x1) * D + d];
                // separable, rows first (ndimage.zoom filters axis by axis)
                const double top = (1.0 - tx) * a + tx * b, bot = (1.0 - tx) * c + tx * e;
                out[p * D + d] = resize_out<OUT>((1.0 - ty) * top + ty * bot);
            }
        }
    }
}

}  // namespace

extern "C" {

int gg_build_warp_map(int device, const gg_distortion *h_dist, int h, int w, int warped_to_ideal, int32_t *d_src_index,
                      float *d_src_rc, void *stream) {
    if (!h_dist || !d_src_index || h < 1 || w < 1 || !(h_dist->f > 0) || !(h_dist->image_scale > 0)) {
        gg_set_error("gg_build_warp_map: bad arguments");
        return GG_ERR_INVALID;
    }
    GG_CUDA(cudaSetDevice(device));
    Dist d;
    d.f = h_dist->f;
    d.cx = h_dist->cx;
    d.cy = h_dist->cy;
    d.halfW = h_dist->W / 2.0;
    d.halfH = h_dist->H / 2.0;
    d.k1 = h_dist->k1;
    d.k2 = h_dist->k2;
    d.k3 = h_dist->k3;
    d.k4 = h_dist->k4;
    d.p1 = h_dist->p1;
    d.p2 = h_dist->p2;
    d.b1 = h_dist->b1;
    d.b2 = h_dist->b2;
    d.s = h_dist->image_scale;
    const double diff = d.s > 1.0 ? d.s - 1.0 : 1.0 - d.s;
    d.one = diff <= 1e-8 + 1e-5;  // numpy.isclose(image_scale, 1.0), cameras.py:1030
    const int64_t n = (int64_t)h * w;
    k_build_warp_map<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, h, w, warped_to_ideal, d_src_index,
                                                                                   d_src_rc);
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}

int gg_gather_i32(int device, const int32_t *d_in, const int32_t *d_src_index, int64_t n_out, int32_t fill,
                  int32_t *d_out, void *stream) {
    if (!d_in || !d_src_index || !d_out || n_out < 1) {
        gg_set_error("gg_gather_i32: bad arguments");
        return GG_ERR_INVALID;
    }
    GG_CUDA(cudaSetDevice(device));
    const int64_t want = (n_out + 255) / 256;
    k_gather_i32<<<(unsigned)(want < 148 * 32 ? want : 148 * 32), 256, 0, (cudaStream_t)stream>>>(d_in, d_src_index, n_out,
                                                                                                 fill, d_out);
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}

int gg_resize_render(int device, const double *d_in, int h_in, int w_in, int D, int h_out, int w_out, int order,
                     void *d_out, int out_dtype, void *stream) {
    if (!d_in || !d_out || h_in < 1 || w_in < 1 || h_out < 1 || w_out < 1 || D < 1 || (order != 0 && order != 1) ||
        (out_dtype != GG_OUT_F64 && out_dtype != GG_OUT_U8)) {
        gg_set_error("gg_resize_render: bad arguments (order 0 | 1; out_dtype GG_OUT_F64 | GG_OUT_U8)");
        return GG_ERR_INVALID;
    }
    GG_CUDA(cudaSetDevice(device));
    const int64_t want = ((int64_t)h_out * w_out + 255) / 256;
    const unsigned g = (unsigned)(want < 148 * 32 ? want : 148 * 32);
    if (out_dtype == GG_OUT_F64)
        k_resize<double><<<g, 256, 0, (cudaStream_t)stream>>>(d_in, h_in, w_in, D, h_out, w_out, order, (double *)d_out);
    else
        k_resize<uint8_t><<<g, 256, 0, (cudaStream_t)stream>>>(d_in, h_in, w_in, D, h_out, w_out, order, (uint8_t *)d_out);
    GG_CUDA(cudaGetLastError());
    return GG_OK;
}

}  // extern "C"
