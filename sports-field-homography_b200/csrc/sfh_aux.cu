// sfh_aux.cu — the remaining consumers of the warp outside the training / predict hot path:
//
//   * utils/transform.py:7-20  Warper.warp(theta, proj): kornia HomographyWarper(mode='nearest') applied to an
//     fp64 multi-channel [H,W,C] projection image, everything in double precision;
//   * dataset_utils/football_dataset.ipynb cell 11 / preparation.py:129-137: the dataset's masks (and UV maps) are
//     rendered with cv2.warpPerspective(template, rescale_theta(...), size, flags=cv2.INTER_NEAREST) — OpenCV's
//     pixel-coordinate convention, its double-precision evaluation order and its round-half-to-even pick.
//
// Both are gathers of whole texels (no interpolation), so they are exact copies of source elements; the only
// arithmetic is the coordinate, evaluated here in fp64 in the respective library's operation order.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sfh_b200.h"

namespace sfh {

// kornia path in double: meshgrid factor -> bmm chain (k = 0,1,2) -> 1/z where |z| > 1e-8 -> grid_sample's
// unnormalise (align_corners=False) -> nearbyint.  One thread per output pixel, channels looped (coalesced per plane).
__global__ void __launch_bounds__(256) k_warp_nearest_f64(const double* __restrict__ theta, const double* __restrict__ tmpl,
                                                          long long tmpl_bstride, const double* __restrict__ xs,
                                                          const double* __restrict__ ys, int C, int Hc, int Wc, int H, int W,
                                                          double* __restrict__ out) {
    const int b = blockIdx.z;
    const int w = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y;
    if (w >= W) return;
    const double* t = theta + 9 * (size_t)b;
    const double u = xs[w], v = ys[h];
    const double X = __dadd_rn(__fma_rn(v, t[1], __dmul_rn(u, t[0])), t[2]);
    const double Y = __dadd_rn(__fma_rn(v, t[4], __dmul_rn(u, t[3])), t[5]);
    const double Z = __dadd_rn(__fma_rn(v, t[7], __dmul_rn(u, t[6])), t[8]);
    const double s = fabs(Z) > 1e-8 ? __drcp_rn(Z) : 1.0;
    const double x = __dmul_rn(s, X), y = __dmul_rn(s, Y);
    double ix = __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, 1.0), (double)Wc), -1.0), 0.5);
    double iy = __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(y, 1.0), (double)Hc), -1.0), 0.5);
    // ATen safe_downgrade_to_int_range: non-finite or beyond the int range -> -100 (out of bounds)
    if (!(fabs(ix) < 2147483647.0)) ix = -100.0;
    if (!(fabs(iy) < 2147483647.0)) iy = -100.0;
    const int xn = (int)nearbyint(ix), yn = (int)nearbyint(iy);
    const bool in = (unsigned)xn < (unsigned)Wc && (unsigned)yn < (unsigned)Hc;
    const double* src = tmpl + (size_t)b * (size_t)tmpl_bstride + (size_t)yn * Wc + xn;
    double* dst = out + (((size_t)b * C) * H + h) * W + w;
    const size_t cs = (size_t)Hc * Wc, cd = (size_t)H * W;
    for (int c = 0; c < C; ++c) dst[c * cd] = in ? __ldg(src + c * cs) : 0.0;
}

// cv::warpPerspective(src, M, dsize, INTER_NEAREST, BORDER_CONSTANT 0) for one or many M.  minv = inverse of M
// (dst -> src), row-major double[9].  OpenCV evaluates a dst row in blocks of 64 columns: with bx the block's first
// column, X0 = minv0*bx + minv1*y + minv2 (left to right), then per pixel W = W0 + minv6*x1, W = W ? 1/W : 0,
// fX = clamp((X0 + minv0*x1) * W) and the source index is cvRound(fX) (round half to even).  Elements are `esz`
// bytes (pixel = C*elem bytes, HWC interleaved like a cv::Mat) and copied verbatim.
template <int ESZ>
__global__ void __launch_bounds__(256) k_warp_perspective_nearest(const double* __restrict__ minv, const unsigned char* __restrict__ src,
                                                                  int Hs, int Ws, int H, int W, unsigned char* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const double* m = minv + 9 * (size_t)b;
    const int bx = x & ~63, x1 = x - bx;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(m[0], (double)bx), __dmul_rn(m[1], (double)y)), m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m[3], (double)bx), __dmul_rn(m[4], (double)y)), m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(m[6], (double)bx), __dmul_rn(m[7], (double)y)), m[8]);
    double Wd = __dadd_rn(W0, __dmul_rn(m[6], (double)x1));
    Wd = Wd != 0.0 ? __ddiv_rn(1.0, Wd) : 0.0;
    const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(m[0], (double)x1)), Wd)));
    const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(m[3], (double)x1)), Wd)));
    const int sx = __double2int_rn(fX), sy = __double2int_rn(fY);
    const bool in = (unsigned)sx < (unsigned)Ws && (unsigned)sy < (unsigned)Hs;
    unsigned char* d = dst + (((size_t)b * H + y) * W + x) * ESZ;
    const unsigned char* s = src + ((size_t)sy * Ws + sx) * ESZ;
#pragma unroll
    for (int k = 0; k < ESZ; ++k) d[k] = in ? __ldg(s + k) : (unsigned char)0;
}

}  // namespace sfh

using namespace sfh;

extern "C" {

int sfh_warp_nearest_f64(const double* theta, const double* tmpl, int64_t tmpl_batch_stride, const double* xs,
                         const double* ys, int B, int C, int Hc, int Wc, int H, int W, double* out, void* stream) {
    if (!theta || !tmpl || !xs || !ys || !out || B <= 0 || C <= 0 || Hc <= 0 || Wc <= 0 || H <= 0 || W <= 0) return SFH_E_BADARG;
    if (B > 65535 || H > 65535) return SFH_E_BADARG;
    dim3 grid((W + 255) / 256, H, B);
    k_warp_nearest_f64<<<grid, 256, 0, (cudaStream_t)stream>>>(theta, tmpl, tmpl_batch_stride, xs, ys, C, Hc, Wc, H, W, out);
    return (int)cudaGetLastError();
}

int sfh_warp_perspective_nearest(const double* minv, int B, const void* src, int Hs, int Ws, int pixel_bytes,
                                 int H, int W, void* dst, void* stream) {
    if (!minv || !src || !dst || B <= 0 || Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0) return SFH_E_BADARG;
    if (B > 65535 || H > 65535) return SFH_E_BADARG;
    dim3 grid((W + 255) / 256, H, B);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned char* s = (const unsigned char*)src;
    unsigned char* d = (unsigned char*)dst;
    switch (pixel_bytes) {
        case 1:  k_warp_perspective_nearest<1><<<grid, 256, 0, st>>>(minv, s, Hs, Ws, H, W, d); break;
        case 2:  k_warp_perspective_nearest<2><<<grid, 256, 0, st>>>(minv, s, Hs, Ws, H, W, d); break;
        case 3:  k_warp_perspective_nearest<3><<<grid, 256, 0, st>>>(minv, s, Hs, Ws, H, W, d); break;
        case 4:  k_warp_perspective_nearest<4><<<grid, 256, 0, st>>>(minv, s, Hs, Ws, H, W, d); break;
        case 8:  k_warp_perspective_nearest<8><<<grid, 256, 0, st>>>(minv, s, Hs, Ws, H, W, d); break;
        case 16: k_warp_perspective_nearest<16><<<grid, 256, 0, st>>>(minv, s, Hs, Ws, H, W, d); break;
        case 24: k_warp_perspective_nearest<24><<<grid, 256, 0, st>>>(minv, s, Hs, Ws, H, W, d); break;
        default: return SFH_E_BADARG;
    }
    return (int)cudaGetLastError();
}

}  // extern "C"
