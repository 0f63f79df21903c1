// sfh_post.cu — GPU post-processing of the predict outputs, sm_100a.
//
// The reference converts its outputs on the CPU after a device->host copy of int32 / fp32 tensors:
//   utils/postprocess.py:7-18   preds_to_masks: argmax over softmax(logits)          -> uint8 class ids
//   predict.py:99               warp_mask.cpu().numpy().astype(np.uint8)
//   predict.py:288-299          mask_type 'rgb' (utils/postprocess.py:21-58 onehot_to_image),
//                               'bin' ((mask > 0) * 255), 'gray' (class ids)
//   predict.py:303-315          cv2.resize(mask, out_size, interpolation=cv2.INTER_NEAREST)
// Doing the same before the copy shrinks the PCIe traffic 4-16x (uint8 at the output size instead of
// 4 logit planes / int32 masks).  One streaming pass: a thread produces 4 consecutive output
// pixels; the cv2 nearest-neighbour source index tables (x_ofs / y_ofs, computed on the host in
// double exactly as cv::resize does) are inputs, NULL = same size.
#include <cuda_runtime.h>
#include <stdint.h>
#include "sfh_device.cuh"

namespace sfh {

struct PostParams {
    const void* src;
    const int* xo;          // [ow] source column of every output column, nullable
    const int* yo;          // [oh] source row of every output row, nullable
    unsigned char* out;     // [B,oh,ow] or [B,oh,ow,3]
    int kind, B, nc, h, w, oh, ow, mask_type;
    unsigned pal[8];        // r | g << 8 | b << 16 per class id (rgb)
};

template <int KIND>
__device__ __forceinline__ int post_class(const PostParams& p, size_t plane, size_t img_off, int sy, int sx) {
    const size_t o = (size_t)sy * p.w + sx;
    if (KIND == SFH_POST_SRC_LOGITS) {
        // argmax(softmax(l)) == argmax(l), first maximum wins (torch.argmax's tie rule)
        const float* lg = (const float*)p.src + img_off * p.nc + o;
        float best = __ldg(lg);
        int arg = 0;
        for (int c = 1; c < p.nc; ++c) {
            const float v = __ldg(lg + (size_t)c * plane);
            if (v > best) { best = v; arg = c; }
        }
        return arg;
    }
    if (KIND == SFH_POST_SRC_MASK_I32) return (int)(unsigned char)__ldg((const int*)p.src + img_off + o);   // .astype(np.uint8) wraps
    return (int)__ldg((const unsigned char*)p.src + img_off + o);
}

__device__ __forceinline__ unsigned post_value(const PostParams& p, int cls) {
    if (p.mask_type == SFH_POST_GRAY) return (unsigned)cls;
    if (p.mask_type == SFH_POST_BIN) return cls > 0 ? 255u : 0u;
    return (unsigned)cls < 8u ? p.pal[cls] : 0u;
}

template <int KIND>
__global__ void __launch_bounds__(256) k_post(const __grid_constant__ PostParams p) {
    const size_t plane = (size_t)p.h * p.w;
    const int qpr = (p.ow + 3) >> 2;                          // quads per output row
    const size_t total = (size_t)p.B * p.oh * qpr;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const size_t rowi = q / qpr;                          // b * oh + y
        const int x0 = (int)(q - rowi * qpr) * 4;
        const int b = (int)(rowi / p.oh), y = (int)(rowi - (size_t)b * p.oh);
        const int sy = p.yo ? __ldg(p.yo + y) : y;
        const size_t img_off = (size_t)b * plane;
        unsigned v[4];
        const bool same = (p.xo == nullptr) && (x0 + 3 < p.ow) && ((p.w & 3) == 0);
        if (same && KIND == SFH_POST_SRC_LOGITS && p.nc == 4) {
            // same-size fast path: 4 x 128-bit loads, one per logit plane
            const float* lg = (const float*)p.src + img_off * 4 + (size_t)sy * p.w + x0;
            const float4 a = __ldcs((const float4*)lg), bb = __ldcs((const float4*)(lg + plane));
            const float4 c = __ldcs((const float4*)(lg + 2 * plane)), d = __ldcs((const float4*)(lg + 3 * plane));
            const float l0[4] = {a.x, a.y, a.z, a.w}, l1[4] = {bb.x, bb.y, bb.z, bb.w};
            const float l2[4] = {c.x, c.y, c.z, c.w}, l3[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float best = l0[j];
                int arg = 0;
                if (l1[j] > best) { best = l1[j]; arg = 1; }
                if (l2[j] > best) { best = l2[j]; arg = 2; }
                if (l3[j] > best) { best = l3[j]; arg = 3; }
                v[j] = post_value(p, arg);
            }
        } else if (same && KIND == SFH_POST_SRC_MASK_I32) {
            const int4 m = __ldcs((const int4*)((const int*)p.src + img_off + (size_t)sy * p.w + x0));
            v[0] = post_value(p, (int)(unsigned char)m.x); v[1] = post_value(p, (int)(unsigned char)m.y);
            v[2] = post_value(p, (int)(unsigned char)m.z); v[3] = post_value(p, (int)(unsigned char)m.w);
        } else {
            int sxs[4];
            if (p.xo && x0 + 3 < p.ow && ((p.ow & 3) == 0)) {
                const int4 t = __ldg(reinterpret_cast<const int4*>(p.xo + x0));
                sxs[0] = t.x; sxs[1] = t.y; sxs[2] = t.z; sxs[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = min(x0 + j, p.ow - 1);
                    sxs[j] = p.xo ? __ldg(p.xo + x) : x;
                }
            }
            int prev_sx = -1;
            unsigned prev_v = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {                    // upscaling repeats source pixels: classify each once
                if (sxs[j] != prev_sx) {
                    prev_sx = sxs[j];
                    prev_v = post_value(p, post_class<KIND>(p, plane, img_off, sy, sxs[j]));
                }
                v[j] = prev_v;
            }
        }
        const size_t opix = rowi * p.ow + x0;
        if (p.mask_type != SFH_POST_RGB) {
            unsigned char* o = p.out + opix;
            if (x0 + 3 < p.ow && ((p.ow & 3) == 0)) {
                __stcs((unsigned*)o, v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24));
            } else {
                for (int j = 0; j < 4 && x0 + j < p.ow; ++j) o[j] = (unsigned char)v[j];
            }
        } else {
            unsigned char* o = p.out + opix * 3;
            if (x0 + 3 < p.ow && ((p.ow & 3) == 0)) {        // 12 bytes = 3 aligned words (a smem transpose to
                unsigned* o32 = (unsigned*)o;                 // 128-bit stores was measured slower)
                __stcs(o32 + 0, (v[0] & 0xffffffu) | (v[1] << 24));
                __stcs(o32 + 1, ((v[1] >> 8) & 0xffffu) | (v[2] << 16));
                __stcs(o32 + 2, ((v[2] >> 16) & 0xffu) | (v[3] << 8));
            } else {
                for (int j = 0; j < 4 && x0 + j < p.ow; ++j) {
                    o[3 * j + 0] = (unsigned char)(v[j] & 0xffu);
                    o[3 * j + 1] = (unsigned char)((v[j] >> 8) & 0xffu);
                    o[3 * j + 2] = (unsigned char)((v[j] >> 16) & 0xffu);
                }
            }
        }
    }
}

}  // namespace sfh

using namespace sfh;

extern "C" {

int sfh_postprocess(const void* src, int src_kind, int B, int nc, int h, int w,
                    int mask_type, const int* x_ofs, const int* y_ofs, int oh, int ow,
                    unsigned char* out, void* stream) {
    if (!src || !out || B <= 0 || h <= 0 || w <= 0 || oh <= 0 || ow <= 0) return SFH_E_BADARG;
    if (src_kind != SFH_POST_SRC_LOGITS && src_kind != SFH_POST_SRC_MASK_I32 && src_kind != SFH_POST_SRC_MASK_U8) return SFH_E_BADARG;
    if (mask_type != SFH_POST_GRAY && mask_type != SFH_POST_BIN && mask_type != SFH_POST_RGB) return SFH_E_BADMODE;
    if (src_kind == SFH_POST_SRC_LOGITS && nc < 2) return SFH_E_BADARG;
    if (mask_type == SFH_POST_RGB && nc != 4 && nc != 7 && nc != 8) return SFH_E_BADARG;   // utils/postprocess.py:57 NotImplementedError
    if ((!x_ofs && ow != w) || (!y_ofs && oh != h)) return SFH_E_BADARG;
    PostParams p = {};
    p.src = src; p.xo = x_ofs; p.yo = y_ofs; p.out = out;
    p.kind = src_kind; p.B = B; p.nc = nc; p.h = h; p.w = w; p.oh = oh; p.ow = ow; p.mask_type = mask_type;
    // utils/postprocess.py:31-55: id -> colour, in the tuple order the reference writes
    static const unsigned char col[8][3] = {{0, 0, 0}, {0, 255, 0}, {255, 0, 0}, {0, 0, 255},
                                            {255, 255, 255}, {255, 0, 255}, {0, 255, 255}, {255, 255, 0}};
    for (int c = 0; c < 8; ++c)
        p.pal[c] = (c < nc) ? ((unsigned)col[c][0] | ((unsigned)col[c][1] << 8) | ((unsigned)col[c][2] << 16)) : 0u;
    const size_t total = (size_t)B * oh * ((ow + 3) / 4);
    size_t ctas = (total + 255) / 256;
    if (ctas > 148 * 8) ctas = 148 * 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (src_kind == SFH_POST_SRC_LOGITS)        k_post<SFH_POST_SRC_LOGITS><<<(int)ctas, 256, 0, st>>>(p);
    else if (src_kind == SFH_POST_SRC_MASK_I32) k_post<SFH_POST_SRC_MASK_I32><<<(int)ctas, 256, 0, st>>>(p);
    else                                        k_post<SFH_POST_SRC_MASK_U8><<<(int)ctas, 256, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // extern "C"
