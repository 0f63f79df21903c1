// sfh_consist.cu — training / evaluation consistency loss, sm_100a.
//
//   train.py:219-223   rec_masks_int = (warp_mask * num_classes).long()
//                      consist_loss  = nn.CrossEntropyLoss()(logits, rec_masks_int) * consist_lambda
//   eval.py:201-203    the same value as a metric (no gradient)
//
// One streaming pass over the logits: a thread owns 4 consecutive pixels (128-bit loads of the
// nc logit planes and of the fp32 warp mask, 128-bit stores of the nc gradient planes), so the
// log_softmax tensor ([B,nc,H,W]) and the int64 class mask ([B,H,W]) of the reference are never
// materialised.  CE = logsumexp(l) - l[cls]; dCE/dl_c = softmax_c - [c == cls], scaled by
// lambda / (B*H*W) (reduction 'mean').  exp/log run on the MUFU ex2/lg2 units (rel. error ~1e-7
// on the probabilities; tests hold loss to 1e-5 relative and gradients to 1e-6 of their scale).
//
// Focal variant (train.py:133-134: kornia.losses.FocalLoss(alpha=1.0, gamma=2.0, reduction='mean') as the
// consistency criterion), restated from kornia 0.5.x/0.6.x focal_loss — eps conventions included:
//   p~_c = softmax_c + 1e-8;  t_c = [c == cls] + 1e-6 (kornia.utils.one_hot adds its own eps);
//   loss_px = sum_c t_c * (-alpha * (1 - p~_c)^gamma * log p~_c);  mean over all pixels
//   dloss_px/dl_k = p_k * (g_k - sum_c g_c p_c),  g_c = t_c * alpha * (gamma (1-p~_c)^(gamma-1) log p~_c - (1-p~_c)^gamma / p~_c)
//
// The loss is reduced without data atomics: thread -> warp shuffle -> CTA -> one fp32 partial per
// CTA; the last CTA (release ticket) adds the partials in fixed order in fp64.  The grid size is a
// function of the problem size only, so results are run-to-run deterministic.
#include <cuda_runtime.h>
#include <stdint.h>
#include "sfh_device.cuh"

namespace sfh {

constexpr int kCsThreads = 256;
constexpr int kCsMaxCtas = 148 * 4;     // one resident wave at 4 CTAs/SM (62 registers)
constexpr int kCsMaxNc = 8;

struct ConsistParams {
    const float* mask;      // [B,1,H,W] fp32 warp mask (values k/nc)
    const float* logits;    // [B,nc,H,W]
    float* dlogits;         // [B,nc,H,W], nullable (evaluation)
    float* loss_out;        // scalar: lambda * mean CE
    float* partials;        // workspace [gridDim.x]
    int* counter;           // workspace, zero between calls
    long long plane;        // H*W
    long long quads;        // B * plane / 4  (plane % 4 == 0)
    long long total_px;     // B * plane
    float ncf, lambda, gscale;   // gscale = lambda / total_px
    int nc;
    int kind;               // 0: cross entropy, 1: focal (kornia 0.5.x focal_loss)
    float alpha, gamma;     // focal parameters
};

constexpr float kFocalEpsP = 1e-8f;     // kornia focal_loss: softmax + eps
constexpr float kFocalEpsT = 1e-6f;     // kornia.utils.one_hot: scatter(1.0) + eps

// One pixel of the focal criterion: e[c] = exp(l_c - max), se = sum e.  Returns the loss and writes the
// gradient w.r.t. the logits (unscaled) into g[].
template <int NC>
__device__ __forceinline__ float focal_pixel(const float (&e)[NC], float se, int cls, float alpha, float gamma, float (&g)[NC]) {
    const float inv = __fdividef(1.0f, se);
    float loss = 0.f, dot = 0.f, pr[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        pr[c] = e[c] * inv;
        const float pt = pr[c] + kFocalEpsP, om = 1.0f - pt;
        const float lg = __log2f(pt) * 0.6931471805599453f;
        const float t = (c == cls ? 1.0f : 0.0f) + kFocalEpsT;
        float w, dw;                         // (1-p~)^gamma and gamma (1-p~)^(gamma-1)
        if (gamma == 2.0f) { w = om * om; dw = 2.0f * om; }
        else { w = __powf(fmaxf(om, 0.f), gamma); dw = gamma * __powf(fmaxf(om, 0.f), gamma - 1.0f); }
        loss = fmaf(t, -alpha * w * lg, loss);
        g[c] = t * alpha * (dw * lg - __fdividef(w, pt));
        dot = fmaf(g[c], pr[c], dot);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) g[c] = pr[c] * (g[c] - dot);
    return loss;
}

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NC, int KIND>
__global__ void __launch_bounds__(kCsThreads, KIND == 0 ? 4 : 3) k_consist(const __grid_constant__ ConsistParams p) {
    __shared__ float s_w[kCsThreads / 32];
    __shared__ int s_last;
    const long long stride = (long long)gridDim.x * kCsThreads;
    const long long qpp = p.plane >> 2;                     // quads per plane
    float acc = 0.f;
    for (long long q = (long long)blockIdx.x * kCsThreads + threadIdx.x; q < p.quads; q += stride) {
        const long long b = q / qpp, r = q - b * qpp;       // sample, quad inside the plane
        const float4 m = __ldg(reinterpret_cast<const float4*>(p.mask) + q);
        const float4* lg = reinterpret_cast<const float4*>(p.logits) + (b * NC) * qpp + r;
        float4 l[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) l[c] = __ldcs(lg + (long long)c * qpp);
        const float mv[4] = {m.x, m.y, m.z, m.w};
        float4 g[NC];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // (warp_mask * nc).long(): truncation toward zero; out-of-range ids are clamped
            int cls = __float2int_rz(__fmul_rn(mv[j], p.ncf));
            cls = min(max(cls, 0), NC - 1);
            float lv[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) lv[c] = j == 0 ? l[c].x : j == 1 ? l[c].y : j == 2 ? l[c].z : l[c].w;
            float mx = lv[0];
#pragma unroll
            for (int c = 1; c < NC; ++c) mx = fmaxf(mx, lv[c]);
            const float k = 1.4426950408889634f, mk = -mx * k;
            float e[NC], se = 0.f, sel = lv[0];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                e[c] = ex2a(fmaf(lv[c], k, mk));
                se += e[c];
                if (c == cls) sel = lv[c];
            }
            if (KIND == 0) {
                acc += fmaf(__log2f(se), 0.6931471805599453f, mx) - sel;
                const float inv = __fdividef(p.gscale, se);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float gv = fmaf(e[c], inv, c == cls ? -p.gscale : 0.f);
                    if (j == 0) g[c].x = gv; else if (j == 1) g[c].y = gv; else if (j == 2) g[c].z = gv; else g[c].w = gv;
                }
            } else {
                float gp[NC];
                acc += focal_pixel<NC>(e, se, cls, p.alpha, p.gamma, gp);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float gv = gp[c] * p.gscale;
                    if (j == 0) g[c].x = gv; else if (j == 1) g[c].y = gv; else if (j == 2) g[c].z = gv; else g[c].w = gv;
                }
            }
        }
        if (p.dlogits) {
            float4* dg = reinterpret_cast<float4*>(p.dlogits) + (b * NC) * qpp + r;
#pragma unroll
            for (int c = 0; c < NC; ++c) __stcs(dg + (long long)c * qpp, g[c]);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsThreads / 32; ++w) s += s_w[w];
        __stcg(p.partials + blockIdx.x, s);
        s_last = (ticket_release(p.counter) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += (double)__ldcg(p.partials + i);
    s = warp_sum(s);
    if (threadIdx.x == 0) {
        *p.loss_out = (float)((double)p.lambda * s / (double)p.total_px);
        *p.counter = 0;
    }
}

// generic class count (1..kCsMaxNc, runtime): one pixel per thread, same arithmetic
__global__ void __launch_bounds__(kCsThreads) k_consist_generic(const __grid_constant__ ConsistParams p) {
    __shared__ float s_w[kCsThreads / 32];
    __shared__ int s_last;
    const long long stride = (long long)gridDim.x * kCsThreads;
    const int nc = p.nc;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * kCsThreads + threadIdx.x; i < p.total_px; i += stride) {
        const long long b = i / p.plane, r = i - b * p.plane;
        int cls = __float2int_rz(__fmul_rn(__ldg(p.mask + i), p.ncf));
        cls = min(max(cls, 0), nc - 1);
        const float* lg = p.logits + (b * nc) * p.plane + r;
        float mx = -INFINITY;
        for (int c = 0; c < nc; ++c) mx = fmaxf(mx, __ldg(lg + (long long)c * p.plane));
        const float k = 1.4426950408889634f, mk = -mx * k;
        float se = 0.f, sel = 0.f;
        for (int c = 0; c < nc; ++c) {
            const float v = __ldg(lg + (long long)c * p.plane);
            se += ex2a(fmaf(v, k, mk));
            if (c == cls) sel = v;
        }
        if (p.kind == 1) {                    // focal: pad the class vector to kCsMaxNc with zero probabilities
            float e[kCsMaxNc], gp[kCsMaxNc];
#pragma unroll
            for (int c = 0; c < kCsMaxNc; ++c) e[c] = c < nc ? ex2a(fmaf(__ldg(lg + (long long)c * p.plane), k, mk)) : 0.f;
            // classes >= nc do not exist in the reference: remove their eps-only terms again
            float loss = focal_pixel<kCsMaxNc>(e, se, cls, p.alpha, p.gamma, gp);
            const float lpad = -p.alpha * __log2f(kFocalEpsP) * 0.6931471805599453f *
                               (p.gamma == 2.0f ? (1.0f - kFocalEpsP) * (1.0f - kFocalEpsP) : __powf(1.0f - kFocalEpsP, p.gamma));
            acc += loss - (float)(kCsMaxNc - nc) * kFocalEpsT * lpad;
            if (p.dlogits) {
                float* dg = p.dlogits + (b * nc) * p.plane + r;
                for (int c = 0; c < nc; ++c) dg[(long long)c * p.plane] = gp[c] * p.gscale;
            }
            continue;
        }
        acc += fmaf(__log2f(se), 0.6931471805599453f, mx) - sel;
        if (p.dlogits) {
            const float inv = __fdividef(p.gscale, se);
            float* dg = p.dlogits + (b * nc) * p.plane + r;
            for (int c = 0; c < nc; ++c)
                dg[(long long)c * p.plane] = fmaf(ex2a(fmaf(__ldg(lg + (long long)c * p.plane), k, mk)), inv, c == cls ? -p.gscale : 0.f);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < kCsThreads / 32; ++w) s += s_w[w];
        __stcg(p.partials + blockIdx.x, s);
        s_last = (ticket_release(p.counter) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += (double)__ldcg(p.partials + i);
    s = warp_sum(s);
    if (threadIdx.x == 0) {
        *p.loss_out = (float)((double)p.lambda * s / (double)p.total_px);
        *p.counter = 0;
    }
}

}  // namespace sfh

using namespace sfh;

extern "C" {

int64_t sfh_consist_workspace_bytes(void) { return 256 + (int64_t)kCsMaxCtas * 4; }

static int consist_launch(const float* warp_mask, const float* logits, int B, int nc, int H, int W, int kind,
                          float alpha, float gamma, float lambda, float* loss_out, float* dlogits,
                          void* workspace, int64_t workspace_bytes, void* stream) {
    if (!warp_mask || !logits || !loss_out || B <= 0 || H <= 0 || W <= 0) return SFH_E_BADARG;
    if (nc < 1 || nc > kCsMaxNc) return SFH_E_BADARG;
    if (!workspace || workspace_bytes < sfh_consist_workspace_bytes()) return SFH_E_WS;
    ConsistParams p = {};
    p.mask = warp_mask; p.logits = logits; p.dlogits = dlogits; p.loss_out = loss_out;
    p.counter = (int*)workspace;
    p.partials = (float*)((char*)workspace + 256);
    p.plane = (long long)H * W;
    p.total_px = (long long)B * p.plane;
    p.quads = p.total_px / 4;
    p.ncf = (float)nc; p.lambda = lambda; p.nc = nc;
    p.kind = kind; p.alpha = alpha; p.gamma = gamma;
    p.gscale = (float)((double)lambda / (double)p.total_px);
    const bool vec = (p.plane % 4 == 0) && (((uintptr_t)warp_mask | (uintptr_t)logits | (uintptr_t)dlogits) & 15u) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long work = (vec && nc == 4) ? p.quads : p.total_px;
    long long ctas = (work + kCsThreads - 1) / kCsThreads;
    if (ctas > kCsMaxCtas) ctas = kCsMaxCtas;
    if (vec && nc == 4 && kind == 0) k_consist<4, 0><<<(int)ctas, kCsThreads, 0, st>>>(p);
    else if (vec && nc == 4)         k_consist<4, 1><<<(int)ctas, kCsThreads, 0, st>>>(p);
    else                             k_consist_generic<<<(int)ctas, kCsThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

int sfh_consist_loss_fwd_bwd(const float* warp_mask, const float* logits, int B, int nc, int H, int W,
                             float lambda, float* loss_out, float* dlogits,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    return consist_launch(warp_mask, logits, B, nc, H, W, 0, 1.0f, 2.0f, lambda, loss_out, dlogits,
                          workspace, workspace_bytes, stream);
}

int sfh_consist_focal_fwd_bwd(const float* warp_mask, const float* logits, int B, int nc, int H, int W,
                              float alpha, float gamma, float lambda, float* loss_out, float* dlogits,
                              void* workspace, int64_t workspace_bytes, void* stream) {
    if (!(gamma >= 0.f)) return SFH_E_BADARG;
    return consist_launch(warp_mask, logits, B, nc, H, W, 1, alpha, gamma, lambda, loss_out, dlogits,
                          workspace, workspace_bytes, stream);
}

}  // extern "C"
