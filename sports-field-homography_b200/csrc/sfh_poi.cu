// sfh_poi.cu — stand-alone launches for the point path: transform_poi fwd/bwd,
// kornia transform_points fwd/bwd (fp32, reference op order) and reprojection_loss.
#include "sfh_device.cuh"
#include "sfh_poi.cuh"

namespace sfh {

__global__ void __launch_bounds__(32) k_poi(const __grid_constant__ PoiParams p) { poi_block(p, blockIdx.x); }

// kornia transform_points on [B,N,2]: bmm(points_h, T^T) then 1/z where |z| > eps.
// Same k-ordered FMA chain as the grid path (sfh_device.cuh).
__device__ __forceinline__ void tp_point(const float* t, float px, float py, float& X, float& Y, float& s, bool& ok) {
    X = __fadd_rn(__fmaf_rn(py, t[1], __fmul_rn(px, t[0])), t[2]);
    Y = __fadd_rn(__fmaf_rn(py, t[4], __fmul_rn(px, t[3])), t[5]);
    const float Z = __fadd_rn(__fmaf_rn(py, t[7], __fmul_rn(px, t[6])), t[8]);
    ok = fabsf(Z) > kEps;
    s = ok ? __frcp_rn(Z) : 1.0f;
}

__global__ void k_tp_fwd(const float* trans, int Bt, const float* pts, int B, int N, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N;
    float t[9];
    const float* tp = trans + (Bt == 1 ? 0 : 9 * b);
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = __ldg(tp + k);
    float X, Y, s; bool ok;
    tp_point(t, pts[2 * i], pts[2 * i + 1], X, Y, s, ok);
    out[2 * i] = __fmul_rn(s, X);
    out[2 * i + 1] = __fmul_rn(s, Y);
}

// one CTA per transform: dtrans[j][k] = sum over the points it applies to of g_j * p_k
__global__ void __launch_bounds__(kThreads) k_tp_bwd(const float* trans, int Bt, const float* pts,
                                                     const float* gout, int B, int N,
                                                     float* dtrans, float* dpts) {
    __shared__ double s_w[kWarps][9];
    const int tb = blockIdx.x;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = __ldg(trans + 9 * tb + k);
    const int first = (Bt == 1) ? 0 : tb * N, count = (Bt == 1) ? B * N : N;
    double G[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) G[k] = 0.0;
    for (int q = threadIdx.x; q < count; q += kThreads) {
        const int i = first + q;
        const float px = pts[2 * i], py = pts[2 * i + 1];
        float X, Y, s; bool ok;
        tp_point(t, px, py, X, Y, s, ok);
        const float gx = gout[2 * i], gy = gout[2 * i + 1];
        const float gX = gx * s, gY = gy * s;
        const float gZ = ok ? -(gx * X + gy * Y) * s * s : 0.f;
        if (dpts) {
            dpts[2 * i] = gX * t[0] + gY * t[3] + gZ * t[6];
            dpts[2 * i + 1] = gX * t[1] + gY * t[4] + gZ * t[7];
        }
        G[0] += (double)gX * px; G[1] += (double)gX * py; G[2] += gX;
        G[3] += (double)gY * px; G[4] += (double)gY * py; G[5] += gY;
        G[6] += (double)gZ * px; G[7] += (double)gZ * py; G[8] += gZ;
    }
    if (!dtrans) return;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double s = warp_sum_d(G[k]);
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += s_w[w][threadIdx.x];
        dtrans[9 * tb + threadIdx.x] = (float)s;
    }
}

// models/losses.py:10-11 per sample, one warp per sample.
__global__ void __launch_bounds__(32) k_reproj(const float* in, const float* tg, const float* nz,
                                               const float* num, int N, float* Rb,
                                               const float* gRb, float* din) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const double nn = (double)num[b];
    const double g = gRb ? (double)gRb[b] : 0.0;
    double loss = 0.0;
    for (int n = lane; n < N; n += 32) {
        const size_t o = ((size_t)b * N + n) * 2;
        const double dx = (double)tg[o] - (double)in[o], dy = (double)tg[o + 1] - (double)in[o + 1];
        const double dist = sqrt(dx * dx + dy * dy);
        const double w = (double)nz[(size_t)b * N + n] / nn;
        loss += dist * w;
        if (din) {
            const double gs = g * w / (2.0 * dist);
            din[o] = (float)(gs * (-2.0 * dx));
            din[o + 1] = (float)(gs * (-2.0 * dy));
        }
    }
    loss = warp_sum_d(loss);
    if (lane == 0 && Rb) Rb[b] = (float)loss;
}

__global__ void k_selftest_rcp(unsigned long long* mismatches) {
    unsigned long long bad = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
        const float z = __uint_as_float((unsigned)i);
        const float az = fabsf(z);
        if (!(az > kEps) || !(az < 1e37f)) continue;
        if (__float_as_uint(rcp_rn_normal(z)) != __float_as_uint(__frcp_rn(z))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// Diagnostic streaming kernel: out[i] = (float)(int)in[i] * 0.25f over n int64 -> fp32, 4 elements
// per thread per iteration (2 x 128-bit loads, 1 x 128-bit store), grid-stride.  Moves exactly the
// bytes of one C2 step with no other work: the practical ceiling for a kernel of that size.
__global__ void __launch_bounds__(256) k_stream_cast(const longlong2* __restrict__ in, float4* __restrict__ out, long long n4) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const longlong2 a = __ldcs(in + 2 * i), b = __ldcs(in + 2 * i + 1);
        __stcs(out + i, make_float4((float)(int)a.x * 0.25f, (float)(int)a.y * 0.25f, (float)(int)b.x * 0.25f, (float)(int)b.y * 0.25f));
    }
}

}  // namespace sfh

using namespace sfh;

extern "C" {

int sfh_poi_fwd(const float* theta, const float* court_poi, int64_t court_poi_bstride,
                int B, int N, int normalize, float* poi_out, void* stream) {
    if (!theta || !court_poi || !poi_out || B <= 0 || N <= 0) return SFH_E_BADARG;
    PoiParams p = {};
    p.theta = theta; p.court_poi = court_poi; p.bstride = court_poi_bstride;
    p.N = N; p.normalize = normalize; p.poi_out = poi_out;
    k_poi<<<B, 32, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

int sfh_poi_bwd(const float* theta, const float* court_poi, int64_t court_poi_bstride,
                const float* grad_poi, int B, int N, int normalize, float* dtheta, void* stream) {
    if (!theta || !court_poi || !grad_poi || !dtheta || B <= 0 || N <= 0) return SFH_E_BADARG;
    PoiParams p = {};
    p.theta = theta; p.court_poi = court_poi; p.bstride = court_poi_bstride;
    p.N = N; p.normalize = normalize; p.grad_poi = grad_poi; p.dtheta = dtheta;
    k_poi<<<B, 32, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

int sfh_transform_points_fwd(const float* trans, int Bt, const float* points, int B, int N,
                             float* out, void* stream) {
    if (!trans || !points || !out || B <= 0 || N <= 0 || (Bt != 1 && Bt != B)) return SFH_E_BADARG;
    const int total = B * N;
    k_tp_fwd<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(trans, Bt, points, B, N, out);
    return (int)cudaGetLastError();
}

int sfh_transform_points_bwd(const float* trans, int Bt, const float* points, const float* grad_out,
                             int B, int N, float* dtrans, float* dpoints, void* stream) {
    if (!trans || !points || !grad_out || B <= 0 || N <= 0 || (Bt != 1 && Bt != B)) return SFH_E_BADARG;
    if (!dtrans && !dpoints) return 0;
    k_tp_bwd<<<Bt, kThreads, 0, (cudaStream_t)stream>>>(trans, Bt, points, grad_out, B, N, dtrans, dpoints);
    return (int)cudaGetLastError();
}

int sfh_reproj_loss(const float* inputs, const float* targets, const float* nonzeros,
                    const float* num_nonzero, int B, int N, float* R_b,
                    const float* grad_Rb, float* dinputs, void* stream) {
    if (!inputs || !targets || !nonzeros || !num_nonzero || B <= 0 || N <= 0) return SFH_E_BADARG;
    if (dinputs && !grad_Rb) return SFH_E_BADARG;
    k_reproj<<<B, 32, 0, (cudaStream_t)stream>>>(inputs, targets, nonzeros, num_nonzero, N, R_b, grad_Rb, dinputs);
    return (int)cudaGetLastError();
}

int sfh_debug_stream_cast(const int64_t* in, float* out, int64_t n, int ctas, void* stream) {
    if (!in || !out || n <= 0 || (n & 3) || ctas <= 0) return SFH_E_BADARG;
    k_stream_cast<<<ctas, 256, 0, (cudaStream_t)stream>>>((const longlong2*)in, (float4*)out, n / 4);
    return (int)cudaGetLastError();
}

int sfh_selftest_rcp(unsigned long long* mismatches, void* stream) {
    if (!mismatches) return SFH_E_BADARG;
    k_selftest_rcp<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(mismatches);
    return (int)cudaGetLastError();
}

}  // extern "C"
