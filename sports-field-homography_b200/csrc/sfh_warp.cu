// sfh_warp.cu — the fused warp kernel family (sm_100a) and its C-ABI launchers.
//
// One kernel template covers the four passes over the output pixels that the reference spreads
// over ~20 ATen launches:
//   kEpiStore   HomographyWarper.forward                     models/reconstructor.py:116
//   kEpiBwd     its autograd w.r.t. theta                    train.py:235
//   kEpiLoss    warp + MSE/SmoothL1 vs gt/nc + dL_b/dtheta   train.py:194-197, models/losses.py:35-38
//   kEpiPredict int32(warp*nc) + CE consistency score        models/reconstructor.py:223-240
//
// Layout of one CTA: 8 warps x 32 lanes on a 128 x 8R pixel tile (R = 1..16 bands of 8 rows) whose gt / logits are
// staged in smem by one TMA load.  The tile's 16x8 patches are classified against a summed-area table of class edges:
// patches that touch an edge are processed one per warp (a lane owns 4 consecutive pixels, exact per-pixel geometry,
// one packed-template load per pixel), the edge-free rest is streamed band by band (a lane owns 32 consecutive bytes
// of a row).  The sampling grid lives in registers only.  Per-sample sums (loss, 9 dtheta terms, score) go lane -> warp
// shuffle -> smem -> one partial per CTA, published as tagged 64-bit words; the sample's last-dispatched CTA adds them
// in fixed order in fp64 inside the same launch (no fences, no data atomics).  See DESIGN.md §4.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "sfh_device.cuh"
#include "sfh_poi.cuh"

namespace sfh {

struct FusedParams {
    alignas(64) CUtensorMap gt_map;   // TMA descriptor: kEpiLoss gt [B,H,W] int64, box 128 x 8R x 1;
                                      // kEpiPredict (ratio 2, nc 4) logits [B,4,h,w] fp32, box 64 x 4R x 4 x 1
    const float* theta;
    const float* xs;
    const float* ys;
    sfh_template t;
    int B, H, W;
    int rows_per_warp;   // R: a CTA covers 128 x 8R output pixels = 8R patches of 16 x 8
    int ntiles;          // CTAs per sample
    int vec4;            // W % 4 == 0 and all row bases 16-byte aligned
    int use_tma;         // the CTA's gt tile (loss) / logits tile (predict) is staged in smem by one TMA load
    int split_finalize;  // 1: CTAs only write their partials, k_train_finalize / k_comp_finalize (next launch) reduce them;
                         // 0: the sample's last-dispatched CTA reduces them inside this launch (tagged slots, no fence)
    int fin_slots;       // partial slots per sample read by k_train_finalize
    int short_last;      // H % (8R) != 0: the short bottom tiles of all samples are dispatched last (shorter tail)
    int lean, fast_free; // development switches (SFH_NO_LEAN / SFH_NO_FAST clear them): guard-free geometry in proven-safe
                         // patches; dedicated loop for the edge-free patches of full tiles
    // kEpiStore / kEpiBwd
    float* out_f;        // [B,C,H,W]
    const float* grad_out;
    float* dtheta;       // [B,9]
    // kEpiLoss
    const long long* gt;     // int64 class ids (reference dtype) ...
    const unsigned char* gt8; // ... or uint8 class ids (narrow surface); exactly one is set
    int nc, kind, nc_pow2;
    float inv_nc, invN;
    float* Lb;
    float* J;
    // optional in-kernel weighting + batch reduction of the training loss (models/losses.py:38-39,
    // train.py:196,213):  loss = rec_lambda*mean(L_b*w) + reproj_lambda*mean(R_b), dtheta = dloss/dtheta
    const void* weights;  // [B] fp32 or fp64, nullable (=> w = 1)
    int w_f64, w_outer;   // w_outer: weights came as [B,1] => the reference's [B,B] broadcast, w_eff = mean(w)
    float rec_lambda, reproj_lambda;
    float* loss_out;      // scalar, nullable: enables the combine stage
    float* dtheta_total;  // [B,9]
    double* contrib;      // workspace [B]
    // kEpiPredict
    const float* logits;
    int lh, lw, ratio;   // ratio: 1 (same size), 2 (H=2h, W=2w), 0 (score not fused)
    int32_t* out_i;          // int32 mask (reference dtype) ...
    unsigned char* out_u8;   // ... or uint8 mask (narrow surface, what predict.py:99 converts to anyway)
    float* score;
    // POI work, done by the first warp of tile 0 of every sample
    PoiParams poi;
    // workspace
    int* counters;
    float* partials;
};

// Geometry of one output pixel, everything the epilogues need.
struct Pix {
    Flow f;
    float ex, wx, sy, ny;   // bilinear factors: ex = x1-ix, wx = ix-x0, sy = y1-iy, ny = iy-y0
    int x0, y0;
};

template <int MODE>
__device__ __forceinline__ Pix pixel_geom(const Homog& Hm, float pu0, float pu3, float pu6, float v,
                                          float Wc_f, float Hc_f) {
    Pix p;
    p.f = flow_at(Hm, pu0, pu3, pu6, v);
    const float ix = unnormalize(p.f.x, Wc_f);
    const float iy = unnormalize(p.f.y, Hc_f);
    if (MODE == SFH_MODE_NEAREST) {
        p.x0 = __float2int_rn(ix);   // nearbyint: half to even
        p.y0 = __float2int_rn(iy);
        p.ex = p.wx = p.sy = p.ny = 0.f;
    } else {
        const float fx = floorf(ix), fy = floorf(iy);
        p.x0 = (int)fx;
        p.y0 = (int)fy;
        // ATen computes (x0+1)-ix and ix-x0; both are exact in fp32, and so is 1-(ix-x0)
        p.wx = __fsub_rn(ix, fx);
        p.ex = __fsub_rn(1.0f, p.wx);
        p.ny = __fsub_rn(iy, fy);
        p.sy = __fsub_rn(1.0f, p.ny);
    }
    return p;
}

// The same arithmetic without the guards (|Z| <= eps -> scale 1, 1/Z subnormal, non-finite / out-of-int-range
// coordinate -> -100).  Valid — and then bit-identical to pixel_geom — for pixels of a patch whose classification
// proved that none of them can fire: Z keeps its sign with |Z| well above eps and below 1e37 at the four corners
// (Z is affine in (u,v)), and the corner coordinates are finite and far inside the int range (x = X/Z and y = Y/Z
// are monotone along rows and columns while Z keeps its sign).  16 instructions shorter per pixel.
template <int MODE>
__device__ __forceinline__ Pix pixel_geom_lean(const Homog& Hm, float u, float v, float Wc_f, float Hc_f) {
    Pix p;
    p.f.X = __fadd_rn(__fmaf_rn(v, Hm.h[1], __fmul_rn(u, Hm.h[0])), Hm.h[2]);
    p.f.Y = __fadd_rn(__fmaf_rn(v, Hm.h[4], __fmul_rn(u, Hm.h[3])), Hm.h[5]);
    const float Z = __fadd_rn(__fmaf_rn(v, Hm.h[7], __fmul_rn(u, Hm.h[6])), Hm.h[8]);
    p.f.zok = true;
    p.f.s = rcp_rn_normal(Z);
    p.f.x = __fmul_rn(p.f.s, p.f.X);
    p.f.y = __fmul_rn(p.f.s, p.f.Y);
    const float ix = __fmul_rn(__fmaf_rn(__fadd_rn(p.f.x, 1.0f), Wc_f, -1.0f), 0.5f);
    const float iy = __fmul_rn(__fmaf_rn(__fadd_rn(p.f.y, 1.0f), Hc_f, -1.0f), 0.5f);
    if (MODE == SFH_MODE_NEAREST) {
        p.x0 = __float2int_rn(ix);
        p.y0 = __float2int_rn(iy);
        p.ex = p.wx = p.sy = p.ny = 0.f;
    } else {
        const float fx = floorf(ix), fy = floorf(iy);
        p.x0 = (int)fx;
        p.y0 = (int)fy;
        p.wx = __fsub_rn(ix, fx);
        p.ex = __fsub_rn(1.0f, p.wx);
        p.ny = __fsub_rn(iy, fy);
        p.sy = __fsub_rn(1.0f, p.ny);
    }
    return p;
}

// ATen accumulation order: nw, ne, sw, se, each step one FMA.
__device__ __forceinline__ float bilerp(const Pix& p, const TapVals& t) {
    float o = __fmul_rn(t.a, __fmul_rn(p.ex, p.sy));
    o = __fmaf_rn(t.b, __fmul_rn(p.wx, p.sy), o);
    o = __fmaf_rn(t.c, __fmul_rn(p.ex, p.ny), o);
    o = __fmaf_rn(t.d, __fmul_rn(p.wx, p.ny), o);
    return o;
}

// d(out)/d(theta) contribution of one pixel given g = dL/d(out):
// grid_sampler_2d_backward -> scale*p backward -> bmm backward.
struct GradAcc {
    float xu, x1, yu, y1, zu, z1;    // sums of gX*u, gX, gY*u, gY, gZ*u, gZ over the current row
    float a[9];                      // running dtheta
    __device__ __forceinline__ void zero() {
        xu = x1 = yu = y1 = zu = z1 = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = 0.f;
    }
    __device__ __forceinline__ void add(const Pix& p, float gix, float giy, float halfWc, float halfHc, float u) {
        const float gx = gix * halfWc, gy = giy * halfHc;
        const float gX = gx * p.f.s, gY = gy * p.f.s;
        const float gZ = p.f.zok ? -(gx * p.f.X + gy * p.f.Y) * p.f.s * p.f.s : 0.f;
        xu = fmaf(gX, u, xu); x1 += gX;
        yu = fmaf(gY, u, yu); y1 += gY;
        zu = fmaf(gZ, u, zu); z1 += gZ;
    }
    __device__ __forceinline__ void end_row(float v) {
        a[0] += xu; a[1] = fmaf(x1, v, a[1]); a[2] += x1;
        a[3] += yu; a[4] = fmaf(y1, v, a[4]); a[5] += y1;
        a[6] += zu; a[7] = fmaf(z1, v, a[7]); a[8] += z1;
        xu = x1 = yu = y1 = zu = z1 = 0.f;
    }
};

// Cross entropy of one pixel with 4 logits in registers: lse(l) - l[cls].  exp/log go through the
// MUFU ex2/lg2 units (rel. error ~1e-6 on terms <= 1, i.e. ~1e-6 absolute on a score of O(1);
// tests hold the per-sample mean to 1e-5 relative against torch's log_softmax).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ce4(float l0, float l1, float l2, float l3, int cls) {
    const float mx = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));
    const float k = 1.4426950408889634f, mk = -mx * k;     // exp(l - mx) = 2^(l*k - mx*k), arguments <= 0
    const float se = ex2_approx(fmaf(l0, k, mk)) + ex2_approx(fmaf(l1, k, mk)) +
                     ex2_approx(fmaf(l2, k, mk)) + ex2_approx(fmaf(l3, k, mk));
    const float sel = cls == 0 ? l0 : cls == 1 ? l1 : cls == 2 ? l2 : l3;
    return fmaf(__log2f(se), 0.6931471805599453f, mx) - sel;
}

// CE of the logits pixel this lane owns inside a patch's 4 x 8 logits block (ratio-2 predict path):
// lane -> (li = lane/8, lj = lane%8); its warp pixel is (row0 + 2*li, col0 + 2*lj).
template <bool FT>
__device__ __forceinline__ float ce_patch_lane(const float* s_lg, int cst, int pr, int pk, int lane, int cls,
                                               int row0, int col0, int H, int W) {
    const int li = lane >> 3, lj = lane & 7;
    if (!FT && !(row0 + 2 * li < H && col0 + 2 * lj < W)) return 0.f;
    const float* sl = s_lg + (pr * 4 + li) * (kTileW / 2) + pk * 8 + lj;
    return ce4(sl[0], sl[cst], sl[2 * cst], sl[3 * cst], cls);
}

// log-sum-exp cross entropy of one pixel, nc logits strided by `cs` (F.cross_entropy, reduction none).
__device__ __forceinline__ float ce_pixel(const float* lg, size_t cs, int nc, int cls) {
    if (nc == 4) return ce4(__ldcs(lg), __ldcs(lg + cs), __ldcs(lg + 2 * cs), __ldcs(lg + 3 * cs), cls);
    float mx = -INFINITY, sel = 0.f;
    for (int c = 0; c < nc; ++c) {
        const float v = __ldg(lg + c * cs);
        mx = fmaxf(mx, v);
        if (c == cls) sel = v;
    }
    float se = 0.f;
    for (int c = 0; c < nc; ++c) se += expf(__ldg(lg + c * cs) - mx);
    return (logf(se) + mx) - sel;
}

struct Gt4 { longlong2 lo, hi; };

template <bool FT>
__device__ __forceinline__ Gt4 load_gt(const long long* gt, size_t rowbase, int col, int W, bool vec, bool row_ok) {
    Gt4 g;
    if (FT || vec) {
        g.lo = __ldcs((const longlong2*)(gt + rowbase));
        g.hi = __ldcs((const longlong2*)(gt + rowbase) + 1);
    } else {
        g.lo.x = (row_ok && col + 0 < W) ? __ldcs(gt + rowbase + 0) : 0;
        g.lo.y = (row_ok && col + 1 < W) ? __ldcs(gt + rowbase + 1) : 0;
        g.hi.x = (row_ok && col + 2 < W) ? __ldcs(gt + rowbase + 2) : 0;
        g.hi.y = (row_ok && col + 3 < W) ? __ldcs(gt + rowbase + 3) : 0;
    }
    return g;
}

// ---- TMA / mbarrier (sm_90+ PTX; SASS: UTMALDG / SYNCS) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One 128-bit shared-memory load that the compiler may not split: with 32 B per lane the four 32-bit words of a
// quarter-warp's 128-bit access cover all banks once, whereas "load only the words that are used" (what nvcc makes
// of a longlong2 whose high halves are dead) puts lanes i, i+4, i+8, ... on the same bank — an 8-way conflict.
template <int OFF>
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int x, int y, int z, int w, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
        :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(w), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// Reduce 16 per-lane values across the warp with 16 shuffles: after the butterfly, lane L holds
// the warp total of value  idx(L) = bit4(L)*8 + bit3(L)*4 + bit2(L)*2 + bit1(L).
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = up ? v[i + half] : v[i];
            const float send = up ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

#ifndef SFH_PIPE_PIX
#define SFH_PIPE_PIX 1
#endif
#ifndef SFH_MINCTAS_HEAVY
#define SFH_MINCTAS_HEAVY 3   // loss / backward epilogues: <= 80 registers
#endif
#ifndef SFH_MINCTAS_LIGHT
#define SFH_MINCTAS_LIGHT 4   // store / predict epilogues: <= 64 registers
#endif
constexpr int min_ctas(int epi) { return (epi == kEpiStore || epi == kEpiPredict) ? SFH_MINCTAS_LIGHT : SFH_MINCTAS_HEAVY; }

constexpr int kEpochIdx = 65590; // launch epoch inside the fixed ticket area of the workspace (kTicketBytes / 4 = 65600 ints)
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// %laneid read again where it is needed: ptxas otherwise keeps `threadIdx.x & 31` alive across the whole patch loop of the
// 80-register epilogues and spills it (one local-memory reload per band / patch)
__device__ __forceinline__ int lane_now() {
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

constexpr int kMaxR = 16;        // bands (8 rows each) per CTA
constexpr int kPatchW = 16;      // a warp works on 16 x 8 pixel patches: lane = (ly 0..7, lx 0..3), 4 px per lane
constexpr float kBoxMargin = 1.0f / 64.0f;   // slack (in texels) on the patch bounding box, >> fp32 rounding of ix

// ------------------------------------------------------------------------------------------
// The fused kernel.
//
// CTA tile: 128 x 8R output pixels of one sample = 8R patches of 16 x 8.  A warp processes one
// patch at a time; a lane owns 4 consecutive pixels of one patch row, so 4 lanes cover 64 B
// (fp32/int32 out) or 128 B (int64 gt) of a row.
//
// Court templates are piecewise-constant class maps, so most patches sample one class only.  The
// CTA first classifies its patches: x(u,v) = X/Z and y(u,v) are monotone in u and in v as long as
// Z keeps its sign, hence a patch's sampling coordinates are bounded by their values at the 4
// corners; a summed-area table of "footprint straddles a class edge" over the packed template
// answers "any edge inside that bounding box?" with 4 loads.  Edge-free patches store the class
// value directly (their theta-gradient is exactly zero — in ATen the four taps cancel); only
// patches on class edges run the per-pixel homography + bilinear + chain rule.  Patches are then
// dealt to the 8 warps round-robin from a list with the expensive (edge) ones first, a static,
// deterministic schedule that keeps the warps of a CTA within one patch of each other.
//
// kEpiLoss: the CTA's int64 gt tile (128 x 8R x 8 B) is fetched by ONE TMA tensor load issued at
// kernel entry, so it streams in behind the classification prologue and the pixel loop never
// waits on a global load.
// ------------------------------------------------------------------------------------------
#ifdef SFH_TIMELINE
// Development aid (tools/timeline.py): per-CTA globaltimer stamps at the phase boundaries.
__device__ long long* g_timeline = nullptr;      // [ctas][8]
__device__ __forceinline__ void tl_stamp(int cta, int phase) {
    if (g_timeline && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_timeline[(size_t)cta * 8 + phase] = (long long)t;
        if (phase == 0) {
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            g_timeline[(size_t)cta * 8 + 7] = (long long)sm;
        }
    }
}
#define SFH_TL(phase) tl_stamp((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x, phase)
#else
#define SFH_TL(phase)
#endif

// (sample, tile row) of this CTA.  CTAs are dispatched x-fastest, then y, then z.  With `short_last`
// (the image height is not a multiple of the tile height) grid.y only spans the full-height tile rows
// and the short bottom row of every sample is appended as extra z-slices (slice B + s/full, row
// s % full holds sample s), so the short tiles form the end of the grid — a shorter tail — without
// any division.  b < 0: an unused slot of the last extra slice.
__device__ __forceinline__ void block_tile(int short_last, int B, int& b, int& ty) {
    b = blockIdx.z; ty = blockIdx.y;
    if (short_last && b >= B) {
        b = (b - B) * (int)gridDim.y + ty;
        ty = (int)gridDim.y;
        if (b >= B) b = -1;
    }
}

template <int FMT, int MODE, int EPI, bool FT>
__global__ void __launch_bounds__(kThreads, min_ctas(EPI)) k_fused(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(128) unsigned char s_dyn[];            // TMA destination (gt tile)
    __shared__ __align__(16) float s_tab[Taps<FMT>::kSmemFloats];
    __shared__ float s_red[kWarps][16];
    __shared__ double s_fin[kNPart][kFinGroup];
    __shared__ float s_gx[kMaxR + 1][kWarps + 1], s_gy[kMaxR + 1][kWarps + 1], s_gz[kMaxR + 1][kWarps + 1];
    __shared__ unsigned short s_items[kMaxR * kWarps];   // patch id | (class+1) << 8 | lean << 15, edge patches first
    __shared__ int s_ecnt[kWarps];
    __shared__ int s_nedge;
    __shared__ signed char s_cls[kMaxR * kWarps];        // class of an edge-free patch, -1: touches an edge
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned s_epoch;

    // the loss epilogue always samples bilinearly; its MODE argument carries the criterion instead
    constexpr int SMODE = (EPI == kEpiLoss) ? SFH_MODE_BILINEAR : MODE;
    constexpr bool kMse = (EPI != kEpiLoss) || (MODE == SFH_LOSS_MSE);
    int b, ty;
    const int tx = blockIdx.x;
    block_tile(p.short_last, p.B, b, ty);
    if (b < 0) return;
    const int tile = ty * gridDim.x + tx;
    const int H = p.H, W = p.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 3, ly = lane >> 2;
    const int R = p.rows_per_warp;
    const int band0 = ty * (8 * R);                       // first row of the CTA tile
    const int nitems = min(R, (H - band0 + 7) >> 3) * kWarps;   // patches with at least one row inside the image
    const float Wc_f = (float)p.t.width, Hc_f = (float)p.t.height;
    const bool tma = (EPI == kEpiLoss || EPI == kEpiPredict) && p.use_tma;

    SFH_TL(0);
    if (EPI != kEpiStore) asm volatile("griddepcontrol.launch_dependents;");   // lets the finalize grid become resident early
    Taps<FMT> taps;
    taps.build_tables(p.t, s_tab);           // from kernel parameters only: may run before the dependency wait
    // Programmatic dependent launch on the consumer side too: this grid may become resident while the
    // previous kernel of the stream (e.g. the previous step's finalize) is still running; everything
    // that touches global memory comes after this wait (no-op when launched without the attribute).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (EPI != kEpiStore && !p.split_finalize && p.counters && threadIdx.x == 32)   // read long before it is needed
        s_epoch = (unsigned)__ldcg(p.counters + kEpochIdx);                          // (the first barrier publishes it)
    if (tma && threadIdx.x == 0) {
        // the tile's streaming input is requested before anything else so that it arrives behind
        // the whole prologue (evict-first in L2: it is read exactly once)
        mbar_init(&s_bar, 1);
        const uint64_t pol = l2_policy_evict_first();
        if (EPI == kEpiLoss) {
            mbar_expect_tx(&s_bar, (uint32_t)(R * 8 * kTileW * (p.gt8 ? 1 : sizeof(long long))));
            tma_load_3d(s_dyn, &p.gt_map, tx * kTileW, band0, b, &s_bar, pol);
        } else {                             // logits [B,4,h,w]: 64 x 4R x 4 floats under this 128 x 8R tile
            mbar_expect_tx(&s_bar, (uint32_t)(R * 4 * (kTileW / 2) * 4 * sizeof(float)));
            tma_load_4d(s_dyn, &p.gt_map, tx * (kTileW / 2), band0 >> 1, 0, b, &s_bar, pol);
        }
    }
    if (EPI != kEpiBwd && tile == 0 && p.poi.court_poi)
        poi_block(p.poi, b);                 // first warp: the 52/33 court points in fp64
    Homog Hm;
    Hm.load(p.theta + 9 * b);

    const bool classify = (FMT != SFH_TMPL_F32) && (p.t.sat != nullptr);
    if (classify && threadIdx.x < (kWarps + 1) * (R + 1)) {
        // sampling coordinates on the (R+1) x 9 grid of patch corners (rows 8r, columns 16k)
        const int r = threadIdx.x / (kWarps + 1), k = threadIdx.x - r * (kWarps + 1);
        const int grow = min(band0 + 8 * r, H - 1), gcol = min(tx * kTileW + kPatchW * k, W - 1);
        const float gu = p.xs ? __ldg(p.xs + gcol) : mesh_coord(gcol, W);
        const float gv = p.ys ? __ldg(p.ys + grow) : mesh_coord(grow, H);
        const Flow f = flow_at(Hm, __fmul_rn(gu, Hm.h[0]), __fmul_rn(gu, Hm.h[3]), __fmul_rn(gu, Hm.h[6]), gv);
        s_gx[r][k] = __fmul_rn(__fmaf_rn(__fadd_rn(f.x, 1.0f), Wc_f, -1.0f), 0.5f);
        s_gy[r][k] = __fmul_rn(__fmaf_rn(__fadd_rn(f.y, 1.0f), Hc_f, -1.0f), 0.5f);
        s_gz[r][k] = f.zok ? __fdividef(1.0f, f.s) : __int_as_float(0x7fc00000);   // sign of Z; NaN poisons the patch
    }
    __syncthreads();                         // s_bar initialised, tables and corner grid written
    SFH_TL(1);
    taps.init(p.t, b, s_tab);
    int cls = -1;
    bool lean = false;                       // guards of the per-pixel geometry provably idle in this patch (pixel_geom_lean)
    if (classify && threadIdx.x < nitems) {
        const int r = threadIdx.x / kWarps, k = threadIdx.x % kWarps;
        const float x00 = s_gx[r][k], x01 = s_gx[r][k + 1], x10 = s_gx[r + 1][k], x11 = s_gx[r + 1][k + 1];
        const float y00 = s_gy[r][k], y01 = s_gy[r][k + 1], y10 = s_gy[r + 1][k], y11 = s_gy[r + 1][k + 1];
        const float z00 = s_gz[r][k], z01 = s_gz[r][k + 1], z10 = s_gz[r + 1][k], z11 = s_gz[r + 1][k + 1];
        const float xmin = fminf(fminf(x00, x01), fminf(x10, x11)), xmax = fmaxf(fmaxf(x00, x01), fmaxf(x10, x11));
        const float ymin = fminf(fminf(y00, y01), fminf(y10, y11)), ymax = fmaxf(fmaxf(y00, y01), fmaxf(y10, y11));
        // Z of one sign at the 4 corners (it is affine in (u,v)) => no horizon inside the patch
        const bool zpos = (z00 > 0.f) & (z01 > 0.f) & (z10 > 0.f) & (z11 > 0.f);
        const bool zneg = (z00 < 0.f) & (z01 < 0.f) & (z10 < 0.f) & (z11 < 0.f);
        const bool fin = (x00 == x00) & (x01 == x01) & (x10 == x10) & (x11 == x11) &
                         (y00 == y00) & (y01 == y01) & (y10 == y10) & (y11 == y11) &
                         (xmin > -1e9f) & (xmax < 1e9f) & (ymin > -1e9f) & (ymax < 1e9f);
        if ((zpos | zneg) & fin) {
            // |Z| over the patch >= the smallest corner |Z| (affine, one sign); the margin covers the fp32 rounding
            // of Z at interior pixels (<= 4 ulp of |h6|+|h7|+|h8|), so no pixel can take the |Z| <= eps or the
            // subnormal-reciprocal branch, and the corner bound keeps every coordinate far inside the int range
            const float Mz = fabsf(Hm.h[6]) + fabsf(Hm.h[7]) + fabsf(Hm.h[8]);
            const float zmin = fminf(fminf(fabsf(z00), fabsf(z01)), fminf(fabsf(z10), fabsf(z11)));
            lean = (zmin > 1e-6f * Mz + 1e-7f) & (Mz < 1e30f) & (p.lean != 0);
            // packed-template entries any pixel of the patch can touch (bilinear: floor+1,
            // nearest: rint+1 <= floor+2), clamped onto the all-zero border like the sampler does
            const int wmax = p.t.width + 1, hmax = p.t.height + 1;
            const int i0 = min(max(__float2int_rd(xmin - kBoxMargin) + 1, 0), wmax);
            const int i1 = min(max(__float2int_rd(xmax + kBoxMargin) + 2, 0), wmax);
            const int j0 = min(max(__float2int_rd(ymin - kBoxMargin) + 1, 0), hmax);
            const int j1 = min(max(__float2int_rd(ymax + kBoxMargin) + 2, 0), hmax);
            const unsigned* S = p.t.sat;
            const int sp = p.t.sat_pitch;
            const unsigned ec = taps.entry_class(i0, j0);          // issued with the 4 table loads
            const unsigned cnt = __ldg(S + (j1 + 1) * sp + (i1 + 1)) - __ldg(S + j0 * sp + (i1 + 1))
                               - __ldg(S + (j1 + 1) * sp + i0) + __ldg(S + j0 * sp + i0);
            if (cnt == 0u) cls = (int)ec;
        }
    }
    // ---- static schedule: list the edge patches first, warps then take entries w, w+8, ... ----
    const bool is_edge = (threadIdx.x < nitems) && (cls < 0);
    const unsigned bal = __ballot_sync(0xffffffffu, is_edge);
    if (lane == 0) s_ecnt[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < nitems) {
        int before = __popc(bal & ((1u << lane) - 1u)), total = 0;
#pragma unroll
        for (int w = 0; w < kMaxR * kWarps / 32; ++w) {
            const int c = s_ecnt[w];
            total += c;
            if (w < warp) before += c;
        }
        const int pos = is_edge ? before : total + ((int)threadIdx.x - before);
        if (threadIdx.x == 0) s_nedge = total;
        s_cls[threadIdx.x] = (signed char)cls;
        s_items[pos] = (unsigned short)(threadIdx.x | ((cls + 1) << 8) | (lean ? 0x8000 : 0));
    }
    __syncthreads();

    const int C = (FMT == SFH_TMPL_F32) ? p.t.channels : 1;
    const size_t base_b = (size_t)b * H * W;
    const float halfWc = 0.5f * Wc_f, halfHc = 0.5f * Hc_f;
    GradAcc acc;
    acc.zero();
    float loss_sum = 0.f, score_sum = 0.f;
    const float gscale = (kMse ? 2.0f : 1.0f) * p.invN;
    const float ncf = (float)p.nc;
    const long long* s_gt = reinterpret_cast<const long long*>(s_dyn);

    SFH_TL(2);
    if (tma) mbar_wait(&s_bar, 0);           // gt tile has landed (it streamed in behind the prologue)
    SFH_TL(3);

    // full tiles of packed templates: the list's leading `s_nedge` entries (the edge patches) go through the patch
    // loop, the edge-free rest is streamed band by band further down
    const bool fast = FT && classify && p.fast_free &&
                      (EPI == kEpiStore || EPI == kEpiLoss ||
                       (EPI == kEpiPredict && (!p.score || p.ratio == 0 || (p.ratio == 2 && tma))));
    const int n_generic = fast ? s_nedge : nitems;
#pragma unroll 1
    for (int it = warp; it < n_generic; it += kWarps) {
        const unsigned item = s_items[it];
        const int pr = (item & 0xffu) >> 3, pk = item & 7u, pc = (int)((item >> 8) & 0x7fu) - 1;
        const bool lean_patch = (item & 0x8000u) != 0u;
        const int row = band0 + pr * 8 + ly;
        const int col = tx * kTileW + pk * kPatchW + lx * 4;
        const bool row_ok = FT || row < H;
        const int rowc = FT ? row : min(row, H - 1);
        const bool vec = FT || (p.vec4 && row_ok && col + 3 < W);
        const size_t rowbase = base_b + (unsigned)(rowc * W + col);     // C == 1 offset (32-bit in-sample part)
#define SFH_PIX_OK(j) (FT || (row_ok && col + (j) < W))

        float tgt[4];
        if (EPI == kEpiLoss) {
            float gf[4];
            if (p.gt8) {
                uchar4 g8;
                if (tma) g8 = *reinterpret_cast<const uchar4*>(s_dyn + (pr * 8 + ly) * kTileW + pk * kPatchW + lx * 4);
                else if (vec) g8 = __ldcs(reinterpret_cast<const uchar4*>(p.gt8 + rowbase));
                else {
                    g8.x = (row_ok && col + 0 < W) ? p.gt8[rowbase + 0] : 0; g8.y = (row_ok && col + 1 < W) ? p.gt8[rowbase + 1] : 0;
                    g8.z = (row_ok && col + 2 < W) ? p.gt8[rowbase + 2] : 0; g8.w = (row_ok && col + 3 < W) ? p.gt8[rowbase + 3] : 0;
                }
                gf[0] = (float)g8.x; gf[1] = (float)g8.y; gf[2] = (float)g8.z; gf[3] = (float)g8.w;
            } else {
                Gt4 g;
                if (tma) {                   // zero-filled outside the image by the TMA unit
                    const longlong2* sp2 = reinterpret_cast<const longlong2*>(s_gt + (pr * 8 + ly) * kTileW + pk * kPatchW + lx * 4);
                    g.lo = sp2[0]; g.hi = sp2[1];
                } else {
                    g = load_gt<FT>(p.gt, rowbase, col, W, vec, row_ok);
                }
                gf[0] = (float)(int)g.lo.x; gf[1] = (float)(int)g.lo.y; gf[2] = (float)(int)g.hi.x; gf[3] = (float)(int)g.hi.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)      // class ids: the low 32 bits carry the value
                tgt[j] = p.nc_pow2 ? __fmul_rn(gf[j], p.inv_nc) : __fdiv_rn(gf[j], ncf);
        }

        if (pc >= 0) {
            // ================= edge-free patch: every pixel samples class `pc` =================
            const float cval = taps.class_value(pc);
            if (EPI == kEpiStore) {
                const size_t off = (((size_t)b * C) * H + rowc) * W + col;
                if (vec) __stcs((float4*)(p.out_f + off), make_float4(cval, cval, cval, cval));
                else
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_f[off + j] = cval;
            }
            if (EPI == kEpiLoss) {
                if (p.out_f) {
                    if (vec) __stcs((float4*)(p.out_f + rowbase), make_float4(cval, cval, cval, cval));
                    else
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_f[rowbase + j] = cval;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = cval - tgt[j];
                    float l = (kMse || fabsf(d) < 1.0f) ? d * d : 2.0f * fabsf(d) - 1.0f;
                    if (!SFH_PIX_OK(j)) l = 0.f;
                    loss_sum += l;
                }
            }
            if (EPI == kEpiPredict) {
                const int ci = __float2int_rz(__fmul_rn(cval, ncf));
                if (p.out_u8) {
                    if (vec) __stcs((uchar4*)(p.out_u8 + rowbase), make_uchar4(ci, ci, ci, ci));
                    else
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_u8[rowbase + j] = (unsigned char)ci;
                } else if (vec) __stcs((int4*)(p.out_i + rowbase), make_int4(ci, ci, ci, ci));
                else
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_i[rowbase + j] = ci;
                if (p.score && p.ratio == 1 && row_ok) {
                    const size_t cs = (size_t)p.lh * p.lw;
                    const float* lg = p.logits + (size_t)b * p.nc * cs + (size_t)row * p.lw + col;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (FT || col + j < W) score_sum += ce_pixel(lg + j, cs, p.nc, ci);
                } else if (p.score && p.ratio == 2 && tma) {
                    // staged logits: the patch's 4 x 8 logits pixels are dealt one per lane
                    score_sum += ce_patch_lane<FT>(reinterpret_cast<const float*>(s_dyn), R * 4 * (kTileW / 2), pr, pk, lane,
                                                   ci, band0 + pr * 8, tx * kTileW + pk * kPatchW, H, W);
                } else if (p.score && p.ratio == 2 && row_ok && !(row & 1)) {
                    {
                        const size_t cs = (size_t)p.lh * p.lw;
                        const float* lg = p.logits + (size_t)b * p.nc * cs + (size_t)(row >> 1) * p.lw + (col >> 1);
                        if (FT || col < W) score_sum += ce_pixel(lg, cs, p.nc, ci);
                        if (FT || col + 2 < W) score_sum += ce_pixel(lg + 1, cs, p.nc, ci);
                    }
                }
            }
            continue;                        // kEpiBwd: zero gradient, grad_out is not even read
        }

        // ======================= per-pixel path (patch touches a class edge) ====================
        const float v = p.ys ? __ldg(p.ys + rowc) : mesh_coord(rowc, H);
        float u[4];
        if (FT && p.xs) {
            const float4 u4 = __ldg(reinterpret_cast<const float4*>(p.xs + col));
            u[0] = u4.x; u[1] = u4.y; u[2] = u4.z; u[3] = u4.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cx = min(col + j, W - 1);
                u[j] = p.xs ? __ldg(p.xs + cx) : mesh_coord(cx, W);
            }
        }
        Pix px[4];
        if (EPI != kEpiLoss) {
            if (lean_patch) {                // warp-uniform: the guards cannot fire anywhere in this patch
#pragma unroll
                for (int j = 0; j < 4; ++j) px[j] = pixel_geom_lean<SMODE>(Hm, u[j], v, Wc_f, Hc_f);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)  // u*h{0,3,6}: the first product of the bmm chain
                    px[j] = pixel_geom<SMODE>(Hm, __fmul_rn(u[j], Hm.h[0]), __fmul_rn(u[j], Hm.h[3]), __fmul_rn(u[j], Hm.h[6]),
                                             v, Wc_f, Hc_f);
            }
        }

        if (EPI == kEpiStore || EPI == kEpiBwd) {
            float gix[4] = {0.f, 0.f, 0.f, 0.f}, giy[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < C; ++c) {
                const size_t off = (((size_t)b * C + c) * H + rowc) * W + col;
                float o[4], go[4];
                if (EPI == kEpiBwd) {
                    if (vec) {
                        const float4 t4 = __ldcs((const float4*)(p.grad_out + off));
                        go[0] = t4.x; go[1] = t4.y; go[2] = t4.z; go[3] = t4.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) go[j] = SFH_PIX_OK(j) ? __ldcs(p.grad_out + off + j) : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (SMODE == SFH_MODE_NEAREST) {
                        o[j] = taps.fetch1(c, px[j].x0, px[j].y0);
                    } else {
                        const TapVals t = taps.fetch4(c, px[j].x0, px[j].y0);
                        if (EPI == kEpiStore) o[j] = bilerp(px[j], t);
                        if (EPI == kEpiBwd) {
                            gix[j] += ((t.b - t.a) * px[j].sy + (t.d - t.c) * px[j].ny) * go[j];
                            giy[j] += ((t.c - t.a) * px[j].ex + (t.d - t.b) * px[j].wx) * go[j];
                        }
                    }
                }
                if (EPI == kEpiStore) {
                    if (vec) __stcs((float4*)(p.out_f + off), make_float4(o[0], o[1], o[2], o[3]));
                    else
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_f[off + j] = o[j];
                }
            }
            if (EPI == kEpiBwd) {
#pragma unroll
                for (int j = 0; j < 4; ++j)      // go is 0 outside the image, so no extra predicate
                    acc.add(px[j], gix[j], giy[j], halfWc, halfHc, u[j]);
                acc.end_row(v);
            }
        }

        if (EPI == kEpiLoss) {
            // one pixel at a time, start to finish (geometry -> taps -> value -> loss -> chain rule):
            // only o[j] crosses iterations, which keeps the live state small enough for 4 CTAs/SM
            float o[4];
            bool any = false;
            // packed templates: the geometry and the packed-entry load of pixel j+1 are issued before
            // pixel j's value / loss / chain rule, so one load is always in flight (SFH_PIPE_PIX)
            constexpr bool kPipe = (FMT != SFH_TMPL_F32) && (SFH_PIPE_PIX != 0);
            // value, loss and chain rule of one pixel, given its geometry and taps
            auto finish = [&](int j, const Pix& q, const TapVals& t) {
                o[j] = bilerp(q, t);
                const float d = o[j] - tgt[j];
                float l, g;
                if (kMse || fabsf(d) < 1.0f) {   // MSE d^2 ; SmoothL1(beta=1) 0.5 d^2
                    l = d * d; g = d;
                } else {
                    l = 2.0f * fabsf(d) - 1.0f; g = d > 0.f ? 1.0f : -1.0f;   // doubled, halved below
                }
                if (!SFH_PIX_OK(j)) { l = 0.f; g = 0.f; }
                loss_sum += l;
                // the gradient lives on footprints that straddle a class edge; a uniform footprint
                // cancels exactly (a*sy - a*sy), so the warp skips the chain rule for every pixel slot
                // on which no lane straddles one (a vertical court line touches 1-2 of the 4 slots)
                if (__any_sync(0xffffffffu, !t.uni)) {
                    g *= gscale;
                    const float gix = ((t.b - t.a) * q.sy + (t.d - t.c) * q.ny) * g;
                    const float giy = ((t.c - t.a) * q.ex + (t.d - t.b) * q.wx) * g;
                    acc.add(q, gix, giy, halfWc, halfHc, u[j]);
                    any = true;
                }
            };
            if (kPipe && lean_patch && p.lean >= 2) {
                // Guard-free patches: the kernel is bound by the latency of the packed-entry loads (60 % L1 hits, the
                // rest from L2), not by issue slots.  So the four texel addresses are computed first and all four
                // loads fly together; the geometry is then evaluated a second time for the weights (24 instructions
                // per pixel — cheaper than keeping four sets of weights alive in registers at 3 CTAs/SM).
                unsigned vq4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const Pix q = pixel_geom_lean<SMODE>(Hm, u[j], v, Wc_f, Hc_f);
                    vq4[j] = taps.quad(q.x0, q.y0);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float uu = u[j];
                    asm volatile("" : "+f"(uu));             // a fresh value for the compiler: recompute, do not keep
                    const Pix q = pixel_geom_lean<SMODE>(Hm, uu, v, Wc_f, Hc_f);
                    finish(j, q, taps.decode(vq4[j]));
                }
            } else {
                Pix qn;
                unsigned vn = 0u;
                if (kPipe) {
                    qn = pixel_geom<SMODE>(Hm, __fmul_rn(u[0], Hm.h[0]), __fmul_rn(u[0], Hm.h[3]), __fmul_rn(u[0], Hm.h[6]), v, Wc_f, Hc_f);
                    vn = taps.quad(qn.x0, qn.y0);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    Pix q;
                    TapVals t;
                    if (kPipe) {
                        q = qn;
                        const unsigned vq = vn;
                        if (j < 3) {
                            qn = pixel_geom<SMODE>(Hm, __fmul_rn(u[j + 1], Hm.h[0]), __fmul_rn(u[j + 1], Hm.h[3]),
                                                   __fmul_rn(u[j + 1], Hm.h[6]), v, Wc_f, Hc_f);
                            vn = taps.quad(qn.x0, qn.y0);
                        }
                        t = taps.decode(vq);
                    } else {
                        q = pixel_geom<SMODE>(Hm, __fmul_rn(u[j], Hm.h[0]), __fmul_rn(u[j], Hm.h[3]),
                                              __fmul_rn(u[j], Hm.h[6]), v, Wc_f, Hc_f);
                        t = taps.fetch4(0, q.x0, q.y0);
                    }
                    finish(j, q, t);
                }
            }
            if (p.out_f) {
                if (vec) __stcs((float4*)(p.out_f + rowbase), make_float4(o[0], o[1], o[2], o[3]));
                else
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_f[rowbase + j] = o[j];
            }
            if (any) acc.end_row(v);
        }

        if (EPI == kEpiPredict) {
            int ci[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o;
                if (SMODE == SFH_MODE_NEAREST) o = taps.fetch1(0, px[j].x0, px[j].y0);
                else o = bilerp(px[j], taps.fetch4(0, px[j].x0, px[j].y0));
                ci[j] = __float2int_rz(__fmul_rn(o, ncf));   // (warp*nc).int()
            }
            if (p.out_u8) {
                if (vec) __stcs((uchar4*)(p.out_u8 + rowbase), make_uchar4(ci[0], ci[1], ci[2], ci[3]));
                else
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_u8[rowbase + j] = (unsigned char)ci[j];
            } else if (vec) __stcs((int4*)(p.out_i + rowbase), make_int4(ci[0], ci[1], ci[2], ci[3]));
            else
#pragma unroll
                for (int j = 0; j < 4; ++j) if (SFH_PIX_OK(j)) p.out_i[rowbase + j] = ci[j];
            if (p.score && p.ratio == 1 && row_ok) {
                const size_t cs = (size_t)p.lh * p.lw;
                const float* lg = p.logits + (size_t)b * p.nc * cs + (size_t)row * p.lw + col;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (FT || col + j < W) score_sum += ce_pixel(lg + j, cs, p.nc, ci[j]);
            } else if (p.score && p.ratio == 2 && tma) {
                // F.interpolate(nearest) to (H/2, W/2) picks source pixel (2i, 2j); with the logits
                // staged in smem the patch's 4 x 8 logits pixels are dealt one per lane, the class of
                // pixel (2i, 2j) comes from its owner lane by shuffle
                const int src = ((lane >> 3) << 3) + ((lane & 7) >> 1);       // owner: ly = 2*li, lx = lj/2
                const int c0 = __shfl_sync(0xffffffffu, ci[0], src), c2 = __shfl_sync(0xffffffffu, ci[2], src);
                score_sum += ce_patch_lane<FT>(reinterpret_cast<const float*>(s_dyn), R * 4 * (kTileW / 2), pr, pk, lane,
                                               (lane & 1) ? c2 : c0, band0 + pr * 8, tx * kTileW + pk * kPatchW, H, W);
            } else if (p.score && p.ratio == 2 && row_ok && !(row & 1)) {
                // F.interpolate(nearest) to (H/2, W/2) picks source pixel (2i, 2j)
                {
                    const size_t cs = (size_t)p.lh * p.lw;
                    const float* lg = p.logits + (size_t)b * p.nc * cs + (size_t)(row >> 1) * p.lw + (col >> 1);
                    if (FT || col < W) score_sum += ce_pixel(lg, cs, p.nc, ci[0]);
                    if (FT || col + 2 < W) score_sum += ce_pixel(lg + 1, cs, p.nc, ci[2]);
                }
            }
        }
    }
#undef SFH_PIX_OK
    // ---- edge-free patches of full tiles, streamed band by band -------------------------------------------
    // A band is one row of 8 patches (8 x 128 px).  A warp walks its bands row by row with lane i on pixels
    // 4i..4i+3 of the row — 32 consecutive bytes of the staged int64 gt per lane, so the two 128-bit smem loads of
    // a row are bank-conflict free (a patch-shaped access hits rows 1 KiB apart: same banks) and the global store is
    // one contiguous 512 B line per row.  A lane's patch is column block i/4; lanes whose patch touches a class edge
    // are predicated off (their patch went through the loop above).  Per patch this costs ~20 instructions instead
    // of ~70: the per-patch bookkeeping is paid once per band.  Placed after the edge-patch loop, where the
    // homography and the chain rule's row sums are dead, so it runs with few live registers.
    if (fast) {
        const int ln = lane_now();
        const int nbands = nitems >> 3;
        const float inv_nc = p.inv_nc;
        const bool pow2 = p.nc_pow2 != 0;
        const bool hot = EPI == kEpiLoss && tma && !p.gt8 && pow2 && p.out_f != nullptr;
#pragma unroll 1
        for (int r = warp; r < nbands; r += kWarps) {
            const int pc = s_cls[r * kWarps + (ln >> 2)];
            const bool act = pc >= 0;
            const float cval = taps.class_value(act ? pc : 0);
            const float4 c4 = make_float4(cval, cval, cval, cval);
            size_t o = base_b + (size_t)(band0 + r * 8) * W + (tx * kTileW + ln * 4);
            if (EPI == kEpiStore) {
#pragma unroll
                for (int k = 0; k < 8; ++k, o += W)
                    if (act) __stcs((float4*)(p.out_f + o), c4);
            }
            if (EPI == kEpiLoss) {
                int so = (r * 8) * kTileW + ln * 4;                    // element offset inside the staged tile
                if (hot) {
                    // the reference's own surface (int64 gt staged by TMA, nc a power of two, warp_mask wanted):
                    // 2 LDS.128 + 4 I2F + 4 FFMA + 4 FFMA + 1 STG.128 per row.  fma(g, -1/nc, cval) == cval - g/nc
                    // bit for bit here: g/nc is exact for a power-of-two nc.
                    // Loads and stores use different ln->pixel maps.  A 128-bit smem load is conflict-free only if
                    // the 8 lanes of a quarter-warp cover 128 contiguous bytes, so ln i loads gt pixels {2i, 2i+1}
                    // and {64+2i, 64+2i+1} (patches i/8 and 4+i/8); the store stays one 128-bit line-friendly write of
                    // pixels 4i..4i+3 (patch i/4).  The loss does not care which ln adds which pixel.
                    const int pa = s_cls[r * kWarps + (ln >> 3)], pb = s_cls[r * kWarps + 4 + (ln >> 3)];
                    const float ca = taps.class_value(pa >= 0 ? pa : 0), cb = taps.class_value(pb >= 0 ? pb : 0);
                    const uint32_t sa = smem_u32(s_gt + (r * 8) * kTileW + ln * 2);
                    float* op = p.out_f + o;
                    const float ninv = -inv_nc;
                    float la = 0.f, lb = 0.f;
#define SFH_HOT_ROW(k) {                                                                                           \
                        const uint4 lo = lds128<(k) * kTileW * 8>(sa), hi = lds128<(k) * kTileW * 8 + 512>(sa);               \
                        if (act) __stcs((float4*)(op + (size_t)(k) * W), c4);                                                 \
                        const float d0 = fmaf((float)(int)lo.x, ninv, ca), d1 = fmaf((float)(int)lo.z, ninv, ca);             \
                        const float d2 = fmaf((float)(int)hi.x, ninv, cb), d3 = fmaf((float)(int)hi.z, ninv, cb);             \
                        if (kMse) {                                                                                           \
                            la = fmaf(d0, d0, la); la = fmaf(d1, d1, la); lb = fmaf(d2, d2, lb); lb = fmaf(d3, d3, lb);       \
                        } else {                                                                                              \
                            la += fabsf(d0) < 1.0f ? d0 * d0 : 2.0f * fabsf(d0) - 1.0f;                                       \
                            la += fabsf(d1) < 1.0f ? d1 * d1 : 2.0f * fabsf(d1) - 1.0f;                                       \
                            lb += fabsf(d2) < 1.0f ? d2 * d2 : 2.0f * fabsf(d2) - 1.0f;                                       \
                            lb += fabsf(d3) < 1.0f ? d3 * d3 : 2.0f * fabsf(d3) - 1.0f;                                       \
                        }                                                                                                     \
                    }
                    SFH_HOT_ROW(0) SFH_HOT_ROW(1) SFH_HOT_ROW(2) SFH_HOT_ROW(3)
                    SFH_HOT_ROW(4) SFH_HOT_ROW(5) SFH_HOT_ROW(6) SFH_HOT_ROW(7)
#undef SFH_HOT_ROW
                    loss_sum += (pa >= 0 ? la : 0.f) + (pb >= 0 ? lb : 0.f);              // edge patches: handled above
                    continue;
                }
#pragma unroll 2
                for (int k = 0; k < 8; ++k, o += W, so += kTileW) {
                    if (!act) continue;
                    float gf[4];
                    if (p.gt8) {
                        const uchar4 g8 = tma ? *reinterpret_cast<const uchar4*>(s_dyn + so)
                                              : __ldcs(reinterpret_cast<const uchar4*>(p.gt8 + o));
                        gf[0] = (float)g8.x; gf[1] = (float)g8.y; gf[2] = (float)g8.z; gf[3] = (float)g8.w;
                    } else {
                        longlong2 lo, hi;
                        if (tma) {
                            const longlong2* sp2 = reinterpret_cast<const longlong2*>(s_gt + so);
                            lo = sp2[0]; hi = sp2[1];
                        } else {
                            lo = __ldcs((const longlong2*)(p.gt + o)); hi = __ldcs((const longlong2*)(p.gt + o) + 1);
                        }
                        gf[0] = (float)(int)lo.x; gf[1] = (float)(int)lo.y; gf[2] = (float)(int)hi.x; gf[3] = (float)(int)hi.y;
                    }
                    if (p.out_f) __stcs((float4*)(p.out_f + o), c4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d = cval - (pow2 ? __fmul_rn(gf[j], inv_nc) : __fdiv_rn(gf[j], ncf));
                        loss_sum += (kMse || fabsf(d) < 1.0f) ? d * d : 2.0f * fabsf(d) - 1.0f;
                    }
                }
            }
            if (EPI == kEpiPredict) {
                const int ci = __float2int_rz(__fmul_rn(cval, ncf));
#pragma unroll
                for (int k = 0; k < 8; ++k, o += W) {
                    if (!act) continue;
                    if (p.out_u8) __stcs((uchar4*)(p.out_u8 + o), make_uchar4(ci, ci, ci, ci));
                    else __stcs((int4*)(p.out_i + o), make_int4(ci, ci, ci, ci));
                }
                if (p.score && p.ratio == 2 && act) {
                    // the band's 4 x 64 staged logits pixels: ln i owns pixels 2i, 2i+1 of each row; both sit
                    // under this ln's patch (F.interpolate(nearest) picks source pixel (2i, 2j))
                    const float* sl = reinterpret_cast<const float*>(s_dyn) + (r * 4) * (kTileW / 2) + ln * 2;
                    const int cst = R * 4 * (kTileW / 2);
#pragma unroll
                    for (int k = 0; k < 4; ++k, sl += kTileW / 2) {
                        const float2 l0 = *reinterpret_cast<const float2*>(sl), l1 = *reinterpret_cast<const float2*>(sl + cst);
                        const float2 l2 = *reinterpret_cast<const float2*>(sl + 2 * cst), l3 = *reinterpret_cast<const float2*>(sl + 3 * cst);
                        score_sum += ce4(l0.x, l1.x, l2.x, l3.x, ci) + ce4(l0.y, l1.y, l2.y, l3.y, ci);
                    }
                }
            }
        }
    }

    if (EPI == kEpiStore) return;
    if (EPI == kEpiPredict && !(p.score && p.ratio != 0)) return;
    SFH_TL(4);

    // ---------------- per-sample reduction: lane -> warp -> CTA partial -> last CTA -----------
    // (sample, tile) are recomputed here so that they are not live across the patch loop
    int bE, tyE;
    {
        int sl = p.short_last;
        asm volatile("" : "+r"(sl));
        block_tile(sl, p.B, bE, tyE);
    }
    const int tileE = tyE * gridDim.x + tx;
    if (EPI == kEpiLoss && !kMse) loss_sum *= 0.5f;
    {
        float vals[16];
        vals[0] = loss_sum;
#pragma unroll
        for (int k = 0; k < 9; ++k) vals[1 + k] = acc.a[k];
        vals[10] = score_sum;
#pragma unroll
        for (int k = 11; k < 16; ++k) vals[k] = 0.f;
        const int ln = lane_now();
        const float tot = warp_reduce16(vals, ln);
        if (!(ln & 1)) s_red[threadIdx.x >> 5][((ln >> 4) & 1) * 8 + ((ln >> 3) & 1) * 4 + ((ln >> 2) & 1) * 2 + ((ln >> 1) & 1)] = tot;
    }
    __syncthreads();
    // One partial per (sample, tile, component) in a fixed slot.  In-launch reduction (split_finalize == 0):
    // a slot is ONE 64-bit word {tag, fp32 value}; the tag is this launch's epoch, so a reader that sees the
    // tag has the value as well (single-copy atomicity of an aligned 8-byte store) and the 1900+ producer
    // CTAs retire without any fence or ticket.  The sample's last-dispatched CTA (highest linear block index
    // among the sample's tiles; CTAs are dispatched in index order, so everything it waits for is already
    // resident or done) polls the sample's slots and adds them in fixed order in fp64.
    const unsigned want = p.split_finalize ? 0u : s_epoch + 1u;
    if (threadIdx.x < kNPart) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += s_red[w][threadIdx.x];
        const size_t slot = ((size_t)bE * p.ntiles + tileE) * kNPart + threadIdx.x;
        if (p.split_finalize) __stcg(p.partials + slot, s);
        else st_relaxed_u64(reinterpret_cast<unsigned long long*>(p.partials) + slot,
                            ((unsigned long long)want << 32) | (unsigned long long)__float_as_uint(s));
    }
    SFH_TL(5); SFH_TL(6);
    if (p.split_finalize) return;                        // reduced by k_train_finalize / k_comp_finalize (next launch)
    if (tileE != p.ntiles - 1) return;
    {
        // components this epilogue produces: loss + 9 dtheta terms / the 9 dtheta terms / the score
        constexpr int kLo = (EPI == kEpiLoss) ? 0 : (EPI == kEpiBwd) ? 1 : 10;
        constexpr int kHi = (EPI == kEpiPredict) ? 11 : 10;
        const int k = threadIdx.x / kFinGroup, jj = threadIdx.x % kFinGroup;
        if (k < kNPart) s_fin[k][jj] = 0.0;
        if (k >= kLo && k < kHi) {
            double s = 0.0;
            const unsigned long long* base = reinterpret_cast<const unsigned long long*>(p.partials) + (size_t)bE * p.ntiles * kNPart + k;
            for (int t0 = jj; t0 < p.ntiles; t0 += 4 * kFinGroup) {
                unsigned long long w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {            // four polls in flight
                    const int tt = t0 + u * kFinGroup;
                    w[u] = (tt < p.ntiles) ? ld_relaxed_u64(base + (size_t)tt * kNPart) : ((unsigned long long)want << 32);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int tt = t0 + u * kFinGroup;
                    while ((unsigned)(w[u] >> 32) != want) w[u] = ld_relaxed_u64(base + (size_t)tt * kNPart);
                    s += (double)__uint_as_float((unsigned)w[u]);
                }
            }
            s_fin[k][jj] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < kNPart) {
        const int k = threadIdx.x;
        double s = 0.0;
#pragma unroll
        for (int jj = 0; jj < kFinGroup; ++jj) s += s_fin[k][jj];
        if (EPI == kEpiBwd) {
            if (k >= 1 && k <= 9) p.dtheta[9 * bE + k - 1] = (float)s;
        } else if (EPI == kEpiLoss) {
            if (k == 0) { s = s / ((double)H * (double)W); p.Lb[bE] = (float)s; }
            else if (k <= 9) p.J[9 * bE + k - 1] = (float)s;
            s_fin[k][0] = s;
        } else if (EPI == kEpiPredict) {
            if (k == 10) p.score[bE] = (float)(s / ((double)p.lh * (double)p.lw));
        }
    }
    bool counted = false;
    if (EPI == kEpiLoss && p.loss_out) {
        // ---- weighting + batch mean + total dtheta, still inside the same launch ------------
        __syncthreads();
        counted = true;
        if (warp == 0) {
            const int B = p.B;
            double w_eff = 1.0;
            if (p.weights) {
                if (p.w_outer) {            // [B]*[B,1] -> [B,B] broadcast quirk: every sample sees mean(w)
                    double sw = 0.0;
                    for (int i = lane; i < B; i += 32)
                        sw += p.w_f64 ? ((const double*)p.weights)[i] : (double)((const float*)p.weights)[i];
                    w_eff = warp_sum(sw) / (double)B;
                } else {
                    w_eff = p.w_f64 ? ((const double*)p.weights)[bE] : (double)((const float*)p.weights)[bE];
                }
            }
            const bool rep = p.poi.gt_poi != nullptr;
            const double cr = (double)p.rec_lambda * w_eff, cp = (double)p.reproj_lambda;
            // dR_b/dtheta (lanes 0-8) and R_b (lane 9) come from the sample's POI warp (tile 0) as tagged words
            double KR = 0.0;
            if (rep && lane < 10) {
                unsigned long long w;
                do { w = ld_relaxed_u64(p.poi.pub + 10 * (size_t)bE + lane); } while ((unsigned)(w >> 32) != want);
                KR = (double)__uint_as_float((unsigned)w);
            }
            const double Rv = __shfl_sync(0xffffffffu, KR, 9);
            if (lane < 9) p.dtheta_total[9 * bE + lane] = (float)((cr * s_fin[1 + lane][0] + cp * KR) / (double)B);
            int last2 = 0;
            if (lane == 0) {
                __stcg(p.contrib + bE, cr * s_fin[0][0] + cp * Rv);
                last2 = (ticket_release(p.counters + B) == B - 1);    // B fences per launch (one per sample), not one per CTA
            }
            last2 = __shfl_sync(0xffffffffu, last2, 0);
            if (last2) {                    // last sample of the batch: fixed-order sum over bE
                __threadfence();
                double s = 0.0;
                for (int i = lane; i < B; i += 32) s += __ldcg(p.contrib + i);
                s = warp_sum(s);
                if (lane == 0) {
                    *p.loss_out = (float)(s / (double)B);
                    p.counters[B] = 0;
                    p.counters[kEpochIdx] = (int)want;       // next launch's tag
                }
            }
        }
    }
    // every sample reduced => advance the epoch (the slots of this launch become stale for the next one)
    if (!counted && threadIdx.x == 0 && atomicAdd(p.counters + p.B, 1) == p.B - 1) {
        p.counters[p.B] = 0;
        p.counters[kEpochIdx] = (int)want;
    }
}

// Second launch of the training tail: one CTA per sample adds that sample's per-tile partials in
// fixed order in fp64, then the optional weighting + batch mean + total dtheta (last CTA by ticket).
constexpr int kFinThreads = 288;                 // 3 float4 columns of a slot x 96 slot lanes
__global__ void __launch_bounds__(kFinThreads) k_train_finalize(const __grid_constant__ FusedParams p) {
    __shared__ double s_w[kFinThreads / 32][4];
    __shared__ double s_out[kNPart];
    const int b = blockIdx.x, H = p.H, W = p.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nslots = p.fin_slots;
    asm volatile("griddepcontrol.launch_dependents;");   // the next kernel of the stream may become resident (it waits itself)
    asm volatile("griddepcontrol.wait;" ::: "memory");   // programmatic dependent launch: the producer grid has completed
    // warp 0 requests what the combine stage needs (weights, dR_b/dtheta, R_b) together with the partials,
    // so the tail of this kernel is one L2 round trip shorter
    const bool rep = p.poi.gt_poi != nullptr;
    double w_eff = 1.0, Kk = 0.0, Rv = 0.0;
    if (p.loss_out && warp == 0) {
        if (p.weights) {
            if (p.w_outer) {            // [B]*[B,1] -> [B,B] broadcast quirk: every sample sees mean(w)
                double sw = 0.0;
                for (int i = lane; i < p.B; i += 32)
                    sw += p.w_f64 ? ((const double*)p.weights)[i] : (double)((const float*)p.weights)[i];
                w_eff = warp_sum(sw) / (double)p.B;
            } else {
                w_eff = p.w_f64 ? ((const double*)p.weights)[b] : (double)((const float*)p.weights)[b];
            }
        }
        if (rep && lane < 9) Kk = (double)__ldcg(p.poi.K + 9 * b + lane);
        if (rep && lane == 0) Rv = (double)__ldcg(p.poi.Rb + b);
    }
    {
        // thread (cq, g): float4 column cq of slots g, g+96, ...; all loads independent (one L2 round trip)
        const int cq = threadIdx.x / 96, g = threadIdx.x % 96;
        const float4* base = reinterpret_cast<const float4*>(p.partials + (size_t)b * nslots * kNPart) + cq;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
        for (int t = g; t < nslots; t += 96) {
            const float4 v = __ldcg(base + (size_t)t * (kNPart / 4));
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
        if (lane == 0) { s_w[warp][0] = a0; s_w[warp][1] = a1; s_w[warp][2] = a2; s_w[warp][3] = a3; }
    }
    __syncthreads();
    if (threadIdx.x < kNPart) {
        const int k = threadIdx.x, cq = k >> 2, c = k & 3;
        double s = (s_w[3 * cq][c] + s_w[3 * cq + 1][c]) + s_w[3 * cq + 2][c];
        if (k == 0) { s = s / ((double)H * (double)W); p.Lb[b] = (float)s; }
        else if (k <= 9) p.J[9 * b + k - 1] = (float)s;
        s_out[k] = s;
    }
    if (!p.loss_out) return;
    __syncthreads();
    if (warp == 0) {
        const int B = p.B;
        const double cr = (double)p.rec_lambda * w_eff, cp = (double)p.reproj_lambda;
        if (lane < 9) p.dtheta_total[9 * b + lane] = (float)((cr * s_out[1 + lane] + cp * Kk) / (double)B);
        int last2 = 0;
        if (lane == 0) {
            __stcg(p.contrib + b, cr * s_out[0] + cp * Rv);
            last2 = (ticket_release(p.counters + B) == B - 1);
        }
        last2 = __shfl_sync(0xffffffffu, last2, 0);
        if (last2) {
            __threadfence();
            double s = 0.0;
            for (int i = lane; i < B; i += 32) s += __ldcg(p.contrib + i);
            s = warp_sum(s);
            if (lane == 0) { *p.loss_out = (float)(s / (double)B); p.counters[B] = 0; }
        }
    }
}

// Consistency score for logits sizes the fused pass does not cover (any h,w): reads the int32
// mask back (L2-hot) with upsample_nearest's index rule.  models/reconstructor.py:230-238.
__global__ void __launch_bounds__(kThreads) k_consistency_generic(const int32_t* mask, const unsigned char* mask8, const float* logits,
                                                                  int nc, int H, int W, int lh, int lw,
                                                                  float* score) {
    __shared__ double s_w[kWarps];
    const int b = blockIdx.x;
    const size_t cs = (size_t)lh * lw;
    const float sy = (float)H / (float)lh, sx = (float)W / (float)lw;
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < cs; i += kThreads) {
        const int r = (int)(i / lw), c = (int)(i % lw);
        int sr = (lh == H) ? r : (lh == 2 * H) ? (r >> 1) : min((int)floorf(r * sy), H - 1);
        int sc = (lw == W) ? c : (lw == 2 * W) ? (c >> 1) : min((int)floorf(c * sx), W - 1);
        const int cls = mask8 ? (int)mask8[((size_t)b * H + sr) * W + sc] : mask[((size_t)b * H + sr) * W + sc];
        acc += (double)ce_pixel(logits + (size_t)b * nc * cs + i, cs, nc, cls);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += s_w[w];
        score[b] = (float)(s / (double)cs);
    }
}

// Build the quad-packed palette-index template (see sfh_template in the header).
template <int BITS, typename T>
__global__ void k_pack(const float* tmpl, int Hc, int Wc, T* q, int pitch, int npal,
                       const __grid_constant__ sfh_template pal, int32_t* err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // packed column 0..Wc+1 (last one all zero)
    const int j = blockIdx.y;                               // packed row    0..Hc+1 (last one all zero)
    if (i > Wc + 1) return;
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int y = j - 1 + (k >> 1), x = i - 1 + (k & 1);
        unsigned idx = 0;
        if (x >= 0 && x < Wc && y >= 0 && y < Hc) {
            const float t = tmpl[(size_t)y * Wc + x];
            int found = -1;
            for (int c = 0; c < npal; ++c) if (pal.palette[c] == t) { found = c; break; }
            if (found < 0) { atomicExch(err, 1); found = 0; }
            idx = (unsigned)found;
        }
        v |= idx << (k * BITS);
    }
    q[(size_t)j * pitch + i] = (T)v;
}

// Summed-area table of edge entries (see sfh_template.sat): row prefix sums, then column sums.
template <int BITS, typename T>
__global__ void __launch_bounds__(32) k_sat_rows(const T* q, int pitch, int Wc, unsigned* S, int sp) {
    constexpr unsigned kMask = (1u << BITS) - 1u, kRep = (BITS == 2) ? 0x55u : 0x1111u;
    const int j = blockIdx.x, lane = threadIdx.x;           // packed row 0..Hc+1
    unsigned run = 0;
    for (int i0 = 0; i0 <= Wc + 1; i0 += 32) {
        const int i = i0 + lane;
        unsigned e = 0;
        if (i <= Wc + 1) {
            const unsigned v = q[(size_t)j * pitch + i];
            e = (v != (v & kMask) * kRep) ? 1u : 0u;
        }
        unsigned inc = e;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (i <= Wc + 1) S[(size_t)(j + 1) * sp + (i + 1)] = run + inc;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

__global__ void k_sat_cols(unsigned* S, int sp, int rows, int cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;   // table column 1..cols
    if (i > cols) return;
    unsigned acc = 0;
    for (int j = 1; j <= rows; ++j) {
        acc += S[(size_t)j * sp + i];
        S[(size_t)j * sp + i] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
constexpr int64_t kTicketBytes = 262400;   // (65535 + 3) int32 tickets, 256-aligned

// partial slots per sample: one per CTA tile of k_fused at R = 1 (upper bound over R)
static int64_t partial_slots(int H, int W) {
    const int64_t tx = (W + kTileW - 1) / kTileW;
    return tx * ((H + kWarps - 1) / kWarps);
}

static int check_template(const sfh_template* t) {
    if (!t || !t->data || t->height <= 0 || t->width <= 0) return SFH_E_BADARG;
    if (t->fmt == SFH_TMPL_F32) return t->channels >= 1 ? 0 : SFH_E_BADARG;
    if (t->fmt == SFH_TMPL_Q2 || t->fmt == SFH_TMPL_Q4) {
        if (t->channels != 1 || t->pitch < t->width + 2) return SFH_E_BADARG;
        if (t->n_palette < 1 || t->n_palette > (t->fmt == SFH_TMPL_Q2 ? 4 : 16)) return SFH_E_BADARG;
        if (t->palette[0] != 0.0f) return SFH_E_BADARG;
        if (t->sat && t->sat_pitch < t->width + 3) return SFH_E_BADARG;
        return 0;
    }
    return SFH_E_BADFMT;
}

static void fill_common(FusedParams& p, const float* theta, const sfh_template* t, const float* xs,
                        const float* ys, int B, int H, int W, int ctas_per_sm = SFH_MINCTAS_LIGHT) {
    memset(&p, 0, sizeof(p));
    p.theta = theta; p.xs = xs; p.ys = ys; p.t = *t;
    p.B = B; p.H = H; p.W = W;
    // bands per CTA: the largest R in {16,8,4,2,1} that still leaves >= 4 waves of CTAs
    // (3 CTAs/SM x 148 SMs); bigger tiles amortise the per-CTA prologue and reduction
    // (measured on B200, C2: R=2 95 us, R=4 70 us, R=8 61 us).
    static const int forced = [] { const char* e = getenv("SFH_ROWS_PER_WARP"); return e ? atoi(e) : 0; }();
    const int tiles_x = (W + kTileW - 1) / kTileW;
    int R = 16;
    if (forced > 0) R = forced;
    else {
        auto ctas = [&](int r) { return (int64_t)tiles_x * ((H + kWarps * r - 1) / (kWarps * r)) * B; };
        while (R > 1 && ctas(R) < 4 * 3 * 148) R >>= 1;
        // small problems: one full wave of big tiles beats four waves of tiny ones (measured, C1 = 16 frames of
        // 640x360: R=2 / 1840 CTAs 10.3 us, R=4 8.4 us, R=8 / 480 CTAs 7.9 us, R=16 9.7 us; at R=1 a CTA is all prologue)
        // (loss / backward epilogues: at most R=8, the tallest tile an int64 gt stage allows, so that int64 and uint8
        // class ids keep the same tiling and therefore bit-identical sums)
        const int rmax = ctas_per_sm == SFH_MINCTAS_HEAVY ? 8 : 16;
        for (int r1 = 2 * R; r1 <= rmax; r1 <<= 1)
            if (ctas(r1) <= (int64_t)SFH_MINCTAS_LIGHT * 148 && ctas(r1) >= 148) { R = r1; break; }
    }
    p.rows_per_warp = R;
    p.ntiles = tiles_x * ((H + kWarps * R - 1) / (kWarps * R));
    static const bool no_lean = getenv("SFH_NO_LEAN") != nullptr, no_fast = getenv("SFH_NO_FAST") != nullptr;
    static const int lean_level = [] { const char* e = getenv("SFH_LEAN_LEVEL"); return e ? atoi(e) : 2; }();
    p.lean = no_lean ? 0 : lean_level;   // 1: guard-free geometry; 2: + the loss epilogue keeps its four entry loads in flight
    p.fast_free = no_fast ? 0 : 1;
}

static inline bool aligned16(const void* q) { return ((uintptr_t)q & 15u) == 0; }

// SFH_TWO_LAUNCH=1: per-sample reductions in a second (programmatically dependent) launch instead of in-kernel
static bool two_launch() {
    static const bool v = getenv("SFH_TWO_LAUNCH") != nullptr;
    return v;
}

static int setup_ws(FusedParams& p, void* ws, int64_t ws_bytes) {
    const int64_t need = sfh_workspace_bytes(p.B, p.H, p.W);
    if (!ws || ws_bytes < need) return SFH_E_WS;
    // tickets live in a FIXED-size area at the front (B <= 65535 samples + 3): only this area has to
    // stay zero between calls, so later calls with another B never read stale contrib/partials as tickets
    p.counters = (int*)ws;                                             // [B+3]: samples, batch, tile ticket, done
    p.contrib = (double*)((char*)ws + kTicketBytes);                   // [B]
    p.partials = (float*)((char*)p.contrib + align_up((int64_t)p.B * 8, 256));    // 8 bytes per value (tagged slots)
    p.poi.pub = (unsigned long long*)((char*)p.partials + align_up((int64_t)p.B * partial_slots(p.H, p.W) * kNPart * 8, 256));
    p.poi.epoch = p.counters + kEpochIdx;
    return 0;
}

template <int FMT, int MODE, int EPI>
static int launch_fmt(const FusedParams& p, dim3 grid, size_t dyn, bool ft, cudaStream_t st) {
    auto kf = k_fused<FMT, MODE, EPI, true>;
    auto kg = k_fused<FMT, MODE, EPI, false>;
    if (dyn > 32 * 1024) {                     // static (~9 KiB) + dynamic above 48 KiB needs the opt-in
        // the attribute is per device: remember which devices of this process already have it
        // (per instantiation; setting it twice is harmless, so a race between threads is too)
        static unsigned long long raised = 0ull;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 64 || !((raised >> dev) & 1ull)) {
            cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8 * kTileW * 8);
            cudaFuncSetAttribute(kg, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8 * kTileW * 8);
            if (dev < 64) raised |= 1ull << dev;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("SFH_NO_PDL_MAIN") != nullptr;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    return (int)(ft ? cudaLaunchKernelEx(&cfg, kf, p) : cudaLaunchKernelEx(&cfg, kg, p));
}

template <int MODE, int EPI>
static int launch_fused(const FusedParams& p_in, cudaStream_t st) {
    if (p_in.B > 65535) return SFH_E_BADARG;
    FusedParams p = p_in;
    const int tiles_x = (p.W + kTileW - 1) / kTileW;
    dim3 grid(tiles_x, p.ntiles / tiles_x, p.B);
    static const bool no_remap = getenv("SFH_NO_REMAP") != nullptr;
    p.short_last = (!no_remap && grid.y > 1 && (p.H % (8 * p.rows_per_warp)) != 0) ? 1 : 0;
    if (p.short_last) {                        // full rows in y, the short rows of all samples as extra z-slices
        grid.y -= 1;
        grid.z = p.B + (p.B + grid.y - 1) / grid.y;
        if (grid.z > 65535) { grid.y += 1; grid.z = p.B; p.short_last = 0; }
    }
    const bool ft = p.vec4 && (p.W % kTileW == 0) && (p.H % 8 == 0) &&
                    (!p.xs || (((uintptr_t)p.xs & 15u) == 0));
    const size_t dyn = !p.use_tma ? 0
                     : (EPI == kEpiLoss) ? (size_t)p.rows_per_warp * 8 * kTileW * (p.gt8 ? 1 : sizeof(long long))
                     : (EPI == kEpiPredict) ? (size_t)p.rows_per_warp * 4 * (kTileW / 2) * 4 * sizeof(float) : 0;
    switch (p.t.fmt) {
        case SFH_TMPL_F32: return launch_fmt<SFH_TMPL_F32, MODE, EPI>(p, grid, dyn, ft, st);
        case SFH_TMPL_Q2:  return launch_fmt<SFH_TMPL_Q2, MODE, EPI>(p, grid, dyn, ft, st);
        case SFH_TMPL_Q4:  return launch_fmt<SFH_TMPL_Q4, MODE, EPI>(p, grid, dyn, ft, st);
        default: return SFH_E_BADFMT;
    }
}

// Second launch of the predict tail and of the generic backward: components [k0, k0+nk) of every sample's
// partials are added in fixed order in fp64 (one warp per component, lanes stride the slots) and written
// as out[b * nk + (k - k0)] = sum * scale.
//   predict:  k0 = 10, nk = 1, scale = 1 / (h*w)  -> consist_score[b]   (models/reconstructor.py:238-239)
//   backward: k0 = 1,  nk = 9, scale = 1          -> dtheta[b]          (autograd of HomographyWarper)
__global__ void __launch_bounds__(128) k_comp_finalize(const __grid_constant__ FusedParams p, int k0, int nk,
                                                       float* out, double scale) {
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int c = warp; c < nk; c += 4) {
        const float* base = p.partials + (size_t)b * p.fin_slots * kNPart + k0 + c;
        double s = 0.0;
        for (int t = lane; t < p.fin_slots; t += 32) s += (double)__ldcg(base + (size_t)t * kNPart);
        s = warp_sum(s);
        if (lane == 0) out[(size_t)b * nk + c] = (float)(s * scale);
    }
}

// k_train_finalize as a programmatically dependent launch: it may become resident while the
// producer grid drains and waits in griddepcontrol.wait, which hides its launch latency.
// comps: 0 = training tail (k_train_finalize), 1 = predict score, 9 = backward dtheta (k_comp_finalize)
static int launch_finalize(const FusedParams& p, cudaStream_t st, int comps = 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.B); cfg.blockDim = dim3(kFinThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    static const bool no_pdl = getenv("SFH_NO_PDL") != nullptr;
    if (no_pdl) cfg.numAttrs = 0;
    if (comps) {
        cfg.blockDim = dim3(128);
        if (comps == 1) return (int)cudaLaunchKernelEx(&cfg, k_comp_finalize, p, 10, 1, p.score, 1.0 / ((double)p.lh * (double)p.lw));
        return (int)cudaLaunchKernelEx(&cfg, k_comp_finalize, p, 1, 9, p.dtheta, 1.0);
    }
    return (int)cudaLaunchKernelEx(&cfg, k_train_finalize, p);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// gt [B,H,W] int64 -> 3-D tensor map, box 128 x 8R x 1 (one TMA load per CTA tile, OOB zero-filled)
static bool make_gt_map(FusedParams& p) {
    static const bool off = getenv("SFH_NO_TMA") != nullptr;
    EncodeTiledFn enc = encode_tiled();
    const void* base = p.gt8 ? (const void*)p.gt8 : (const void*)p.gt;
    const cuuint64_t esz = p.gt8 ? 1 : 8;
    if (off || !enc || !aligned16(base) || p.rows_per_warp > (p.gt8 ? 16 : 8) || ((p.W * esz) % 16) != 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[2] = {(cuuint64_t)p.W * esz, (cuuint64_t)p.W * p.H * esz};
    const cuuint32_t box[3] = {(cuuint32_t)kTileW, (cuuint32_t)(8 * p.rows_per_warp), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(&p.gt_map, p.gt8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT64, 3, (void*)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// logits [B,4,h,w] fp32 (h = H/2, w = W/2) -> 4-D tensor map, box 64 x 4R x 4 x 1
static bool make_logits_map(FusedParams& p) {
    static const bool off = getenv("SFH_NO_TMA") != nullptr;
    EncodeTiledFn enc = encode_tiled();
    if (off || !enc || !aligned16(p.logits) || p.nc != 4 || p.ratio != 2 || p.rows_per_warp > 16 || (p.lw % 4) != 0) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)p.lw, (cuuint64_t)p.lh, 4, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.lw * 4, (cuuint64_t)p.lw * p.lh * 4, (cuuint64_t)p.lw * p.lh * 16};
    const cuuint32_t box[4] = {(cuuint32_t)(kTileW / 2), (cuuint32_t)(4 * p.rows_per_warp), 4, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(&p.gt_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p.logits, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace sfh

using namespace sfh;

extern "C" {

#ifdef SFH_TIMELINE
int sfh_debug_set_timeline(long long* buf) {     // debug builds only; not part of the shipped ABI
    return (int)cudaMemcpyToSymbol(g_timeline, &buf, sizeof(buf));
}
#endif

int sfh_abi_version(void) { return SFH_ABI_VERSION; }

const char* sfh_build_info(void) {
    return "sfh_b200 abi " "1" " sm_100a tile 128x(8R) threads 256 (" __DATE__ " " __TIME__ ")";
}

const char* sfh_error_string(int code) {
    if (code == 0) return "success";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
        case SFH_E_BADARG: return "sfh: bad argument";
        case SFH_E_BADFMT: return "sfh: unknown template format";
        case SFH_E_BADMODE: return "sfh: unknown interpolation mode / loss kind";
        case SFH_E_WS: return "sfh: workspace missing or too small";
    }
    return "sfh: unknown error";
}

int64_t sfh_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return kTicketBytes + align_up((int64_t)B * 8, 256) + align_up((int64_t)B * partial_slots(H, W) * kNPart * 8, 256) +
           align_up((int64_t)B * 10 * 8, 256);
}

int sfh_template_pack(const float* tmpl, int Hc, int Wc, const float* palette_host, int n_palette,
                      void* packed, int pitch, int fmt, int32_t* err_flag,
                      uint32_t* sat, int sat_pitch, void* stream) {
    if (sat && sat_pitch < Wc + 3) return SFH_E_BADARG;
    if (!tmpl || !packed || !palette_host || !err_flag || Hc <= 0 || Wc <= 0 || pitch < Wc + 2) return SFH_E_BADARG;
    const int cap = fmt == SFH_TMPL_Q2 ? 4 : fmt == SFH_TMPL_Q4 ? 16 : 0;
    if (!cap) return SFH_E_BADFMT;
    if (n_palette < 1 || n_palette > cap || palette_host[0] != 0.0f) return SFH_E_BADARG;
    sfh_template pal;
    memset(&pal, 0, sizeof(pal));
    for (int i = 0; i < n_palette; ++i) pal.palette[i] = palette_host[i];
    dim3 block(128), grid((Wc + 2 + 127) / 128, Hc + 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (fmt == SFH_TMPL_Q2) k_pack<2, uint8_t><<<grid, block, 0, st>>>(tmpl, Hc, Wc, (uint8_t*)packed, pitch, n_palette, pal, err_flag);
    else                    k_pack<4, uint16_t><<<grid, block, 0, st>>>(tmpl, Hc, Wc, (uint16_t*)packed, pitch, n_palette, pal, err_flag);
    if (sat) {
        if (fmt == SFH_TMPL_Q2) k_sat_rows<2, uint8_t><<<Hc + 2, 32, 0, st>>>((const uint8_t*)packed, pitch, Wc, sat, sat_pitch);
        else                    k_sat_rows<4, uint16_t><<<Hc + 2, 32, 0, st>>>((const uint16_t*)packed, pitch, Wc, sat, sat_pitch);
        k_sat_cols<<<(Wc + 2 + 127) / 128, 128, 0, st>>>(sat, sat_pitch, Hc + 2, Wc + 2);
    }
    return (int)cudaGetLastError();
}

int sfh_warp_fwd(const float* theta, const sfh_template* tmpl, const float* xs, const float* ys,
                 int B, int H, int W, int mode, float* out, void* stream) {
    int rc = check_template(tmpl);
    if (rc) return rc;
    if (!theta || !out || B <= 0 || H <= 0 || W <= 0) return SFH_E_BADARG;
    FusedParams p;
    fill_common(p, theta, tmpl, xs, ys, B, H, W);
    p.out_f = out;
    p.vec4 = (W % 4 == 0) && aligned16(out);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SFH_MODE_BILINEAR) return launch_fused<SFH_MODE_BILINEAR, kEpiStore>(p, st);
    if (mode == SFH_MODE_NEAREST) return launch_fused<SFH_MODE_NEAREST, kEpiStore>(p, st);
    return SFH_E_BADMODE;
}

int sfh_forward_tail(const float* theta, const sfh_template* tmpl, const float* xs, const float* ys,
                     int B, int H, int W, int mode, float* out,
                     const float* court_poi, int64_t court_poi_bstride, int N, float* poi_out, void* stream) {
    int rc = check_template(tmpl);
    if (rc) return rc;
    if (!theta || !out || B <= 0 || H <= 0 || W <= 0) return SFH_E_BADARG;
    if (court_poi && (N <= 0 || !poi_out)) return SFH_E_BADARG;
    FusedParams p;
    fill_common(p, theta, tmpl, xs, ys, B, H, W);
    p.out_f = out;
    p.vec4 = (W % 4 == 0) && aligned16(out);
    if (court_poi) {
        p.poi.theta = theta; p.poi.court_poi = court_poi; p.poi.bstride = court_poi_bstride;
        p.poi.N = N; p.poi.normalize = 1; p.poi.poi_out = poi_out;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SFH_MODE_BILINEAR) return launch_fused<SFH_MODE_BILINEAR, kEpiStore>(p, st);
    if (mode == SFH_MODE_NEAREST) return launch_fused<SFH_MODE_NEAREST, kEpiStore>(p, st);
    return SFH_E_BADMODE;
}

int sfh_warp_bwd(const float* theta, const sfh_template* tmpl, const float* xs, const float* ys,
                 const float* grad_out, int B, int H, int W, float* dtheta,
                 void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_template(tmpl);
    if (rc) return rc;
    if (!theta || !grad_out || !dtheta || B <= 0 || H <= 0 || W <= 0) return SFH_E_BADARG;
    FusedParams p;
    fill_common(p, theta, tmpl, xs, ys, B, H, W, SFH_MINCTAS_HEAVY);
    if ((rc = setup_ws(p, workspace, workspace_bytes))) return rc;
    p.grad_out = grad_out; p.dtheta = dtheta;
    p.vec4 = (W % 4 == 0) && aligned16(grad_out);
    p.split_finalize = two_launch() ? 1 : 0;
    p.fin_slots = p.ntiles;
    rc = launch_fused<SFH_MODE_BILINEAR, kEpiBwd>(p, (cudaStream_t)stream);
    if (rc || !p.split_finalize) return rc;
    return launch_finalize(p, (cudaStream_t)stream, 9);
}

int sfh_warp_loss_fwd_bwd(const sfh_template* tmpl, const sfh_train_tail_args* a, void* stream) {
    int rc = check_template(tmpl);
    if (rc) return rc;
    if (!a || !a->theta || !a->gt || !a->L_b || !a->dLb_dtheta || a->B <= 0 || a->H <= 0 || a->W <= 0 || a->nc <= 0)
        return SFH_E_BADARG;
    if (tmpl->channels != 1) return SFH_E_BADARG;
    if (a->kind != SFH_LOSS_MSE && a->kind != SFH_LOSS_SMOOTHL1) return SFH_E_BADMODE;
    if (a->loss_out && !a->dtheta_total) return SFH_E_BADARG;
    if (a->gt_dtype != SFH_GT_I64 && a->gt_dtype != SFH_GT_U8) return SFH_E_BADARG;
    FusedParams p;
    fill_common(p, a->theta, tmpl, a->xs, a->ys, a->B, a->H, a->W, SFH_MINCTAS_HEAVY);
    if ((rc = setup_ws(p, a->workspace, a->workspace_bytes))) return rc;
    if (a->gt_dtype == SFH_GT_U8) p.gt8 = (const unsigned char*)a->gt; else p.gt = (const long long*)a->gt;
    p.nc = a->nc; p.kind = a->kind;
    p.nc_pow2 = (a->nc & (a->nc - 1)) == 0;
    p.inv_nc = 1.0f / (float)a->nc;
    p.invN = 1.0f / ((float)a->H * (float)a->W);
    p.out_f = a->warp_out; p.Lb = a->L_b; p.J = a->dLb_dtheta;
    p.vec4 = (a->W % 4 == 0) && (a->gt_dtype == SFH_GT_U8 ? (((uintptr_t)a->gt & 3u) == 0) : aligned16(a->gt)) &&
             (!a->warp_out || aligned16(a->warp_out));
    const bool poi_tail = a->court_poi != nullptr;
    if (poi_tail) {
        if (a->N <= 0 || !a->poi_out) return SFH_E_BADARG;
        if (a->gt_poi && (!a->nonzeros || !a->num_nonzero || !a->R_b || !a->dRb_dtheta)) return SFH_E_BADARG;
        p.poi.theta = a->theta; p.poi.court_poi = a->court_poi; p.poi.bstride = a->court_poi_bstride;
        p.poi.N = a->N; p.poi.normalize = 1; p.poi.poi_out = a->poi_out;
        p.poi.gt_poi = a->gt_poi; p.poi.nonzeros = a->nonzeros; p.poi.num_nonzero = a->num_nonzero;
        p.poi.Rb = a->R_b; p.poi.K = a->dRb_dtheta;
    }
    p.weights = a->weights; p.w_f64 = a->weights_f64; p.w_outer = a->weights_outer;
    p.rec_lambda = a->rec_lambda; p.reproj_lambda = a->reproj_lambda;
    p.loss_out = a->loss_out; p.dtheta_total = a->dtheta_total;
    if (p.rows_per_warp > 8 && !p.gt8) {     // keep the staged int64 gt tile <= 64 KiB per CTA
        p.rows_per_warp = 8;
        p.ntiles = ((a->W + kTileW - 1) / kTileW) * ((a->H + 63) / 64);
    }
    p.use_tma = make_gt_map(p) ? 1 : 0;
    // the per-sample / batch reduction runs inside the launch (tagged slots, see k_fused) or, with
    // SFH_TWO_LAUNCH=1, as a second, programmatically dependent launch
    p.split_finalize = two_launch() ? 1 : 0;
    p.fin_slots = p.ntiles;
    rc = a->kind == SFH_LOSS_MSE ? launch_fused<SFH_LOSS_MSE, kEpiLoss>(p, (cudaStream_t)stream)
                                 : launch_fused<SFH_LOSS_SMOOTHL1, kEpiLoss>(p, (cudaStream_t)stream);
    if (rc || !p.split_finalize) return rc;
    return launch_finalize(p, (cudaStream_t)stream);
}

int sfh_predict_tail(const sfh_template* tmpl, const sfh_predict_tail_args* a, void* stream) {
    int rc = check_template(tmpl);
    if (rc) return rc;
    if (!a || !a->theta || !a->warp_out || a->B <= 0 || a->H <= 0 || a->W <= 0 || a->nc <= 0) return SFH_E_BADARG;
    if (tmpl->channels != 1) return SFH_E_BADARG;
    if (a->score && (!a->logits || a->h <= 0 || a->w <= 0)) return SFH_E_BADARG;
    FusedParams p;
    fill_common(p, a->theta, tmpl, a->xs, a->ys, a->B, a->H, a->W);
    p.nc = a->nc;
    if (a->mask_dtype == SFH_MASK_U8) p.out_u8 = (unsigned char*)a->warp_out; else p.out_i = (int32_t*)a->warp_out;
    if (a->mask_dtype != SFH_MASK_I32 && a->mask_dtype != SFH_MASK_U8) return SFH_E_BADARG;
    p.vec4 = (a->W % 4 == 0) && (p.out_u8 ? (((uintptr_t)a->warp_out & 3u) == 0) : aligned16(a->warp_out));
    if (a->score) {
        if ((rc = setup_ws(p, a->workspace, a->workspace_bytes))) return rc;
        p.logits = a->logits; p.lh = a->h; p.lw = a->w; p.score = a->score;
        p.ratio = (a->h == a->H && a->w == a->W) ? 1 : (2 * a->h == a->H && 2 * a->w == a->W) ? 2 : 0;
        if (p.ratio == 2 && p.rows_per_warp > 8) {
            // a staged logits tile of 64 x 4R x 4 floats: R = 8 (32 KiB) keeps 4 CTAs per SM resident, R = 16 only 3
            // (measured on the C5 micro-batch, 256 frames of 1280x720: 391.5 us at R = 16, 378.9 us at R = 8)
            p.rows_per_warp = 8;
            p.ntiles = ((a->W + kTileW - 1) / kTileW) * ((a->H + 63) / 64);
        }
        p.use_tma = make_logits_map(p) ? 1 : 0;
    }
    const bool poi_tail = a->court_poi != nullptr;
    if (poi_tail) {
        if (a->N <= 0 || !a->poi_out) return SFH_E_BADARG;
        p.poi.theta = a->theta; p.poi.court_poi = a->court_poi; p.poi.bstride = a->court_poi_bstride;
        p.poi.N = a->N; p.poi.normalize = 1; p.poi.poi_out = a->poi_out;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool split = a->score && p.ratio != 0 && two_launch();     // the score is reduced by k_comp_finalize
    p.split_finalize = split ? 1 : 0;
    p.fin_slots = p.ntiles;
    if (a->mode == SFH_MODE_NEAREST) rc = launch_fused<SFH_MODE_NEAREST, kEpiPredict>(p, st);
    else if (a->mode == SFH_MODE_BILINEAR) rc = launch_fused<SFH_MODE_BILINEAR, kEpiPredict>(p, st);
    else return SFH_E_BADMODE;
    if (rc) return rc;
    if (split && (rc = launch_finalize(p, st, 1))) return rc;
    if (a->score && p.ratio == 0) {
        k_consistency_generic<<<a->B, kThreads, 0, st>>>(p.out_i, p.out_u8, a->logits, a->nc, a->H, a->W, a->h, a->w, a->score);
        rc = (int)cudaGetLastError();
    }
    return rc;
}

}  // extern "C"
