/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's STN warp stage with the
 * fp32 operation order written out explicitly (one rounding per line).  Not a product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it.
 *
 * Parity status: unpinned at the kornia boundary (kornia >= 0.5.0 is a requirements.txt:1
 * dependency of the reference and is not vendored; its algorithm is restated from SURVEY.md
 * Appendix A).  This file is pinned against the torch restatement oracle/kornia_restated.py,
 * which runs the real ATen ops (bmm, grid_sample): the flow field is bit-identical and the
 * warped mask is bit-identical on CPU (tests/test_oracle.py).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fPIC -shared)
 * -ffp-contract=off + explicit fmaf() calls make every rounding below intentional.
 *
 * What each function follows (files under /root/reference):
 *   sfh_oracle_meshgrid      kornia create_meshgrid, used by HomographyWarper
 *                            (models/reconstructor.py:105,107)
 *   sfh_oracle_flow1         kornia transform_points + convert_points_from_homogeneous
 *                            (models/reconstructor.py:4,116,124)
 *   sfh_oracle_warp_fwd/bwd  HomographyWarper.forward -> F.grid_sample(zeros, align_corners=False)
 *                            (models/reconstructor.py:109-118) and its autograd
 *   sfh_oracle_warp_loss     train.py:194-197 + models/losses.py:33-41 (per-sample part)
 *   sfh_oracle_predict_tail  models/reconstructor.py:221-245
 *   sfh_oracle_poi_*         models/reconstructor.py:120-130, models/losses.py:6-18
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SFH_MODE_BILINEAR 0
#define SFH_MODE_NEAREST 1
#define SFH_LOSS_MSE 0
#define SFH_LOSS_SMOOTHL1 1

/* (i/(n-1) - 0.5) * 2 with IEEE division: torch.linspace(0,n-1,n) is exactly i. */
void sfh_oracle_meshgrid(int n, float *out) {
    for (int i = 0; i < n; ++i) {
        float q = (float)i / (float)(n - 1);
        float c = q - 0.5f;
        out[i] = c * 2.0f;
    }
}

/* One grid point through theta: bmm order measured on torch CPU (MKL sgemm, K=3):
 * acc = u*h0; acc = fma(v,h1,acc); acc = acc + h2   (tests/test_oracle.py pins this). */
static inline void flow1(const float *t, float u, float v, float *x, float *y,
                         float *Xo, float *Yo, float *so, int *zok) {
    float X = fmaf(v, t[1], u * t[0]) + t[2];
    float Y = fmaf(v, t[4], u * t[3]) + t[5];
    float Z = fmaf(v, t[7], u * t[6]) + t[8];
    int ok = fabsf(Z) > 1e-8f;
    float s = ok ? 1.0f / Z : 1.0f;
    *x = s * X;
    *y = s * Y;
    if (Xo) { *Xo = X; *Yo = Y; *so = s; *zok = ok; }
}

void sfh_oracle_flow1(const float *theta9, float u, float v, float *xy) {
    flow1(theta9, u, v, &xy[0], &xy[1], 0, 0, 0, 0);
}

/* GridSampler.cuh:29 (align_corners=False) as torch compiles it: one FMA then * 0.5.
 * Identical bits to the CPU kernel's fma(x+1, size/2, -0.5). */
static inline float unnormalize(float c, int size) {
    float r = fmaf(c + 1.0f, (float)size, -1.0f) * 0.5f;
    /* safe_downgrade_to_int_range (GridSampler.cuh:140-147) */
    if (!(r <= 2147483647.0f - 1.0f) || !(r >= -2147483648.0f) || !isfinite(r)) r = -100.0f;
    return r;
}

static inline float tap(const float *img, int Hc, int Wc, int y, int x) {
    return (x >= 0 && x < Wc && y >= 0 && y < Hc) ? img[(size_t)y * Wc + x] : 0.0f;
}

/* theta [B,9]; tmpl [Bt,C,Hc,Wc] with batch stride tmpl_bstride (0 => one shared template);
 * xs[W], ys[H] may be NULL (built with sfh_oracle_meshgrid); out [B,C,H,W]. */
void sfh_oracle_warp_fwd(const float *theta, const float *tmpl, long tmpl_bstride,
                         int B, int C, int Hc, int Wc, int H, int W, int mode,
                         const float *xs_in, const float *ys_in, float *out) {
    float *xs = (float *)malloc(sizeof(float) * W), *ys = (float *)malloc(sizeof(float) * H);
    if (xs_in) memcpy(xs, xs_in, sizeof(float) * W); else sfh_oracle_meshgrid(W, xs);
    if (ys_in) memcpy(ys, ys_in, sizeof(float) * H); else sfh_oracle_meshgrid(H, ys);
    for (int b = 0; b < B; ++b) {
        const float *t = theta + 9 * b;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                float x, y;
                flow1(t, xs[w], ys[h], &x, &y, 0, 0, 0, 0);
                float ix = unnormalize(x, Wc), iy = unnormalize(y, Hc);
                for (int c = 0; c < C; ++c) {
                    const float *img = tmpl + (size_t)b * tmpl_bstride + (size_t)c * Hc * Wc;
                    float o;
                    if (mode == SFH_MODE_NEAREST) {
                        int xn = (int)nearbyintf(ix), yn = (int)nearbyintf(iy);
                        o = tap(img, Hc, Wc, yn, xn);
                    } else {
                        float fx = floorf(ix), fy = floorf(iy);
                        int x0 = (int)fx, y0 = (int)fy;
                        float ex = (fx + 1.0f) - ix, wx = ix - fx;
                        float sy = (fy + 1.0f) - iy, ny = iy - fy;
                        float nw = ex * sy, ne = wx * sy, sw = ex * ny, se = wx * ny;
                        o = tap(img, Hc, Wc, y0, x0) * nw;
                        o = fmaf(tap(img, Hc, Wc, y0, x0 + 1), ne, o);
                        o = fmaf(tap(img, Hc, Wc, y0 + 1, x0), sw, o);
                        o = fmaf(tap(img, Hc, Wc, y0 + 1, x0 + 1), se, o);
                    }
                    out[(((size_t)b * C + c) * H + h) * W + w] = o;
                }
            }
    }
    free(xs); free(ys);
}

/* Per-pixel gradient of the bilinear sample wrt theta, times gout; accumulated in double.
 * grid_sampler_2d_backward (GridSampler.cu) -> scale*p backward -> bmm backward
 * ("chain to theta"). */
static inline void bwd_pixel(const float *t, const float *img, int C, int Hc, int Wc,
                             float u, float v, const float *gout, size_t gstride, double *acc) {
    float x, y, X, Y, s; int ok;
    flow1(t, u, v, &x, &y, &X, &Y, &s, &ok);
    float ix = unnormalize(x, Wc), iy = unnormalize(y, Hc);
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy;
    float ex = (fx + 1.0f) - ix, wx = ix - fx, sy = (fy + 1.0f) - iy, ny = iy - fy;
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
        const float *im = img + (size_t)c * Hc * Wc;
        float g = gout[c * gstride];
        float a = tap(im, Hc, Wc, y0, x0), b = tap(im, Hc, Wc, y0, x0 + 1);
        float cc = tap(im, Hc, Wc, y0 + 1, x0), d = tap(im, Hc, Wc, y0 + 1, x0 + 1);
        gix -= a * sy * g; giy -= a * ex * g;
        gix += b * sy * g; giy -= b * wx * g;
        gix -= cc * ny * g; giy += cc * ex * g;
        gix += d * ny * g; giy += d * wx * g;
    }
    float gx = (0.5f * (float)Wc) * gix, gy = (0.5f * (float)Hc) * giy;
    float gX = gx * s, gY = gy * s;
    float gZ = ok ? -(gx * X + gy * Y) * s * s : 0.0f;
    acc[0] += (double)gX * u; acc[1] += (double)gX * v; acc[2] += (double)gX;
    acc[3] += (double)gY * u; acc[4] += (double)gY * v; acc[5] += (double)gY;
    acc[6] += (double)gZ * u; acc[7] += (double)gZ * v; acc[8] += (double)gZ;
}

/* grad_out [B,C,H,W] -> dtheta [B,9] (bilinear only; nearest has zero gradient). */
void sfh_oracle_warp_bwd(const float *theta, const float *tmpl, long tmpl_bstride,
                         const float *grad_out, int B, int C, int Hc, int Wc, int H, int W,
                         const float *xs_in, const float *ys_in, float *dtheta) {
    float *xs = (float *)malloc(sizeof(float) * W), *ys = (float *)malloc(sizeof(float) * H);
    if (xs_in) memcpy(xs, xs_in, sizeof(float) * W); else sfh_oracle_meshgrid(W, xs);
    if (ys_in) memcpy(ys, ys_in, sizeof(float) * H); else sfh_oracle_meshgrid(H, ys);
    for (int b = 0; b < B; ++b) {
        double acc[9] = {0};
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                bwd_pixel(theta + 9 * b, tmpl + (size_t)b * tmpl_bstride, C, Hc, Wc, xs[w], ys[h],
                          grad_out + ((size_t)b * C * H + h) * W + w, (size_t)H * W, acc);
        for (int k = 0; k < 9; ++k) dtheta[9 * b + k] = (float)acc[k];
    }
    free(xs); free(ys);
}

/* Fused training tail: warp (C=1, bilinear) + per-sample MSE / SmoothL1(beta=1) against
 * gt/nc (train.py:194-197, models/losses.py:35-38 before the weight multiply) and
 * J_b = dL_b/dtheta_b.  warp_out may be NULL.  Sums in double. */
void sfh_oracle_warp_loss(const float *theta, const float *tmpl, long tmpl_bstride,
                          const int64_t *gt, int nc, int kind, int B, int Hc, int Wc, int H, int W,
                          const float *xs_in, const float *ys_in,
                          float *warp_out, float *Lb, float *dLb_dtheta) {
    float *xs = (float *)malloc(sizeof(float) * W), *ys = (float *)malloc(sizeof(float) * H);
    if (xs_in) memcpy(xs, xs_in, sizeof(float) * W); else sfh_oracle_meshgrid(W, xs);
    if (ys_in) memcpy(ys, ys_in, sizeof(float) * H); else sfh_oracle_meshgrid(H, ys);
    float *row = (float *)malloc(sizeof(float) * W);
    const float invN = 1.0f / ((float)H * (float)W);
    for (int b = 0; b < B; ++b) {
        double acc[9] = {0}, lsum = 0.0;
        const float *img = tmpl + (size_t)b * tmpl_bstride;
        for (int h = 0; h < H; ++h) {
            sfh_oracle_warp_fwd(theta + 9 * b, img, 0, 1, 1, Hc, Wc, 1, W, SFH_MODE_BILINEAR,
                                xs, ys + h, row);
            for (int w = 0; w < W; ++w) {
                size_t i = ((size_t)b * H + h) * W + w;
                float tgt = (float)gt[i] / (float)nc;
                float d = row[w] - tgt, g;
                if (kind == SFH_LOSS_MSE) { lsum += (double)(d * d); g = 2.0f * d; }
                else if (fabsf(d) < 1.0f) { lsum += (double)(0.5f * d * d); g = d; }
                else { lsum += (double)(fabsf(d) - 0.5f); g = d > 0 ? 1.0f : -1.0f; }
                g *= invN;
                if (warp_out) warp_out[i] = row[w];
                bwd_pixel(theta + 9 * b, img, 1, Hc, Wc, xs[w], ys[h], &g, 0, acc);
            }
        }
        Lb[b] = (float)(lsum / ((double)H * (double)W));
        for (int k = 0; k < 9; ++k) dLb_dtheta[9 * b + k] = (float)acc[k];
    }
    free(row); free(xs); free(ys);
}

/* upsample_nearest source index (UpSample.h nearest_idx). */
static inline int nearest_idx(int dst, int in_size, int out_size) {
    if (out_size == in_size) return dst;
    if (out_size == 2 * in_size) return dst >> 1;
    float scale = (float)in_size / (float)out_size;
    int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

/* Reconstructor.predict tail (models/reconstructor.py:221-245), C=1:
 * m = warp(theta)*nc (float); score_b = mean_{i,j} CE(logits[b,:,i,j], int64(m'[i,j])) with
 * m' = nearest-resize of m to (h,w); warp_out = int32(m). logits [B,nc,h,w]. */
void sfh_oracle_predict_tail(const float *theta, const float *tmpl, long tmpl_bstride,
                             const float *logits, int nc, int h, int w,
                             int B, int Hc, int Wc, int H, int W, int mode,
                             const float *xs_in, const float *ys_in,
                             int32_t *warp_out, float *score) {
    float *m = (float *)malloc(sizeof(float) * (size_t)H * W);
    for (int b = 0; b < B; ++b) {
        sfh_oracle_warp_fwd(theta + 9 * b, tmpl + (size_t)b * tmpl_bstride, 0, 1, 1, Hc, Wc, H, W,
                            mode, xs_in, ys_in, m);
        for (size_t i = 0; i < (size_t)H * W; ++i) {
            m[i] = m[i] * (float)nc;
            warp_out[(size_t)b * H * W + i] = (int32_t)m[i];
        }
        if (!score) continue;
        double ssum = 0.0;
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j) {
                int si = nearest_idx(i, H, h), sj = nearest_idx(j, W, w);
                int64_t cls = (int64_t)m[(size_t)si * W + sj];
                const float *lg = logits + (size_t)b * nc * h * w + (size_t)i * w + j;
                float mx = lg[0];
                for (int c = 1; c < nc; ++c) mx = fmaxf(mx, lg[(size_t)c * h * w]);
                float se = 0.f;
                for (int c = 0; c < nc; ++c) se += expf(lg[(size_t)c * h * w] - mx);
                float lse = logf(se) + mx;
                ssum += (double)(lse - lg[(size_t)cls * h * w]);
            }
        score[b] = (float)(ssum / ((double)h * (double)w));
    }
    free(m);
}

/* transform_poi (models/reconstructor.py:120-130) evaluated in double: adjugate inverse,
 * transform_points, eps rule, /2+0.5.  theta [B,9] fp32, court_poi [B,N,2] fp32 -> poi [B,N,2]. */
static void inv3(const double *a, double *inv) {
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02, r = 1.0 / det;
    inv[0] = c00 * r; inv[1] = (a[2] * a[7] - a[1] * a[8]) * r; inv[2] = (a[1] * a[5] - a[2] * a[4]) * r;
    inv[3] = c01 * r; inv[4] = (a[0] * a[8] - a[2] * a[6]) * r; inv[5] = (a[2] * a[3] - a[0] * a[5]) * r;
    inv[6] = c02 * r; inv[7] = (a[1] * a[6] - a[0] * a[7]) * r; inv[8] = (a[0] * a[4] - a[1] * a[3]) * r;
}

void sfh_oracle_poi_fwd(const float *theta, const float *court_poi, int B, int N, int normalize,
                        float *poi) {
    for (int b = 0; b < B; ++b) {
        double a[9], iv[9];
        for (int k = 0; k < 9; ++k) a[k] = theta[9 * b + k];
        inv3(a, iv);
        for (int n = 0; n < N; ++n) {
            double px = court_poi[((size_t)b * N + n) * 2], py = court_poi[((size_t)b * N + n) * 2 + 1];
            double X = iv[0] * px + iv[1] * py + iv[2], Y = iv[3] * px + iv[4] * py + iv[5];
            double Z = iv[6] * px + iv[7] * py + iv[8];
            double s = fabs(Z) > 1e-8 ? 1.0 / Z : 1.0;
            double x = s * X, y = s * Y;
            if (normalize) { x = x / 2.0 + 0.5; y = y / 2.0 + 0.5; }
            poi[((size_t)b * N + n) * 2] = (float)x;
            poi[((size_t)b * N + n) * 2 + 1] = (float)y;
        }
    }
}

/* reprojection_loss per sample (models/losses.py:10-11): L_b = sum_n ||gt-poi|| * nz / num. */
void sfh_oracle_reproj_per_sample(const float *poi, const float *gt_poi, const float *nonzeros,
                                  const float *num_nonzero, int B, int N, float *Lb) {
    for (int b = 0; b < B; ++b) {
        double s = 0.0;
        for (int n = 0; n < N; ++n) {
            size_t i = ((size_t)b * N + n) * 2;
            double dx = (double)gt_poi[i] - poi[i], dy = (double)gt_poi[i + 1] - poi[i + 1];
            s += sqrt(dx * dx + dy * dy) * nonzeros[(size_t)b * N + n];
        }
        Lb[b] = (float)(s / num_nonzero[b]);
    }
}
