/*
 * sfh_b200.h — C ABI of the B200-native STN warp stage (libsfh_b200.so).
 *
 * The reference (darkAlert/sports-field-homography) is pure Python: the "FFI" this library
 * replaces is the Python module boundary of the Reconstructor's warp stage.  Each entry point
 * names the reference interface it stands in for (file:line under the reference tree).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator in the
 *     shipped host code); the library never allocates, frees or retains device memory;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     nothing here synchronises;
 *   - return value: 0 on success, a positive cudaError_t for CUDA failures, a negative
 *     SFH_E_* code for argument errors.  No exceptions cross this boundary;
 *   - theta is [B,3,3] row-major fp32 and maps OUTPUT-frame normalised coords to TEMPLATE
 *     normalised coords (no inversion), exactly as kornia's HomographyWarper consumes it;
 *   - xs[W] / ys[H] are the 1-D factors of kornia's create_meshgrid; NULL means "compute
 *     (i/(n-1) - 0.5) * 2 with IEEE division in the kernel".
 */
#ifndef SFH_B200_H_
#define SFH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFH_ABI_VERSION 1

/* argument-error codes (negative) */
#define SFH_E_BADARG   (-1)
#define SFH_E_BADFMT   (-2)
#define SFH_E_BADMODE  (-3)
#define SFH_E_WS       (-4)

/* interpolation mode: HomographyWarper(mode=...) — models/reconstructor.py:105,107 */
#define SFH_MODE_BILINEAR 0
#define SFH_MODE_NEAREST  1

/* reconstruction criterion — train.py:113-116 (nn.MSELoss / nn.SmoothL1Loss, reduction='none') */
#define SFH_LOSS_MSE      0
#define SFH_LOSS_SMOOTHL1 1

/* dtype of the ground-truth masks handed to sfh_warp_loss_fwd_bwd */
#define SFH_GT_I64 0   /* int64, what the reference's dataset yields (utils/dataset.py:167) */
#define SFH_GT_U8  1   /* uint8 class ids: the narrow surface of SURVEY.md §8 f-1 (8x less PCIe / HBM traffic) */

/* dtype of the integer mask written by sfh_predict_tail */
#define SFH_MASK_I32 0  /* int32, what Reconstructor.predict returns (models/reconstructor.py:240) */
#define SFH_MASK_U8  1  /* uint8, what predict.py:99 converts it to before it leaves the GPU worker (§8 f-1) */

/* template storage formats */
#define SFH_TMPL_F32 0  /* [Bt,C,Hc,Wc] fp32 as open_court_template returns it (utils/dataset.py:47-61) */
#define SFH_TMPL_Q2  1  /* quad-packed palette indices, 2 bits/tap, uint8  [(Hc+2) x pitch] */
#define SFH_TMPL_Q4  2  /* quad-packed palette indices, 4 bits/tap, uint16 [(Hc+2) x pitch] */

/* Template descriptor.  For Q2/Q4, entry (j,i), 0 <= j <= Hc+1, 0 <= i <= Wc+1, holds the palette
 * indices of the 2x2 neighbourhood whose top-left texel is (y=j-1, x=i-1); texels outside the
 * image are palette index 0 and palette[0] must be 0.0f (= padding_mode='zeros'), so the last
 * row and column are all zero and serve as the clamp target.  C must be 1 for Q2/Q4. */
typedef struct sfh_template {
    const void* data;        /* device pointer */
    int32_t     fmt;         /* SFH_TMPL_* */
    int32_t     channels;    /* C */
    int32_t     height;      /* Hc */
    int32_t     width;       /* Wc */
    int32_t     pitch;       /* Q2/Q4: elements per packed row (>= Wc+2); F32: ignored */
    int32_t     n_palette;   /* Q2/Q4: number of valid palette entries */
    int64_t     batch_stride;/* F32: elements between consecutive samples, 0 = one shared template */
    float       palette[16]; /* Q2/Q4: texel value of each palette index */
    /* Q2/Q4, optional: summed-area table of "packed entry straddles a class edge",
     * S[(j+1)*sat_pitch + (i+1)] = #{edge entries (j',i') : j' <= j, i' <= i}, row 0 / column 0
     * zero, (Hc+3) x sat_pitch uint32.  With it, 16x8 output patches whose sampling footprint
     * contains no class edge are written as a constant (nearest: bit-exact; bilinear: within
     * 2 ulp of the interpolated constant, and the exactly-zero gradient is kept).  NULL disables
     * the shortcut: every pixel is then evaluated in ATen's exact operation order. */
    const uint32_t* sat;
    int32_t     sat_pitch;
} sfh_template;

int sfh_abi_version(void);
const char* sfh_build_info(void);
const char* sfh_error_string(int code);

/* Bytes of scratch the fused entry points need for a batch of B samples warped to HxW.
 * The workspace must be zero-filled once when it is allocated and must not be written by anyone else
 * afterwards: it carries a launch epoch (the tag of the in-launch reductions) besides the per-call partials.
 * One workspace serves one stream at a time. */
int64_t sfh_workspace_bytes(int B, int H, int W);

/* Pack a class-index template (fp32 [Hc,Wc], every texel equal to one of palette[0..n)) into
 * the Q2 (n<=4) or Q4 (n<=16) layout and, optionally, its edge summed-area table.  *err_flag (device int32, pre-zeroed) is set to 1 if a
 * texel matches no palette entry.  Replaces nothing in the reference: it is the "stage the
 * template once" step; source data is open_court_template's tensor (utils/dataset.py:47-61). */
int sfh_template_pack(const float* tmpl, int Hc, int Wc, const float* palette_host, int n_palette,
                      void* packed, int pitch, int fmt, int32_t* err_flag,
                      uint32_t* sat /* nullable, zero-filled (Hc+3) x sat_pitch */, int sat_pitch,
                      void* stream);

/* kornia HomographyWarper.forward(patch_src, src_homo_dst) — models/reconstructor.py:116
 * (Reconstructor.warp, :109-118).  out: [B,C,H,W] fp32. */
int sfh_warp_fwd(const float* theta, const sfh_template* tmpl, const float* xs, const float* ys,
                 int B, int H, int W, int mode, float* out, void* stream);

/* Warp-stage part of Reconstructor.forward — models/reconstructor.py:185-192 — in ONE launch:
 *   out     = warper(court_img, theta)                      [B,C,H,W] fp32   (ret['warp_mask'], :191)
 *   poi_out = transform_poi(theta, court_poi)               [B,N,2]  fp32   (ret['poi'], :186; skipped if court_poi == NULL) */
int sfh_forward_tail(const float* theta, const sfh_template* tmpl, const float* xs, const float* ys,
                     int B, int H, int W, int mode, float* out,
                     const float* court_poi, int64_t court_poi_bstride, int N, float* poi_out, void* stream);

/* autograd of the above w.r.t. theta (train.py:235 loss.backward() through grid_sample/bmm).
 * grad_out [B,C,H,W] fp32 -> dtheta [B,3,3] fp32.  Bilinear only (nearest has zero gradient). */
int sfh_warp_bwd(const float* theta, const sfh_template* tmpl, const float* xs, const float* ys,
                 const float* grad_out, int B, int H, int W, float* dtheta,
                 void* workspace, int64_t workspace_bytes, void* stream);

/* Fused training tail, one pass over the pixels, ONE launch (the per-sample and batch reductions run
 * inside it; with the environment switch SFH_TWO_LAUNCH=1 they run as a second, programmatically
 * dependent launch instead).  int64 class ids are read through their low 32 bits: ids outside
 * [-2^31, 2^31) differ from the reference's gt.float() (mask class ids are 0..nc-1).
 *   warp_mask = warp(theta)                               models/reconstructor.py:191
 *   L_b = mean_{h,w} crit(warp_mask, gt/nc)               train.py:194-197, models/losses.py:35-38
 *   J_b = dL_b/dtheta_b
 * when court_poi != NULL also
 *   poi = transform_poi(theta, court_poi)                 models/reconstructor.py:186,120-130
 *   R_b = sum_n ||gt_poi-poi|| * nonzeros / num_nonzero   models/losses.py:10-11   (if gt_poi)
 *   K_b = dR_b/dtheta_b
 * and when loss_out != NULL the reference's weighting and batch means as well
 *   loss   = rec_lambda * mean(L_b * w) + reproj_lambda * mean(R_b)   models/losses.py:38-39,13-14
 *   dtheta = dloss/dtheta                                            train.py:196,213,235
 * with w = weights[b] (weights_outer == 0) or, for weights that arrived as [B,1], the
 * reference's [B]*[B,1] -> [B,B] broadcast, i.e. w = mean(weights) (weights_outer == 1). */
typedef struct sfh_train_tail_args {
    const float*   theta;            /* [B,3,3] */
    const float*   xs;               /* [W] nullable */
    const float*   ys;               /* [H] nullable */
    const void*    gt;               /* [B,H,W] class ids, int64 (utils/dataset.py:167) or uint8, see gt_dtype */
    int32_t B, H, W;
    int32_t nc;                      /* mask_classes */
    int32_t kind;                    /* SFH_LOSS_* */
    int32_t N;                       /* points per sample */
    float*  warp_out;                /* [B,H,W] fp32, nullable */
    float*  L_b;                     /* [B] */
    float*  dLb_dtheta;              /* [B,3,3] */
    const float* court_poi;          /* [*,N,2] in [-1,1], nullable */
    int64_t court_poi_bstride;       /* elements between samples, 0 = shared */
    const float* gt_poi;             /* [B,N,2] nullable */
    const float* nonzeros;           /* [B,N] */
    const float* num_nonzero;        /* [B] */
    float*  poi_out;                 /* [B,N,2] */
    float*  R_b;                     /* [B] */
    float*  dRb_dtheta;              /* [B,3,3] */
    const void* weights;             /* [B] fp32/fp64, nullable (w = 1) */
    int32_t weights_f64;
    int32_t weights_outer;
    int32_t gt_dtype;                /* SFH_GT_I64 (default 0) | SFH_GT_U8 */
    float   rec_lambda;
    float   reproj_lambda;
    float*  loss_out;                /* scalar, nullable */
    float*  dtheta_total;            /* [B,3,3], required with loss_out */
    void*   workspace;
    int64_t workspace_bytes;
} sfh_train_tail_args;

int sfh_warp_loss_fwd_bwd(const sfh_template* tmpl, const sfh_train_tail_args* args, void* stream);

/* Reconstructor.predict tail — models/reconstructor.py:221-245, ONE launch:
 *   warp_out = int32(warp(theta) * nc)            [B,H,W]
 *   score_b  = mean CE(logits, int64(resize_nearest(warp*nc)))   logits [B,nc,h,w]  (nullable)
 *   poi_out  = transform_poi(theta, court_poi)    (nullable) */
typedef struct sfh_predict_tail_args {
    const float* theta;
    const float* xs;
    const float* ys;
    int32_t B, H, W;
    int32_t mode;                    /* SFH_MODE_* */
    int32_t nc;
    int32_t h, w;                    /* logits spatial size */
    int32_t N;
    const float* logits;             /* [B,nc,h,w] nullable */
    void*    warp_out;               /* [B,H,W] int32 (default) or uint8, see mask_dtype */
    float*  score;                   /* [B] nullable */
    const float* court_poi;          /* nullable */
    int64_t court_poi_bstride;
    float*  poi_out;                 /* [B,N,2] */
    void*   workspace;
    int64_t workspace_bytes;
    int32_t mask_dtype;              /* SFH_MASK_I32 (default 0) | SFH_MASK_U8 */
} sfh_predict_tail_args;

int sfh_predict_tail(const sfh_template* tmpl, const sfh_predict_tail_args* args, void* stream);

/* Reconstructor.transform_poi — models/reconstructor.py:120-130:
 * poi = transform_points(inverse(theta), court_poi) [/2 + 0.5 if normalize].  fp64 inside. */
int sfh_poi_fwd(const float* theta, const float* court_poi, int64_t court_poi_bstride,
                int B, int N, int normalize, float* poi_out, void* stream);

/* autograd of sfh_poi_fwd: grad_poi [B,N,2] -> dtheta [B,3,3]. */
int sfh_poi_bwd(const float* theta, const float* court_poi, int64_t court_poi_bstride,
                const float* grad_poi, int B, int N, int normalize, float* dtheta, void* stream);

/* kornia.geometry.linalg.transform_points(trans_01, points_1) — models/reconstructor.py:124.
 * trans [Bt,3,3] with Bt in {1,B}; points [B,N,2] -> out [B,N,2] (fp32 op order of the
 * reference: bmm chain, 1/z where |z|>1e-8). */
int sfh_transform_points_fwd(const float* trans, int Bt, const float* points, int B, int N,
                             float* out, void* stream);

/* autograd of the above: grad_out [B,N,2] -> dtrans [Bt,3,3] (nullable), dpoints [B,N,2] (nullable) */
int sfh_transform_points_bwd(const float* trans, int Bt, const float* points, const float* grad_out,
                             int B, int N, float* dtrans, float* dpoints, void* stream);

/* models/losses.py:6-18 reprojection_loss per-sample part and its gradient w.r.t. `inputs`:
 * R_b = sum_n sqrt(sum_xy (targets-inputs)^2) * nonzeros / num_nonzero;
 * dinputs (nullable) = grad_Rb[b] * dR_b/dinputs. */
int sfh_reproj_loss(const float* inputs, const float* targets, const float* nonzeros,
                    const float* num_nonzero, int B, int N, float* R_b,
                    const float* grad_Rb, float* dinputs, void* stream);

/* Training / evaluation consistency loss — train.py:219-223, eval.py:201-203:
 *   loss = lambda * CrossEntropyLoss(reduction='mean')(logits, (warp_mask * nc).long())
 * warp_mask [B,1,H,W] fp32 (what sfh_warp_fwd / sfh_warp_loss_fwd_bwd wrote), logits [B,nc,H,W]
 * fp32 of the SAME size (the training configuration), 1 <= nc <= 8.  loss_out: device scalar.
 * dlogits [B,nc,H,W] (nullable) = d loss / d logits = lambda/(B*H*W) * (softmax - onehot).
 * No gradient flows to theta (the integer cast cuts it, as in the reference).  Class ids outside
 * [0,nc) are clamped.  workspace: >= sfh_consist_workspace_bytes() bytes, zero before the first
 * call (the library leaves it zeroed). */
int64_t sfh_consist_workspace_bytes(void);
int sfh_consist_loss_fwd_bwd(const float* warp_mask, const float* logits, int B, int nc, int H, int W,
                             float lambda, float* loss_out, float* dlogits,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* The same pass with the reference's other consistency criterion — train.py:133-134:
 *   kornia.losses.FocalLoss(alpha=1.0, gamma=2.0, reduction='mean')(logits, (warp_mask * nc).long()) * lambda
 * following kornia 0.5.x/0.6.x focal_loss including its eps conventions: probabilities softmax + 1e-8, one-hot
 * targets + 1e-6 (kornia.utils.one_hot), loss_px = sum_c t_c * (-alpha (1 - p_c)^gamma log p_c).  gamma >= 0. */
int sfh_consist_focal_fwd_bwd(const float* warp_mask, const float* logits, int B, int nc, int H, int W,
                              float alpha, float gamma, float lambda, float* loss_out, float* dlogits,
                              void* workspace, int64_t workspace_bytes, void* stream);

/* GPU post-processing of the predict outputs, done BEFORE the device->host copy:
 *   src_kind LOGITS:   [B,nc,h,w] fp32 -> class ids by argmax (utils/postprocess.py:7-18 preds_to_masks;
 *                      argmax(softmax(l)) == argmax(l), first maximum wins)
 *   src_kind MASK_I32: [B,h,w] int32 as Reconstructor.predict returns it, `.astype(np.uint8)` (predict.py:99)
 *   src_kind MASK_U8:  [B,h,w] uint8 (the narrow surface of sfh_predict_tail)
 *   mask_type GRAY: class ids; BIN: (id > 0) * 255 (predict.py:293-297); RGB: id -> colour of
 *                   utils/postprocess.py:21-58 (nc in {4,7,8}), out [B,oh,ow,3]
 *   resize: cv2.resize(..., interpolation=cv2.INTER_NEAREST) (predict.py:303-315) through the source
 *           index tables x_ofs[ow] / y_ofs[oh] (device int32; x_ofs[x] = min(floor(x * w/ow), w-1) evaluated
 *           in double on the host as cv::resize does); NULL = same size.
 * out: uint8 [B,oh,ow] (GRAY/BIN) or [B,oh,ow,3] (RGB). */
#define SFH_POST_SRC_LOGITS   0
#define SFH_POST_SRC_MASK_I32 1
#define SFH_POST_SRC_MASK_U8  2
#define SFH_POST_GRAY 0
#define SFH_POST_BIN  1
#define SFH_POST_RGB  2
int sfh_postprocess(const void* src, int src_kind, int B, int nc, int h, int w,
                    int mask_type, const int* x_ofs, const int* y_ofs, int oh, int ow,
                    unsigned char* out, void* stream);

/* utils/transform.py:7-20 — Warper.warp(theta, proj): kornia HomographyWarper(mode='nearest',
 * normalized_coordinates=True) on a DOUBLE-precision multi-channel image.
 * theta [B,3,3] fp64 (output-frame -> template, normalised, as everywhere in this header), tmpl [Bt,C,Hc,Wc] fp64
 * (tmpl_batch_stride elements between samples, 0 = shared), xs[W] / ys[H] the fp64 meshgrid factors (required: which
 * dtype kornia builds them in is version dependent, so the caller states it), out [B,C,H,W] fp64.  Coordinates in
 * fp64 in kornia's operation order, nearbyint pick, zeros outside. */
int sfh_warp_nearest_f64(const double* theta, const double* tmpl, int64_t tmpl_batch_stride, const double* xs,
                         const double* ys, int B, int C, int Hc, int Wc, int H, int W, double* out, void* stream);

/* dataset_utils/football_dataset.ipynb cell 11 with preparation.py:129-137 (rescale_theta) — the dataset's mask /
 * UV-map rendering: cv2.warpPerspective(src, M, (W,H), flags=cv2.INTER_NEAREST), border constant 0, for B matrices
 * at once.  minv [B,3,3] fp64 = inverse of each M (destination pixel -> source pixel; cv2 inverts M itself),
 * src [Hs,Ws] pixels of pixel_bytes bytes each (channels interleaved like a cv::Mat; 1,2,3,4,8,16 or 24 bytes),
 * dst [B,H,W] pixels.  OpenCV's evaluation order (64-column blocks, double precision) and cvRound are reproduced. */
int sfh_warp_perspective_nearest(const double* minv, int B, const void* src, int Hs, int Ws, int pixel_bytes,
                                 int H, int W, void* dst, void* stream);

/* Diagnostic: exhaustively compares the kernels' fast correctly-rounded reciprocal with IEEE
 * rcp.rn over every fp32 value with |z| in (1e-8, 1e37) (the range it is used on; the warp path
 * falls back to rcp.rn outside).  *mismatches (device uint64, pre-zeroed) receives the count. */
int sfh_selftest_rcp(unsigned long long* mismatches, void* stream);

/* Diagnostic: plain streaming kernel out[i] = (float)in[i] / 4 over n (multiple of 4) int64 -> fp32
 * elements with `ctas` CTAs of 256 threads; moves the bytes of a training step with no other work
 * (tools/stream_ceiling.py uses it to measure the size-matched streaming ceiling). */
int sfh_debug_stream_cast(const int64_t* in, float* out, int64_t n, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFH_B200_H_ */
