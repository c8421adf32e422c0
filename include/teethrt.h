/* libteethrt — C ABI of the B200-native (sm_100a) hot path of
 * ahmedmajid92/multimodal-teeth-restoration-selection.
 *
 * The reference has no FFI/plugin interface: its boundary is the PyTorch nn.Module surface
 * (MMJointDualHead, MMNet, MILNet, MMEnsemble, MILEnsemble) plus OpenCV calls.  This header is the NEW boundary that sits
 * directly under that surface; every entry point names the reference call site it replaces.  The Python package
 * `teethrt` binds it with ctypes (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned storage unless the name ends in `_host`
 *   - the callee never allocates, frees or synchronises; work is enqueued on `stream`
 *   - returns 0 (TRT_OK) or a negative trt_status; the message is in trt_last_error_string() (thread-local)
 *   - activations are NHWC bf16 ("rows" = N*H*W pixels, "C" channels, C % 8 == 0); parameters/gradients fp32
 *   - compiled for sm_100a only; there is no CPU or other-arch fallback
 */
#ifndef TEETHRT_H
#define TEETHRT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEETHRT_VERSION 100

typedef struct CUstream_st* trt_stream_t; /* == cudaStream_t */

enum { TRT_OK = 0, TRT_ERR_INVALID = -1, TRT_ERR_CUDA = -2, TRT_ERR_UNSUPPORTED = -3 };

/* BatchNorm statistics buffers (`stats`, `bstats` below) are REPLICATED: [TRT_STAT_REPLICAS][2][C] fp64, zeroed by the caller.
 * Same-address fp64 atomics serialise at the L2 (~23 ns each, measured: a 444-block grid spends 10 us draining one chain), so
 * each block adds into replica (block index mod TRT_STAT_REPLICAS) and the finalise step sums the replicas. */
#define TRT_STAT_REPLICAS 8
int trt_stat_replicas(void);

/* Lazy BatchNorm records (consumer-side finalisation).  The kernel that PRODUCES a train-mode BatchNorm's input only
 * accumulates its statistics; the first kernel that CONSUMES the BatchNorm derives scale/shift for its own channel slice
 * from those statistics in its prologue, and one designated block also publishes the record (and the running statistics)
 * for every later consumer.  That removes the 4-5 us finalise launch that used to sit between producer and consumer where
 * the tensor is large; on small tensors the redundant per-block reads of the statistics would cost more than the launch, so
 * the streaming entry points fall back to enqueueing the finalise kernel themselves (same arithmetic, same results).
 * Host structs; every pointer inside is a device pointer. */
typedef struct {
  const double* stats;             /* [TRT_STAT_REPLICAS][2][C] {sum, sum^2} of the BatchNorm's input, complete */
  const float* gamma;              /* [C] */
  const float* beta;               /* [C] */
  float* running_mean;             /* [C] updated in place, may be NULL */
  float* running_var;              /* [C] */
  long long* num_batches_tracked;  /* may be NULL */
  float* rec;                      /* [4][C] out: scale, shift, mean, rstd */
  double count;                    /* elements per channel */
  float eps, momentum;
} trt_bn_fin_t;
typedef struct {
  const double* bstats;            /* [TRT_STAT_REPLICAS][2][C] {sum dy, sum dy*xhat}, complete */
  const float* rec;                /* [4][C] forward record of this BatchNorm */
  const float* gamma;              /* [C] */
  float* dgamma;                   /* [C] written */
  float* dbeta;                    /* [C] written */
  float* coef;                     /* [3][C] scratch (optional): lets the callee finalise with a launch of its own when the
                                      in-prologue form would cost more than it saves (small tensors, many blocks) */
  double count;
  int raw_x;                       /* 1: bstats[1] holds sum dy*x (trt_gemm_bf16_bnbwd); converted with the record's mean / rstd */
} trt_bn_bwd_fin_t;

int trt_version(void);
const char* trt_last_error_string(void);
/* Select the device, verify it is sm_100, resolve cuTensorMapEncodeTiled. Call once per process/device. */
int trt_init(int device);
/* Launch mode of the kernels that support programmatic dependent launch (the inference chain): on != 0 lets each start
 * while its predecessor in the stream drains and wait on the device (griddepcontrol).  Returns the previous setting; the
 * Python host turns it on around eval forwards only (it measured slower inside the train step).  Process-wide. */
int trt_set_pdl(int on);
/* number of kernels this library has launched in this process (bench.py reports the delta as gpu_launches) */
unsigned long long trt_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * 1x1 convolutions as tcgen05 GEMMs.
 * Replaces: conv_pw / conv_pwl / conv_head (+ BatchNormAct2d + residual) of timm's EfficientNet, reached through
 * `self.backbone(x_img)` experiments/multimodal_v1/train_mm_joint_dualtask.py:154, ui/gradio_app/infer_mm.py:36,
 * experiments/vision_v2/train_mil_attention_v1.py:143 — and their autograd backward (:248).
 * ------------------------------------------------------------------------------------------------------------------ */
#define TRT_EPI_SCALE_SHIFT 1 /* y = acc * scale[n] + shift[n]  (eval-mode BN folded) */
#define TRT_EPI_SILU 2        /* y = silu(y) */
#define TRT_EPI_RESIDUAL 4    /* y += residual[m, n]  (bf16) */
#define TRT_EPI_STATS 8       /* stats[0][n] += sum_m y, stats[1][n] += sum_m y^2  (fp64; train-mode BN batch statistics) */
#define TRT_EPI_MILGATE 16    /* internal: gated-attention score epilogue of trt_mil_attn_fwd_tc (no C output) */
#define TRT_EPI_BNBWD 32      /* with TRT_EPI_STATS: stats[1][n] += sum_m y * bn_x[m,n] instead of sum_m y^2 (trt_gemm_bf16_bnbwd) */

/* C[M,N] (bf16) = epi( A[M,K] (bf16, row-major) . B[N,K]^T (bf16, row-major) ), fp32 accumulate in TMEM.
 * block_n_override: 0 = choose. */
int trt_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int flags, const float* scale,
                  const float* shift, const void* residual, double* stats, int block_n_override, trt_stream_t stream);

/* Data-gradient GEMM that also accumulates the backward sums of the BatchNorm its OUTPUT feeds next in the backward pass:
 * C = A . B^T (+ residual); bstats[0][n] += sum_m C[m,n], bstats[1][n] += sum_m C[m,n] * bn_x[m,n]  (bn_x = the raw input
 * of that BatchNorm; the sums are over the bf16-rounded C that is stored).  Saves the separate trt_bn_bwd_reduce pass over
 * (C, bn_x); trt_affine2 is told through trt_bn_bwd_fin_t.raw_x that the second sum is sum dy*x, not sum dy*xhat. */
int trt_gemm_bf16_bnbwd(const void* A, const void* B, void* C, int M, int N, int K, int flags, const void* residual,
                        const void* bn_x, double* bstats, trt_stream_t stream);

/* out[p*so_p + q*so_q] (fp32) += sum_m P[m,p] * Q[m,q]   (P:[M,Cp], Q:[M,Cq] bf16 row-major; weight gradient).
 * q_store: 0 = Cq; otherwise only columns q < q_store are stored (Q zero-padded beyond, e.g. the stem's 27 of 32 taps).
 * lbo/sbo/kstep_bytes: 0 = canonical MN-major 128B-swizzle descriptor values (bring-up knobs). */
int trt_gemm_wgrad_bf16(const void* P, const void* Q, float* out, int M, int Cp, int Cq, long long so_p, long long so_q,
                        int q_store, int lbo, int sbo, int kstep_bytes, trt_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * On-device input stage (byte-exact vs OpenCV 4.x).
 * Replaces: src/preprocessing/normalise.py:10-16 `apply_clahe` (cv2.cvtColor BGR2LAB -> createCLAHE(3.0,(8,8)).apply(L)
 * -> cv2.cvtColor LAB2BGR), src/preprocessing/pipeline.py:23-29 `centre_crop_resize` (cv2.resize INTER_LINEAR), and
 * ToTensor/Normalize/torch.flip of experiments/multimodal_v1/train_mm_joint_dualtask.py:83-84,91-92,328-333.
 * ------------------------------------------------------------------------------------------------------------------ */
/* Packed Lab lookup tables (built on the host by teethrt.lab_tables): byte offsets */
#define TRT_TAB_GAMMA_OFF 0          /* uint16[256]   sRGBGammaTab_b   */
#define TRT_TAB_CBRT_OFF 512         /* uint16[3072]  LabCbrtTab_b     */
#define TRT_TAB_YF_OFF 6656          /* int32[512]    LabToYF_b        */
#define TRT_TAB_ABXZ_OFF 8704        /* int32[36864]  abToXZ_b         */
#define TRT_TAB_INVGAMMA_OFF 156160  /* uint8[4096]   sRGBInvGammaTab_b */
#define TRT_TAB_BYTES 160256

size_t trt_clahe_workspace_bytes(int n);
/* src/dst: uint8 [n,h,w,3] BGR (HWC). clip = CLAHE_CLIP (3.0), grid fixed at 8x8 (src/config.py:15-16). */
int trt_clahe_bgr_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, float clip, const void* tables,
                     void* workspace, size_t workspace_bytes, trt_stream_t stream);
/* uint8 [n,h,w,3] -> [n,dh,dw,3]; centre_crop != 0 crops the centred min(h,w) square first (pipeline.py:25-28). */
int trt_resize_linear_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int centre_crop, int dh, int dw,
                         trt_stream_t stream);
/* uint8 BGR [n,h,w,3] -> RGB planar [n,3,h,w] = (u8/255 - mean)/std; flip: 0 none, 1 W-reverse, 2 H-reverse.
 * out_bf16 != 0 writes bf16, else fp32. */
int trt_normalize_flip_u8(const uint8_t* src, void* dst, int n, int h, int w, int flip, int out_bf16, trt_stream_t stream);
/* Pillow-exact antialiased resampling pass over uint8 (one separable pass per call; horizontal first, then vertical, each
 * rounding and clipping to uint8 like Pillow's ImagingResample).  Replaces the PIL resize inside the reference's eval
 * transforms: timm create_transform(..., interpolation='bicubic') at ui/gradio_app/infer_mm.py:12-17 and
 * torchvision Resize(512) at ui/gradio_app/infer_mil.py:116-119.  bounds[o] = {first tap, tap count}, coeffs[o][ksize] =
 * 22-bit fixed-point weights (host-built, teethrt.preproc.pil_coeffs).  horizontal: out[y][o] from in row y;
 * vertical: out[o][x] from in column x.  `in` points at the first row/column the tables refer to. */
int trt_resample_u8(const uint8_t* in, size_t in_pitch_bytes, int channels, uint8_t* out, int out_rows, int out_cols,
                    const int* bounds, const int* coeffs, int ksize, int vertical, int swap_channels, trt_stream_t stream);
/* Orientation normalisation, src/preprocessing/normalise.py:19-57 `deskew`, OpenCV-bit-exact stages:
 * trt_canny_nms_bgr_u8      cvtColor(BGR2GRAY) + Sobel 3x3 + |dx|+|dy| + non-maximum suppression of cv2.Canny(gray, low, high)
 *                           -> map[h*w]: 0 none, 1 candidate (> low), 2 strong (> high)
 * trt_canny_hysteresis_pass one flood pass of the strong label through 8-connected candidates (complete inside 32x32 tiles,
 *                           one tile further per pass); sets *changed (device int, caller zeroes it) if anything was promoted;
 *                           repeat until it stays 0
 * trt_canny_finish          edges[h*w] = 255 where map == 2 (may be NULL) and moments = {N, Sum y, Sum x, Sum yy, Sum xy,
 *                           Sum xx} over the edge pixels (the PCA of normalise.py:31-38 needs nothing else)
 * trt_warp_affine_linear_u8 cv2.warpAffine(src, M, (dw, dh), INTER_LINEAR, BORDER_REPLICATE); inverse_map_host = the 2x3
 *                           dst->src matrix (HOST doubles; cv2 inverts M itself, teethrt.preproc.invert_affine does the same) */
int trt_canny_nms_bgr_u8(const uint8_t* bgr, int h, int w, int low, int high, uint8_t* map, trt_stream_t stream);
int trt_canny_hysteresis_pass(uint8_t* map, int h, int w, int* changed, trt_stream_t stream);
int trt_canny_finish(const uint8_t* map, int h, int w, uint8_t* edges, long long* moments, trt_stream_t stream);
int trt_warp_affine_linear_u8(const uint8_t* src, int h, int w, int channels, uint8_t* dst, int dh, int dw,
                              const double* inverse_map_host, trt_stream_t stream);

/* Pillow-exact operations behind timm's RandAugment ('rand-m9-mstd0.5-inc1', the train transform of
 * experiments/multimodal_v1/train_mm_joint_dualtask.py:75-84), one uint8 HWC image per call:
 * trt_hist_u8 + trt_lut_build_u8  per-channel histogram -> lookup table of ImageOps.autocontrast (mode 0) / equalize (mode 1)
 * trt_lut_apply_u8                out = lut[c][img]  (also serves invert / posterize / solarize / solarize_add with a
 *                                 host-built table)
 * trt_enhance_rgb_u8              ImageEnhance.{Brightness 0, Color 1, Contrast 2, Sharpness 3}(img).enhance(factor), RGB;
 *                                 scratch = one 8-byte device word (contrast's luma sum)
 * trt_affine_pil_u8               Image.transform(size, AFFINE, matrix, BILINEAR | BICUBIC, fillcolor) with Pillow's
 *                                 sampler (double coordinates, truncating store); matrix / fill are HOST arrays.
 *                                 Serves rotate, shear_x/y, translate_x/y. */
int trt_hist_u8(const uint8_t* img, size_t n_px, int channels, long long* hist, trt_stream_t stream);
int trt_lut_build_u8(const long long* hist, int channels, int mode, uint8_t* lut, trt_stream_t stream);
int trt_lut_apply_u8(const uint8_t* img, const uint8_t* lut, size_t n_px, int channels, uint8_t* out, trt_stream_t stream);
int trt_enhance_rgb_u8(const uint8_t* img, int h, int w, int mode, float factor, long long* scratch, uint8_t* out,
                       trt_stream_t stream);
int trt_affine_pil_u8(const uint8_t* img, int h, int w, int channels, const double* matrix_host, int bicubic,
                      const uint8_t* fill_host, uint8_t* out, trt_stream_t stream);

/* Batched train transform (one DataLoader batch per call; experiments/multimodal_v1/train_mm_joint_dualtask.py:72-85 runs
 * the same transform image by image inside DataLoader workers).  Job tables are DEVICE arrays the host fills after sampling:
 *   crop job  : int32[8]  {top, left, h, w, flip, 0, 0, 0}  crop box inside the source image; flip mirrors the output
 *   aug job   : 88 bytes  {i64 src, i64 dst, i32 op, i32 mode, f32 factor, i32 bicubic, f64 m[6], i32 slot, u8 fill[4]}
 *               op 0 = lookup table luts[slot] (host-built: invert, posterize, solarize, solarize_add), 1 = histogram LUT
 *               (mode 0 autocontrast, 1 equalize; built on the device into luts[slot]), 2 = ImageEnhance blend (mode 0..3 =
 *               brightness, color, contrast, sharpness), 3 = Image.transform(AFFINE, m)
 *   fin job   : 32 bytes  {i64 src, i32 top, left, eh, ew, i64 noise}  erase box (eh = 0: none) filled from fp32 noise[3][S][S]
 * trt_crop_resize_batch_u8: src [n,h,w,3] -> out [n,size,size,3] = img.crop(box).resize(size) (+ horizontal flip), Pillow-exact;
 *   bounds int32 [n][2][size][2], coeffs int32 [n][2][size][kmax], tmp uint8 [n][hmax][size][3] are caller scratch
 *   (kmax = 2*ceil(support * max(1, hmax_or_wmax / size)) + 1, hmax >= every crop height).
 * trt_aug_layer_batch_u8: one RandAugment layer over `njobs` images (S x S x 3 each); hist u64 [njobs][768] and luma u64
 *   [njobs] must be zeroed by the caller when need_stats (some job is op 1 or contrast).
 * trt_normalize_erase_batch: ToTensor + Normalize (+ RandomErasing 'pixel') -> dst [n,3,S,S] fp32 or bf16. */
int trt_crop_resize_batch_u8(const uint8_t* src, int n, int h, int w, const void* crop_jobs, int size, int bicubic, int kmax,
                             int hmax, int* bounds, int* coeffs, uint8_t* tmp, uint8_t* out, trt_stream_t stream);
int trt_aug_layer_batch_u8(const void* aug_jobs, int njobs, int size, int need_stats, unsigned long long* hist,
                           unsigned long long* luma, uint8_t* luts, trt_stream_t stream);
int trt_normalize_erase_batch(const void* fin_jobs, int n, int size, void* dst, int out_bf16, trt_stream_t stream);
int trt_aug_job_bytes(void);

/* ------------------------------------------------------------------------------------------------------------------
 * BatchNorm / squeeze-excite / pooling kernels around the GEMMs (NHWC bf16, rows = N*H*W, C % 8 == 0).
 * Replace timm's BatchNormAct2d, SqueezeExcite and global_pool inside `self.backbone(x_img)`
 * (experiments/multimodal_v1/train_mm_joint_dualtask.py:154; experiments/vision_v2/train_mil_attention_v1.py:143)
 * and their autograd backward (:248 / :184).
 * A "rec" is a per-BatchNorm record float[4][C] = {scale, shift, mean, rstd} with y = x*scale + shift.
 * ------------------------------------------------------------------------------------------------------------------ */
/* train mode: stats = {sum, sum^2} (fp64[2][C]) over `count` values -> rec; updates running stats (momentum, unbiased var) */
int trt_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* rec, int C, double count, float eps, float momentum,
                    trt_stream_t stream);
/* eval mode: rec from the running statistics */
int trt_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float* rec,
                     int C, float eps, trt_stream_t stream);
/* bstats = {sum dy, sum dy*xhat} -> coef float[3][C] with dx = a*dy + b*x + c; writes dgamma, dbeta */
int trt_bn_bwd_finalize(const double* bstats, const float* rec, const float* gamma, float* coef, float* dgamma, float* dbeta,
                        int C, double count, trt_stream_t stream);
/* out = act(x*scale+shift) (+ residual); act: 0 none, 1 SiLU.  fin_host (optional): `rec` is not final yet - derive
 * scale/shift from fin_host->stats and publish the record (lazy BatchNorm, above) */
int trt_bn_apply(const void* x, const float* rec, const void* residual, void* out, const trt_bn_fin_t* fin_host, int rows, int C,
                 int act, trt_stream_t stream);
/* pooled_sum[n,c] = sum_hw act(bn(x)) (zeroed here; rec may be NULL = plain sum) — SE squeeze and global average pool.
 * zeroed != 0: pooled_sum was cleared by the caller (one arena memset per pass); fin_host: lazy BatchNorm as in trt_bn_apply */
int trt_pool_act(const void* x, const float* rec, float* pooled_sum, int zeroed, const trt_bn_fin_t* fin_host, int N, int HW,
                 int C, int act, trt_stream_t stream);
/* gate[n,c] = sigmoid(We . silu(Wr . mean + br) + be); s1 [N,rd] receives the pre-activation of the reduce conv.
 * apply_x (optional, bf16 [N*HW, C], in place): x[n,hw,c] *= gate[n,c] in the same launch - for the small feature maps of a
 * batch-1 / TTA inference forward, where a separate trt_gate_apply launch is pure latency (x must already be activated) */
int trt_se_fwd(const float* pooled_sum, float inv_hw, const float* Wr, const float* br, const float* We, const float* be,
               float* s1, float* gate, void* apply_x, int HW, int N, int C, int rd, trt_stream_t stream);
/* out = (rec ? silu(bn(x)) : x) * gate[n,c] */
int trt_gate_apply(const void* x, const float* rec, const float* gate, void* out, int N, int HW, int C, trt_stream_t stream);
/* bstats += {sum dy, sum dy*xhat} of a BatchNorm without activation (the project conv's) */
int trt_bn_bwd_reduce(const void* dy, const void* x, const float* rec, double* bstats, int rows, int C, trt_stream_t stream);
/* out = a*dy + b*x + c.  Either coef float[3][C] = {a, b, c} is given, or fin_host: the BatchNorm-backward coefficients are
 * derived from fin_host->bstats in the prologue (lazy, above) and dgamma / dbeta are written by one block */
int trt_affine2(const void* dy, const void* x, const float* coef, void* out, const trt_bn_bwd_fin_t* fin_host, int rows, int C,
                trt_stream_t stream);
/* sums[0][n,c] = dgate_pre = sum_hw dA * silu(bn(x)) (zeroed here unless `zeroed`).  full != 0: sums is [5][N][C] and also
 * receives S1 = sum dA*s', S2 = sum s', S3 = sum dA*s'*x, S4 = sum s'*x (s' = silu'(bn(x))): with them the BatchNorm-backward
 * sums of the gated activation follow from [N,C]-sized arrays once the SE MLP backward has produced gate-gradient terms,
 * so the tensor is read twice in the backward pass instead of three times (trt_se_bwd + trt_act_bwd_apply below). */
int trt_se_bwd_reduce(const void* dA, const void* x, const float* rec, float* sums, int zeroed, int full, int N, int HW, int C,
                      trt_stream_t stream);
/* BatchNorm backward of the depthwise output folded into the SE MLP backward (host struct, device pointers) */
typedef struct {
  const float* sums;               /* [5][N][C] from trt_se_bwd_reduce(full) */
  const float* rec;                /* [4][C] record of the BatchNorm in front of the SiLU */
  const float* gamma;              /* [C] */
  float* coef;                     /* [3][C] out: dx = a*g + b*x + c */
  float* dgamma;                   /* [C] written */
  float* dbeta;                    /* [C] written */
  double count;                    /* N*HW */
} trt_se_bn_t;
/* SE MLP backward: ds2 [N,C], ds1 [N,rd], dmean [N,C] scratch/outputs; parameter gradients are WRITTEN (=).
 * bn_host (optional): also derive the BatchNorm-backward coefficients / dgamma / dbeta of the gated activation */
int trt_se_bwd(const float* dgate_pre, const float* gate, const float* s1, const float* pooled_sum, float inv_hw,
               const float* Wr, const float* We, float* ds2, float* ds1, float* dmean, float* dWr, float* dbr, float* dWe,
               float* dbe, int ds1_zeroed, const trt_se_bn_t* bn_host, int N, int C, int rd, trt_stream_t stream);
/* One-launch versions of trt_se_fwd / trt_se_bwd for training batches (timm SqueezeExcite inside `self.backbone(x_img)`,
 * experiments/multimodal_v1/train_mm_joint_dualtask.py:154, and its backward, :248).  One block per channel chunk (at most one
 * per SM); the contraction over channels is split across the blocks and combined behind a grid-wide barrier, so nothing is
 * accumulated atomically.  `workspace`: trt_se_workspace_bytes(N, C, rd) bytes, 16-byte aligned, ZEROED once by the caller
 * before its first use (the first 256 bytes hold the barrier state) and owned by one stream at a time.  Results equal the
 * two-launch entry points up to fp32 summation order; ds2 may be NULL. */
size_t trt_se_workspace_bytes(int N, int C, int rd);
int trt_se_fwd_fused(const float* pooled_sum, float inv_hw, const float* Wr, const float* br, const float* We, const float* be,
                     float* s1, float* gate, void* workspace, size_t ws_bytes, int N, int C, int rd, trt_stream_t stream);
int trt_se_bwd_fused(const float* dgate_pre, const float* gate, const float* s1, const float* pooled_sum, float inv_hw,
                     const float* Wr, const float* We, float* ds2, float* ds1, float* dmean, float* dWr, float* dbr, float* dWe,
                     float* dbe, const trt_se_bn_t* bn_host, void* workspace, size_t ws_bytes, int N, int C, int rd,
                     trt_stream_t stream);
/* out = a*g + b*x + c with g = (dA*gate[n,c] + dmean[n,c]*inv_hw) * silu'(bn(x)) formed on the fly: the gradient w.r.t. the
 * raw depthwise output in one read of (dA, x) and one write */
int trt_act_bwd_apply(const void* dA, const float* gate, const float* dmean, float inv_hw, const void* x, const float* rec,
                      const float* coef, void* out, int N, int HW, int C, trt_stream_t stream);
/* g = (dA*gate[n,c] + dmean[n,c]*inv_hw) * (act ? silu'(bn(x)) : 1); bstats += {sum g, sum g*xhat}. dA/gate/dmean may be NULL */
int trt_act_bwd(const void* dA, const float* gate, const float* dmean, float inv_hw, const void* x, const float* rec,
                void* g_out, double* bstats, int N, int HW, int C, int act, trt_stream_t stream);
/* x *= alpha (fp32; turns pooled sums into the global-average-pool features) */
int trt_scale_f32(float* x, size_t n, float alpha, trt_stream_t stream);
/* fp32 [N,K] -> bf16 [N,K] and (optional) bf16 [K,N] */
int trt_pack_w1x1(const float* w, void* w_bf16, void* wt_bf16, int N, int K, trt_stream_t stream);
/* the same for `count` weights in one launch.  table_dev: device array of count x 6 int64 {w, w_bf16, wt_bf16 (or 0), N, K,
 * first_tile}, first_tile = running sum of ceil(N/32)*ceil(K/32) over the preceding entries; total_tiles = the full sum. */
int trt_pack_w1x1_batch(const long long* table_dev, int count, int total_tiles, trt_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Depthwise and stem convolutions (TF 'same' padding), forward + backward.  Replace timm's Conv2dSame / depthwise
 * conv_dw / conv_stem inside `self.backbone(x_img)` (train_mm_joint_dualtask.py:154) and autograd (:248).
 * Weights and weight gradients stay in torch layout ([C,1,k,k] / [CS,3,3,3], fp32).
 * ------------------------------------------------------------------------------------------------------------------ */
/* x: [N,H,W,C]; in_rec != NULL: input = silu(bn(x)) applied on load; out_rec != NULL (eval): out = silu(bn_out(conv)),
 * pooled_sum[n,c] (optional; zeroed here unless pooled_zeroed: one arena memset per forward instead of one per block)
 * += sum_hw out; stats != NULL (train): fp64 {sum, sum^2} of the raw output.
 * in_fin_host (optional): the INPUT's BatchNorm is lazy - in_rec is derived from in_fin_host->stats and published.
 * act_out (optional, stride 1 with in_rec, [N,H,W,C] bf16): receives silu(bn(x)), the activated input the kernel forms in
 * shared memory anyway; trt_dwconv_bwd's weight gradient then takes it as x_raw with x_rec = NULL and skips the activation. */
int trt_dwconv_fwd(const void* x, const float* in_rec, const float* w, void* out, const float* out_rec, float* pooled_sum,
                   int pooled_zeroed, double* stats, const trt_bn_fin_t* in_fin_host, void* act_out, int N, int H, int W, int C,
                   int k, int s, trt_stream_t stream);
/* gy = dD, the gradient w.r.t. the RAW depthwise output (the BN-backward affine of the following BatchNorm has already been
 * applied by trt_affine2).  g_out (NULL = skip) = convT(dD) * (x_rec ? silu'(bn(x_raw)) : 1), bstats += {sum g, sum g*xhat};
 * dw[C,1,k,k] (NULL = skip) += correlation of dD with act(x) (act = silu(bn) when x_rec).  The two halves are independent:
 * a caller may issue them as two calls on two streams. */
int trt_dwconv_bwd(const void* gy, const float* w, const void* x_raw, const float* x_rec, void* g_out, double* bstats,
                   float* dw, int N, int H, int W, int C, int k, int s, trt_stream_t stream);
/* x: NCHW [N,3,H,W] fp32 or bf16 -> out NHWC bf16 [N,ceil(H/2),ceil(W/2),CS]; CS in {32, 48} */
int trt_stem_fwd(const void* x, int x_is_bf16, const float* w, void* out, const float* out_rec, double* stats, int N, int H,
                 int W, int CS, trt_stream_t stream);
int trt_stem_wgrad(const void* x, int x_is_bf16, const void* ds, float* dw, int N, int H, int W, int CS, trt_stream_t stream);
/* Train-mode stem as tcgen05 GEMMs: patches [N*OH*OW, 32] bf16 = im2col(x) (column ci*9+kh*3+kw, 5 zero columns);
 * forward = trt_gemm_bf16(patches, w_bf16[CS,32]) with BN statistics, weight gradient = trt_gemm_wgrad_bf16(dS, patches,
 * q_store = 27).  trt_stem_pack_w: fp32 [CS,3,3,3] -> bf16 [CS,32]. */
int trt_stem_im2col(const void* x, int x_is_bf16, void* patches, int N, int H, int W, trt_stream_t stream);
int trt_stem_pack_w(const float* w, void* w_bf16, int CS, trt_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * MIL gated-attention pooling.  Replaces AttentionMIL.forward (experiments/vision_v2/train_mil_attention_v1.py:124-130)
 * and MILAttention.forward (ui/gradio_app/infer_mil.py:62-68).  H [B,K,D] fp32, V/U [hid,D], w [hid].
 * ------------------------------------------------------------------------------------------------------------------ */
size_t trt_mil_attn_smem_bytes(int K, int D, int hid, int backward);
int trt_mil_attn_fwd(const float* H, const float* Vw, const float* Vb, const float* Uw, const float* Ub, const float* ww,
                     const float* wb, float* M, float* A, float* gV, float* gU, int B, int K, int D, int hid,
                     trt_stream_t stream);
/* parameter gradients are ACCUMULATED (+=): zero them first.  gV / gU (the gate activations saved by the forward) are
 * CONSUMED: they are overwritten with the gate gradients dv / du, which the weight-gradient and dH kernels then read. */
/* The same forward with the score projection on the tensor cores (tcgen05 GEMM over split-bf16 operands, fp32-level
 * accuracy: hi.Whi + lo.Whi + hi.Wlo; gate math in the epilogue).  workspace: trt_mil_attn_tc_workspace_bytes(B, K, D, hid)
 * bytes, 256-byte aligned; needs D % 64 == 0 and hid % 8 == 0 (1280 / 128 and 256 in the reference). */
size_t trt_mil_attn_tc_workspace_bytes(int B, int K, int D, int hid);
int trt_mil_attn_fwd_tc(const float* H, const float* Vw, const float* Vb, const float* Uw, const float* Ub, const float* ww,
                        const float* wb, float* M, float* A, float* gV, float* gU, int B, int K, int D, int hid, void* workspace,
                        size_t workspace_bytes, trt_stream_t stream);
int trt_mil_attn_bwd(const float* dM, const float* H, const float* A, float* gV, float* gU, const float* Vw,
                     const float* Uw, const float* ww, float* dH, float* dVw, float* dVb, float* dUw, float* dUb, float* dww,
                     float* dwb, int B, int K, int D, int hid, trt_stream_t stream);

/* MIL bag head: logit[b] = <dropout(M[b]), w> + bias (MILNet.drop + MILNet.head, train_mil_attention_v1.py:146-147;
 * infer_mil.py:95) and its backward (dw, db WRITTEN); BCE-with-logits mean loss + gradient (train_mil_attention_v1.py:183). */
int trt_linear1_fwd(const float* M, const float* w, const float* bias, float* logit, int B, int D, float drop_p,
                    unsigned long long seed, const unsigned long long* step, trt_stream_t stream);
int trt_linear1_bwd(const float* dlogit, const float* M, const float* w, float* dM, float* dw, float* db, int B, int D,
                    float drop_p, unsigned long long seed, const unsigned long long* step, trt_stream_t stream);
int trt_bce_logits(const float* logit, const float* y, const float* sample_w, float* loss, float* dlogit, int B,
                   trt_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Tabular MLP + late fusion + dual heads + dual BCE loss.  Replaces MMJointDualHead.tab / fusion / cls_head / reg_head
 * (experiments/multimodal_v1/train_mm_joint_dualtask.py:140-159) and the loss (:176-179, :244-247).
 * params_host / grads_host: HOST arrays of 10 DEVICE pointers in the order
 *   tab.0.weight, tab.0.bias, tab.1.weight, tab.1.bias, tab.4.weight, tab.4.bias, cls_head.weight, cls_head.bias,
 *   reg_head.weight, reg_head.bias.
 * ------------------------------------------------------------------------------------------------------------------ */
size_t trt_tab_heads_scratch_floats(int B, int Hd);
int trt_tab_heads_fwd(const float* feat, const float* xtab, const float* const* params_host, float* bn_rm, float* bn_rv,
                      long long* bn_nbt, const float* y_hard, const float* y_soft, const float* sample_w, float* logit,
                      float* reg, float* loss, float* dlogit, float* dreg, float* scratch, int B, int T, int Hd, int F,
                      int train, float drop_p, float alpha, float beta, unsigned long long seed,
                      const unsigned long long* step, trt_stream_t stream);
int trt_tab_heads_bwd(const float* feat, const float* xtab, const float* const* params_host, const float* bn_rm,
                      const float* bn_rv, const float* dlogit, const float* dreg, float* dfeat, float* const* grads_host,
                      float* scratch, int B, int T, int Hd, int F, int train, float drop_p, unsigned long long seed,
                      const unsigned long long* step, trt_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused optimiser over flat fp32 buffers.  Replaces unscale_/clip_grad_norm_/AdamW.step/CosineAnnealingLR.step,
 * experiments/multimodal_v1/train_mm_joint_dualtask.py:217-220,249-254.
 * state (device, trt_optim_state_bytes()): {u64 step; f64 lr0, t_max, beta1, beta2; f32 lr, bc1, bc2, pad; f64 reserved}
 * ------------------------------------------------------------------------------------------------------------------ */
size_t trt_optim_state_bytes(void);
int trt_optim_advance(void* state, trt_stream_t stream);                       /* ++step, lr(step), bias corrections */
int trt_grad_sumsq(const float* g, size_t n, double* out, trt_stream_t stream); /* out[0] = sum g^2 */
/* g_eff = g * grad_scale * min(1, max_norm / (|g*grad_scale| + 1e-6)); AdamW(lr, betas, eps, weight_decay) */
int trt_adamw_step(float* p, const float* g, float* m, float* v, size_t n, const void* state, const double* normsq,
                   float* norm_out, float grad_scale, float max_norm, float eps, float weight_decay, trt_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fold calibration.  Replaces TemperatureScaler + the LBFGS closure
 * (experiments/multimodal_v1/train_mm_joint_dualtask.py:162-174, 271-287), compute_metrics (:181-186) and the 61-point
 * F1 threshold sweep (:289-295).
 * trt_temperature_nll : loss_grad[0] = mean BCE(logits / exp(log_T), targets), loss_grad[1] = d loss / d log_T
 * trt_scaled_sigmoid  : prob = sigmoid(logits / T)                                  (:286-287, :336)
 * trt_binary_metrics  : counts[t] = {tp, fp, fn, tn} of (double)prob >= thr[t] for every threshold in one launch
 *                       (prob fp32, or fp64 when prob_is_f64);
 *                       auc = {2*#(pos > neg) + #(pos == neg), #pos, #neg, #labels outside {0,1}}; ROC-AUC = auc[0] /
 *                       (2 * auc[1] * auc[2]) (the Mann-Whitney form of sklearn.metrics.roc_auc_score, ties = 1/2)
 * ------------------------------------------------------------------------------------------------------------------ */
int trt_temperature_nll(const float* logits, const float* targets, const float* log_T, float* loss_grad, int n,
                        trt_stream_t stream);
int trt_scaled_sigmoid(const float* logits, float T, float* prob, int n, trt_stream_t stream);
int trt_binary_metrics(const void* prob, int prob_is_f64, const float* y, int n, const double* thr, int nthr,
                       long long* counts, long long* auc, trt_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Late-fusion stacker.  Replaces sklearn LogisticRegression(max_iter=1000).fit / predict_proba over the
 * stream probabilities (experiments/fusion_v1/stack_blend.py:245-252, ui/gradio_app/stack_meta.py:55-57,114-116); the
 * threshold modes of choose_threshold (:49-86) run on trt_binary_metrics with fp64 scores.
 * trt_logreg_fit: min 0.5|w|^2 + C sum log(1+exp(-s(w.x+b))), X [n,d] fp64 row-major, y in {0,1}; coef = {w[0..d), b};
 *                 info = {Newton passes, |grad|_inf at exit, objective}.  d <= 4.
 * ------------------------------------------------------------------------------------------------------------------ */
int trt_logreg_fit(const double* X, const float* y, int n, int d, double C, int max_iter, double tol, double* coef,
                   double* info, trt_stream_t stream);
int trt_logreg_predict(const double* X, int n, int d, const double* coef, double* prob, trt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TEETHRT_H */
