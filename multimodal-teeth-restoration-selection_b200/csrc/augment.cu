// Pillow-exact image operations behind timm's RandAugment policy 'rand-m9-mstd0.5-inc1' (the train
// transform of experiments/multimodal_v1/train_mm_joint_dualtask.py:75-84 builds it with timm.data.create_transform).
// The per-pixel arithmetic lives in augment_core.h (shared with the CPU harness that pins it to Pillow); this file only maps
// it onto the grid.  Everything is uint8 HWC streaming work on one image (a data-loader transform): latency-sized launches.
// Compiled with --fmad=false: Pillow's float / double expressions must not be contracted into fused multiply-adds.
#include "common.cuh"
#include "augment_core.h"

namespace {

__global__ void __launch_bounds__(256) hist_u8_kernel(const uint8_t* __restrict__ img, size_t n_px, int ch, long long* __restrict__ hist) {
  __shared__ unsigned int sh[3 * 256];
  for (int i = threadIdx.x; i < ch * 256; i += 256) sh[i] = 0;
  __syncthreads();
  for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < n_px; p += (size_t)gridDim.x * 256)
    for (int c = 0; c < ch; ++c) atomicAdd(&sh[c * 256 + img[p * ch + c]], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < ch * 256; i += 256)
    if (sh[i]) atomicAdd(reinterpret_cast<unsigned long long*>(hist) + i, (unsigned long long)sh[i]);
}

__global__ void lut_build_kernel(const long long* __restrict__ hist, int mode, uint8_t* __restrict__ lut) {
  const int c = threadIdx.x;
  if (mode == 0) pil_autocontrast_lut(hist + c * 256, lut + c * 256);
  else pil_equalize_lut(hist + c * 256, lut + c * 256);
}

__global__ void __launch_bounds__(256) lut_apply_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ lut, size_t n,
                                                        int ch, uint8_t* __restrict__ out) {
  __shared__ uint8_t sl[3 * 256];
  for (int i = threadIdx.x; i < ch * 256; i += 256) sl[i] = lut[i];
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = sl[(i % ch) * 256 + img[i]];
}

__global__ void __launch_bounds__(256) luma_sum_kernel(const uint8_t* __restrict__ img, size_t n_px, unsigned long long* __restrict__ sum) {
  __shared__ unsigned long long red[8];
  unsigned long long s = 0;
  for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < n_px; p += (size_t)gridDim.x * 256)
    s += (unsigned long long)pil_luma(img[p * 3], img[p * 3 + 1], img[p * 3 + 2]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    if (t) atomicAdd(sum, t);
  }
}

// ImageEnhance.{Brightness, Color, Contrast, Sharpness}(img).enhance(factor) on an RGB image
__global__ void __launch_bounds__(256) enhance_kernel(const uint8_t* __restrict__ img, int H, int W, int mode, float factor,
                                                      const unsigned long long* __restrict__ luma_sum, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t o = ((size_t)y * W + x) * 3;
  uint8_t deg[3] = {0, 0, 0};
  if (mode == 1) {
    const uint8_t l = (uint8_t)pil_luma(img[o], img[o + 1], img[o + 2]);
    deg[0] = deg[1] = deg[2] = l;
  } else if (mode == 2) {
    // mean = int(ImageStat.Stat(img.convert("L")).mean[0] + 0.5)
    const uint8_t m = (uint8_t)(int)((double)*luma_sum / (double)((size_t)H * W) + 0.5);
    deg[0] = deg[1] = deg[2] = m;
  } else if (mode == 3) {
    const bool border = x == 0 || y == 0 || x == W - 1 || y == H - 1;
    for (int c = 0; c < 3; ++c) deg[c] = border ? img[o + c] : pil_smooth3x3(img + o + c, 3, (long)W * 3);
  }
  for (int c = 0; c < 3; ++c) out[o + c] = pil_blend(deg[c], img[o + c], factor);
}

struct AffineArgs { double m[6]; uint8_t fill[4]; };

template <int CH>
__global__ void __launch_bounds__(256) affine_pil_kernel(const uint8_t* __restrict__ img, int H, int W, const AffineArgs a, int bicubic,
                                                         uint8_t* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  uint8_t px[CH];
  const int ok = pil_affine_pixel<CH>(img, H, W, a.m, bicubic, x, y, px);
  uint8_t* q = out + ((size_t)y * W + x) * CH;
  for (int c = 0; c < CH; ++c) q[c] = ok ? px[c] : a.fill[c];
}

// ================================================================================================= batched train transform
// One DataLoader batch per call: every stage is ONE launch over all the images that need it, driven by
// small per-image job tables the host samples in one vectorised call and uploads once.  The per-pixel code is the same
// augment_core.h code as the single-image kernels above, so the batch is bit-identical to running them image by image.

// ---- RandomResizedCrop: per-image Pillow coefficient tables built on the device, then the two separable passes
struct CropJob { int top, left, h, w, flip, pad0, pad1, pad2; };    // crop box inside the source image; flip: mirror the output columns

// grid (B, 2): axis 0 = horizontal (crop w -> S), 1 = vertical (crop h -> S); thread = output index
__global__ void pil_coeffs_batch_kernel(const CropJob* __restrict__ jobs, int S, int bicubic, int kmax, int* __restrict__ bounds,
                                        int* __restrict__ coeffs) {
  const int b = blockIdx.x, axis = blockIdx.y;
  const CropJob j = jobs[b];
  const int in_size = axis ? j.h : j.w;
  for (int o = threadIdx.x; o < S; o += blockDim.x) {
    int first;
    int* k = coeffs + (((size_t)b * 2 + axis) * S + o) * kmax;
    const int n = pil_resample_taps(in_size, S, bicubic, o, kmax, &first, k);
    bounds[(((size_t)b * 2 + axis) * S + o) * 2] = first;
    bounds[(((size_t)b * 2 + axis) * S + o) * 2 + 1] = n;
  }
}

// horizontal pass: src [B,H,W,3] crop rows -> tmp [B,Hmax,S,3] (rows 0..crop_h), columns mirrored when the job flips
__global__ void __launch_bounds__(256) resample_h_batch_kernel(const uint8_t* __restrict__ src, int H, int W, const CropJob* __restrict__ jobs,
                                                               const int* __restrict__ bounds, const int* __restrict__ coeffs, int kmax,
                                                               int S, int Hmax, uint8_t* __restrict__ tmp) {
  const int x = blockIdx.x * 256 + threadIdx.x, r = blockIdx.y, b = blockIdx.z;
  const CropJob j = jobs[b];
  if (x >= S || r >= j.h) return;
  const size_t t = ((size_t)b * 2 + 0) * S + x;
  const int first = bounds[2 * t], taps = bounds[2 * t + 1];
  const int* k = coeffs + t * kmax;
  const uint8_t* p = src + (((size_t)b * H + j.top + r) * W + j.left + first) * 3;
  int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
  for (int i = 0; i < taps; ++i, p += 3) {
    const int w = __ldg(k + i);
    a0 += (int)__ldg(p) * w; a1 += (int)__ldg(p + 1) * w; a2 += (int)__ldg(p + 2) * w;
  }
  const int xo = j.flip ? S - 1 - x : x;
  uint8_t* q = tmp + (((size_t)b * Hmax + r) * S + xo) * 3;
  a0 >>= 22; a1 >>= 22; a2 >>= 22;
  q[0] = (uint8_t)(a0 < 0 ? 0 : (a0 > 255 ? 255 : a0));
  q[1] = (uint8_t)(a1 < 0 ? 0 : (a1 > 255 ? 255 : a1));
  q[2] = (uint8_t)(a2 < 0 ? 0 : (a2 > 255 ? 255 : a2));
}

// vertical pass: tmp [B,Hmax,S,3] -> out [B,S,S,3]
__global__ void __launch_bounds__(256) resample_v_batch_kernel(const uint8_t* __restrict__ tmp, const int* __restrict__ bounds,
                                                               const int* __restrict__ coeffs, int kmax, int S, int Hmax,
                                                               uint8_t* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= S) return;
  const size_t t = ((size_t)b * 2 + 1) * S + y;
  const int first = bounds[2 * t], taps = bounds[2 * t + 1];
  const int* k = coeffs + t * kmax;
  const uint8_t* p = tmp + (((size_t)b * Hmax + first) * S + x) * 3;
  int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
  for (int i = 0; i < taps; ++i, p += (size_t)S * 3) {
    const int w = __ldg(k + i);
    a0 += (int)__ldg(p) * w; a1 += (int)__ldg(p + 1) * w; a2 += (int)__ldg(p + 2) * w;
  }
  uint8_t* q = out + (((size_t)b * S + y) * S + x) * 3;
  a0 >>= 22; a1 >>= 22; a2 >>= 22;
  q[0] = (uint8_t)(a0 < 0 ? 0 : (a0 > 255 ? 255 : a0));
  q[1] = (uint8_t)(a1 < 0 ? 0 : (a1 > 255 ? 255 : a1));
  q[2] = (uint8_t)(a2 < 0 ? 0 : (a2 > 255 ? 255 : a2));
}

// ---- RandAugment: one job per (image, layer) that drew an operation
enum { AUG_LUT = 0, AUG_HISTLUT = 1, AUG_ENHANCE = 2, AUG_AFFINE = 3 };
struct AugJob {                 // 88 bytes, mirrored by teethrt.augment (numpy structured dtype; checked against trt_aug_job_bytes)
  long long src, dst;           // device pointers of the image's current / next S x S x 3 buffer
  int op, mode;                 // AUG_*; mode: HISTLUT 0 = autocontrast, 1 = equalize; ENHANCE 0..3 = brightness, color, contrast, sharpness
  float factor;                 // ENHANCE blend factor
  int bicubic;                  // AFFINE sampler
  double m[6];                  // AFFINE inverse matrix
  int slot;                     // LUT / histogram / luma-sum slot of this job
  unsigned char fill[4];        // AFFINE fill colour
};

// pre-pass of a layer: histograms (HISTLUT) or the luma sum (ENHANCE contrast) of the jobs that need them.
// hist [J][3][256] u64 and luma [J] u64 are zeroed by the caller.  grid (blocks, J)
__global__ void __launch_bounds__(256) aug_stats_batch_kernel(const AugJob* __restrict__ jobs, int S, unsigned long long* __restrict__ hist,
                                                              unsigned long long* __restrict__ luma) {
  const AugJob j = jobs[blockIdx.y];
  const uint8_t* img = reinterpret_cast<const uint8_t*>(j.src);
  const size_t n_px = (size_t)S * S;
  if (j.op == AUG_HISTLUT) {
    __shared__ unsigned int sh[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += 256) sh[i] = 0;
    __syncthreads();
    for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < n_px; p += (size_t)gridDim.x * 256)
      for (int c = 0; c < 3; ++c) atomicAdd(&sh[c * 256 + img[p * 3 + c]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 256; i += 256)
      if (sh[i]) atomicAdd(hist + (size_t)j.slot * 768 + i, (unsigned long long)sh[i]);
  } else if (j.op == AUG_ENHANCE && j.mode == 2) {
    __shared__ unsigned long long red[8];
    unsigned long long s = 0;
    for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < n_px; p += (size_t)gridDim.x * 256)
      s += (unsigned long long)pil_luma(img[p * 3], img[p * 3 + 1], img[p * 3 + 2]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int w = 0; w < 8; ++w) t += red[w];
      if (t) atomicAdd(luma + j.slot, t);
    }
  }
}

// grid J, 3 threads: the HISTLUT jobs' tables into the LUT pool
__global__ void aug_lut_build_batch_kernel(const AugJob* __restrict__ jobs, const unsigned long long* __restrict__ hist, uint8_t* __restrict__ luts) {
  const AugJob j = jobs[blockIdx.x];
  if (j.op != AUG_HISTLUT) return;
  const int c = threadIdx.x;
  const long long* h = reinterpret_cast<const long long*>(hist) + (size_t)j.slot * 768 + c * 256;
  if (j.mode == 0) pil_autocontrast_lut(h, luts + (size_t)j.slot * 768 + c * 256);
  else pil_equalize_lut(h, luts + (size_t)j.slot * 768 + c * 256);
}

// the layer itself: grid (ceil(S*S/256), J), thread = pixel
__global__ void __launch_bounds__(256) aug_apply_batch_kernel(const AugJob* __restrict__ jobs, int S, const uint8_t* __restrict__ luts,
                                                              const unsigned long long* __restrict__ luma) {
  const AugJob j = jobs[blockIdx.y];
  const uint8_t* img = reinterpret_cast<const uint8_t*>(j.src);
  uint8_t* out = reinterpret_cast<uint8_t*>(j.dst);
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= S * S) return;
  const size_t o = (size_t)p * 3;
  if (j.op == AUG_LUT || j.op == AUG_HISTLUT) {
    const uint8_t* lut = luts + (size_t)j.slot * 768;
    for (int c = 0; c < 3; ++c) out[o + c] = __ldg(lut + c * 256 + img[o + c]);
  } else if (j.op == AUG_ENHANCE) {
    const int x = p % S, y = p / S;
    uint8_t deg[3] = {0, 0, 0};
    if (j.mode == 1) {
      const uint8_t l = (uint8_t)pil_luma(img[o], img[o + 1], img[o + 2]);
      deg[0] = deg[1] = deg[2] = l;
    } else if (j.mode == 2) {
      const uint8_t m = (uint8_t)(int)((double)luma[j.slot] / (double)((size_t)S * S) + 0.5);
      deg[0] = deg[1] = deg[2] = m;
    } else if (j.mode == 3) {
      const bool border = x == 0 || y == 0 || x == S - 1 || y == S - 1;
      for (int c = 0; c < 3; ++c) deg[c] = border ? img[o + c] : pil_smooth3x3(img + o + c, 3, (long)S * 3);
    }
    for (int c = 0; c < 3; ++c) out[o + c] = pil_blend(deg[c], img[o + c], j.factor);
  } else {
    uint8_t px[3];
    const int ok = pil_affine_pixel<3>(img, S, S, j.m, j.bicubic, p % S, p / S, px);
    for (int c = 0; c < 3; ++c) out[o + c] = ok ? px[c] : j.fill[c];
  }
}

// ---- ToTensor + Normalize + RandomErasing: per-image source pointer (the image's final buffer) and erase box
struct FinJob { long long src; int top, left, eh, ew; long long noise; };     // eh == 0: no erase; noise: fp32 [3][S][S] device pointer
template <typename OutT>
__global__ void __launch_bounds__(256) normalize_erase_batch_kernel(const FinJob* __restrict__ jobs, int S, OutT* __restrict__ dst) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= S) return;
  const FinJob j = jobs[b];
  const uint8_t* px = reinterpret_cast<const uint8_t*>(j.src) + ((size_t)y * S + x) * 3;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const bool erase = j.eh > 0 && y >= j.top && y < j.top + j.eh && x >= j.left && x < j.left + j.ew;
  OutT* o = dst + (size_t)b * 3 * S * S + (size_t)y * S + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v;
    if (erase) v = reinterpret_cast<const float*>(j.noise)[((size_t)c * S + y) * S + x];
    else v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[c], 255.f), mean[c]), stdv[c]);
    if constexpr (sizeof(OutT) == 2) o[(size_t)c * S * S] = __float2bfloat16_rn(v);
    else o[(size_t)c * S * S] = v;
  }
}

inline int stream_blocks(size_t n) {
  size_t b = (n + 256 * 8 - 1) / (256 * 8);
  const size_t cap = (size_t)4 * trt_num_sms();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int trt_hist_u8(const uint8_t* img, size_t n_px, int channels, long long* hist, cudaStream_t stream) {
  TRT_REQUIRE(img && hist && n_px > 0 && (channels == 1 || channels == 3), "trt_hist_u8: bad argument");
  TRT_CUDA(cudaMemsetAsync(hist, 0, (size_t)channels * 256 * sizeof(long long), stream));
  hist_u8_kernel<<<stream_blocks(n_px), 256, 0, stream>>>(img, n_px, channels, hist);
  return trt_check_launch("trt_hist_u8");
}

extern "C" int trt_lut_build_u8(const long long* hist, int channels, int mode, uint8_t* lut, cudaStream_t stream) {
  TRT_REQUIRE(hist && lut && (channels == 1 || channels == 3) && (mode == 0 || mode == 1), "trt_lut_build_u8: bad argument");
  lut_build_kernel<<<1, channels, 0, stream>>>(hist, mode, lut);
  return trt_check_launch("trt_lut_build_u8");
}

extern "C" int trt_lut_apply_u8(const uint8_t* img, const uint8_t* lut, size_t n_px, int channels, uint8_t* out, cudaStream_t stream) {
  TRT_REQUIRE(img && lut && out && n_px > 0 && (channels == 1 || channels == 3), "trt_lut_apply_u8: bad argument");
  lut_apply_kernel<<<stream_blocks(n_px * channels), 256, 0, stream>>>(img, lut, n_px * channels, channels, out);
  return trt_check_launch("trt_lut_apply_u8");
}

extern "C" int trt_enhance_rgb_u8(const uint8_t* img, int h, int w, int mode, float factor, long long* scratch, uint8_t* out,
                                  cudaStream_t stream) {
  TRT_REQUIRE(img && out && h > 0 && w > 0 && mode >= 0 && mode <= 3, "trt_enhance_rgb_u8: bad argument");
  TRT_REQUIRE(mode != 2 || scratch, "trt_enhance_rgb_u8: contrast needs an 8-byte scratch word");
  if (mode == 2) {
    TRT_CUDA(cudaMemsetAsync(scratch, 0, sizeof(long long), stream));
    luma_sum_kernel<<<stream_blocks((size_t)h * w), 256, 0, stream>>>(img, (size_t)h * w, reinterpret_cast<unsigned long long*>(scratch));
    trt_count_launch(1);
  }
  enhance_kernel<<<dim3((w + 255) / 256, h), 256, 0, stream>>>(img, h, w, mode, factor,
                                                                reinterpret_cast<const unsigned long long*>(scratch), out);
  return trt_check_launch("trt_enhance_rgb_u8");
}

extern "C" int trt_affine_pil_u8(const uint8_t* img, int h, int w, int channels, const double* matrix_host, int bicubic,
                                 const uint8_t* fill_host, uint8_t* out, cudaStream_t stream) {
  TRT_REQUIRE(img && out && matrix_host && fill_host && h > 0 && w > 0, "trt_affine_pil_u8: bad argument");
  TRT_REQUIRE(channels == 1 || channels == 3, "trt_affine_pil_u8: %d channels not built (1 or 3)", channels);
  AffineArgs a;
  for (int i = 0; i < 6; ++i) a.m[i] = matrix_host[i];
  for (int i = 0; i < 4; ++i) a.fill[i] = i < channels ? fill_host[i] : 0;
  dim3 grid((w + 255) / 256, h);
  if (channels == 3) affine_pil_kernel<3><<<grid, 256, 0, stream>>>(img, h, w, a, bicubic, out);
  else affine_pil_kernel<1><<<grid, 256, 0, stream>>>(img, h, w, a, bicubic, out);
  return trt_check_launch("trt_affine_pil_u8");
}

// ------------------------------------------------------------------------------------------------ batched entry points
extern "C" int trt_crop_resize_batch_u8(const uint8_t* src, int n, int h, int w, const void* crop_jobs, int size, int bicubic,
                                        int kmax, int hmax, int* bounds, int* coeffs, uint8_t* tmp, uint8_t* out,
                                        cudaStream_t stream) {
  TRT_REQUIRE(src && crop_jobs && bounds && coeffs && tmp && out, "trt_crop_resize_batch_u8: null pointer");
  TRT_REQUIRE(n > 0 && h > 0 && w > 0 && size > 0 && kmax > 0 && hmax > 0 && hmax <= h, "trt_crop_resize_batch_u8: bad shape");
  const CropJob* jobs = reinterpret_cast<const CropJob*>(crop_jobs);
  pil_coeffs_batch_kernel<<<dim3(n, 2), 256, 0, stream>>>(jobs, size, bicubic, kmax, bounds, coeffs);
  trt_count_launch(1);
  resample_h_batch_kernel<<<dim3((size + 255) / 256, hmax, n), 256, 0, stream>>>(src, h, w, jobs, bounds, coeffs, kmax, size, hmax, tmp);
  trt_count_launch(1);
  resample_v_batch_kernel<<<dim3((size + 255) / 256, size, n), 256, 0, stream>>>(tmp, bounds, coeffs, kmax, size, hmax, out);
  return trt_check_launch("trt_crop_resize_batch_u8");
}

extern "C" int trt_aug_layer_batch_u8(const void* aug_jobs, int njobs, int size, int need_stats, unsigned long long* hist,
                                      unsigned long long* luma, uint8_t* luts, cudaStream_t stream) {
  TRT_REQUIRE(aug_jobs && njobs > 0 && size > 0 && luts, "trt_aug_layer_batch_u8: bad argument");
  TRT_REQUIRE(!need_stats || (hist && luma), "trt_aug_layer_batch_u8: statistics buffers missing");
  const AugJob* jobs = reinterpret_cast<const AugJob*>(aug_jobs);
  if (need_stats) {
    aug_stats_batch_kernel<<<dim3(8, njobs), 256, 0, stream>>>(jobs, size, hist, luma);
    trt_count_launch(1);
    aug_lut_build_batch_kernel<<<njobs, 3, 0, stream>>>(jobs, hist, luts);
    trt_count_launch(1);
  }
  aug_apply_batch_kernel<<<dim3((size * size + 255) / 256, njobs), 256, 0, stream>>>(jobs, size, luts, luma);
  return trt_check_launch("trt_aug_layer_batch_u8");
}

extern "C" int trt_normalize_erase_batch(const void* fin_jobs, int n, int size, void* dst, int out_bf16, cudaStream_t stream) {
  TRT_REQUIRE(fin_jobs && dst && n > 0 && size > 0, "trt_normalize_erase_batch: bad argument");
  const FinJob* jobs = reinterpret_cast<const FinJob*>(fin_jobs);
  dim3 grid((size + 255) / 256, size, n);
  if (out_bf16) normalize_erase_batch_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(jobs, size, reinterpret_cast<__nv_bfloat16*>(dst));
  else normalize_erase_batch_kernel<float><<<grid, 256, 0, stream>>>(jobs, size, reinterpret_cast<float*>(dst));
  return trt_check_launch("trt_normalize_erase_batch");
}

extern "C" int trt_aug_job_bytes(void) { return (int)sizeof(AugJob); }
