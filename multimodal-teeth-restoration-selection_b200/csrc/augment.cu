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
