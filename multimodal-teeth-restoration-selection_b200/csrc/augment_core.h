// Per-pixel arithmetic of the Pillow operations behind timm's RandAugment (train transform of
// experiments/multimodal_v1/train_mm_joint_dualtask.py:75-84), written once for host and device: augment.cu wraps these in
// kernels, tests/cpu_harness/augment_host.cpp runs the very same functions on the CPU against Pillow.  Every expression
// keeps Pillow's evaluation order in its own precision (float for blend / 3x3 filter, double for the affine sampler); this
// file must be compiled WITHOUT fused multiply-add contraction (nvcc --fmad=false; x86-64 g++ has no FMA by default).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TRT_HD __host__ __device__ __forceinline__
#else
#define TRT_HD inline
#endif

// Pillow convert("L") of an RGB pixel: ITU-R 601-2 in 16-bit fixed point
TRT_HD int pil_luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Image.blend(degenerate, image, factor) on one uint8 sample (ImageEnhance.*.enhance)
TRT_HD uint8_t pil_blend(uint8_t deg, uint8_t img, float factor) {
  const float t = (float)deg + factor * ((float)img - (float)deg);
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : (uint8_t)t);
}

// ImageFilter.SMOOTH (3x3, (1,1,1,1,5,1,1,1,1)/13) at an INTERIOR pixel; border pixels are copied by the caller.
// p points at the centre sample, `xs` / `ys` are the sample strides of one pixel / one row.
TRT_HD uint8_t pil_smooth3x3(const uint8_t* p, long xs, long ys) {
  const float k1 = 1.0f / 13.0f, k5 = 5.0f / 13.0f;
  float ss = 0.5f;
  ss = ss + (((float)p[ys - xs] * k1 + (float)p[ys] * k1) + (float)p[ys + xs] * k1);
  ss = ss + (((float)p[-xs] * k1 + (float)p[0] * k5) + (float)p[xs] * k1);
  ss = ss + (((float)p[-ys - xs] * k1 + (float)p[-ys] * k1) + (float)p[-ys + xs] * k1);
  return ss <= 0.f ? 0 : (ss >= 255.f ? 255 : (uint8_t)ss);
}

TRT_HD int pil_floor(double v) { return v >= 0.0 ? (int)v : (int)floor(v); }
TRT_HD int pil_clipi(int v, int n) { return v < 0 ? 0 : (v < n ? v : n - 1); }
TRT_HD double pil_cubic(double v1, double v2, double v3, double v4, double d) {
  const double p1 = v2, p2 = -v1 + v3, p3 = 2 * (v1 - v2) + v3 - v4, p4 = -v1 + v2 - v3 + v4;
  return p1 + d * (p2 + d * (p3 + d * p4));
}

// Image.transform(size, AFFINE, m, resample, fillcolor): one output pixel (x, y) of a CH-channel interleaved uint8 image.
// bicubic != 0 selects Pillow's bicubic sampler, else bilinear.  Returns 0 when the source point falls outside the image
// (the caller writes the fill colour), 1 otherwise.
template <int CH>
TRT_HD int pil_affine_pixel(const uint8_t* img, int H, int W, const double* m, int bicubic, int x, int y, uint8_t* out) {
  double xin = m[0] * (x + 0.5) + m[1] * (y + 0.5) + m[2];
  double yin = m[3] * (x + 0.5) + m[4] * (y + 0.5) + m[5];
  if (xin < 0.0 || xin >= W || yin < 0.0 || yin >= H) return 0;
  xin -= 0.5; yin -= 0.5;
  int sx = pil_floor(xin), sy = pil_floor(yin);
  const double dx = xin - sx, dy = yin - sy;
  const long pitch = (long)W * CH;
  if (!bicubic) {
    const int x0 = pil_clipi(sx, W) * CH, x1 = pil_clipi(sx + 1, W) * CH;
    const uint8_t* r0 = img + (long)pil_clipi(sy, H) * pitch;
    const bool has1 = sy + 1 >= 0 && sy + 1 < H;
    const uint8_t* r1 = img + (long)(sy + 1) * pitch;
    for (int c = 0; c < CH; ++c) {
      double v1 = (double)r0[x0 + c] + ((double)r0[x1 + c] - (double)r0[x0 + c]) * dx;
      double v2 = v1;
      if (has1) v2 = (double)r1[x0 + c] + ((double)r1[x1 + c] - (double)r1[x0 + c]) * dx;
      v1 = v1 + (v2 - v1) * dy;
      out[c] = (uint8_t)v1;
    }
    return 1;
  }
  sx -= 1; sy -= 1;
  const int x0 = pil_clipi(sx, W) * CH, x1 = pil_clipi(sx + 1, W) * CH, x2 = pil_clipi(sx + 2, W) * CH, x3 = pil_clipi(sx + 3, W) * CH;
  for (int c = 0; c < CH; ++c) {
    const uint8_t* r = img + (long)pil_clipi(sy, H) * pitch;
    double v1 = pil_cubic(r[x0 + c], r[x1 + c], r[x2 + c], r[x3 + c], dx), v2 = v1, v3, v4;
    if (sy + 1 >= 0 && sy + 1 < H) { r = img + (long)(sy + 1) * pitch; v2 = pil_cubic(r[x0 + c], r[x1 + c], r[x2 + c], r[x3 + c], dx); }
    v3 = v2;
    if (sy + 2 >= 0 && sy + 2 < H) { r = img + (long)(sy + 2) * pitch; v3 = pil_cubic(r[x0 + c], r[x1 + c], r[x2 + c], r[x3 + c], dx); }
    v4 = v3;
    if (sy + 3 >= 0 && sy + 3 < H) { r = img + (long)(sy + 3) * pitch; v4 = pil_cubic(r[x0 + c], r[x1 + c], r[x2 + c], r[x3 + c], dx); }
    const double v = pil_cubic(v1, v2, v3, v4, dy);
    out[c] = v <= 0.0 ? 0 : (v >= 255.0 ? 255 : (uint8_t)v);
  }
  return 1;
}

// Lookup tables of the histogram-driven operations, one channel: ImageOps.autocontrast (cutoff 0) and ImageOps.equalize.
TRT_HD void pil_autocontrast_lut(const long long* h, uint8_t* lut) {
  int lo = 0, hi = 255;
  while (lo < 256 && !h[lo]) ++lo;
  while (hi >= 0 && !h[hi]) --hi;
  if (hi <= lo) { for (int i = 0; i < 256; ++i) lut[i] = (uint8_t)i; return; }
  const double scale = 255.0 / (hi - lo), offset = -lo * scale;
  for (int i = 0; i < 256; ++i) {
    const long long ix = (long long)(i * scale + offset);       // Python int(): truncation toward zero
    lut[i] = (uint8_t)(ix < 0 ? 0 : (ix > 255 ? 255 : ix));
  }
}
TRT_HD void pil_equalize_lut(const long long* h, uint8_t* lut) {
  long long total = 0, last = 0;
  int nonzero = 0;
  for (int i = 0; i < 256; ++i) { total += h[i]; if (h[i]) { ++nonzero; last = h[i]; } }
  const long long step = nonzero <= 1 ? 0 : (total - last) / 255;
  if (!step) { for (int i = 0; i < 256; ++i) lut[i] = (uint8_t)i; return; }
  long long n = step / 2;
  for (int i = 0; i < 256; ++i) {
    const long long v = n / step;
    lut[i] = (uint8_t)(v > 255 ? 255 : v);
    n += h[i];
  }
}

// Pillow's precompute_coeffs + normalize_coeffs_8bpc (Resample.c) for ONE output index of an axis resize in_size -> out_size:
// the taps [*first, *first + n) and their 22-bit fixed-point weights k[0..n), k[n..kmax) = 0.  Same double arithmetic and
// evaluation order as the C source (and as teethrt.preproc.pil_coeffs, which is pinned to Pillow).  Returns n.
TRT_HD double pil_filter(double x, int bicubic) {
  if (x < 0.0) x = -x;
  if (bicubic) {
    const double a = -0.5;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
  }
  return x < 1.0 ? 1.0 - x : 0.0;
}
TRT_HD int pil_resample_taps(int in_size, int out_size, int bicubic, int xx, int kmax, int* first, int* k) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = (bicubic ? 2.0 : 1.0) * filterscale;
  const double ss = 1.0 / filterscale;
  const double center = (xx + 0.5) * scale;
  int xmin = (int)(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kmax) xmax = kmax;                     // cannot happen when kmax = ceil(support) * 2 + 1
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) ww += pil_filter((x + xmin - center + 0.5) * ss, bicubic);
  const double one = (double)(1 << 22);
  for (int x = 0; x < kmax; ++x) {
    if (x < xmax) {
      double w = pil_filter((x + xmin - center + 0.5) * ss, bicubic);
      if (ww != 0.0) w = w / ww;
      k[x] = w < 0 ? (int)(-0.5 + w * one) : (int)(0.5 + w * one);
    } else {
      k[x] = 0;
    }
  }
  *first = xmin;
  return xmax;
}
