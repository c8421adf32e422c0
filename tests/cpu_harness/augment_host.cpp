// CPU harness (test infrastructure): runs the per-pixel functions of csrc/augment_core.h — the code the CUDA kernels wrap —
// over whole images on the host, so the CPU suite can pin that arithmetic to Pillow without a GPU.
#include "augment_core.h"
#include <string.h>

extern "C" {
void h_enhance_rgb(const uint8_t* img, int H, int W, int mode, float factor, uint8_t* out) {
  unsigned long long sum = 0;
  for (long p = 0; p < (long)H * W; ++p) sum += pil_luma(img[p * 3], img[p * 3 + 1], img[p * 3 + 2]);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const long o = ((long)y * W + x) * 3;
      uint8_t deg[3] = {0, 0, 0};
      if (mode == 1) deg[0] = deg[1] = deg[2] = (uint8_t)pil_luma(img[o], img[o + 1], img[o + 2]);
      else if (mode == 2) deg[0] = deg[1] = deg[2] = (uint8_t)(int)((double)sum / (double)((long)H * W) + 0.5);
      else if (mode == 3) {
        const bool border = x == 0 || y == 0 || x == W - 1 || y == H - 1;
        for (int c = 0; c < 3; ++c) deg[c] = border ? img[o + c] : pil_smooth3x3(img + o + c, 3, (long)W * 3);
      }
      for (int c = 0; c < 3; ++c) out[o + c] = pil_blend(deg[c], img[o + c], factor);
    }
}
void h_affine(const uint8_t* img, int H, int W, int ch, const double* m, int bicubic, const uint8_t* fill, uint8_t* out) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      uint8_t px[3];
      const int ok = ch == 3 ? pil_affine_pixel<3>(img, H, W, m, bicubic, x, y, px) : pil_affine_pixel<1>(img, H, W, m, bicubic, x, y, px);
      for (int c = 0; c < ch; ++c) out[((long)y * W + x) * ch + c] = ok ? px[c] : fill[c];
    }
}
void h_hist_lut(const uint8_t* img, long n_px, int ch, int mode, uint8_t* lut) {
  for (int c = 0; c < ch; ++c) {
    long long h[256];
    memset(h, 0, sizeof(h));
    for (long p = 0; p < n_px; ++p) h[img[p * ch + c]]++;
    if (mode == 0) pil_autocontrast_lut(h, lut + c * 256); else pil_equalize_lut(h, lut + c * 256);
  }
}
}

// Pillow's resampling taps of one output index (the device builds the batched crop tables with this function)
extern "C" int h_resample_taps(int in_size, int out_size, int bicubic, int xx, int kmax, int* first, int* k) {
  return pil_resample_taps(in_size, out_size, bicubic, xx, kmax, first, k);
}
