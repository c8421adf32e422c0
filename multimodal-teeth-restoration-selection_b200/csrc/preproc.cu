// On-device input stage: CLAHE in Lab space (byte-exact vs OpenCV), centre-crop + bilinear resize (byte-exact),
// ToTensor + Normalize + flip.  Integer/byte work, HBM-bound: shared-memory histograms, LUTs staged in shared memory,
// coalesced 4-byte global accesses.  Arithmetic follows SURVEY.md App. A (OpenCV 4.x color_lab.cpp / clahe.cpp /
// resize.cpp fixed-point paths).
//
// Replaces: src/preprocessing/normalise.py:10-16 (apply_clahe), src/preprocessing/pipeline.py:23-29
// (centre_crop_resize), and ToTensor/Normalize/flip of experiments/multimodal_v1/train_mm_joint_dualtask.py:83-84,328-333.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int GRID = 8;        // CLAHE_TILEGR = (8, 8)  (src/config.py:16)
constexpr int NT = GRID * GRID;

struct LabTables {             // views into the packed device table buffer (layout in teethrt.h)
  const uint16_t* gamma;       // [256]
  const uint16_t* cbrt;        // [3072]
  const int32_t* yf;           // [512]  (y, ify) interleaved
  const int32_t* abxz;         // [36864]
  const uint8_t* invgamma;     // [4096]
};
__host__ __device__ inline LabTables table_views(const void* packed) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(packed);
  LabTables t;
  t.gamma = reinterpret_cast<const uint16_t*>(b + TRT_TAB_GAMMA_OFF);
  t.cbrt = reinterpret_cast<const uint16_t*>(b + TRT_TAB_CBRT_OFF);
  t.yf = reinterpret_cast<const int32_t*>(b + TRT_TAB_YF_OFF);
  t.abxz = reinterpret_cast<const int32_t*>(b + TRT_TAB_ABXZ_OFF);
  t.invgamma = b + TRT_TAB_INVGAMMA_OFF;
  return t;
}

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ int clamp_u8(int v) { return min(max(v, 0), 255); }

// BGR -> (L, a, b), all uint8 (App. A.1).  s_gamma/s_cbrt are shared-memory copies.
__device__ __forceinline__ void bgr2lab(int b8, int g8, int r8, const uint16_t* s_gamma, const uint16_t* s_cbrt, int& L,
                                        int& a, int& bb) {
  const int R = s_gamma[r8], G = s_gamma[g8], B = s_gamma[b8];
  const int fX = s_cbrt[descale(R * 1777 + G * 1541 + B * 778, 12)];
  const int fY = s_cbrt[descale(R * 871 + G * 2929 + B * 296, 12)];
  const int fZ = s_cbrt[descale(R * 73 + G * 448 + B * 3575, 12)];
  L = clamp_u8(descale(296 * fY - 1336934, 15));
  a = clamp_u8(descale(500 * (fX - fY) + 128 * 32768, 15));
  bb = clamp_u8(descale(200 * (fY - fZ) + 128 * 32768, 15));
}

// (L, a, b) -> BGR (App. A.3)
__device__ __forceinline__ void lab2bgr(int L, int a, int b, const int32_t* s_yf, const int32_t* __restrict__ g_abxz,
                                        const uint8_t* s_inv, int& ob, int& og, int& orr) {
  constexpr int BASE = 16384, minAB = -8145;
  const int y = s_yf[2 * L], ify = s_yf[2 * L + 1];
  const int adiv = ((5 * a * 53687 + 128) >> 13) - 128 * BASE / 500;
  const int bdiv = ((b * 41943 + 16) >> 9) - 128 * BASE / 200 + 1;
  const int X = __ldg(g_abxz + (ify + adiv - minAB));
  const int Z = __ldg(g_abxz + (ify - bdiv - minAB));
  const int ro = min(max(descale(12615 * X - 6296 * y - 2223 * Z, 14), 0), 4095);
  const int go = min(max(descale(-3773 * X + 7684 * y + 185 * Z, 14), 0), 4095);
  const int bo = min(max(descale(217 * X - 836 * y + 4715 * Z, 14), 0), 4095);
  ob = s_inv[bo]; og = s_inv[go]; orr = s_inv[ro];
}

__device__ __forceinline__ int reflect101(int p, int n) { return p < n ? p : 2 * (n - 1) - p; }

// ---------------------------------------------------------------- pass A: per-tile histograms of L over the padded image
// grid = (NT * bands, N); each block histograms a horizontal band of one tile with per-warp shared sub-histograms.
__global__ void __launch_bounds__(256) clahe_hist_kernel(const uint8_t* __restrict__ src, uint32_t* __restrict__ hist,
                                                         const void* __restrict__ tables, int H, int W, int th, int tw,
                                                         int bands) {
  __shared__ uint16_t s_gamma[256];
  __shared__ uint16_t s_cbrt[3072];
  __shared__ uint32_t s_hist[8][256];
  const LabTables T = table_views(tables);
  for (int i = threadIdx.x; i < 256; i += 256) s_gamma[i] = T.gamma[i];
  for (int i = threadIdx.x; i < 3072; i += 256) s_cbrt[i] = T.cbrt[i];
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int tile = blockIdx.x / bands, band = blockIdx.x % bands;
  const int ty = tile / GRID, tx = tile % GRID;
  const int rows_per_band = (th + bands - 1) / bands;
  const int y_begin = band * rows_per_band, y_end = min(th, y_begin + rows_per_band);
  const uint8_t* img = src + (size_t)blockIdx.y * H * W * 3;
  const int warp = threadIdx.x >> 5;
  const int npx = (y_end - y_begin) * tw;
  for (int i = threadIdx.x; i < npx; i += 256) {
    const int yy = y_begin + i / tw, xx = i % tw;
    const int y = reflect101(ty * th + yy, H), x = reflect101(tx * tw + xx, W);
    const uint8_t* px = img + ((size_t)y * W + x) * 3;
    int L, a, b;
    bgr2lab(px[0], px[1], px[2], s_gamma, s_cbrt, L, a, b);
    atomicAdd(&s_hist[warp][L], 1u);
  }
  __syncthreads();
  uint32_t* out = hist + ((size_t)blockIdx.y * NT + tile) * 256;
  uint32_t v = 0;
#pragma unroll
  for (int wv = 0; wv < 8; ++wv) v += s_hist[wv][threadIdx.x];
  if (v) atomicAdd(out + threadIdx.x, v);
}

// ---------------------------------------------------------------- pass A2: clip, redistribute, cumulative LUT (App. A.2)
// grid = (NT, N), 256 threads = one per bin.
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint32_t* __restrict__ hist, uint8_t* __restrict__ luts,
                                                        int clip_limit, float lut_scale) {
  __shared__ int s_scan[256];
  __shared__ int s_red[8];
  const int i = threadIdx.x;
  const size_t base = ((size_t)blockIdx.y * NT + blockIdx.x) * 256;
  int h = (int)hist[base + i];
  int over = max(h - clip_limit, 0);
  int ws = over;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
  if ((i & 31) == 0) s_red[i >> 5] = ws;
  __syncthreads();
  int clipped = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) clipped += s_red[k];
  h = min(h, clip_limit) + clipped / 256;
  const int residual = clipped % 256;
  if (residual) {
    const int step = max(256 / residual, 1);
    if (i % step == 0 && i / step < residual) h += 1;
  }
  // inclusive scan over 256 bins
  s_scan[i] = h;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int add = i >= o ? s_scan[i - o] : 0;
    __syncthreads();
    s_scan[i] += add;
    __syncthreads();
  }
  const int v = __float2int_rn(__fmul_rn((float)s_scan[i], lut_scale));
  luts[base + i] = (uint8_t)clamp_u8(v);
}

// ---------------------------------------------------------------- pass B: Lab -> CLAHE(L) blend -> BGR
__device__ __forceinline__ void clahe_pixel(int b8, int g8, int r8, int ty1, int ty2, float ya, float ya1, int tx1, int tx2,
                                            float xa, float xa1, const uint16_t* s_gamma, const uint16_t* s_cbrt,
                                            const int32_t* s_yf, const int32_t* g_abxz, const uint8_t* s_inv,
                                            const uint8_t* s_luts, int& ob, int& og, int& orr) {
  int L, a, b;
  bgr2lab(b8, g8, r8, s_gamma, s_cbrt, L, a, b);
  const float l11 = (float)s_luts[(ty1 * GRID + tx1) * 256 + L], l12 = (float)s_luts[(ty1 * GRID + tx2) * 256 + L];
  const float l21 = (float)s_luts[(ty2 * GRID + tx1) * 256 + L], l22 = (float)s_luts[(ty2 * GRID + tx2) * 256 + L];
  // fp32, written order, no FMA contraction (OpenCV CLAHE_Interpolation_Body)
  const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
  const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
  const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
  const int L2 = clamp_u8(__float2int_rn(res));
  lab2bgr(L2, a, b, s_yf, g_abxz, s_inv, ob, og, orr);
}

__device__ __forceinline__ void tile_coord(int p, float inv_t, int& t1, int& t2, float& a, float& a1) {
  const float tf = __fsub_rn(__fmul_rn((float)p, inv_t), 0.5f);
  const int f = (int)floorf(tf);
  a = __fsub_rn(tf, (float)f);
  a1 = __fsub_rn(1.0f, a);
  t1 = max(f, 0);
  t2 = min(f + 1, GRID - 1);
}

// grid = (ceil(W/ (256*4)) , rows/ROWS_PER_BLOCK, N); each thread handles 4 consecutive pixels of ROWS rows.
constexpr int APPLY_ROWS = 16;
__global__ void __launch_bounds__(256) clahe_apply_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          const uint8_t* __restrict__ luts, const void* __restrict__ tables,
                                                          int H, int W, float inv_th, float inv_tw, int vec_ok) {
  __shared__ uint16_t s_gamma[256];
  __shared__ uint16_t s_cbrt[3072];
  __shared__ int32_t s_yf[512];
  __shared__ uint8_t s_inv[4096];
  __shared__ __align__(16) uint8_t s_luts[NT * 256];
  const LabTables T = table_views(tables);
  for (int i = threadIdx.x; i < 256; i += 256) s_gamma[i] = T.gamma[i];
  for (int i = threadIdx.x; i < 3072; i += 256) s_cbrt[i] = T.cbrt[i];
  for (int i = threadIdx.x; i < 512; i += 256) s_yf[i] = T.yf[i];
  for (int i = threadIdx.x; i < 1024; i += 256)
    reinterpret_cast<uint32_t*>(s_inv)[i] = reinterpret_cast<const uint32_t*>(T.invgamma)[i];
  const uint8_t* lut_img = luts + (size_t)blockIdx.z * NT * 256;
  for (int i = threadIdx.x; i < NT * 256 / 16; i += 256)
    reinterpret_cast<uint4*>(s_luts)[i] = reinterpret_cast<const uint4*>(lut_img)[i];
  __syncthreads();
  const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (x0 >= W) return;
  const size_t img_off = (size_t)blockIdx.z * H * W * 3;
  int tx1[4], tx2[4];
  float xa[4], xa1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tile_coord(x0 + j, inv_tw, tx1[j], tx2[j], xa[j], xa1[j]);
  const int y_begin = blockIdx.y * APPLY_ROWS;
  for (int y = y_begin; y < min(H, y_begin + APPLY_ROWS); ++y) {
    int ty1, ty2;
    float ya, ya1;
    tile_coord(y, inv_th, ty1, ty2, ya, ya1);
    const size_t off = img_off + ((size_t)y * W + x0) * 3;
    if (vec_ok && x0 + 4 <= W) {
      const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + off);
      const uint32_t w0 = __ldg(s4), w1 = __ldg(s4 + 1), w2 = __ldg(s4 + 2);
      uint8_t in[12], out[12];
      *reinterpret_cast<uint32_t*>(in) = w0;
      *reinterpret_cast<uint32_t*>(in + 4) = w1;
      *reinterpret_cast<uint32_t*>(in + 8) = w2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ob, og, orr;
        clahe_pixel(in[3 * j], in[3 * j + 1], in[3 * j + 2], ty1, ty2, ya, ya1, tx1[j], tx2[j], xa[j], xa1[j], s_gamma,
                    s_cbrt, s_yf, T.abxz, s_inv, s_luts, ob, og, orr);
        out[3 * j] = (uint8_t)ob; out[3 * j + 1] = (uint8_t)og; out[3 * j + 2] = (uint8_t)orr;
      }
      uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + off);
      d4[0] = *reinterpret_cast<uint32_t*>(out);
      d4[1] = *reinterpret_cast<uint32_t*>(out + 4);
      d4[2] = *reinterpret_cast<uint32_t*>(out + 8);
    } else {
      for (int j = 0; j < 4 && x0 + j < W; ++j) {
        const uint8_t* px = src + off + 3 * j;
        int ob, og, orr;
        clahe_pixel(px[0], px[1], px[2], ty1, ty2, ya, ya1, tx1[j], tx2[j], xa[j], xa1[j], s_gamma, s_cbrt, s_yf, T.abxz,
                    s_inv, s_luts, ob, og, orr);
        uint8_t* q = dst + off + 3 * j;
        q[0] = (uint8_t)ob; q[1] = (uint8_t)og; q[2] = (uint8_t)orr;
      }
    }
  }
}

// ---------------------------------------------------------------- fast path (no padding, power-of-two tiles)
// Round-2 rewrite of both passes for the shapes the pipeline really sees (512^2 / 1024^2 radiographs).  ncu on the kernels
// above (profiles/r02_ncu_full_small_kernels.csv, 64 x 1024^2): histogram pass 72 thread-instructions per pixel at 72 %
// issue, apply pass 129 per pixel at 60 % issue with 46 M shared-memory bank-conflict cycles and two random global gathers
// per pixel - the path is INTEGER-ISSUE-bound (201 instructions per pixel = 0.45 ms of issue slots at 100 %), nowhere near
// HBM.  So the rewrite removes instructions, not bytes:
//   (1) BGR -> Lab is computed ONCE: pass A writes the Lab image into `dst` while it histograms L, pass B converts `dst` in
//       place (6 more bytes per pixel of traffic, ~50 fewer instructions per pixel);
//   (2) abToXZ (147 KB, gathered from global memory twice per pixel) is pure integer arithmetic and is computed (abxz_at);
//   (3) a block of pass B works inside ONE interpolation cell (the square between four tile centres), where the four LUTs
//       are fixed, and packs them into one 256-entry uint32 table: one lookup per pixel instead of four, 1 KB instead of 16;
//   (4) a thread owns 16 consecutive pixels = three 16-byte loads / stores, and its 16 horizontal interpolation weights live
//       in registers across all its rows;
//   (5) the sRGB gamma table (3 lookups per pixel) is replicated per lane (32 KB) so those lookups never conflict;
//   (6) constant tables are staged once per persistent block.
// Bit-exactness is unchanged: per pixel it is the same sequence of integer / non-contracted fp32 operations.
constexpr int MIN_AB = -8145;
// abToXZ_b[i - MIN_AB] of OpenCV's Lab2RGBinteger, computed instead of looked up (checked against the
// digest-verified table entry by entry: tests/test_oracle_preproc.py)
__device__ __forceinline__ int abxz_at(int i) {
  if (i <= 3390) return (i * 108) / 841 - 290;                  // C division truncates toward zero, like the table builder
  return (int)((((unsigned)(i * i)) >> 14) * (unsigned)i >> 14);
}

// lane-private gamma table: entry v of lane l at word v * 32 + l, so a warp's 32 lookups hit 32 different banks whatever v is
__device__ __forceinline__ void stage_gamma32(uint32_t* g32, const uint16_t* gamma) {
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) g32[i] = __ldg(gamma + (i >> 5));
}

// pass A: grid = (NT * bands, N); block = one band of one tile, thread = 16 consecutive pixels of a row.
// Writes the Lab image (L, a, b interleaved like the input) to `lab` and the per-tile histogram of L.
__global__ void __launch_bounds__(256) clahe_lab_hist_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ lab,
                                                             uint32_t* __restrict__ hist, const void* __restrict__ tables, int H,
                                                             int W, int th, int tw, int bands) {
  __shared__ uint32_t s_gamma32[256 * 32];
  __shared__ uint16_t s_cbrt[3072];
  __shared__ uint32_t s_hist[8][256];
  const LabTables T = table_views(tables);
  stage_gamma32(s_gamma32, T.gamma);
  for (int i = threadIdx.x; i < 3072 / 2; i += 256) reinterpret_cast<uint32_t*>(s_cbrt)[i] = __ldg(reinterpret_cast<const uint32_t*>(T.cbrt) + i);
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const int tile = blockIdx.x / bands, band = blockIdx.x % bands;
  const int ty = tile / GRID, tx = tile % GRID;
  const int rows_per_band = (th + bands - 1) / bands;
  const int y_begin = band * rows_per_band, y_end = min(th, y_begin + rows_per_band);
  const size_t img_off = (size_t)blockIdx.y * H * W * 3;
  const int warp = threadIdx.x >> 5;
  const uint32_t* gl = s_gamma32 + (threadIdx.x & 31);
  const int groups = tw >> 4;                                   // 16-pixel groups per tile row
  const int items = (y_end - y_begin) * groups;
  for (int i = threadIdx.x; i < items; i += 256) {
    const int yy = y_begin + i / groups, g = i % groups;
    const size_t off = img_off + ((size_t)(ty * th + yy) * W + tx * tw + g * 16) * 3;
    const uint4* p = reinterpret_cast<const uint4*>(src + off);
    uint32_t w[12], o[12];
    *reinterpret_cast<uint4*>(w) = __ldg(p);
    *reinterpret_cast<uint4*>(w + 4) = __ldg(p + 1);
    *reinterpret_cast<uint4*>(w + 8) = __ldg(p + 2);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(w);
    uint8_t* ob = reinterpret_cast<uint8_t*>(o);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int B = (int)gl[b[3 * j] * 32], G = (int)gl[b[3 * j + 1] * 32], R = (int)gl[b[3 * j + 2] * 32];
      const int fX = s_cbrt[descale(R * 1777 + G * 1541 + B * 778, 12)];
      const int fY = s_cbrt[descale(R * 871 + G * 2929 + B * 296, 12)];
      const int fZ = s_cbrt[descale(R * 73 + G * 448 + B * 3575, 12)];
      const int L = clamp_u8(descale(296 * fY - 1336934, 15));
      ob[3 * j] = (uint8_t)L;
      ob[3 * j + 1] = (uint8_t)clamp_u8(descale(500 * (fX - fY) + 128 * 32768, 15));
      ob[3 * j + 2] = (uint8_t)clamp_u8(descale(200 * (fY - fZ) + 128 * 32768, 15));
      atomicAdd(&s_hist[warp][L], 1u);
    }
    uint4* q = reinterpret_cast<uint4*>(lab + off);
    q[0] = *reinterpret_cast<uint4*>(o);
    q[1] = *reinterpret_cast<uint4*>(o + 4);
    q[2] = *reinterpret_cast<uint4*>(o + 8);
  }
  __syncthreads();
  uint32_t* out = hist + ((size_t)blockIdx.y * NT + tile) * 256;
  uint32_t v = 0;
#pragma unroll
  for (int wv = 0; wv < 8; ++wv) v += s_hist[wv][threadIdx.x];
  if (v) atomicAdd(out + threadIdx.x, v);
}

// pass B: persistent blocks over (image, cell, 32-row band) items; converts the Lab image IN PLACE to the CLAHE'd BGR image
constexpr int CELLS = GRID + 1;
constexpr int BAND = 32;      // rows per item: 8 threads x 16 pixels cover a 128-pixel cell row, 32 rows per 256 threads
__global__ void __launch_bounds__(256) clahe_apply_lab_kernel(uint8_t* __restrict__ img, const uint8_t* __restrict__ luts,
                                                              const void* __restrict__ tables, int H, int W, int th, int tw,
                                                              float inv_th, float inv_tw, int bands_per_cell, int n_items) {
  __shared__ int2 s_yf[256];
  __shared__ __align__(16) uint8_t s_inv[4096];
  __shared__ uint32_t s_comb[2][256];
  const LabTables T = table_views(tables);
  s_yf[threadIdx.x] = make_int2(T.yf[2 * threadIdx.x], T.yf[2 * threadIdx.x + 1]);
  for (int i = threadIdx.x; i < 1024; i += 256) reinterpret_cast<uint32_t*>(s_inv)[i] = __ldg(reinterpret_cast<const uint32_t*>(T.invgamma) + i);
  const int per_img = CELLS * CELLS * bands_per_cell;
  int buf = 0;
  for (int it = blockIdx.x; it < n_items; it += gridDim.x, buf ^= 1) {
    const int n = it / per_img, r0 = it - n * per_img;
    const int cell = r0 / bands_per_cell, band = r0 - cell * bands_per_cell;
    const int cy = cell / CELLS, cx = cell - cy * CELLS;
    // cell geometry: cells 1..GRID-1 span [half + (c-1)*t, half + c*t); cell 0 = [0, half); cell GRID = [size - half, size)
    const int hx = tw >> 1, hy = th >> 1;
    const int x_lo = cx == 0 ? 0 : hx + (cx - 1) * tw, x_hi = cx == 0 ? hx : min(W, hx + cx * tw);
    const int y_lo = cy == 0 ? 0 : hy + (cy - 1) * th, y_hi = cy == 0 ? hy : min(H, hy + cy * th);
    const int tx1 = max(cx - 1, 0), tx2 = min(cx, GRID - 1), ty1 = max(cy - 1, 0), ty2 = min(cy, GRID - 1);
    // the four LUTs of the cell packed per grey level (double-buffered: the previous item's readers may still be in flight)
    {
      const uint8_t* L = luts + (size_t)n * NT * 256 + threadIdx.x;
      s_comb[buf][threadIdx.x] = (uint32_t)__ldg(L + (ty1 * GRID + tx1) * 256) | ((uint32_t)__ldg(L + (ty1 * GRID + tx2) * 256) << 8) |
                                 ((uint32_t)__ldg(L + (ty2 * GRID + tx1) * 256) << 16) | ((uint32_t)__ldg(L + (ty2 * GRID + tx2) * 256) << 24);
    }
    __syncthreads();
    const uint32_t* comb = s_comb[buf];
    const int groups = (x_hi - x_lo) >> 4;                      // 16-pixel groups per cell row: a power of two <= 256
    const int rows = min(BAND, y_hi - (y_lo + band * BAND));
    const size_t img_off = (size_t)n * H * W * 3;
    // 256 % groups == 0: a thread keeps its column group for all its rows, so its horizontal weights are computed once
    const int g = threadIdx.x & (groups - 1), x0 = x_lo + g * 16;
    float xa[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int u1, u2;
      float a1;
      tile_coord(x0 + j, inv_tw, u1, u2, xa[j], a1);
    }
    for (int yy = threadIdx.x / groups; yy < rows; yy += 256 / groups) {
      const int y = y_lo + band * BAND + yy;
      int t1, t2;
      float ya, ya1;
      tile_coord(y, inv_th, t1, t2, ya, ya1);
      uint4* p = reinterpret_cast<uint4*>(img + img_off + ((size_t)y * W + x0) * 3);
      uint32_t w[12], o[12];
      *reinterpret_cast<uint4*>(w) = p[0];
      *reinterpret_cast<uint4*>(w + 4) = p[1];
      *reinterpret_cast<uint4*>(w + 8) = p[2];
      const uint8_t* b = reinterpret_cast<const uint8_t*>(w);
      uint8_t* ob = reinterpret_cast<uint8_t*>(o);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int L = b[3 * j], a = b[3 * j + 1], bb = b[3 * j + 2];
        // CLAHE interpolation (fp32, written order, no FMA contraction - OpenCV CLAHE_Interpolation_Body)
        const float xa1 = __fsub_rn(1.0f, xa[j]);
        const uint32_t c4 = comb[L];
        const float l11 = (float)(c4 & 255u), l12 = (float)((c4 >> 8) & 255u), l21 = (float)((c4 >> 16) & 255u), l22 = (float)(c4 >> 24);
        const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa[j]));
        const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa[j]));
        const int L2 = clamp_u8(__float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya))));
        // Lab -> BGR
        constexpr int BASE = 16384;
        const int2 yv = s_yf[L2];
        const int adiv = ((5 * a * 53687 + 128) >> 13) - 128 * BASE / 500;
        const int bdiv = ((bb * 41943 + 16) >> 9) - 128 * BASE / 200 + 1;
        const int X = abxz_at(yv.y + adiv), Z = abxz_at(yv.y - bdiv);
        const int ro = min(max(descale(12615 * X - 6296 * yv.x - 2223 * Z, 14), 0), 4095);
        const int go = min(max(descale(-3773 * X + 7684 * yv.x + 185 * Z, 14), 0), 4095);
        const int bo = min(max(descale(217 * X - 836 * yv.x + 4715 * Z, 14), 0), 4095);
        ob[3 * j] = s_inv[bo]; ob[3 * j + 1] = s_inv[go]; ob[3 * j + 2] = s_inv[ro];
      }
      p[0] = *reinterpret_cast<uint4*>(o);
      p[1] = *reinterpret_cast<uint4*>(o + 4);
      p[2] = *reinterpret_cast<uint4*>(o + 8);
    }
  }
}

// ---------------------------------------------------------------- centre crop + cv2.resize(INTER_LINEAR) (App. A.4)
__device__ __forceinline__ void axis_coef(int d, double scale, int ssize, bool zero_edges, int& s, int& w0, int& w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int si = (int)floorf(f);
  f = __fsub_rn(f, (float)si);
  if (zero_edges) {
    if (si < 0) { f = 0.f; si = 0; }
    if (si >= ssize - 1) { f = 0.f; si = ssize - 1; }
  }
  s = si;
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.f));
}

__global__ void __launch_bounds__(256) resize_linear_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H,
                                                            int W, int y_off, int x_off, int sh, int sw, int dh, int dw) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= dw) return;
  const double scale_x = (double)sw / (double)dw, scale_y = (double)sh / (double)dh;
  int sx, wx0, wx1, sy, wy0, wy1;
  axis_coef(dx, scale_x, sw, true, sx, wx0, wx1);
  axis_coef(dy, scale_y, sh, false, sy, wy0, wy1);
  const int sx1 = min(sx + 1, sw - 1);
  const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
  const uint8_t* img = src + (size_t)blockIdx.z * H * W * 3;
  const uint8_t* r0 = img + ((size_t)(y0 + y_off) * W + x_off) * 3;
  const uint8_t* r1 = img + ((size_t)(y1 + y_off) * W + x_off) * 3;
  uint8_t* o = dst + (((size_t)blockIdx.z * dh + dy) * dw + dx) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int a0 = r0[sx * 3 + c] * wx0 + r0[sx1 * 3 + c] * wx1;
    const int a1 = r1[sx * 3 + c] * wx0 + r1[sx1 * 3 + c] * wx1;
    const int v = (((wy0 * (a0 >> 4)) >> 16) + ((wy1 * (a1 >> 4)) >> 16) + 2) >> 2;
    o[c] = (uint8_t)clamp_u8(v);
  }
}

// ---------------------------------------------------------------- ToTensor + Normalize + flip: u8 HWC BGR -> CHW RGB
template <typename OutT>
__global__ void __launch_bounds__(256) normalize_flip_kernel(const uint8_t* __restrict__ src, OutT* __restrict__ dst, int H,
                                                             int W, int flip) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int sy = flip == 2 ? H - 1 - y : y, sx = flip == 1 ? W - 1 - x : x;
  const uint8_t* px = src + (((size_t)blockIdx.z * H + sy) * W + sx) * 3;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  OutT* o = dst + (size_t)blockIdx.z * 3 * H * W + (size_t)y * W + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[2 - c], 255.f), mean[c]), stdv[c]);
    if constexpr (sizeof(OutT) == 2) o[(size_t)c * H * W] = __float2bfloat16_rn(v);
    else o[(size_t)c * H * W] = v;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Pillow-exact antialiased resampling of 8-bit images (the eval transform the reference builds with
// timm.data.create_transform / torchvision Resize on PIL images, ui/gradio_app/infer_mm.py:12-17, infer_mil.py:116-119).
// Pillow resamples in two separable passes (horizontal, then vertical) over uint8 with 22-bit fixed-point coefficients and
// round-half-up + clip after EACH pass; the coefficient tables (bounds = {first tap, tap count}, coeffs = ksize ints per
// output index) are built on the host in double precision exactly as Pillow's precompute_coeffs does (preproc.py).
// One thread per output pixel (all channels); the tables are tiny and stay in L1/L2.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8_fixed(int v) {
  v >>= RS_PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

template <int CH, bool VERTICAL>
__global__ void __launch_bounds__(256) resample_u8_kernel(const uint8_t* __restrict__ in, size_t in_pitch, uint8_t* __restrict__ out,
                                                          int out_rows, int out_cols, const int* __restrict__ bounds,
                                                          const int* __restrict__ coeffs, int ksize, int swap_channels) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= out_cols) return;
  const int o = VERTICAL ? y : x;
  const int first = __ldg(bounds + 2 * o), taps = __ldg(bounds + 2 * o + 1);
  const int* k = coeffs + (size_t)o * ksize;
  int acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 1 << (RS_PRECISION_BITS - 1);
  const uint8_t* p = VERTICAL ? in + (size_t)first * in_pitch + (size_t)x * CH : in + (size_t)y * in_pitch + (size_t)first * CH;
  const size_t step = VERTICAL ? in_pitch : (size_t)CH;
  for (int t = 0; t < taps; ++t, p += step) {
    const int w = __ldg(k + t);
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] += (int)__ldg(p + c) * w;
  }
  uint8_t* q = out + ((size_t)y * out_cols + x) * CH;
#pragma unroll
  for (int c = 0; c < CH; ++c) q[swap_channels ? CH - 1 - c : c] = clip8_fixed(acc[c]);
}

}  // namespace

extern "C" int trt_resample_u8(const uint8_t* in, size_t in_pitch_bytes, int channels, uint8_t* out, int out_rows, int out_cols,
                               const int* bounds, const int* coeffs, int ksize, int vertical, int swap_channels,
                               cudaStream_t stream) {
  TRT_REQUIRE(in && out && bounds && coeffs, "trt_resample_u8: null pointer");
  TRT_REQUIRE(out_rows > 0 && out_cols > 0 && ksize > 0, "trt_resample_u8: bad shape %d x %d, ksize %d", out_rows, out_cols, ksize);
  TRT_REQUIRE(channels == 1 || channels == 3, "trt_resample_u8: %d channels not built (1 or 3)", channels);
  dim3 grid((out_cols + 255) / 256, out_rows);
#define TRT_RS(CH, V) resample_u8_kernel<CH, V><<<grid, 256, 0, stream>>>(in, in_pitch_bytes, out, out_rows, out_cols, bounds, \
                                                                         coeffs, ksize, swap_channels)
  if (channels == 3) { if (vertical) TRT_RS(3, true); else TRT_RS(3, false); }
  else               { if (vertical) TRT_RS(1, true); else TRT_RS(1, false); }
#undef TRT_RS
  return trt_check_launch("trt_resample_u8");
}


extern "C" size_t trt_clahe_workspace_bytes(int n) { return (size_t)n * NT * 256 * (sizeof(uint32_t) + 1); }

extern "C" int trt_clahe_bgr_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, float clip, const void* tables,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  TRT_REQUIRE(src && dst && tables && workspace, "trt_clahe_bgr_u8: null pointer");
  TRT_REQUIRE(n > 0 && h >= GRID && w >= GRID, "trt_clahe_bgr_u8: bad shape n=%d h=%d w=%d", n, h, w);
  TRT_REQUIRE(workspace_bytes >= trt_clahe_workspace_bytes(n), "trt_clahe_bgr_u8: workspace too small");
  // OpenCV pads to a multiple of the grid with a FULL extra tile-size step when only one axis is off (App. A.2)
  int hp = h, wp = w;
  if (h % GRID || w % GRID) { hp = h + (GRID - h % GRID); wp = w + (GRID - w % GRID); }
  const int th = hp / GRID, tw = wp / GRID;
  TRT_REQUIRE(hp - h < h && wp - w < w, "trt_clahe_bgr_u8: image too small for reflect-101 padding");
  const int area = th * tw;
  int clip_limit = (int)((double)clip * (double)area / 256.0);
  if (clip_limit < 1) clip_limit = 1;
  const float lut_scale = 255.0f / (float)area;
  const float inv_th = 1.0f / (float)th, inv_tw = 1.0f / (float)tw;
  uint32_t* hist = reinterpret_cast<uint32_t*>(workspace);
  uint8_t* luts = reinterpret_cast<uint8_t*>(workspace) + (size_t)n * NT * 256 * sizeof(uint32_t);
  TRT_CUDA(cudaMemsetAsync(hist, 0, (size_t)n * NT * 256 * sizeof(uint32_t), stream));
  int bands = 1;
  while (bands < 8 && n * NT * bands < 4 * trt_num_sms() && th / (bands * 2) >= 8) bands *= 2;
  // fast path: no reflect padding, 16-pixel groups never straddle a tile or cell boundary, 16-byte aligned rows
  const char* slow = getenv("TEETHRT_CLAHE_SLOW");
  // tile sizes must be powers of two: then p * (1 / tile) is exact in fp32 and OpenCV's floor(p / tile - 0.5) changes exactly
  // at the integer cell boundaries the fast kernel uses (with e.g. 96-pixel tiles the rounded product can put a boundary
  // pixel in the neighbouring cell)
  const bool pow2 = (tw & (tw - 1)) == 0 && (th & (th - 1)) == 0;
  const bool fast = hp == h && wp == w && pow2 && tw >= 32 && tw <= 4096 && th >= 2 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0 &&
                    !(slow && *slow == '1');
  if (fast) {
    clahe_lab_hist_kernel<<<dim3(NT * bands, n), 256, 0, stream>>>(src, dst, hist, tables, h, w, th, tw, bands);
    clahe_lut_kernel<<<dim3(NT, n), 256, 0, stream>>>(hist, luts, clip_limit, lut_scale);
    const int bands_per_cell = (th + BAND - 1) / BAND;            // edge cells are half as tall: their upper bands are empty
    const long long items = (long long)n * CELLS * CELLS * bands_per_cell;
    TRT_REQUIRE(items < (1ll << 31), "trt_clahe_bgr_u8: too many items");
    const int blocks = (int)(items < 6ll * trt_num_sms() ? items : 6ll * trt_num_sms());
    clahe_apply_lab_kernel<<<blocks, 256, 0, stream>>>(dst, luts, tables, h, w, th, tw, inv_th, inv_tw, bands_per_cell, (int)items);
  } else {
    clahe_hist_kernel<<<dim3(NT * bands, n), 256, 0, stream>>>(src, hist, tables, h, w, th, tw, bands);
    clahe_lut_kernel<<<dim3(NT, n), 256, 0, stream>>>(hist, luts, clip_limit, lut_scale);
    const int vec_ok = (w % 4 == 0) && (((uintptr_t)src & 3) == 0) && (((uintptr_t)dst & 3) == 0);
    dim3 grid((w + 1023) / 1024, (h + APPLY_ROWS - 1) / APPLY_ROWS, n);
    clahe_apply_kernel<<<grid, 256, 0, stream>>>(src, dst, luts, tables, h, w, inv_th, inv_tw, vec_ok);
  }
  trt_count_launch(2);
  return trt_check_launch("trt_clahe_bgr_u8");
}

extern "C" int trt_resize_linear_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int centre_crop, int dh, int dw,
                                    cudaStream_t stream) {
  TRT_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && dh > 0 && dw > 0, "trt_resize_linear_u8: bad argument");
  int sh = h, sw = w, y_off = 0, x_off = 0;
  if (centre_crop) {
    const int d = h < w ? h : w;
    y_off = (h - d) / 2; x_off = (w - d) / 2; sh = sw = d;
  }
  dim3 grid((dw + 255) / 256, dh, n);
  resize_linear_kernel<<<grid, 256, 0, stream>>>(src, dst, h, w, y_off, x_off, sh, sw, dh, dw);
  return trt_check_launch("trt_resize_linear_u8");
}

extern "C" int trt_normalize_flip_u8(const uint8_t* src, void* dst, int n, int h, int w, int flip, int out_bf16,
                                     cudaStream_t stream) {
  TRT_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && flip >= 0 && flip <= 2, "trt_normalize_flip_u8: bad argument");
  dim3 grid((w + 255) / 256, h, n);
  if (out_bf16)
    normalize_flip_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), h, w, flip);
  else
    normalize_flip_kernel<float><<<grid, 256, 0, stream>>>(src, reinterpret_cast<float*>(dst), h, w, flip);
  return trt_check_launch("trt_normalize_flip_u8");
}
