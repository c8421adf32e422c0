// Depthwise k3/k5 stride-1/2 convolutions, forward and backward, TF-'same' padding, NHWC bf16.
// Memory-bound: every block stages one input tile (with halo) in shared memory — all global loads of the tile are issued
// before any is consumed (memory-level parallelism), the producer's BatchNorm+SiLU (forward) or the BatchNorm-backward
// affine (backward) is applied once per element on the way in — then each thread keeps an input row segment in registers
// and slides the filter over it.  Outputs carry BN batch statistics (train) or folded BN + SiLU + SE pooling (eval).
//
// Replaces cuDNN/ATen depthwise conv launches inside `self.backbone(x_img)`
// (experiments/multimodal_v1/train_mm_joint_dualtask.py:154) and their autograd backward (:248).
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int CL = 8;     // channel lanes per block (8 lanes x 8 channels = 64 channels)
constexpr int NPT = TPB / CL;   // 32 pixel-threads per channel lane
constexpr int TOH = 8;    // output tile height

struct DwGeom {
  int N, H, W, C, OH, OW, S, pad_t, pad_l, tiles_x, tiles_y;
};

__device__ __forceinline__ uint4 zero4() { return make_uint4(0, 0, 0, 0); }

// Blackwell mixed-precision FMA (FHFMA.BF16): fp32 acc += bf16 * bf16, operands picked straight out of register halves, so
// the packed tiles in shared memory are never unpacked.  The product of two bf16 is exact in fp32 (one rounding, at the add).
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void fma8(float (&acc)[8], const uint4& a, const uint4& b) {
  const uint32_t as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[2 * i] = fhfma((uint16_t)(as[i] & 0xffffu), (uint16_t)(bs[i] & 0xffffu), acc[2 * i]);
    acc[2 * i + 1] = fhfma((uint16_t)(as[i] >> 16), (uint16_t)(bs[i] >> 16), acc[2 * i + 1]);
  }
}

// Block reduction of per-thread channel partials over the 32 pixel-threads of each channel lane: every thread parks its
// NV x 8 values in shared memory, then thread t < NV*64 sums the 32 entries of (k = t/64, channel c = t%64) — consecutive
// threads read consecutive floats (conflict-free) — and returns the total; other threads return 0.
template <int NV>
__device__ __forceinline__ float reduce_over_pt(float (&acc)[NV][8], float* s_red, int lane, int pt) {
  static_assert(NV * 64 <= TPB, "one thread per reduced value");
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[(k * NPT + pt) * 64 + lane * 8 + i] = acc[k][i];
  __syncthreads();
  float total = 0.f;
  if (threadIdx.x < NV * 64) {
    const int k = threadIdx.x / 64, c = threadIdx.x % 64;
#pragma unroll 8
    for (int q = 0; q < NPT; ++q) total += s_red[(k * NPT + q) * 64 + c];
  }
  return total;
}

// ---- tile loaders: all global loads first, then transform + store to shared memory
// act tile: v = in_rec ? silu(x*scale+shift) : x ; zero outside the image (conv padding applies to the ACTIVATED tensor)
template <int TH, int TW>
__device__ __forceinline__ void load_act_tile(uint4* s_tile, const uint4* __restrict__ x, const float* __restrict__ in_rec,
                                              int n, int H, int W, int V, int C, int gy0, int gx0, int cv, bool cvalid,
                                              int lane, int pt) {
  constexpr int NPIX = TH * TW, NL = (NPIX + NPT - 1) / NPT, CH = 6;
  f8 sc, sh;
  if (in_rec && cvalid) { sc = ldf8(in_rec + 8 * cv); sh = ldf8(in_rec + C + 8 * cv); }
#pragma unroll
  for (int j0 = 0; j0 < NL; j0 += CH) {
    uint4 v[CH];
    uint32_t ok = 0;
#pragma unroll
    for (int jj = 0; jj < CH; ++jj) {
      const int i = pt + NPT * (j0 + jj);
      const int iy = i / TW, ix = i - iy * TW;
      const int gy = gy0 + iy, gx = gx0 + ix;
      v[jj] = zero4();
      if (j0 + jj < NL && i < NPIX && cvalid && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        v[jj] = __ldg(x + ((size_t)(n * H + gy) * W + gx) * V + cv);
        ok |= 1u << jj;
      }
    }
#pragma unroll
    for (int jj = 0; jj < CH; ++jj) {
      const int i = pt + NPT * (j0 + jj);
      if (j0 + jj < NL && i < NPIX) {
        uint4 o = v[jj];
        if (in_rec && ((ok >> jj) & 1u)) {
          f8 a = unpack8(o);
#pragma unroll
          for (int k = 0; k < 8; ++k) a.v[k] = siluf_(fmaf(a.v[k], sc.v[k], sh.v[k]));
          o = pack8(a);
        }
        s_tile[i * CL + lane] = o;
      }
    }
  }
}

// gradient tile: v = coef ? a*gy + b*y_raw + c : gy ; zero outside [0,OH)x[0,OW)
template <int TH, int TW>
__device__ __forceinline__ void load_grad_tile(uint4* s_tile, const uint4* __restrict__ gy_, const uint4* __restrict__ y_raw,
                                               const float* __restrict__ coef, int n, int OH, int OW, int V, int C, int oy0,
                                               int ox0, int cv, bool cvalid, int lane, int pt) {
  constexpr int NPIX = TH * TW, NL = (NPIX + NPT - 1) / NPT, CH = 4;
  f8 ca, cb, cc;
  if (coef && cvalid) { ca = ldf8(coef + 8 * cv); cb = ldf8(coef + C + 8 * cv); cc = ldf8(coef + 2 * C + 8 * cv); }
#pragma unroll
  for (int j0 = 0; j0 < NL; j0 += CH) {
    uint4 v[CH], y[CH];
    uint32_t ok = 0;
#pragma unroll
    for (int jj = 0; jj < CH; ++jj) {
      const int i = pt + NPT * (j0 + jj);
      const int dy = i / TW, dx = i - dy * TW;
      const int oy = oy0 + dy, ox = ox0 + dx;
      v[jj] = zero4();
      y[jj] = zero4();
      if (j0 + jj < NL && i < NPIX && cvalid && oy >= 0 && oy < OH && ox >= 0 && ox < OW) {
        const size_t idx = ((size_t)(n * OH + oy) * OW + ox) * V + cv;
        v[jj] = __ldg(gy_ + idx);
        if (coef) y[jj] = __ldg(y_raw + idx);
        ok |= 1u << jj;
      }
    }
#pragma unroll
    for (int jj = 0; jj < CH; ++jj) {
      const int i = pt + NPT * (j0 + jj);
      if (j0 + jj < NL && i < NPIX) {
        uint4 o = v[jj];
        if (coef && ((ok >> jj) & 1u)) {
          f8 a = unpack8(o);
          const f8 yr = unpack8(y[jj]);
#pragma unroll
          for (int k = 0; k < 8; ++k) a.v[k] = fmaf(ca.v[k], a.v[k], fmaf(cb.v[k], yr.v[k], cc.v[k]));
          o = pack8(a);
        }
        s_tile[i * CL + lane] = o;
      }
    }
  }
}

// weights of the block's 64 channels -> shared bf16 [K*K][64] (one uint4 per tap and channel lane); flip = rotate the
// filter by 180 degrees (data gradient).  bf16 weights: the same rounding the 1x1 convs apply to theirs.
template <int K>
__device__ __forceinline__ void load_weights(uint4* s_w4, const float* __restrict__ w, int cb, int C, bool flip) {
  __nv_bfloat16* s_w = reinterpret_cast<__nv_bfloat16*>(s_w4);
  for (int i = threadIdx.x; i < K * K * 64; i += TPB) {
    const int tap = i / 64, c = cb * 64 + (i % 64);
    const int src_tap = flip ? (K * K - 1 - tap) : tap;
    s_w[i] = __float2bfloat16_rn(c < C ? __ldg(w + (size_t)c * K * K + src_tap) : 0.f);
  }
}

// sliding-window tile convolution: thread (lane, pt) computes P outputs of row oy starting at column oxb
template <int K, int S, int P, int IW>
__device__ __forceinline__ void conv_rows(const uint4* s_in, const uint4* s_w, int oy, int oxb, int lane, float (&acc)[P][8]) {
  constexpr int ROWV = (P - 1) * S + K;
#pragma unroll
  for (int kh = 0; kh < K; ++kh) {
    uint4 row[ROWV];
#pragma unroll
    for (int j = 0; j < ROWV; ++j) row[j] = s_in[((oy * S + kh) * IW + oxb * S + j) * CL + lane];
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const uint4 wv = s_w[(kh * K + kw) * CL + lane];
#pragma unroll
      for (int p = 0; p < P; ++p) fma8(acc[p], row[p * S + kw], wv);
    }
  }
}

template <int K, int S> struct FwdTile {
  static constexpr int TOW = S == 1 ? 16 : 8;
  static constexpr int P = TOW / 4;
  static constexpr int IH = (TOH - 1) * S + K, IW = (TOW - 1) * S + K;
  static constexpr size_t smem = (size_t)IH * IW * CL * 16 + (size_t)K * K * CL * 16;
};

// ------------------------------------------------------------------------------------------------ forward
template <int K, int S>
__global__ void __launch_bounds__(TPB, 3) dwconv_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ in_rec,
                                                            const float* __restrict__ w, uint4* __restrict__ out,
                                                            const float* __restrict__ out_rec, float* __restrict__ pooled,
                                                            double* __restrict__ stats, const DwGeom g) {
  using T = FwdTile<K, S>;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* s_in = reinterpret_cast<uint4*>(smem);
  uint4* s_w = reinterpret_cast<uint4*>(smem + (size_t)T::IH * T::IW * CL * 16);
  float* s_red = reinterpret_cast<float*>(smem);
  const int V = g.C / 8;
  const int tile = blockIdx.x, cb = blockIdx.y, n = blockIdx.z;
  const int oy0 = (tile / g.tiles_x) * TOH, ox0 = (tile % g.tiles_x) * T::TOW;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  load_weights<K>(s_w, w, cb, g.C, false);
  load_act_tile<T::IH, T::IW>(s_in, x, in_rec, n, g.H, g.W, V, g.C, oy0 * S - g.pad_t, ox0 * S - g.pad_l, cv, cvalid, lane, pt);
  __syncthreads();
  const int oy = pt / 4, oxb = (pt % 4) * T::P;
  float acc[T::P][8];
#pragma unroll
  for (int p = 0; p < T::P; ++p)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
  conv_rows<K, S, T::P, T::IW>(s_in, s_w, oy, oxb, lane, acc);

  float red[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
  f8 osc, osh;
  if (out_rec && cvalid) { osc = ldf8(out_rec + 8 * cv); osh = ldf8(out_rec + g.C + 8 * cv); }
  const int gy = oy0 + oy;
#pragma unroll
  for (int p = 0; p < T::P; ++p) {
    const int gx = ox0 + oxb + p;
    if (cvalid && gy < g.OH && gx < g.OW) {
      f8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = out_rec ? siluf_(fmaf(acc[p][i], osc.v[i], osh.v[i])) : acc[p][i];
      const uint4 q = pack8(o);
      out[((size_t)(n * g.OH + gy) * g.OW + gx) * V + cv] = q;
      const f8 r = unpack8(q);
#pragma unroll
      for (int i = 0; i < 8; ++i) { red[0][i] += r.v[i]; red[1][i] = fmaf(r.v[i], r.v[i], red[1][i]); }
    }
  }
  if (stats || pooled) {
    const float total = reduce_over_pt<2>(red, s_red, lane, pt);
    if (threadIdx.x < 128) {
      const int k = threadIdx.x / 64, c = cb * 64 + (threadIdx.x % 64);
      if (c < g.C) {
        if (stats) atomicAdd(stats + k * g.C + c, (double)total);
        if (pooled && k == 0) atomicAdd(pooled + (size_t)n * g.C + c, total);
      }
    }
  }
}

// epilogue shared by both data-gradient kernels: g = dIn * silu'(bn(x_raw)) + BN-backward sums, or plain dIn
template <int P>
__device__ __forceinline__ void prefetch_x(uint4 (&xr4)[P], const uint4* __restrict__ x_raw, const float* x_rec,
                                           const DwGeom& g, int n, int iy, int ixb, int cv, bool cvalid, int V) {
#pragma unroll
  for (int p = 0; p < P; ++p) {
    xr4[p] = zero4();
    if (x_rec && cvalid && iy < g.H && ixb + p < g.W) xr4[p] = __ldg(x_raw + ((size_t)(n * g.H + iy) * g.W + ixb + p) * V + cv);
  }
}

template <int P>
__device__ __forceinline__ void bwd_data_epilogue(float (&acc)[P][8], const uint4 (&xr4)[P], const uint4* __restrict__ x_raw,
                                                  const float* __restrict__ x_rec, uint4* __restrict__ g_out,
                                                  double* __restrict__ bstats, const DwGeom& g, int n, int iy, int ixb, int cv,
                                                  bool cvalid, int lane, int pt, float* s_red, int V, int cb) {
  float red[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
  f8 sc, sh, mu, rs;
  if (x_rec && cvalid) {
    sc = ldf8(x_rec + 8 * cv); sh = ldf8(x_rec + g.C + 8 * cv);
    mu = ldf8(x_rec + 2 * g.C + 8 * cv); rs = ldf8(x_rec + 3 * g.C + 8 * cv);
  }
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int ix = ixb + p;
    if (cvalid && iy < g.H && ix < g.W) {
      const size_t idx = ((size_t)(n * g.H + iy) * g.W + ix) * V + cv;
      f8 o;
      if (x_rec) {
        const f8 xr = unpack8(xr4[p]);
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = acc[p][i] * silu_gradf_(fmaf(xr.v[i], sc.v[i], sh.v[i]));
        const uint4 q = pack8(o);
        g_out[idx] = q;
        const f8 r = unpack8(q);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          red[0][i] += r.v[i];
          red[1][i] = fmaf(r.v[i], (xr.v[i] - mu.v[i]) * rs.v[i], red[1][i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = acc[p][i];
        g_out[idx] = pack8(o);
      }
    }
  }
  if (x_rec && bstats) {
    const float total = reduce_over_pt<2>(red, s_red, lane, pt);
    if (threadIdx.x < 128) {
      const int k = threadIdx.x / 64, c = cb * 64 + (threadIdx.x % 64);
      if (c < g.C) atomicAdd(bstats + k * g.C + c, (double)total);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward data, stride 1
// stride 1 + symmetric 'same' padding: dIn = conv(dD, rot180(w)) with the same padding -> the forward machinery
template <int K>
__global__ void __launch_bounds__(TPB, 2) dwconv_bwd_data_s1_kernel(const uint4* __restrict__ gy_, const uint4* __restrict__ y_raw,
                                                                    const float* __restrict__ coef, const float* __restrict__ w,
                                                                    const uint4* __restrict__ x_raw, const float* __restrict__ x_rec,
                                                                    uint4* __restrict__ g_out, double* __restrict__ bstats,
                                                                    const DwGeom g) {
  using T = FwdTile<K, 1>;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* s_in = reinterpret_cast<uint4*>(smem);
  uint4* s_w = reinterpret_cast<uint4*>(smem + (size_t)T::IH * T::IW * CL * 16);
  float* s_red = reinterpret_cast<float*>(smem);
  const int V = g.C / 8;
  const int tile = blockIdx.x, cb = blockIdx.y, n = blockIdx.z;
  const int iy0 = (tile / g.tiles_x) * TOH, ix0 = (tile % g.tiles_x) * T::TOW;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  constexpr int PAD = (K - 1) / 2;
  load_weights<K>(s_w, w, cb, g.C, true);
  const int oy = pt / 4, oxb = (pt % 4) * T::P;
  uint4 xr4[T::P];
  prefetch_x<T::P>(xr4, x_raw, x_rec, g, n, iy0 + oy, ix0 + oxb, cv, cvalid, V);      // needed only by the epilogue
  load_grad_tile<T::IH, T::IW>(s_in, gy_, y_raw, coef, n, g.OH, g.OW, V, g.C, iy0 - PAD, ix0 - PAD, cv, cvalid, lane, pt);
  __syncthreads();
  float acc[T::P][8];
#pragma unroll
  for (int p = 0; p < T::P; ++p)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
  conv_rows<K, 1, T::P, T::IW>(s_in, s_w, oy, oxb, lane, acc);
  bwd_data_epilogue<T::P>(acc, xr4, x_raw, x_rec, g_out, bstats, g, n, iy0 + oy, ix0 + oxb, cv, cvalid, lane, pt, s_red, V, cb);
}

// ------------------------------------------------------------------------------------------------ backward data, generic stride
template <int K>
__global__ void __launch_bounds__(TPB, 2) dwconv_bwd_data_kernel(const uint4* __restrict__ gy_, const uint4* __restrict__ y_raw,
                                                                 const float* __restrict__ coef, const float* __restrict__ w,
                                                                 const uint4* __restrict__ x_raw, const float* __restrict__ x_rec,
                                                                 uint4* __restrict__ g_out, double* __restrict__ bstats,
                                                                 const DwGeom g) {
  constexpr int TIW = 16, P = 4;
  constexpr int DH = (TOH + K - 2) / 2 + 2, DW = (TIW + K - 2) / 2 + 2;      // stride 2
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* s_d = reinterpret_cast<uint4*>(smem);
  uint4* s_w = reinterpret_cast<uint4*>(smem + (size_t)DH * DW * CL * 16);
  float* s_red = reinterpret_cast<float*>(smem);
  const int V = g.C / 8, S = g.S;
  const int tile = blockIdx.x, cb = blockIdx.y, n = blockIdx.z;
  const int iy0 = (tile / g.tiles_x) * TOH, ix0 = (tile % g.tiles_x) * TIW;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  const int ny = iy0 + g.pad_t - (K - 1), nx = ix0 + g.pad_l - (K - 1);
  const int oyb = ny <= 0 ? 0 : (ny + S - 1) / S, oxb0 = nx <= 0 ? 0 : (nx + S - 1) / S;
  load_weights<K>(s_w, w, cb, g.C, false);
  const int iy = iy0 + pt / 4, ixb = ix0 + (pt % 4) * P;
  uint4 xr4[P];
  prefetch_x<P>(xr4, x_raw, x_rec, g, n, iy, ixb, cv, cvalid, V);
  load_grad_tile<DH, DW>(s_d, gy_, y_raw, coef, n, g.OH, g.OW, V, g.C, oyb, oxb0, cv, cvalid, lane, pt);
  __syncthreads();
  float acc[P][8];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
  for (int kh = 0; kh < K; ++kh) {
    const int ty = iy + g.pad_t - kh;
    if (ty < 0 || (ty % S) != 0) continue;
    const int oy = ty / S;
    if (oy >= g.OH) continue;
    for (int kw = 0; kw < K; ++kw) {
      const uint4 wv = s_w[(kh * K + kw) * CL + lane];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const int tx = ixb + p + g.pad_l - kw;
        if (tx < 0 || (tx % S) != 0) continue;
        const int ox = tx / S;
        if (ox >= g.OW) continue;
        fma8(acc[p], s_d[((oy - oyb) * DW + (ox - oxb0)) * CL + lane], wv);
      }
    }
  }
  bwd_data_epilogue<P>(acc, xr4, x_raw, x_rec, g_out, bstats, g, n, iy, ixb, cv, cvalid, lane, pt, s_red, V, cb);
}

// ------------------------------------------------------------------------------------------------ backward weight
// dW[c][kh][kw] += sum_{n,oy,ox} dD[n,oy,ox,c] * act(x)[n, oy*S-pad_t+kh, ox*S-pad_l+kw, c]
// thread = (channel lane, filter row kh, output-row subset); the filter row slides over an input row kept in registers.
template <int K, int S>
__global__ void __launch_bounds__(TPB, 2) dwconv_bwd_weight_kernel(const uint4* __restrict__ gy_, const uint4* __restrict__ y_raw,
                                                                   const float* __restrict__ coef, const uint4* __restrict__ x,
                                                                   const float* __restrict__ in_rec, float* __restrict__ dw,
                                                                   const DwGeom g) {
  constexpr int TOW = 8;
  constexpr int IH = (TOH - 1) * S + K, IW = (TOW - 1) * S + K;
  constexpr int SUBS = NPT / K;                 // row subsets per filter row (K=3: 10, K=5: 6)
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* s_in = reinterpret_cast<uint4*>(smem);                       // [IH][IW][CL]
  uint4* s_d = s_in + (size_t)IH * IW * CL;                           // [TOH][TOW][CL]
  float* s_acc = reinterpret_cast<float*>(s_d + (size_t)TOH * TOW * CL);   // [K*K][64]
  const int V = g.C / 8;
  const int tile = blockIdx.x, cb = blockIdx.y;
  const int oy0 = (tile / g.tiles_x) * TOH, ox0 = (tile % g.tiles_x) * TOW;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  const int kh = pt % K, sub = pt / K;
  const bool worker = sub < SUBS;
  for (int i = threadIdx.x; i < K * K * 64; i += TPB) s_acc[i] = 0.f;
  float acc[K][8];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
  for (int n = blockIdx.z; n < g.N; n += gridDim.z) {
    __syncthreads();
    load_act_tile<IH, IW>(s_in, x, in_rec, n, g.H, g.W, V, g.C, oy0 * S - g.pad_t, ox0 * S - g.pad_l, cv, cvalid, lane, pt);
    load_grad_tile<TOH, TOW>(s_d, gy_, y_raw, coef, n, g.OH, g.OW, V, g.C, oy0, ox0, cv, cvalid, lane, pt);
    __syncthreads();
    if (worker) {
      for (int oy = sub; oy < TOH; oy += SUBS) {
        const uint4* xrow = s_in + ((oy * S + kh) * IW) * CL + lane;
        const uint4* drow = s_d + (oy * TOW) * CL + lane;
        if (S == 1) {
          uint4 win[K];     // sliding window over the input row; the shifts below are register renames after unrolling
#pragma unroll
          for (int j = 0; j < K - 1; ++j) win[j] = xrow[j * CL];
#pragma unroll
          for (int ox = 0; ox < TOW; ++ox) {
            win[K - 1] = xrow[(ox + K - 1) * CL];
            const uint4 d = drow[ox * CL];
#pragma unroll
            for (int kw = 0; kw < K; ++kw) fma8(acc[kw], d, win[kw]);
#pragma unroll
            for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
          }
        } else {
#pragma unroll
          for (int ox = 0; ox < TOW; ++ox) {
            const uint4 d = drow[ox * CL];
#pragma unroll
            for (int kw = 0; kw < K; ++kw) fma8(acc[kw], d, xrow[(ox * S + kw) * CL]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int kw = 0; kw < K; ++kw)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(s_acc + (kh * K + kw) * 64 + lane * 8 + i, acc[kw][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * 64; i += TPB) {
    const int tap = i / 64, c = cb * 64 + (i % 64);
    if (c < g.C) atomicAdd(dw + (size_t)c * K * K + tap, s_acc[i]);
  }
}

inline void same_pad(int i, int k, int s, int& out, int& pad_before) {
  out = (i + s - 1) / s;
  int total = (out - 1) * s + k - i;
  if (total < 0) total = 0;
  pad_before = total / 2;
}

template <typename Kern>
int set_smem_attr(Kern kern) {
  TRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  return TRT_OK;
}

}  // namespace

extern "C" int trt_dwconv_fwd(const void* x, const float* in_rec, const float* w, void* out, const float* out_rec,
                              float* pooled_sum, double* stats, int N, int H, int W, int C, int k, int s,
                              cudaStream_t stream) {
  TRT_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "trt_dwconv_fwd: bad argument");
  TRT_REQUIRE((k == 3 || k == 5) && (s == 1 || s == 2), "trt_dwconv_fwd: only k in {3,5}, s in {1,2}");
  DwGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.S = s;
  same_pad(H, k, s, g.OH, g.pad_t);
  same_pad(W, k, s, g.OW, g.pad_l);
  const int tow = s == 1 ? 16 : 8;
  g.tiles_x = (g.OW + tow - 1) / tow;
  g.tiles_y = (g.OH + TOH - 1) / TOH;
  if (pooled_sum) TRT_CUDA(cudaMemsetAsync(pooled_sum, 0, (size_t)N * C * sizeof(float), stream));
  const size_t red_bytes = 2 * NPT * 64 * 4;
  dim3 grid(g.tiles_x * g.tiles_y, (C / 8 + CL - 1) / CL, N);
#define LAUNCH_DW(KK, SS)                                                                                          \
  do {                                                                                                             \
    static bool attr = false;                                                                                      \
    if (!attr) { int rc = set_smem_attr(dwconv_fwd_kernel<KK, SS>); if (rc) return rc; attr = true; }               \
    size_t smem = FwdTile<KK, SS>::smem < red_bytes ? red_bytes : FwdTile<KK, SS>::smem;                            \
    dwconv_fwd_kernel<KK, SS><<<grid, TPB, smem, stream>>>((const uint4*)x, in_rec, w, (uint4*)out, out_rec, pooled_sum, stats, g); \
  } while (0)
  if (k == 3 && s == 1) LAUNCH_DW(3, 1);
  else if (k == 3 && s == 2) LAUNCH_DW(3, 2);
  else if (k == 5 && s == 1) LAUNCH_DW(5, 1);
  else LAUNCH_DW(5, 2);
#undef LAUNCH_DW
  return trt_check_launch("trt_dwconv_fwd");
}

extern "C" int trt_dwconv_bwd(const void* gy, const void* y_raw, const float* coef, const float* w, const void* x_raw,
                              const float* x_rec, void* g_out, double* bstats, float* dw, int N, int H, int W, int C, int k,
                              int s, cudaStream_t stream) {
  TRT_REQUIRE(gy && w && x_raw && dw && N > 0 && C > 0 && C % 8 == 0, "trt_dwconv_bwd: bad argument");
  TRT_REQUIRE(!coef || y_raw, "trt_dwconv_bwd: coef needs y_raw");
  TRT_REQUIRE((k == 3 || k == 5) && (s == 1 || s == 2), "trt_dwconv_bwd: only k in {3,5}, s in {1,2}");
  DwGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.S = s;
  same_pad(H, k, s, g.OH, g.pad_t);
  same_pad(W, k, s, g.OW, g.pad_l);
  const int cblocks = (C / 8 + CL - 1) / CL;
  const size_t red_bytes = 2 * NPT * 64 * 4;
  if (g_out) {   // data gradient (skipped for a first layer whose input needs no gradient)
    g.tiles_x = (W + 15) / 16;
    g.tiles_y = (H + TOH - 1) / TOH;
    dim3 grid(g.tiles_x * g.tiles_y, cblocks, N);
#define LAUNCH_BD_S1(KK)                                                                                           \
  do {                                                                                                             \
    static bool attr = false;                                                                                      \
    if (!attr) { int rc = set_smem_attr(dwconv_bwd_data_s1_kernel<KK>); if (rc) return rc; attr = true; }           \
    size_t smem = FwdTile<KK, 1>::smem < red_bytes ? red_bytes : FwdTile<KK, 1>::smem;                              \
    dwconv_bwd_data_s1_kernel<KK><<<grid, TPB, smem, stream>>>((const uint4*)gy, (const uint4*)y_raw, coef, w, (const uint4*)x_raw, x_rec, (uint4*)g_out, bstats, g); \
  } while (0)
#define LAUNCH_BD(KK)                                                                                              \
  do {                                                                                                             \
    static bool attr = false;                                                                                      \
    if (!attr) { int rc = set_smem_attr(dwconv_bwd_data_kernel<KK>); if (rc) return rc; attr = true; }              \
    const int DH = (TOH + KK - 2) / 2 + 2, DW = (16 + KK - 2) / 2 + 2;                                             \
    size_t smem = (size_t)DH * DW * CL * 16 + (size_t)KK * KK * CL * 16;                                            \
    if (smem < red_bytes) smem = red_bytes;                                                                        \
    dwconv_bwd_data_kernel<KK><<<grid, TPB, smem, stream>>>((const uint4*)gy, (const uint4*)y_raw, coef, w, (const uint4*)x_raw, x_rec, (uint4*)g_out, bstats, g); \
  } while (0)
    if (s == 1) { if (k == 3) LAUNCH_BD_S1(3); else LAUNCH_BD_S1(5); }
    else { if (k == 3) LAUNCH_BD(3); else LAUNCH_BD(5); }
#undef LAUNCH_BD_S1
#undef LAUNCH_BD
    int rc = trt_check_launch("trt_dwconv_bwd(data)");
    if (rc) return rc;
  }
  {
    g.tiles_x = (g.OW + 7) / 8;
    g.tiles_y = (g.OH + TOH - 1) / TOH;
    int zsplit = (4 * trt_num_sms()) / (g.tiles_x * g.tiles_y * cblocks);
    if (zsplit < 1) zsplit = 1;
    if (zsplit > N) zsplit = N;
    dim3 grid(g.tiles_x * g.tiles_y, cblocks, zsplit);
#define LAUNCH_BW(KK, SS)                                                                                          \
  do {                                                                                                             \
    static bool attr = false;                                                                                      \
    if (!attr) { int rc = set_smem_attr(dwconv_bwd_weight_kernel<KK, SS>); if (rc) return rc; attr = true; }        \
    const int IH = (TOH - 1) * SS + KK, IW = 7 * SS + KK;                                                          \
    const size_t smem = ((size_t)IH * IW + TOH * 8) * CL * 16 + (size_t)KK * KK * 64 * 4;                           \
    dwconv_bwd_weight_kernel<KK, SS><<<grid, TPB, smem, stream>>>((const uint4*)gy, (const uint4*)y_raw, coef, (const uint4*)x_raw, x_rec, dw, g); \
  } while (0)
    if (k == 3 && s == 1) LAUNCH_BW(3, 1);
    else if (k == 3 && s == 2) LAUNCH_BW(3, 2);
    else if (k == 5 && s == 1) LAUNCH_BW(5, 1);
    else LAUNCH_BW(5, 2);
#undef LAUNCH_BW
  }
  return trt_check_launch("trt_dwconv_bwd(weight)");
}
