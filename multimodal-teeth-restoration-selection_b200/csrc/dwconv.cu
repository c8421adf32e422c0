// Depthwise k3/k5 stride-1/2 convolutions, forward and backward, TF-'same' padding, NHWC bf16.
//
// B200 design: PERSISTENT blocks fed by 4D TMA.  A block owns one 64-channel group and walks a strided list of
// (image, tile) items.  For every item one thread issues `cp.async.bulk.tensor.4d` for the input tile WITH its halo —
// coordinates run negative / past the edge and the TMA unit zero-fills, which is the convolution's padding — into a
// double-buffered shared-memory tile, so the next tile streams in while the current one is computed.  The producer's
// BatchNorm+SiLU is applied in place in shared memory (once per element), then every thread slides the filter over packed
// bf16 rows with FHFMA.BF16 (`fma.rn.f32.bf16`: fp32 accumulate, operands taken from register halves, no unpack).
// Filter taps are staged once per block; BN batch statistics and weight gradients accumulate in registers across all the
// block's items and are reduced / published once per block.
//
// Replaces cuDNN/ATen depthwise conv launches inside `self.backbone(x_img)`
// (experiments/multimodal_v1/train_mm_joint_dualtask.py:154) and their autograd backward (:248).
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int CL = 8;     // channel lanes per block (8 lanes x 8 channels = 64 channels)
constexpr int NPT = TPB / CL;   // 32 pixel-threads per channel lane
constexpr int TOH = 8;    // output tile height
#ifndef DW_MINB
#define DW_MINB 2         // resident blocks per SM the persistent kernels are compiled for (register cap 65536 / (256 * DW_MINB))
#endif
constexpr int PX_BYTES = CL * 16;   // one pixel of a 64-channel tile

struct DwGeom {
  int N, H, W, C, OH, OW, S, pad_t, pad_l, tiles_x, tiles_y;
  uint32_t magic_tiles, magic_tx;      // ceil(2^32 / d): exact quotients by __umulhi for the item counts that occur
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

// weights of the block's 64 channels -> shared bf16 [K*K][64] (one uint4 per tap and channel lane); flip = rotate the
// filter by 180 degrees (data gradient).  bf16 weights: the same rounding the 1x1 convs apply to theirs.
template <int K>
__device__ __forceinline__ void load_weights(uint4* s_w4, const float* __restrict__ w, int cb, int C, bool flip) {
  __nv_bfloat16* s_w = reinterpret_cast<__nv_bfloat16*>(s_w4);
  for (int i = threadIdx.x; i < K * K * 64; i += TPB) {
    const int cl = i / (K * K), tap = i % (K * K);        // consecutive threads read consecutive floats
    const int c = cb * 64 + cl;
    const int dst_tap = flip ? (K * K - 1 - tap) : tap;
    s_w[dst_tap * 64 + cl] = __float2bfloat16_rn(c < C ? __ldg(w + (size_t)c * K * K + tap) : 0.f);
  }
}

// in-place producer activation of a staged tile: v = silu(x*scale+shift) for the pixels inside the image (the zero-filled
// halo must stay zero: padding applies to the ACTIVATED tensor)
template <int TH, int TW>
__device__ __forceinline__ void activate_tile(uint4* s_tile, const f8& sc, const f8& sh, int gy0, int gx0, int H, int W,
                                              int lane, int pt) {
  constexpr int NPIX = TH * TW, NL = (NPIX + NPT - 1) / NPT;
  f8 sc2, sh2;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc2.v[k] = sc.v[k] * -1.4426950408889634f; sh2.v[k] = sh.v[k] * -1.4426950408889634f; }
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    const int i = pt + NPT * j;
    const int iy = i / TW, ix = i - iy * TW;
    const int gy = gy0 + iy, gx = gx0 + ix;
    if (i < NPIX && gy >= 0 && gy < H && gx >= 0 && gx < W) {
      f8 a = unpack8(s_tile[i * CL + lane]);
#pragma unroll
      for (int k = 0; k < 8; ++k) a.v[k] = silu_affine_(a.v[k], sc.v[k], sh.v[k], sc2.v[k], sh2.v[k]);
      s_tile[i * CL + lane] = pack8(a);
    }
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

// output tile = TOH x (4*P); P = outputs per thread along x
template <int K, int S, int P> struct FwdTile {
  static constexpr int TOW = 4 * P;
  static constexpr int IH = (TOH - 1) * S + K, IW = (TOW - 1) * S + K;
  static constexpr int IN_BYTES = IH * IW * PX_BYTES;           // multiple of 128
  static constexpr int OUT_BYTES = TOH * TOW * PX_BYTES;
};

struct Item { int n, ty, tx; };
__device__ __forceinline__ Item decode_item(int it, int tiles, const DwGeom& g) {
  Item r;
  r.n = tiles == 1 ? it : (int)__umulhi((uint32_t)it, g.magic_tiles);
  const int t = it - r.n * tiles;
  r.ty = g.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, g.magic_tx);
  r.tx = t - r.ty * g.tiles_x;
  return r;
}

// ------------------------------------------------------------------------------------------------ forward
template <int K, int S, int P>
__global__ void __launch_bounds__(TPB, DW_MINB) dwconv_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ in_rec,
                                                            const float* __restrict__ w, uint4* __restrict__ out,
                                                            const float* __restrict__ out_rec, float* __restrict__ pooled,
                                                            double* __restrict__ stats, const DwGeom g, const int has_fin,
                                                            const trt_bn_fin_t fin, uint4* __restrict__ act_out) {
  using T = FwdTile<K, S, P>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = TRT_ALIGNED_SMEM(smem_raw, 128);
  uint4* s_w = reinterpret_cast<uint4*>(smem + 2 * T::IN_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * T::IN_BYTES + K * K * PX_BYTES);
  const int V = g.C / 8, cb = blockIdx.y;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  pdl_launch_dependents();
  pdl_wait();
  const int tiles = g.tiles_x * g.tiles_y, items = g.N * tiles, G = gridDim.x;
  auto issue = [&](int it, int st) {
    const Item q = decode_item(it, tiles, g);
    ptx::mbar_expect_tx(&bar[st], T::IN_BYTES);
    ptx::tma_load_4d(smem + st * T::IN_BYTES, &tm_x, &bar[st], cb * 64, q.tx * T::TOW * S - g.pad_l, q.ty * TOH * S - g.pad_t, q.n);
  };
  // the first tile is requested BEFORE the prologue (filter staging, lazy BatchNorm record): its HBM round trip overlaps them
  // (the issuing thread initialised the barriers itself; the other threads meet them after the __syncthreads below)
  if (threadIdx.x == 0 && (int)blockIdx.x < items) issue(blockIdx.x, 0);
  load_weights<K>(s_w, w, cb, g.C, false);
  __syncthreads();
  f8 sc, sh, osc, osh;
  if (has_fin) {
    // lazy BatchNorm on the input: scale/shift of this block's 64 channels straight from the producer's statistics; the
    // first block of each channel group publishes the record (the backward pass reads it)
    __shared__ __align__(16) float s_ss[2][64];
    bn_lazy_block(fin, g.C, cb * 64, 64, s_ss[0], s_ss[1], blockIdx.x == 0);
    if (cvalid) { sc = lds8(s_ss[0] + 8 * lane); sh = lds8(s_ss[1] + 8 * lane); }
  } else if (in_rec && cvalid) { sc = ldf8(in_rec + 8 * cv); sh = ldf8(in_rec + g.C + 8 * cv); }
  if (out_rec && cvalid) { osc = ldf8(out_rec + 8 * cv); osh = ldf8(out_rec + g.C + 8 * cv); }
  float red[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
  const int oy = pt / 4, oxb = (pt % 4) * P;
  int it = blockIdx.x, st = 0;
  uint32_t phase = 0;
  for (; it < items; it += G, st ^= 1) {
    const Item q = decode_item(it, tiles, g);
    const int oy0 = q.ty * TOH, ox0 = q.tx * T::TOW;
    if (threadIdx.x == 0 && it + G < items) {
      ptx::fence_proxy_async();          // the other buffer was last touched by generic-proxy stores (in-place activation)
      issue(it + G, st ^ 1);
    }
    uint4* s_in = reinterpret_cast<uint4*>(smem + st * T::IN_BYTES);
    ptx::mbar_wait(&bar[st], (phase >> st) & 1u);
    phase ^= 1u << st;
    if (in_rec) {
      if (cvalid) activate_tile<T::IH, T::IW>(s_in, sc, sh, oy0 * S - g.pad_t, ox0 * S - g.pad_l, g.H, g.W, lane, pt);
      __syncthreads();
    }
    float acc[P][8];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
    conv_rows<K, S, P, T::IW>(s_in, s_w, oy, oxb, lane, acc);
    const int gy = oy0 + oy;
    if (S == 1 && act_out) {
      // training, stride 1: the activated input of this tile's centre goes to HBM once, so the weight-gradient kernel reads
      // it instead of recomputing silu(bn(x)) over its own halo'd tiles (two MUFU per element, 1.6-2.3x redundant there)
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const int gx = ox0 + oxb + p;
        if (cvalid && gy < g.H && gx < g.W)
          act_out[((size_t)(q.n * g.H + gy) * g.W + gx) * V + cv] = s_in[((oy + g.pad_t) * T::IW + oxb + p + g.pad_l) * CL + lane];
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int gx = ox0 + oxb + p;
      if (cvalid && gy < g.OH && gx < g.OW) {
        f8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = out_rec ? siluf_(fmaf(acc[p][i], osc.v[i], osh.v[i])) : acc[p][i];
        const uint4 qv = pack8(o);
        out[((size_t)(q.n * g.OH + gy) * g.OW + gx) * V + cv] = qv;
        const f8 r = unpack8(qv);
#pragma unroll
        for (int i = 0; i < 8; ++i) { red[0][i] += r.v[i]; red[1][i] = fmaf(r.v[i], r.v[i], red[1][i]); }
      }
    }
    if (pooled) {      // eval: SE pooling is per image -> publish per tile (few tiles at inference batch sizes)
      float red1[1][8];
#pragma unroll
      for (int i = 0; i < 8; ++i) red1[0][i] = red[0][i];
      const float total = reduce_over_pt<1>(red1, reinterpret_cast<float*>(s_in), lane, pt);    // 8 KB <= any tile
      if (threadIdx.x < 64) {
        const int c = cb * 64 + threadIdx.x;
        if (c < g.C) atomicAdd(pooled + (size_t)q.n * g.C + c, total);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
    }
    __syncthreads();                     // everyone is done reading this buffer: the next-but-one tile may land in it
  }
  if (stats) {
    const float total = reduce_over_pt<2>(red, reinterpret_cast<float*>(smem), lane, pt);
    if (threadIdx.x < 128) {
      const int k = threadIdx.x / 64, c = cb * 64 + (threadIdx.x % 64);
      if (c < g.C) atomicAdd(stats + ((size_t)(blockIdx.x % TRT_STAT_REPLICAS) * 2 + k) * g.C + c, (double)total);
    }
  }
}

// epilogue shared by both data-gradient kernels: g = dIn * silu'(bn(x_raw)) + BN-backward sums, or plain dIn
template <int P>
__device__ __forceinline__ void bwd_data_epilogue(float (&acc)[P][8], const uint4 (&xr4)[P], const float* __restrict__ x_rec,
                                                  uint4* __restrict__ g_out, float (&red)[2][8], const f8& sc, const f8& sh,
                                                  const DwGeom& g, int n, int iy, int ixb, int cv, bool cvalid, int V) {
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
          red[1][i] = fmaf(r.v[i], xr.v[i], red[1][i]);       // sum g*x; the caller turns it into sum g*xhat at the end
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = acc[p][i];
        g_out[idx] = pack8(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward data, stride 1
// stride 1 + symmetric 'same' padding: dIn = conv(dD, rot180(w)) with the same padding -> the forward machinery.
// Per item two TMA tiles: dD with halo, and (when the input carried a BatchNorm+SiLU) the raw input tile for silu'.
template <int K, int P>
__global__ void __launch_bounds__(TPB, DW_MINB) dwconv_bwd_data_s1_kernel(const __grid_constant__ CUtensorMap tm_d,
                                                                    const __grid_constant__ CUtensorMap tm_x,
                                                                    const float* __restrict__ w, const float* __restrict__ x_rec,
                                                                    uint4* __restrict__ g_out, double* __restrict__ bstats,
                                                                    const DwGeom g) {
  using T = FwdTile<K, 1, P>;
  constexpr int STAGE = T::IN_BYTES + T::OUT_BYTES;
  constexpr int PAD = (K - 1) / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = TRT_ALIGNED_SMEM(smem_raw, 128);
  uint4* s_w = reinterpret_cast<uint4*>(smem + 2 * STAGE);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * STAGE + K * K * PX_BYTES);
  const int V = g.C / 8, cb = blockIdx.y;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_d);
    ptx::prefetch_tmap(&tm_x);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  const int tiles = g.tiles_x * g.tiles_y, items = g.N * tiles, G = gridDim.x;
  auto issue = [&](int it, int st) {
    const Item q = decode_item(it, tiles, g);
    ptx::mbar_expect_tx(&bar[st], x_rec ? STAGE : T::IN_BYTES);
    ptx::tma_load_4d(smem + st * STAGE, &tm_d, &bar[st], cb * 64, q.tx * T::TOW - PAD, q.ty * TOH - PAD, q.n);
    if (x_rec) ptx::tma_load_4d(smem + st * STAGE + T::IN_BYTES, &tm_x, &bar[st], cb * 64, q.tx * T::TOW, q.ty * TOH, q.n);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < items) issue(blockIdx.x, 0);      // first tile in flight during the filter staging
  load_weights<K>(s_w, w, cb, g.C, true);
  __syncthreads();
  f8 sc, sh;
  if (x_rec && cvalid) { sc = ldf8(x_rec + 8 * cv); sh = ldf8(x_rec + g.C + 8 * cv); }
  float red[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
  const int oy = pt / 4, oxb = (pt % 4) * P;
  int it = blockIdx.x, st = 0;
  uint32_t phase = 0;
  for (; it < items; it += G, st ^= 1) {
    const Item q = decode_item(it, tiles, g);
    if (threadIdx.x == 0 && it + G < items) issue(it + G, st ^ 1);
    const uint4* s_in = reinterpret_cast<const uint4*>(smem + st * STAGE);
    const uint4* s_x = reinterpret_cast<const uint4*>(smem + st * STAGE + T::IN_BYTES);
    ptx::mbar_wait(&bar[st], (phase >> st) & 1u);
    phase ^= 1u << st;
    float acc[P][8];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
    conv_rows<K, 1, P, T::IW>(s_in, s_w, oy, oxb, lane, acc);
    uint4 xr4[P];
#pragma unroll
    for (int p = 0; p < P; ++p) xr4[p] = x_rec ? s_x[(oy * T::TOW + oxb + p) * CL + lane] : zero4();
    bwd_data_epilogue<P>(acc, xr4, x_rec, g_out, red, sc, sh, g, q.n, q.ty * TOH + oy, q.tx * T::TOW + oxb, cv, cvalid, V);
    __syncthreads();
  }
  if (x_rec && bstats) {
    if (cvalid) {     // sum g*xhat = rstd * (sum g*x - mean * sum g): mean / rstd are not live across the tile loop
      const f8 mu = ldf8(x_rec + 2 * g.C + 8 * cv), rs = ldf8(x_rec + 3 * g.C + 8 * cv);
#pragma unroll
      for (int i = 0; i < 8; ++i) red[1][i] = rs.v[i] * (red[1][i] - mu.v[i] * red[0][i]);
    }
    const float total = reduce_over_pt<2>(red, reinterpret_cast<float*>(smem), lane, pt);
    if (threadIdx.x < 128) {
      const int k = threadIdx.x / 64, c = cb * 64 + (threadIdx.x % 64);
      if (c < g.C) atomicAdd(bstats + ((size_t)(blockIdx.x % TRT_STAT_REPLICAS) * 2 + k) * g.C + c, (double)total);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward data, stride 2
// Parity decomposition: an input pixel (iy, ix) only sees the taps kh = (iy+pad_t) mod 2 + 2a, kw = (ix+pad_l) mod 2 + 2b, and
// dIn[iy,ix] = sum_{a,b} dD[(iy+pad_t-kh)/2, (ix+pad_l-kw)/2] * w[kh,kw]: each of the four parity classes is a small
// stride-1 correlation over dD, so a thread that owns 8 same-parity pixels of one row slides its sub-filter over one packed
// row of dD.  A warp is one parity class (no divergence); tile = 16x16 input pixels, dD tile + raw input tile via TMA.
template <int K> struct S2Tile {
  static constexpr int TI = 16;                                   // input tile (rows and columns)
  static constexpr int DH = TI / 2 + (K - 1) / 2 + 1, DW = DH;    // dD rows/cols that can touch the tile
  static constexpr int D_BYTES = DH * DW * PX_BYTES, X_BYTES = TI * TI * PX_BYTES, STAGE = D_BYTES + X_BYTES;
};

// NP = pixels per call: the thread's 8 same-parity pixels are computed as two halves of 4 (accumulators for all 8 at once
// need 64 registers on top of the row window: the k5 kernel spilled 800 bytes per thread under the 128-register cap)
template <int K, int PY, int PX, int NP>
__device__ __forceinline__ void s2_class_rows(const uint4* s_d, const uint4* s_w, int dy, int dx, int lane, float (&acc)[NP][8]) {
  // taps of this class: kh = PY + 2a (a < NA), kw = PX + 2b (b < NB); dD row = dy - a, dD col = dx - b + p
  constexpr int NA = (K - PY + 1) / 2, NB = (K - PX + 1) / 2, DW = S2Tile<K>::DW;
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    uint4 row[NP + NB - 1];
#pragma unroll
    for (int j = 0; j < NP + NB - 1; ++j) row[j] = s_d[((dy - a) * DW + dx - (NB - 1) + j) * CL + lane];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const uint4 wv = s_w[((PY + 2 * a) * K + PX + 2 * b) * CL + lane];
#pragma unroll
      for (int p = 0; p < NP; ++p) fma8(acc[p], row[p + (NB - 1) - b], wv);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(TPB, 2) dwconv_bwd_data_s2_kernel(const __grid_constant__ CUtensorMap tm_d,
                                                                    const __grid_constant__ CUtensorMap tm_x,
                                                                    const float* __restrict__ w, const float* __restrict__ x_rec,
                                                                    uint4* __restrict__ g_out, double* __restrict__ bstats,
                                                                    const DwGeom g) {
  using T = S2Tile<K>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = TRT_ALIGNED_SMEM(smem_raw, 128);
  uint4* s_w = reinterpret_cast<uint4*>(smem + 2 * T::STAGE);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * T::STAGE + K * K * PX_BYTES);
  const int V = g.C / 8, cb = blockIdx.y;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_d);
    ptx::prefetch_tmap(&tm_x);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  const int tiles = g.tiles_x * g.tiles_y, items = g.N * tiles, G = gridDim.x;
  // first dD row/col a tile can touch: floor((i0 + pad - (K-1)) / 2); i0 is a multiple of 16 and pad <= K-1, so the
  // numerator is >= -(K-1) and the floor is taken on a shifted non-negative value
  auto d_origin = [&](int i0, int pad) { return (i0 + pad - (K - 1) + 16) / 2 - 8; };
  auto issue = [&](int it, int st) {
    const Item q = decode_item(it, tiles, g);
    ptx::mbar_expect_tx(&bar[st], x_rec ? T::STAGE : T::D_BYTES);
    ptx::tma_load_4d(smem + st * T::STAGE, &tm_d, &bar[st], cb * 64, d_origin(q.tx * T::TI, g.pad_l), d_origin(q.ty * T::TI, g.pad_t), q.n);
    if (x_rec) ptx::tma_load_4d(smem + st * T::STAGE + T::D_BYTES, &tm_x, &bar[st], cb * 64, q.tx * T::TI, q.ty * T::TI, q.n);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < items) issue(blockIdx.x, 0);      // first tile in flight during the filter staging
  load_weights<K>(s_w, w, cb, g.C, false);
  __syncthreads();
  f8 sc, sh;
  if (x_rec && cvalid) { sc = ldf8(x_rec + 8 * cv); sh = ldf8(x_rec + g.C + 8 * cv); }
  float red[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
  // warp -> parity class (py, px) of (iy + pad_t, ix + pad_l); the warp's 4 pixel-threads and its twin warp cover 8 rows
  const int warp = threadIdx.x >> 5;
  const int py = (warp >> 1) & 1, px = warp & 1;
  const int u = (warp >> 2) * 4 + (pt & 3);                       // 0..7: which row of parity py
  const int ry = (py + g.pad_t) & 1, rx = (px + g.pad_l) & 1;     // tile-local offset of the class's first row / column
  const int r = 2 * u + ry;                                       // tile-local input row; the thread's columns are rx + 2p
  int it = blockIdx.x, st = 0;
  uint32_t phase = 0;
  for (; it < items; it += G, st ^= 1) {
    const Item q = decode_item(it, tiles, g);
    if (threadIdx.x == 0 && it + G < items) issue(it + G, st ^ 1);
    const uint4* s_d = reinterpret_cast<const uint4*>(smem + st * T::STAGE);
    const uint4* s_x = reinterpret_cast<const uint4*>(smem + st * T::STAGE + T::D_BYTES);
    ptx::mbar_wait(&bar[st], (phase >> st) & 1u);
    phase ^= 1u << st;
    const int iy0 = q.ty * T::TI, ix0 = q.tx * T::TI;
    // dD coordinates (tile-local) of tap (a = 0, b = 0) for this thread's first pixel
    const int dy = (iy0 + r + g.pad_t - py) / 2 - d_origin(iy0, g.pad_t);
    const int dx = (ix0 + rx + g.pad_l - px) / 2 - d_origin(ix0, g.pad_l);
    const int iy = iy0 + r;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      constexpr int NP = 4;
      float acc[NP][8];
#pragma unroll
      for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
      const int dxh = dx + NP * half;
      if (py == 0) {
        if (px == 0) s2_class_rows<K, 0, 0, NP>(s_d, s_w, dy, dxh, lane, acc);
        else s2_class_rows<K, 0, 1, NP>(s_d, s_w, dy, dxh, lane, acc);
      } else {
        if (px == 0) s2_class_rows<K, 1, 0, NP>(s_d, s_w, dy, dxh, lane, acc);
        else s2_class_rows<K, 1, 1, NP>(s_d, s_w, dy, dxh, lane, acc);
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int pp = NP * half + p;
        const int ix = ix0 + rx + 2 * pp;
        if (cvalid && iy < g.H && ix < g.W) {
          const size_t idx = ((size_t)(q.n * g.H + iy) * g.W + ix) * V + cv;
          f8 o;
          if (x_rec) {
            const f8 xr = unpack8(s_x[(r * T::TI + rx + 2 * pp) * CL + lane]);
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] = acc[p][i] * silu_gradf_(fmaf(xr.v[i], sc.v[i], sh.v[i]));
            const uint4 qv = pack8(o);
            g_out[idx] = qv;
            const f8 rr = unpack8(qv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              red[0][i] += rr.v[i];
              red[1][i] = fmaf(rr.v[i], xr.v[i], red[1][i]);     // sum g*x, fixed up after the loop
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] = acc[p][i];
            g_out[idx] = pack8(o);
          }
        }
      }
    }
    __syncthreads();
  }
  if (x_rec && bstats) {
    if (cvalid) {     // sum g*xhat = rstd * (sum g*x - mean * sum g): mean / rstd are not live across the tile loop
      const f8 mu = ldf8(x_rec + 2 * g.C + 8 * cv), rs = ldf8(x_rec + 3 * g.C + 8 * cv);
#pragma unroll
      for (int i = 0; i < 8; ++i) red[1][i] = rs.v[i] * (red[1][i] - mu.v[i] * red[0][i]);
    }
    const float total = reduce_over_pt<2>(red, reinterpret_cast<float*>(smem), lane, pt);
    if (threadIdx.x < 128) {
      const int k = threadIdx.x / 64, c = cb * 64 + (threadIdx.x % 64);
      if (c < g.C) atomicAdd(bstats + ((size_t)(blockIdx.x % TRT_STAT_REPLICAS) * 2 + k) * g.C + c, (double)total);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward weight
// dW[c][kh][kw] += sum_{n,oy,ox} dD[n,oy,ox,c] * act(x)[n, oy*S-pad_t+kh, ox*S-pad_l+kw, c]
// thread = (channel lane, filter row kh, output-row subset); the filter row slides over an input row kept in registers.
// The gradient does not depend on the tile position, so the accumulators live in registers across ALL the block's items.
// Tile height of the weight-gradient kernel: NPT / K row lanes share a tile's rows (K = 3: 10 lanes, K = 5: 6), so stride-1
// tiles are 10 / 12 rows tall - every pass over the lanes is full (with the 8-row tile of the other kernels the K = 5 lanes ran
// 6 + 2 and the K = 3 lanes 8 of 10).  Stride 2 keeps 8 rows: a taller input tile would not leave room for two blocks per SM.
__host__ __device__ constexpr int wgrad_tile_h(int K, int S) { return S == 1 ? (K == 5 ? 12 : 10) : TOH; }

template <int K, int S, int WH>
__global__ void __launch_bounds__(TPB, DW_MINB) dwconv_bwd_weight_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                   const __grid_constant__ CUtensorMap tm_d,
                                                                   const float* __restrict__ in_rec, float* __restrict__ dw,
                                                                   const DwGeom g) {
  constexpr int TOW = 8;
  constexpr int IH = (WH - 1) * S + K, IW = (TOW - 1) * S + K;
  constexpr int X_BYTES = IH * IW * PX_BYTES, D_BYTES = WH * TOW * PX_BYTES, STAGE = X_BYTES + D_BYTES;
  constexpr int SUBS = NPT / K;                 // row subsets per filter row (K=3: 10, K=5: 6)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = TRT_ALIGNED_SMEM(smem_raw, 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  float* s_acc = reinterpret_cast<float*>(smem);                       // [SUBS][K*K][64], overlays the tiles after the loop
  const int V = g.C / 8, cb = blockIdx.y;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  const int kh = pt % K, sub = pt / K;
  const bool worker = sub < SUBS;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_d);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int tiles = g.tiles_x * g.tiles_y, items = g.N * tiles, G = gridDim.x;
  auto issue = [&](int it, int st) {
    const Item q = decode_item(it, tiles, g);
    ptx::mbar_expect_tx(&bar[st], STAGE);
    ptx::tma_load_4d(smem + st * STAGE, &tm_x, &bar[st], cb * 64, q.tx * TOW * S - g.pad_l, q.ty * WH * S - g.pad_t, q.n);
    ptx::tma_load_4d(smem + st * STAGE + X_BYTES, &tm_d, &bar[st], cb * 64, q.tx * TOW, q.ty * WH, q.n);
  };
  f8 sc, sh;
  if (in_rec && cvalid) { sc = ldf8(in_rec + 8 * cv); sh = ldf8(in_rec + g.C + 8 * cv); }
  float acc[K][8];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
  int it = blockIdx.x, st = 0;
  uint32_t phase = 0;
  if (threadIdx.x == 0 && it < items) issue(it, 0);
  for (; it < items; it += G, st ^= 1) {
    const Item q = decode_item(it, tiles, g);
    if (threadIdx.x == 0 && it + G < items) {
      ptx::fence_proxy_async();
      issue(it + G, st ^ 1);
    }
    uint4* s_in = reinterpret_cast<uint4*>(smem + st * STAGE);
    const uint4* s_d = reinterpret_cast<const uint4*>(smem + st * STAGE + X_BYTES);
    ptx::mbar_wait(&bar[st], (phase >> st) & 1u);
    phase ^= 1u << st;
    if (in_rec) {
      if (cvalid) activate_tile<IH, IW>(s_in, sc, sh, q.ty * WH * S - g.pad_t, q.tx * TOW * S - g.pad_l, g.H, g.W, lane, pt);
      __syncthreads();
    }
    if (worker) {
      // only the tile's rows inside the map: a short bottom tile needs fewer passes of the row lanes (the rows past the
      // edge are zero-filled dD: correct, but up to a third of the kernel's time on the small maps)
      const int rows_here = min(WH, g.OH - q.ty * WH);
      for (int oy = sub; oy < rows_here; oy += SUBS) {
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
    __syncthreads();
  }
  // block reduction over the row lanes: every lane parks its K x 8 partials ([sub][tap][64 channels], over the tiles - all
  // loads have landed), then thread i sums the SUBS entries of (tap, channel) i - consecutive threads read consecutive floats.
  // (fp32 atomicAdd on shared memory compiles to an ATOMS.CAST.SPIN loop: 24 / 40 of them per thread, SUBS-way contended,
  // were the tail of every block - a fixed cost that weighed most on the 7x7 / 14x14 layers with ~6 tiles per block.)
  static_assert((size_t)SUBS * K * K * 64 * sizeof(float) <= 2 * (size_t)STAGE, "partials must fit over the tile buffers");
  if (worker) {
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      float4* dst = reinterpret_cast<float4*>(s_acc + ((size_t)(sub * K + kh) * K + kw) * 64 + lane * 8);
      dst[0] = make_float4(acc[kw][0], acc[kw][1], acc[kw][2], acc[kw][3]);
      dst[1] = make_float4(acc[kw][4], acc[kw][5], acc[kw][6], acc[kw][7]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * 64; i += TPB) {
    float total = 0.f;
#pragma unroll
    for (int sb = 0; sb < SUBS; ++sb) total += s_acc[sb * K * K * 64 + i];
    const int tap = i / 64, c = cb * 64 + (i % 64);
    if (c < g.C) atomicAdd(dw + (size_t)c * K * K + tap, total);
  }
}

inline uint32_t magic_div(int d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); }
inline void set_magic(DwGeom& g) {     // exact while item_index * divisor < 2^32 (asserted by the entry points)
  g.magic_tiles = magic_div(g.tiles_x * g.tiles_y);
  g.magic_tx = magic_div(g.tiles_x);
}

inline void same_pad(int i, int k, int s, int& out, int& pad_before) {
  out = (i + s - 1) / s;
  int total = (out - 1) * s + k - i;
  if (total < 0) total = 0;
  pad_before = total / 2;
}

// persistent grid: every block resident at once (one wave), a fixed channel group per blockIdx.y
struct OccEntry { const void* fn; int occ; };
OccEntry g_occ[32];
int g_occ_n = 0;

template <typename Kern>
int persistent_blocks(Kern kern, size_t smem, int cblocks, int items, int* out) {
  const void* key = reinterpret_cast<const void*>(kern);     // smem is a compile-time function of the instantiation
  int occ = -1;
  for (int i = 0; i < g_occ_n; ++i)
    if (g_occ[i].fn == key) occ = g_occ[i].occ;
  TRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: every call
  if (occ < 0) {
    TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, smem));
    if (occ < 1) return trt_set_error(TRT_ERR_CUDA, "depthwise kernel does not fit on an SM (smem %zu)", smem);
    if (g_occ_n < 32) { g_occ[g_occ_n].fn = key; g_occ[g_occ_n].occ = occ; ++g_occ_n; }
  }
  int G = (trt_num_sms() * occ) / cblocks;
  if (G < 1) G = 1;
  if (G > items) G = items;
  *out = G;
  return TRT_OK;
}

}  // namespace

extern "C" int trt_dwconv_fwd(const void* x, const float* in_rec, const float* w, void* out, const float* out_rec,
                              float* pooled_sum, int pooled_zeroed, double* stats, const trt_bn_fin_t* fin_host, void* act_out,
                              int N, int H, int W, int C, int k, int s, cudaStream_t stream) {
  TRT_REQUIRE(!act_out || (s == 1 && in_rec), "trt_dwconv_fwd: act_out needs stride 1 and an input BatchNorm record");
  TRT_REQUIRE(!fin_host || (in_rec && fin_host->stats && fin_host->gamma && fin_host->beta && fin_host->rec == in_rec && fin_host->count > 0),
              "trt_dwconv_fwd: incomplete lazy BatchNorm record for the input (rec must be in_rec)");
  trt_bn_fin_t fin = {};
  if (fin_host) fin = *fin_host;
  const int has_fin = fin_host ? 1 : 0;
  TRT_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "trt_dwconv_fwd: bad argument");
  TRT_REQUIRE((k == 3 || k == 5) && (s == 1 || s == 2), "trt_dwconv_fwd: only k in {3,5}, s in {1,2}");
  DwGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.S = s;
  same_pad(H, k, s, g.OH, g.pad_t);
  same_pad(W, k, s, g.OW, g.pad_l);
  const int p = (s == 1 && g.OW > 8) ? 4 : 2;        // outputs per thread along x: tile width 4*p
  g.tiles_x = (g.OW + 4 * p - 1) / (4 * p);
  g.tiles_y = (g.OH + TOH - 1) / TOH;
  set_magic(g);
  if (pooled_sum && !pooled_zeroed) TRT_CUDA(cudaMemsetAsync(pooled_sum, 0, (size_t)N * C * sizeof(float), stream));
  const int cblocks = (C / 8 + CL - 1) / CL;
  const int items = N * g.tiles_x * g.tiles_y;
  TRT_REQUIRE((long long)items * g.tiles_x * g.tiles_y < (1ll << 32), "trt_dwconv_fwd: too many tiles");
  int rc;
#define LAUNCH_DW(KK, SS, PP)                                                                                      \
  do {                                                                                                             \
    using T = FwdTile<KK, SS, PP>;                                                                                 \
    const size_t smem = 2 * T::IN_BYTES + KK * KK * PX_BYTES + 16 + 128;                                           \
    CUtensorMap tm;                                                                                                \
    if ((rc = trt_make_tmap_nhwc(&tm, x, N, H, W, C, 64, T::IW, T::IH))) return rc;                                \
    int G;                                                                                                         \
    if ((rc = persistent_blocks(dwconv_fwd_kernel<KK, SS, PP>, smem, cblocks, items, &G))) return rc;              \
    TRT_CUDA(trt_launch(dwconv_fwd_kernel<KK, SS, PP>, dim3(G, cblocks), dim3(TPB), smem, stream, tm, in_rec, w, (uint4*)out, out_rec, pooled_sum, stats, g, has_fin, fin, (uint4*)act_out)); \
  } while (0)
  if (k == 3 && s == 1) { if (p == 4) LAUNCH_DW(3, 1, 4); else LAUNCH_DW(3, 1, 2); }
  else if (k == 5 && s == 1) { if (p == 4) LAUNCH_DW(5, 1, 4); else LAUNCH_DW(5, 1, 2); }
  else if (k == 3) LAUNCH_DW(3, 2, 2);
  else LAUNCH_DW(5, 2, 2);
#undef LAUNCH_DW
  return trt_check_launch("trt_dwconv_fwd");
}

extern "C" int trt_dwconv_bwd(const void* gy, const float* w, const void* x_raw, const float* x_rec, void* g_out,
                              double* bstats, float* dw, int N, int H, int W, int C, int k, int s, cudaStream_t stream) {
  TRT_REQUIRE(gy && w && x_raw && (dw || g_out) && N > 0 && C > 0 && C % 8 == 0, "trt_dwconv_bwd: bad argument");
  TRT_REQUIRE((k == 3 || k == 5) && (s == 1 || s == 2), "trt_dwconv_bwd: only k in {3,5}, s in {1,2}");
  DwGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.S = s;
  same_pad(H, k, s, g.OH, g.pad_t);
  same_pad(W, k, s, g.OW, g.pad_l);
  const int cblocks = (C / 8 + CL - 1) / CL;
  int rc;
  if (g_out) {   // data gradient (skipped for a first layer whose input needs no gradient)
    if (s == 1) {
      const int p = W <= 8 ? 2 : 4;
      g.tiles_x = (W + 4 * p - 1) / (4 * p);
      g.tiles_y = (H + TOH - 1) / TOH;
      set_magic(g);
      const int items = N * g.tiles_x * g.tiles_y;
#define LAUNCH_BD_S1(KK, PP)                                                                                       \
  do {                                                                                                             \
    using T = FwdTile<KK, 1, PP>;                                                                                  \
    const size_t smem = 2 * (T::IN_BYTES + T::OUT_BYTES) + KK * KK * PX_BYTES + 16 + 128;                          \
    CUtensorMap td, tx;                                                                                            \
    if ((rc = trt_make_tmap_nhwc(&td, gy, N, g.OH, g.OW, C, 64, T::IW, T::IH))) return rc;                         \
    if ((rc = trt_make_tmap_nhwc(&tx, x_raw, N, H, W, C, 64, T::TOW, TOH))) return rc;                             \
    int G;                                                                                                         \
    if ((rc = persistent_blocks(dwconv_bwd_data_s1_kernel<KK, PP>, smem, cblocks, items, &G))) return rc;          \
    dwconv_bwd_data_s1_kernel<KK, PP><<<dim3(G, cblocks), TPB, smem, stream>>>(td, tx, w, x_rec, (uint4*)g_out, bstats, g); \
  } while (0)
      if (k == 3) { if (p == 4) LAUNCH_BD_S1(3, 4); else LAUNCH_BD_S1(3, 2); }
      else { if (p == 4) LAUNCH_BD_S1(5, 4); else LAUNCH_BD_S1(5, 2); }
#undef LAUNCH_BD_S1
    } else {
      g.tiles_x = (W + 15) / 16;
      g.tiles_y = (H + 15) / 16;
      set_magic(g);
      const int items = N * g.tiles_x * g.tiles_y;
#define LAUNCH_BD_S2(KK)                                                                                           \
  do {                                                                                                             \
    using T = S2Tile<KK>;                                                                                          \
    const size_t smem = 2 * T::STAGE + KK * KK * PX_BYTES + 16 + 128;                                              \
    CUtensorMap td, tx;                                                                                            \
    if ((rc = trt_make_tmap_nhwc(&td, gy, N, g.OH, g.OW, C, 64, T::DW, T::DH))) return rc;                         \
    if ((rc = trt_make_tmap_nhwc(&tx, x_raw, N, H, W, C, 64, T::TI, T::TI))) return rc;                            \
    int G;                                                                                                         \
    if ((rc = persistent_blocks(dwconv_bwd_data_s2_kernel<KK>, smem, cblocks, items, &G))) return rc;              \
    dwconv_bwd_data_s2_kernel<KK><<<dim3(G, cblocks), TPB, smem, stream>>>(td, tx, w, x_rec, (uint4*)g_out, bstats, g); \
  } while (0)
      if (k == 3) LAUNCH_BD_S2(3); else LAUNCH_BD_S2(5);
#undef LAUNCH_BD_S2
    }
    rc = trt_check_launch("trt_dwconv_bwd(data)");
    if (rc) return rc;
  }
  if (!dw) return TRT_OK;            // data gradient only (the weight gradient may run as its own call on another stream)
  {
    // tall tiles only where they save passes (not on the 14x14 / 7x7 maps: same pass count, larger boxes);
    // TEETHRT_DW_WGRAD_TALL=0: the 8-row tile everywhere (A/B)
    static const int tall_on = [] { const char* e = getenv("TEETHRT_DW_WGRAD_TALL"); return (e && *e == '0') ? 0 : 1; }();
    auto lane_passes = [&](int th) {          // passes of the NPT / k row lanes over one column of tiles
      const int subs = NPT / k;
      int n = 0;
      for (int y = 0; y < g.OH; y += th) n += ((g.OH - y < th ? g.OH - y : th) + subs - 1) / subs;
      return n;
    };
    const int wh = (tall_on && lane_passes(wgrad_tile_h(k, s)) < lane_passes(TOH)) ? wgrad_tile_h(k, s) : TOH;
    g.tiles_x = (g.OW + 7) / 8;
    g.tiles_y = (g.OH + wh - 1) / wh;
    set_magic(g);
    const int items = N * g.tiles_x * g.tiles_y;
    TRT_REQUIRE((long long)items * g.tiles_x * g.tiles_y < (1ll << 32), "trt_dwconv_bwd: too many tiles");
#define LAUNCH_BW(KK, SS, WH)                                                                                      \
  do {                                                                                                             \
    const int IH = (WH - 1) * SS + KK, IW = 7 * SS + KK;                                                           \
    const size_t smem = 2 * ((size_t)IH * IW + WH * 8) * PX_BYTES + 16 + 128;                                      \
    CUtensorMap tx, td;                                                                                            \
    if ((rc = trt_make_tmap_nhwc(&tx, x_raw, N, H, W, C, 64, IW, IH))) return rc;                                  \
    if ((rc = trt_make_tmap_nhwc(&td, gy, N, g.OH, g.OW, C, 64, 8, WH))) return rc;                                \
    int G;                                                                                                         \
    if ((rc = persistent_blocks(dwconv_bwd_weight_kernel<KK, SS, WH>, smem, cblocks, items, &G))) return rc;       \
    dwconv_bwd_weight_kernel<KK, SS, WH><<<dim3(G, cblocks), TPB, smem, stream>>>(tx, td, x_rec, dw, g);           \
  } while (0)
    if (k == 3 && s == 1) { if (wh == TOH) LAUNCH_BW(3, 1, TOH); else LAUNCH_BW(3, 1, wgrad_tile_h(3, 1)); }
    else if (k == 3 && s == 2) LAUNCH_BW(3, 2, TOH);
    else if (k == 5 && s == 1) { if (wh == TOH) LAUNCH_BW(5, 1, TOH); else LAUNCH_BW(5, 1, wgrad_tile_h(5, 1)); }
    else LAUNCH_BW(5, 2, TOH);
#undef LAUNCH_BW
  }
  return trt_check_launch("trt_dwconv_bwd(weight)");
}
