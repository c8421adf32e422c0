// Channel-wise NHWC kernels around the GEMMs: BatchNorm statistics finalisation / application / backward,
// squeeze-excite (pool, MLP, gate) forward and backward, global average pooling.  All HBM-bound: bf16 activations,
// 16-byte vector accesses, one thread = one 8-channel vector of one pixel row, fp32 math, fp64 channel sums.
//
// Replaces the ATen/cuDNN launches behind timm's BatchNormAct2d / SqueezeExcite / global_pool inside
// `self.backbone(x_img)` (experiments/multimodal_v1/train_mm_joint_dualtask.py:154) and their autograd backward (:248).
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int UNR = 4;    // rows in flight per thread (memory-level parallelism of the streaming kernels)

// ---- thread mapping: rows x (C/8) vectors; a thread keeps a fixed channel vector so it can own channel accumulators
struct RowMap {
  int v;       // channel-vector index (channels 8v .. 8v+7)
  int vl, ry;  // position inside the block
  int VX, RY;
  bool active;
};
__device__ __forceinline__ RowMap row_map(int V, int VX, int RY, int slab) {
  RowMap m;
  m.VX = VX; m.RY = RY;
  m.ry = threadIdx.x / VX;
  m.vl = threadIdx.x - m.ry * VX;
  m.v = slab * VX + m.vl;
  m.active = (m.ry < RY) && (m.v < V);
  return m;
}
struct Launch { int V, slabs, VX, RY; };
inline Launch plan(int C) {
  Launch L;
  L.V = C / 8;
  L.slabs = (L.V + TPB - 1) / TPB;
  L.VX = (L.V + L.slabs - 1) / L.slabs;
  L.RY = TPB / L.VX;
  return L;
}
// Per-image reductions (SE squeeze, SE backward sums) on SMALL feature maps: a block takes a narrow channel slab and ALL rows of
// one image in a single pass (RY row-lanes x UNR rows in flight >= HW), so the per-(image, channel) results are plain stores
// - no atomics, no zeroed target, no serial row loop.  Measured need (profiles/r02_step_per_op.txt): with full-width slabs
// the 7x7 / 14x14 layers spent 15-28 us on 10-50 MB tensors, most of it in fp32 atomics (five sums x every row block).
// Large maps keep full-width slabs with several row blocks per image.
inline Launch plan_img(int C, int HW, bool auto_narrow = false) {
  Launch L = plan(C);
  const int lanes = (HW + UNR - 1) / UNR;
  // TEETHRT_NARROW_SLABS=1 / 0 forces the narrow geometry on / off wherever it applies.  Unset: only where the isolated
  // per-layer timings (tools/elt_probe.py, HBM-cold, inside a CUDA graph) show a win - the SE squeeze on 7x7 maps (6.6 -> 4.9,
  // 9.6 -> 6.9, 12.5 -> 9.7 us at 960 / 1632 / 2688 channels); on 14x14 maps and for the five-sum backward pass it is a wash
  // or a loss (more blocks, each with its own reduction tail).
  const char* on = getenv("TEETHRT_NARROW_SLABS");
  const bool want = (on && *on) ? (*on == '1') : (auto_narrow && HW <= 64);
  if (lanes > TPB / 4 || !want) return L;
  int ry = 1;
  while (ry < lanes) ry <<= 1;
  const int vx = TPB / ry;                 // 4 .. 256, a power of two
  if (vx >= L.V) return L;                 // the whole channel range fits one narrow slab anyway
  L.VX = vx;
  L.RY = ry;
  L.slabs = (L.V + vx - 1) / vx;
  return L;
}
// grid-size multipliers (blocks per SM the per-image kernels aim for); environment overrides are bring-up knobs.  The defaults
// are the minima of the isolated sweeps (tools/gpu_elt_sweep.sh -> profiles/r02_elt_sweep.txt): 4 blocks per SM = one full
// wave of these 64-register kernels; 6 (1.4 waves) cost 17 / 25 / 40 / 100 us per step more for the squeeze, the gate, the
// activation backward and the five-sum pass.
inline int env_mult(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
inline bool one_pass(const Launch& L, int HW) { return (size_t)L.RY * UNR >= (size_t)HW; }

inline int row_blocks(int rows, int RY, int slabs, int target_blocks) {
  int nb = (rows + RY - 1) / RY;
  int cap = target_blocks / slabs;
  if (cap < 1) cap = 1;
  return nb < cap ? nb : cap;
}

// reduce NV per-thread values (per channel of the thread's vector) over the block's RY row-lanes; result valid for ry==0
template <int NV>
__device__ __forceinline__ void block_reduce_rows(float (&acc)[NV], float* s_red, const RowMap& m) {
  __syncthreads();
  if (m.ry < m.RY) {
#pragma unroll
    for (int i = 0; i < NV; ++i) s_red[(size_t)i * TPB + threadIdx.x] = acc[i];
  }
  __syncthreads();
  if (m.ry == 0 && m.active) {
    for (int r = 1; r < m.RY; ++r) {
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += s_red[(size_t)i * TPB + r * m.VX + m.vl];
    }
  }
}

// The same for 8 values when the block was planned by plan_img (VX a power of two <= 32: a warp holds 32 / VX whole rows):
// rows inside a warp meet in shuffles, the 8 warps meet in 8 KB of shared memory - two barriers instead of an RY-long walk.
// Falls back to block_reduce_rows for full-width slabs.  Result valid for ry == 0.
__device__ __forceinline__ void reduce_rows8(float (&acc)[8], float* s_red, const RowMap& m) {
  if (m.VX > 32 || (m.VX & (m.VX - 1)) != 0) {
    block_reduce_rows<8>(acc, s_red, m);
    return;
  }
  for (int o = 16; o >= m.VX; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane < m.VX) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[(warp * m.VX + lane) * 8 + i] = acc[i];
  }
  __syncthreads();
  if (m.ry == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
      for (int w = 0; w < TPB / 32; ++w) t += s_red[(w * m.VX + m.vl) * 8 + i];
      acc[i] = t;
    }
  }
}

__device__ __forceinline__ float act_apply(float x, int act) { return act ? siluf_(x) : x; }

// ------------------------------------------------------------------------------------------------ BN finalize
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   long long* nbt, float* __restrict__ rec, int C, double count, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  bn_publish(bn_channel(stats, gamma, beta, C, c, count, eps), rec, running_mean, running_var, C, c, count, momentum);
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ rm, const float* __restrict__ rv, float* __restrict__ rec, int C,
                               float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rstd = 1.0f / sqrtf(rv[c] + eps);
  const float sc = gamma[c] * rstd;
  rec[c] = sc;
  rec[C + c] = beta[c] - rm[c] * sc;
  rec[2 * C + c] = rm[c];
  rec[3 * C + c] = rstd;
}

// coef[0]=a, [1]=b, [2]=c with dx = a*dy + b*x + c ; also dgamma / dbeta (written, not accumulated)
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ bstats, const float* __restrict__ rec,
                                       const float* __restrict__ gamma, float* __restrict__ coef, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int C, double count, int raw_x) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const BnBwdChannel b = bn_bwd_channel(bstats, rec, gamma, C, c, count, raw_x);
  coef[c] = b.a;
  coef[C + c] = b.b;
  coef[2 * C + c] = b.c;
  dgamma[c] = b.dgamma;
  dbeta[c] = b.dbeta;
}

// shared-memory scale/shift of a block's channel slab (VX vectors = up to TPB*8 channels): filled by bn_lazy_block when the
// BatchNorm is lazy, otherwise the record is read directly
constexpr int SLAB_CH = TPB * 8;

// ------------------------------------------------------------------------------------------------ y = act(bn(x)) (+res)
__global__ void __launch_bounds__(TPB) bn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ rec,
                                                       const uint4* __restrict__ res, uint4* __restrict__ out, int rows,
                                                       int C, int V, int VX, int RY, int act, const int has_fin,
                                                       const trt_bn_fin_t fin) {
  __shared__ __align__(16) float s_ss[2][SLAB_CH];
  const RowMap m = row_map(V, VX, RY, blockIdx.y);
  if (has_fin) bn_lazy_block(fin, C, blockIdx.y * VX * 8, VX * 8, s_ss[0], s_ss[1], blockIdx.x == 0);
  if (!m.active) return;
  f8 sc, sh;
  if (has_fin) { sc = lds8(s_ss[0] + 8 * m.vl); sh = lds8(s_ss[1] + 8 * m.vl); }
  else { sc = ldf8(rec + 8 * m.v); sh = ldf8(rec + C + 8 * m.v); }
  const int step = gridDim.x * RY;
  for (int r = blockIdx.x * RY + m.ry; r < rows; r += UNR * step) {
    uint4 xa[UNR], xr[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int rr = r + u * step;
      if (rr < rows) {
        xa[u] = __ldg(x + (size_t)rr * V + m.v);
        if (res) xr[u] = __ldg(res + (size_t)rr * V + m.v);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int rr = r + u * step;
      if (rr < rows) {
        f8 a = unpack8(xa[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) a.v[i] = act_apply(fmaf(a.v[i], sc.v[i], sh.v[i]), act);
        if (res) {
          const f8 b = unpack8(xr[u]);
#pragma unroll
          for (int i = 0; i < 8; ++i) a.v[i] += b.v[i];
        }
        out[(size_t)rr * V + m.v] = pack8(a);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ pooled[n,c] += sum_hw act(bn(x))
__global__ void __launch_bounds__(TPB) pool_act_kernel(const uint4* __restrict__ x, const float* __restrict__ rec,
                                                       float* __restrict__ pooled, int HW, int C, int V, int VX, int RY,
                                                       int act, const int has_fin, const trt_bn_fin_t fin) {
  __shared__ float s_red[8 * TPB];
  __shared__ __align__(16) float s_ss[2][SLAB_CH];
  const RowMap m = row_map(V, VX, RY, blockIdx.z);
  const int n = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  if (has_fin) bn_lazy_block(fin, C, blockIdx.z * VX * 8, VX * 8, s_ss[0], s_ss[1], blockIdx.x == 0 && blockIdx.y == 0);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    f8 sc, sh;
    if (has_fin) { sc = lds8(s_ss[0] + 8 * m.vl); sh = lds8(s_ss[1] + 8 * m.vl); }
    else if (rec) { sc = ldf8(rec + 8 * m.v); sh = ldf8(rec + C + 8 * m.v); }
    const int step = gridDim.x * RY;
    for (int r = blockIdx.x * RY + m.ry; r < HW; r += UNR * step) {
      uint4 xa[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (r + u * step < HW) xa[u] = __ldg(x + ((size_t)n * HW + r + u * step) * V + m.v);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < HW) {
          const f8 a = unpack8(xa[u]);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += rec ? act_apply(fmaf(a.v[i], sc.v[i], sh.v[i]), act) : a.v[i];
        }
      }
    }
  }
  reduce_rows8(acc, s_red, m);
  if (m.ry == 0 && m.active) {
    float* o = pooled + (size_t)n * C + 8 * m.v;
    if (gridDim.x == 1) {                  // this block saw every row of the image: plain stores
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = acc[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(o + i, acc[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ SE MLP forward
// Two batched micro-GEMMs over the whole batch so every weight is read once:
//   (1) s1[n,r] = br[r] + <mean[n,:], Wr[r,:]>   grid = rd blocks, one warp per image, lanes over channels
//       (eight rows of Wr per block - an eighth of the L2 reads of the pooled means - measured SLOWER: 17.9 vs 11.4 us at
//       1632 channels, 27.9 vs 23.2 at 2688; the launch is latency-bound and wants the 8x more blocks)
//   (2) gate[n,c] = sigmoid(be[c] + <silu(s1[n,:]), We[c,:]>)   grid = (C/64, image splits), We chunk + silu(s1) in smem
__global__ void __launch_bounds__(TPB) se_reduce_kernel(const float* __restrict__ pooled, float inv_hw,
                                                        const float* __restrict__ Wr, const float* __restrict__ br,
                                                        float* __restrict__ s1, int N, int C, int rd) {
  extern __shared__ float s_mem[];   // [C] one row of Wr, pre-scaled by 1/HW
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  for (int c = threadIdx.x; c < C; c += TPB) s_mem[c] = __ldg(Wr + (size_t)r * C + c) * inv_hw;
  __syncthreads();
  const float bias = br[r];
  for (int n = blockIdx.y * (TPB / 32) + warp; n < N; n += gridDim.y * (TPB / 32)) {
    float acc = 0.f;
    const float* p = pooled + (size_t)n * C;
    for (int c = lane; c < C; c += 128) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (c + 32 * u < C) ? __ldg(p + c + 32 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + 32 * u < C) acc = fmaf(v[u], s_mem[c + 32 * u], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s1[(size_t)n * rd + r] = acc + bias;
  }
}

constexpr int SE_CC = 64;   // channels per block in the expand kernels
__global__ void __launch_bounds__(TPB) se_expand_kernel(const float* __restrict__ s1, const float* __restrict__ We,
                                                        const float* __restrict__ be, float* __restrict__ gate, int N, int C,
                                                        int rd, int n_per_block, uint4* __restrict__ apply_x, int HW) {
  extern __shared__ float s_mem[];
  float* s_we = s_mem;                         // [SE_CC][rd+1]
  float* s_a1 = s_mem + SE_CC * (rd + 1);      // [n_per_block][rd]
  float* s_g = s_a1 + n_per_block * rd;        // [n_per_block][SE_CC] gates of this block (only with apply_x)
  const int c0 = blockIdx.x * SE_CC, n0 = blockIdx.y * n_per_block;
  const int nn = min(n_per_block, N - n0);
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < SE_CC * rd; i += TPB) {
    const int cl = i / rd, r = i - cl * rd;
    s_we[cl * (rd + 1) + r] = (c0 + cl < C) ? __ldg(We + (size_t)(c0 + cl) * rd + r) : 0.f;
  }
  for (int i = threadIdx.x; i < nn * rd; i += TPB) s_a1[i] = siluf_(s1[(size_t)n0 * rd + i]);
  __syncthreads();
  const int cl = threadIdx.x % SE_CC, nsub = threadIdx.x / SE_CC;
  const int c = c0 + cl;
  if (c < C) {
    const float bias = be[c];
    const float* wrow = s_we + cl * (rd + 1);
    // four images per pass: one dependent FMA chain per image was latency-bound (rd steps x 4 cycles, one after the other)
    constexpr int NS = TPB / SE_CC;
    for (int nb = nsub; nb < nn; nb += 4 * NS) {
      float acc[4] = {bias, bias, bias, bias};
      const float* a0 = s_a1 + nb * rd;
      const float* a1 = s_a1 + min(nb + NS, nn - 1) * rd;
      const float* a2 = s_a1 + min(nb + 2 * NS, nn - 1) * rd;
      const float* a3 = s_a1 + min(nb + 3 * NS, nn - 1) * rd;
      for (int r = 0; r < rd; ++r) {
        const float w = wrow[r];
        acc[0] = fmaf(a0[r], w, acc[0]);
        acc[1] = fmaf(a1[r], w, acc[1]);
        acc[2] = fmaf(a2[r], w, acc[2]);
        acc[3] = fmaf(a3[r], w, acc[3]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = nb + u * NS;
        if (n < nn) {
          const float g = sigmoidf_(acc[u]);
          gate[(size_t)(n0 + n) * C + c] = g;
          if (apply_x) s_g[n * SE_CC + cl] = g;
        }
      }
    }
  }
  if (apply_x == nullptr) return;
  // Small feature maps at inference (a few images x <= 196 pixels): this block also applies its 64 gates to the activation in
  // place, x[n, hw, c0..c0+63] *= gate - the separate gate_apply launch is pure latency there.  One thread = one 16-byte
  // vector (8 channels) of one pixel row.
  __syncthreads();                                     // the block's gates are staged
  const int vec = SE_CC / 8, V = C / 8;
  for (int i = threadIdx.x; i < nn * HW * vec; i += TPB) {
    const int v = i % vec, row = i / vec;              // row = n_local * HW + hw
    const int n = row / HW;
    if (c0 / 8 + v >= V) continue;
    uint4* px = apply_x + ((size_t)(n0 * HW + row)) * V + c0 / 8 + v;
    f8 a = unpack8(*px);
    const float* g = s_g + n * SE_CC + v * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) a.v[k] *= g[k];
    *px = pack8(a);
  }
}

// SE MLP forward for an inference batch (N <= SE_SMALL_N images: batch 1, three TTA flips).  The batched kernels above give
// one WARP to an image and 64 threads to a channel chunk, so at one image 7 of 8 warps idle and every dot product is one long
// dependent chain: 8.7 us per block inside the captured batch-1 forward (tools/knockout.py --infer 1: 0.28 ms of 1.09 ms).
// Here the whole block works on every dot product: the reduce kernel splits the channels over its 256 threads, the expand
// kernel gives each output channel 8 lanes over the reduced dimension.
constexpr int SE_SMALL_N = 4;
__global__ void __launch_bounds__(TPB) se_reduce_small_kernel(const float* __restrict__ pooled, float inv_hw,
                                                              const float* __restrict__ Wr, const float* __restrict__ br,
                                                              float* __restrict__ s1, int N, int C, int rd) {
  __shared__ float s_part[SE_SMALL_N][TPB / 32];
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  float acc[SE_SMALL_N];
#pragma unroll
  for (int n = 0; n < SE_SMALL_N; ++n) acc[n] = 0.f;
  for (int c = threadIdx.x; c < C; c += TPB) {
    const float w = __ldg(Wr + (size_t)r * C + c);
#pragma unroll
    for (int n = 0; n < SE_SMALL_N; ++n)
      if (n < N) acc[n] = fmaf(__ldg(pooled + (size_t)n * C + c), w, acc[n]);
  }
#pragma unroll
  for (int n = 0; n < SE_SMALL_N; ++n) {
    acc[n] = warp_sum(acc[n]);
    if (lane == 0) s_part[n][warp] = acc[n];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < TPB / 32; ++w) t += s_part[threadIdx.x][w];
    s1[(size_t)threadIdx.x * rd + r] = fmaf(t, inv_hw, __ldg(br + r));
  }
}

constexpr int SE_SCC = 32;    // channels per block of the small-batch expand kernel (8 lanes per channel)
__global__ void __launch_bounds__(TPB) se_expand_small_kernel(const float* __restrict__ s1, const float* __restrict__ We,
                                                              const float* __restrict__ be, float* __restrict__ gate, int N,
                                                              int C, int rd) {
  extern __shared__ float s_mem[];
  float* s_we = s_mem;                          // [SE_SCC][rd + 1]
  float* s_a1 = s_mem + SE_SCC * (rd + 1);      // [SE_SMALL_N][rd]
  const int c0 = blockIdx.x * SE_SCC;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < SE_SCC * rd; i += TPB) {       // the weights do not depend on the previous kernel
    const int cl = i / rd, r = i - cl * rd;
    s_we[cl * (rd + 1) + r] = (c0 + cl < C) ? __ldg(We + (size_t)(c0 + cl) * rd + r) : 0.f;
  }
  pdl_wait();
  for (int i = threadIdx.x; i < N * rd; i += TPB) s_a1[i] = siluf_(s1[i]);
  __syncthreads();
  const int cl = threadIdx.x >> 3, q = threadIdx.x & 7;        // 8 consecutive lanes share an output channel
  float acc[SE_SMALL_N];
#pragma unroll
  for (int n = 0; n < SE_SMALL_N; ++n) acc[n] = 0.f;
  const float* wrow = s_we + cl * (rd + 1);
  for (int r = q; r < rd; r += 8) {
    const float w = wrow[r];
#pragma unroll
    for (int n = 0; n < SE_SMALL_N; ++n)
      if (n < N) acc[n] = fmaf(s_a1[n * rd + r], w, acc[n]);
  }
#pragma unroll
  for (int n = 0; n < SE_SMALL_N; ++n) {
    acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 4);
    acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 2);
    acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 1);
  }
  const int c = c0 + cl;
  if (q < N && c < C) {
    float mine = acc[0];
#pragma unroll
    for (int n = 1; n < SE_SMALL_N; ++n) mine = q == n ? acc[n] : mine;
    gate[(size_t)q * C + c] = sigmoidf_(mine + __ldg(be + c));
  }
}

// ------------------------------------------------------------------------------------------------ y = act(bn(x)) * gate[n,c]
__global__ void __launch_bounds__(TPB) gate_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ rec,
                                                         const float* __restrict__ gate, uint4* __restrict__ out, int HW,
                                                         int C, int V, int VX, int RY) {
  const RowMap m = row_map(V, VX, RY, blockIdx.z);
  pdl_launch_dependents();
  pdl_wait();
  if (!m.active) return;
  const int n = blockIdx.y;
  f8 sc, sh;
  if (rec) { sc = ldf8(rec + 8 * m.v); sh = ldf8(rec + C + 8 * m.v); }
  const f8 g = ldf8(gate + (size_t)n * C + 8 * m.v);
  const int step = gridDim.x * RY;
  for (int r = blockIdx.x * RY + m.ry; r < HW; r += UNR * step) {
    uint4 xa[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (r + u * step < HW) xa[u] = __ldg(x + ((size_t)n * HW + r + u * step) * V + m.v);
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (r + u * step < HW) {
        f8 a = unpack8(xa[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) a.v[i] = (rec ? siluf_(fmaf(a.v[i], sc.v[i], sh.v[i])) : a.v[i]) * g.v[i];
        out[((size_t)n * HW + r + u * step) * V + m.v] = pack8(a);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ BN backward: sums
// bstats[0][c] += sum dy ; bstats[1][c] += sum dy * xhat
__global__ void __launch_bounds__(TPB) bn_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                            const float* __restrict__ rec, double* __restrict__ bstats,
                                                            int rows, int C, int V, int VX, int RY) {
  __shared__ float s_red[16 * TPB];
  const RowMap m = row_map(V, VX, RY, blockIdx.y);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (m.active) {
    const int step = gridDim.x * RY;
    for (int r = blockIdx.x * RY + m.ry; r < rows; r += UNR * step) {
      uint4 da[UNR], xa[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < rows) {
          da[u] = __ldg(dy + (size_t)(r + u * step) * V + m.v);
          xa[u] = __ldg(x + (size_t)(r + u * step) * V + m.v);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < rows) {
          const f8 d = unpack8(da[u]), a = unpack8(xa[u]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i] += d.v[i];
            acc[8 + i] = fmaf(d.v[i], a.v[i], acc[8 + i]);          // sum dy*x, turned into sum dy*xhat below
          }
        }
      }
    }
    const f8 mu = ldf8(rec + 2 * C + 8 * m.v), rs = ldf8(rec + 3 * C + 8 * m.v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[8 + i] = rs.v[i] * (acc[8 + i] - mu.v[i] * acc[i]);
  }
  block_reduce_rows<16>(acc, s_red, m);
  if (m.ry == 0 && m.active) {
    double* rep = bstats + (size_t)((blockIdx.x + blockIdx.y * gridDim.x) % TRT_STAT_REPLICAS) * 2 * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(rep + 8 * m.v + i, (double)acc[i]);
      atomicAdd(rep + C + 8 * m.v + i, (double)acc[8 + i]);
    }
  }
}

// out = a*dy + b*x + c (per channel)   [BN backward apply]
__global__ void __launch_bounds__(TPB) affine2_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                      const float* __restrict__ coef, uint4* __restrict__ out, int rows,
                                                      int C, int V, int VX, int RY, const int has_fin,
                                                      const trt_bn_bwd_fin_t fin) {
  __shared__ __align__(16) float s_abc[3][SLAB_CH];
  const RowMap m = row_map(V, VX, RY, blockIdx.y);
  if (has_fin) {
    // lazy BatchNorm backward: the producer of `dy` only accumulated {sum dy, sum dy*xhat}; every block turns them into the
    // coefficients of its channel slab, block 0 of the slab also writes dgamma / dbeta
    const int c0 = blockIdx.y * VX * 8;
    for (int i = threadIdx.x; i < VX * 8; i += TPB) {
      const int c = c0 + i;
      if (c < C) {
        const BnBwdChannel b = bn_bwd_channel(fin.bstats, fin.rec, fin.gamma, C, c, fin.count, fin.raw_x);
        s_abc[0][i] = b.a; s_abc[1][i] = b.b; s_abc[2][i] = b.c;
        if (blockIdx.x == 0) { fin.dgamma[c] = b.dgamma; fin.dbeta[c] = b.dbeta; }
      }
    }
    __syncthreads();
  }
  if (!m.active) return;
  f8 ca, cb, cc;
  if (has_fin) { ca = lds8(s_abc[0] + 8 * m.vl); cb = lds8(s_abc[1] + 8 * m.vl); cc = lds8(s_abc[2] + 8 * m.vl); }
  else { ca = ldf8(coef + 8 * m.v); cb = ldf8(coef + C + 8 * m.v); cc = ldf8(coef + 2 * C + 8 * m.v); }
  const int step = gridDim.x * RY;
  for (int r = blockIdx.x * RY + m.ry; r < rows; r += UNR * step) {
    uint4 da[UNR], xa[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (r + u * step < rows) {
        da[u] = __ldg(dy + (size_t)(r + u * step) * V + m.v);
        xa[u] = __ldg(x + (size_t)(r + u * step) * V + m.v);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (r + u * step < rows) {
        const f8 d = unpack8(da[u]), a = unpack8(xa[u]);
        f8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = fmaf(ca.v[i], d.v[i], fmaf(cb.v[i], a.v[i], cc.v[i]));
        out[(size_t)(r + u * step) * V + m.v] = pack8(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ SE backward, pass 1
// dgate_pre[n,c] += sum_hw dA * silu(bn(x)).
// FULL: four more per-(image, channel) sums in the same pass, from which the BatchNorm-backward sums of the gated
// activation follow WITHOUT another pass over the tensor.  With z = bn(x), s' = silu'(z) and the upstream gradient of the
// BatchNorm output g = (dA*gate[n,c] + dmean[n,c]/HW) * s':
//     sum_{n,hw} g     = sum_n gate*S1 + dmean/HW*S2        S1 = sum_hw dA*s'      S2 = sum_hw s'
//     sum_{n,hw} g*x   = sum_n gate*S3 + dmean/HW*S4        S3 = sum_hw dA*s'*x    S4 = sum_hw s'*x
// (gate and dmean only exist after the SE MLP backward, which needs dgate_pre first; the old path made a second pass to
// write g and a third to apply the BatchNorm-backward affine).  sums layout: [5][N][C], slot 0 = dgate_pre.
template <bool FULL>
__global__ void __launch_bounds__(TPB) se_bwd_reduce_kernel(const uint4* __restrict__ dA, const uint4* __restrict__ x,
                                                            const float* __restrict__ rec, float* __restrict__ sums,
                                                            int HW, int C, int V, int VX, int RY, size_t NC) {
  constexpr int NS = FULL ? 5 : 1;
  __shared__ float s_red[8 * TPB];
  const RowMap m = row_map(V, VX, RY, blockIdx.z);
  const int n = blockIdx.y;
  float acc[NS][8];
#pragma unroll
  for (int k = 0; k < NS; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  if (m.active) {
    const f8 sc = ldf8(rec + 8 * m.v), sh = ldf8(rec + C + 8 * m.v);
    const int step = gridDim.x * RY;
    for (int r = blockIdx.x * RY + m.ry; r < HW; r += UNR * step) {
      uint4 da[UNR], xa[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < HW) {
          da[u] = __ldg(dA + ((size_t)n * HW + r + u * step) * V + m.v);
          xa[u] = __ldg(x + ((size_t)n * HW + r + u * step) * V + m.v);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < HW) {
          const f8 d = unpack8(da[u]), a = unpack8(xa[u]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float z = fmaf(a.v[i], sc.v[i], sh.v[i]);
            const float sg = sigmoidf_(z);
            acc[0][i] = fmaf(d.v[i], z * sg, acc[0][i]);
            if (FULL) {
              const float sp = sg * (1.0f + z * (1.0f - sg));      // silu'(z)
              const float t = d.v[i] * sp;
              acc[1][i] += t;
              acc[2][i] += sp;
              acc[3][i] = fmaf(t, a.v[i], acc[3][i]);
              acc[4][i] = fmaf(sp, a.v[i], acc[4][i]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    reduce_rows8(acc[k], s_red, m);
    if (m.ry == 0 && m.active) {
      float* o = sums + k * NC + (size_t)n * C + 8 * m.v;
      if (gridDim.x == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = acc[k][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(o + i, acc[k][i]);
      }
    }
  }
}

// The FULL pass again with FOUR channels per thread (8-byte loads).  ncu on the 8-channel version: 127 registers, 25 %
// occupancy, ~30 % issue utilisation although the instruction stream alone would need a third of the time - five FMA chains
// behind a two-MUFU sigmoid per element need more warps in flight, and 40 accumulators per thread leave room for only 16.
// 20 accumulators per thread fit four blocks per SM.  Same sums, same layout ([5][N][C]).
__global__ void __launch_bounds__(TPB, 4) se_bwd_reduce5_kernel(const uint2* __restrict__ dA, const uint2* __restrict__ x,
                                                                const float* __restrict__ rec, float* __restrict__ sums,
                                                                int HW, int C, int V4, int VX, int RY, size_t NC) {
  __shared__ float s_red[4 * TPB];
  const RowMap m = row_map(V4, VX, RY, blockIdx.z);
  const int n = blockIdx.y;
  float acc[5][4];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
  if (m.active) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(rec) + m.v), sh = __ldg(reinterpret_cast<const float4*>(rec + C) + m.v);
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
    const int step = gridDim.x * RY;
    for (int r = blockIdx.x * RY + m.ry; r < HW; r += UNR * step) {
      uint2 da[UNR], xa[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < HW) {
          da[u] = __ldg(dA + ((size_t)n * HW + r + u * step) * V4 + m.v);
          xa[u] = __ldg(x + ((size_t)n * HW + r + u * step) * V4 + m.v);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < HW) {
          const float d[4] = {bf16_lo(da[u].x), bf16_hi(da[u].x), bf16_lo(da[u].y), bf16_hi(da[u].y)};
          const float a[4] = {bf16_lo(xa[u].x), bf16_hi(xa[u].x), bf16_lo(xa[u].y), bf16_hi(xa[u].y)};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float z = fmaf(a[i], scv[i], shv[i]);
            const float sg = sigmoidf_(z);
            const float sl = z * sg;
            const float sp = fmaf(sl, 1.0f - sg, sg);              // silu'(z) = s + silu * (1 - s)
            const float t = d[i] * sp;
            acc[0][i] = fmaf(d[i], sl, acc[0][i]);
            acc[1][i] += t;
            acc[2][i] += sp;
            acc[3][i] = fmaf(t, a[i], acc[3][i]);
            acc[4][i] = fmaf(sp, a[i], acc[4][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    block_reduce_rows<4>(acc[k], s_red, m);
    if (m.ry == 0 && m.active) {
      float* o = sums + k * NC + (size_t)n * C + 4 * m.v;
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(o + i, acc[k][i]);
    }
  }
}

// pass 2 of the merged path: dD = a*g + b*x + c with g = (dA*gate[n,c] + dmean[n,c]/HW) * silu'(bn(x)) formed on the fly -
// the gradient w.r.t. the raw depthwise output in ONE read of (dA, x) and one write (no g tensor, no separate affine pass)
__global__ void __launch_bounds__(TPB) act_bwd_apply_kernel(const uint4* __restrict__ dA, const float* __restrict__ gate,
                                                            const float* __restrict__ dmean, float inv_hw,
                                                            const uint4* __restrict__ x, const float* __restrict__ rec,
                                                            const float* __restrict__ coef, uint4* __restrict__ out, int HW,
                                                            int C, int V, int VX, int RY) {
  const RowMap m = row_map(V, VX, RY, blockIdx.z);
  if (!m.active) return;
  const int n = blockIdx.y;
  const f8 sc = ldf8(rec + 8 * m.v), sh = ldf8(rec + C + 8 * m.v);
  const f8 ca = ldf8(coef + 8 * m.v), cb = ldf8(coef + C + 8 * m.v), cc = ldf8(coef + 2 * C + 8 * m.v);
  f8 gt = ldf8(gate + (size_t)n * C + 8 * m.v), dm = ldf8(dmean + (size_t)n * C + 8 * m.v);
#pragma unroll
  for (int i = 0; i < 8; ++i) { gt.v[i] *= ca.v[i]; dm.v[i] *= inv_hw * ca.v[i]; }      // a folded into the upstream factors
  const int step = gridDim.x * RY;
  for (int r = blockIdx.x * RY + m.ry; r < HW; r += UNR * step) {
    uint4 da[UNR], xa[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (r + u * step < HW) {
        const size_t idx = ((size_t)n * HW + r + u * step) * V + m.v;
        xa[u] = __ldg(x + idx);
        da[u] = __ldg(dA + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (r + u * step < HW) {
        const f8 a = unpack8(xa[u]), d = unpack8(da[u]);
        f8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float up = fmaf(d.v[i], gt.v[i], dm.v[i]);
          o.v[i] = fmaf(up, silu_gradf_(fmaf(a.v[i], sc.v[i], sh.v[i])), fmaf(cb.v[i], a.v[i], cc.v[i]));
        }
        out[((size_t)n * HW + r + u * step) * V + m.v] = pack8(o);
      }
    }
  }
}

// SE MLP backward, batched like the forward:
//   (a) ds2[n,c] = dgate_pre*g*(1-g);  ds1[n,r] = silu'(s1[n,r]) * sum_c ds2[n,c] We[c,r]     grid = rd blocks
//   (b) dmean[n,c] = sum_r ds1[n,r] Wr[r,c]                                                   grid = (C/256, image splits)
__global__ void __launch_bounds__(TPB) se_bwd_a_kernel(const float* __restrict__ dgate_pre, const float* __restrict__ gate,
                                                       const float* __restrict__ s1, const float* __restrict__ We,
                                                       float* __restrict__ ds2_out, float* __restrict__ ds1_out, int N, int C,
                                                       int rd) {
  extern __shared__ float s_mem[];   // [C] column r of We
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += TPB) s_mem[c] = __ldg(We + (size_t)c * rd + r);
  __syncthreads();
  for (int n = blockIdx.y * (TPB / 32) + warp; n < N; n += gridDim.y * (TPB / 32)) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 128) {
      float gv[4], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cc = c + 32 * u;
        gv[u] = cc < C ? __ldg(gate + (size_t)n * C + cc) : 0.f;
        dv[u] = cc < C ? __ldg(dgate_pre + (size_t)n * C + cc) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cc = c + 32 * u;
        if (cc < C) {
          const float d = dv[u] * gv[u] * (1.f - gv[u]);
          if (r == 0) ds2_out[(size_t)n * C + cc] = d;
          acc = fmaf(d, s_mem[cc], acc);
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) ds1_out[(size_t)n * rd + r] = acc * silu_gradf_(s1[(size_t)n * rd + r]);
  }
}

__global__ void __launch_bounds__(TPB) se_bwd_b_kernel(const float* __restrict__ ds1, const float* __restrict__ Wr,
                                                       float* __restrict__ dmean, int N, int C, int rd, int n_per_block) {
  extern __shared__ float s_mem[];   // [n_per_block][rd]
  const int n0 = blockIdx.y * n_per_block;
  const int nn = min(n_per_block, N - n0);
  for (int i = threadIdx.x; i < nn * rd; i += TPB) s_mem[i] = ds1[(size_t)n0 * rd + i];
  __syncthreads();
  const int c = blockIdx.x * TPB + threadIdx.x;
  if (c >= C) return;
  for (int nb = 0; nb < nn; nb += 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < rd; ++r) {
      const float w = __ldg(Wr + (size_t)r * C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (nb + j < nn) acc[j] = fmaf(s_mem[(nb + j) * rd + r], w, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (nb + j < nn) dmean[(size_t)(n0 + nb + j) * C + c] = acc[j];
  }
}

// SE parameter gradients: reduction over the batch without atomics.  A warp owns 4 weight elements x 8 batch slices
// (lane = 8*e + slice); each lane walks every 8th image, then the 8 slices are combined with shuffles.
__global__ void __launch_bounds__(TPB) se_bwd_w_kernel(const float* __restrict__ ds2, const float* __restrict__ ds1,
                                                       const float* __restrict__ s1, const float* __restrict__ pooled,
                                                       float inv_hw, float* __restrict__ dWr, float* __restrict__ dbr,
                                                       float* __restrict__ dWe, float* __restrict__ dbe, int N, int C,
                                                       int rd) {
  const int lane = threadIdx.x & 31, slice = lane & 7;
  const int i = (blockIdx.x * TPB + threadIdx.x) / 8;
  const bool valid = i < C * rd;
  float a_we = 0.f, a_be = 0.f, a_wr = 0.f, a_br = 0.f;
  if (valid) {
    const int c1 = i / rd, r1 = i - c1 * rd;     // dWe[c1][r1]
    const int r2 = i / C, c2 = i - r2 * C;       // dWr[r2][c2]
    for (int n = slice; n < N; n += 8) {
      const float d2 = __ldg(ds2 + (size_t)n * C + c1), sv = __ldg(s1 + (size_t)n * rd + r1);
      const float d1 = __ldg(ds1 + (size_t)n * rd + r2), pv = __ldg(pooled + (size_t)n * C + c2);
      a_we = fmaf(d2, siluf_(sv), a_we);
      a_be += d2;
      a_wr = fmaf(d1, pv * inv_hw, a_wr);
      a_br += d1;
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    a_we += __shfl_xor_sync(0xffffffffu, a_we, o);
    a_be += __shfl_xor_sync(0xffffffffu, a_be, o);
    a_wr += __shfl_xor_sync(0xffffffffu, a_wr, o);
    a_br += __shfl_xor_sync(0xffffffffu, a_br, o);
  }
  if (valid && slice == 0) {
    const int c1 = i / rd, r1 = i - c1 * rd, r2 = i / C, c2 = i - r2 * C;
    dWe[i] = a_we;
    if (r1 == 0) dbe[c1] = a_be;
    dWr[i] = a_wr;
    if (c2 == 0) dbr[r2] = a_br;
  }
}


// ------------------------------------------------------------------------------------------------ SE MLP backward, fused
// Two launches, each a grid over 16-channel chunks with 4x4 register tiles (the old three kernels were latency-bound chains
// of dependent L2 loads, ~37 us per block for ~20 MFLOP):
//   K1  ds2 = dgate_pre*g*(1-g);  dWe[c,r] = sum_n ds2[n,c]*silu(s1[n,r]);  dbe[c];  ds1_acc[n,r] += sum_{c in chunk} ds2[n,c]*We[c,r]
//   K2  ds1 = silu'(s1)*ds1_acc;  dmean[n,c] = sum_r ds1[n,r]*Wr[r,c];  dWr[r,c] = inv_hw*sum_n ds1[n,r]*pooled[n,c];  dbr[r]
constexpr int SEB_CC = 16;     // channels per block (small chunks: the launch is latency-bound, so spread it over all SMs)
constexpr int SEB_NT = 64;     // images per pass
constexpr int SEB_RD = 128;    // max rd handled by the fused kernels

__global__ void __launch_bounds__(TPB) se_bwd_k1_kernel(const float* __restrict__ dgate_pre, const float* __restrict__ gate,
                                                        const float* __restrict__ s1, const float* __restrict__ We,
                                                        float* __restrict__ ds2_out, float* __restrict__ ds1_acc,
                                                        float* __restrict__ dWe, float* __restrict__ dbe, int N, int C, int rd) {
  extern __shared__ __align__(16) float s_mem[];
  const int rdp = (rd + 3) & ~3;
  float* s_a1 = s_mem;                         // [SEB_NT][rdp]  silu(s1)
  float* s_d2 = s_a1 + SEB_NT * rdp;           // [SEB_NT][SEB_CC]
  float* s_we = s_d2 + SEB_NT * SEB_CC;        // [SEB_CC][rdp]
  const int c0 = blockIdx.x * SEB_CC, t = threadIdx.x;
  for (int i = t; i < SEB_CC * rdp; i += TPB) {
    const int cl = i / rdp, r = i - cl * rdp;
    s_we[i] = (c0 + cl < C && r < rd) ? __ldg(We + (size_t)(c0 + cl) * rd + r) : 0.f;
  }
  const int rgroups = rdp >> 2;
  // dWe tile: thread = (4 channels, 4 r)
  const int wc = (t % (SEB_CC / 4)) * 4, wr = (t / (SEB_CC / 4)) * 4;
  const bool w_active = t / (SEB_CC / 4) < rgroups;
  float acc_w[4][4] = {};
  float acc_b = 0.f;                            // dbe: threads 0..31
  for (int n0 = 0; n0 < N; n0 += SEB_NT) {
    const int nn = min(SEB_NT, N - n0);
    __syncthreads();
    for (int i = t; i < SEB_NT * rdp; i += TPB) {
      const int n = i / rdp, r = i - n * rdp;
      s_a1[i] = (n < nn && r < rd) ? siluf_(__ldg(s1 + (size_t)(n0 + n) * rd + r)) : 0.f;
    }
    for (int i = t; i < SEB_NT * SEB_CC; i += TPB) {
      const int n = i / SEB_CC, cl = i % SEB_CC;
      float d = 0.f;
      if (n < nn && c0 + cl < C) {
        const size_t idx = (size_t)(n0 + n) * C + c0 + cl;
        const float g = __ldg(gate + idx);
        d = __ldg(dgate_pre + idx) * g * (1.f - g);
        ds2_out[idx] = d;
      }
      s_d2[i] = d;
    }
    __syncthreads();
    if (w_active) {
#pragma unroll 4
      for (int n = 0; n < SEB_NT; ++n) {
        const float4 d = *reinterpret_cast<const float4*>(s_d2 + n * SEB_CC + wc);
        const float4 a = *reinterpret_cast<const float4*>(s_a1 + n * rdp + wr);
        const float dv[4] = {d.x, d.y, d.z, d.w}, av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc_w[i][j] = fmaf(dv[i], av[j], acc_w[i][j]);
      }
    }
    if (t < SEB_CC)
      for (int n = 0; n < SEB_NT; ++n) acc_b += s_d2[n * SEB_CC + t];
    // ds1 partial: item = (4 images, 4 r); sum over the chunk's channels
    for (int item = t; item < (SEB_NT / 4) * rgroups; item += TPB) {
      const int nb = (item % (SEB_NT / 4)) * 4, rb = (item / (SEB_NT / 4)) * 4;
      float acc[4][4] = {};
#pragma unroll 4
      for (int cl = 0; cl < SEB_CC; ++cl) {
        const float4 w = *reinterpret_cast<const float4*>(s_we + cl * rdp + rb);
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d = s_d2[(nb + i) * SEB_CC + cl];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d, wv[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + i < nn && rb + j < rd) atomicAdd(ds1_acc + (size_t)(n0 + nb + i) * rd + rb + j, acc[i][j]);
    }
  }
  if (w_active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + wc + i < C && wr + j < rd) dWe[(size_t)(c0 + wc + i) * rd + wr + j] = acc_w[i][j];
  }
  if (t < SEB_CC && c0 + t < C) dbe[c0 + t] = acc_b;
}

struct SeBn {                  // device view of trt_se_bn_t (+ the gate): BatchNorm backward of the gated activation
  const float* sums;           // [5][N][C] from se_bwd_reduce_kernel<true>; null = not requested
  const float* gate;           // [N][C]
  const float* rec;            // [4][C]
  const float* gamma;
  float *coef, *dgamma, *dbeta;
  double count;
};

__global__ void __launch_bounds__(TPB) se_bwd_k2_kernel(const float* __restrict__ ds1_acc, const float* __restrict__ s1,
                                                        const float* __restrict__ pooled, float inv_hw,
                                                        const float* __restrict__ Wr, float* __restrict__ dmean, float* __restrict__ dWr,
                                                        float* __restrict__ dbr, int N, int C, int rd, const SeBn bn) {
  extern __shared__ __align__(16) float s_mem[];
  const int rdp = (rd + 3) & ~3;
  float* s_d1 = s_mem;                         // [SEB_NT][rdp]
  float* s_wr = s_d1 + SEB_NT * rdp;           // [rdp][SEB_CC]
  float* s_po = s_wr + rdp * SEB_CC;           // [SEB_NT][SEB_CC]
  const int c0 = blockIdx.x * SEB_CC, t = threadIdx.x;
  for (int i = t; i < rdp * SEB_CC; i += TPB) {
    const int r = i / SEB_CC, cl = i % SEB_CC;
    s_wr[i] = (r < rd && c0 + cl < C) ? __ldg(Wr + (size_t)r * C + c0 + cl) : 0.f;
  }
  const int rgroups = rdp >> 2;
  const int wc = (t % (SEB_CC / 4)) * 4, wr = (t / (SEB_CC / 4)) * 4;      // dWr tile: (4 r, 4 channels)
  const bool w_active = t / (SEB_CC / 4) < rgroups;
  float acc_w[4][4] = {};
  float acc_b = 0.f;                            // dbr: block 0, threads 0..rd-1
  double bn_g[4] = {0, 0, 0, 0}, bn_gx[4] = {0, 0, 0, 0};     // sum g, sum g*x of this thread's 4 channels over its images
  for (int n0 = 0; n0 < N; n0 += SEB_NT) {
    const int nn = min(SEB_NT, N - n0);
    __syncthreads();
    for (int i = t; i < SEB_NT * rdp; i += TPB) {
      const int n = i / rdp, r = i - n * rdp;
      float d = 0.f;
      if (n < nn && r < rd) {
        const size_t idx = (size_t)(n0 + n) * rd + r;
        d = __ldg(ds1_acc + idx) * silu_gradf_(__ldg(s1 + idx));
      }
      s_d1[i] = d;
    }
    for (int i = t; i < SEB_NT * SEB_CC; i += TPB) {
      const int n = i / SEB_CC, cl = i % SEB_CC;
      s_po[i] = (n < nn && c0 + cl < C) ? __ldg(pooled + (size_t)(n0 + n) * C + c0 + cl) * inv_hw : 0.f;
    }
    __syncthreads();
    // dmean: thread = (2 images, 4 channels)
    if ((t / (SEB_CC / 4)) * 2 < SEB_NT) {
      const int dc = (t % (SEB_CC / 4)) * 4, dn = (t / (SEB_CC / 4)) * 2;
      float acc[2][4] = {};
#pragma unroll 4
      for (int r = 0; r < rd; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(s_wr + r * SEB_CC + dc);
        const float a0 = s_d1[dn * rdp + r], a1 = s_d1[(dn + 1) * rdp + r];
        acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]);
        acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
        acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]);
        acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (dn + i < nn && c0 + dc + j < C) dmean[(size_t)(n0 + dn + i) * C + c0 + dc + j] = acc[i][j];
      if (bn.sums) {
        // BatchNorm-backward sums of the gated activation from the per-(image, channel) sums of pass 1 (C % 8 == 0 and dc is
        // a multiple of 4, so the four channels are all inside or all outside the tensor)
        const size_t NC = (size_t)N * C;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (dn + i < nn && c0 + dc < C) {
            const size_t o = (size_t)(n0 + dn + i) * C + c0 + dc;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(bn.gate + o));
            const float4 q1 = __ldcg(reinterpret_cast<const float4*>(bn.sums + NC + o));
            const float4 q2 = __ldcg(reinterpret_cast<const float4*>(bn.sums + 2 * NC + o));
            const float4 q3 = __ldcg(reinterpret_cast<const float4*>(bn.sums + 3 * NC + o));
            const float4 q4 = __ldcg(reinterpret_cast<const float4*>(bn.sums + 4 * NC + o));
            const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, v1[4] = {q1.x, q1.y, q1.z, q1.w}, v2[4] = {q2.x, q2.y, q2.z, q2.w};
            const float v3[4] = {q3.x, q3.y, q3.z, q3.w}, v4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float dm = acc[i][j] * inv_hw;
              bn_g[j] += (double)fmaf(gv[j], v1[j], dm * v2[j]);
              bn_gx[j] += (double)fmaf(gv[j], v3[j], dm * v4[j]);
            }
          }
        }
      }
    }
    if (w_active) {
#pragma unroll 4
      for (int n = 0; n < SEB_NT; ++n) {
        const float4 d = *reinterpret_cast<const float4*>(s_d1 + n * rdp + wr);
        const float4 q = *reinterpret_cast<const float4*>(s_po + n * SEB_CC + wc);
        const float dv[4] = {d.x, d.y, d.z, d.w}, pv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc_w[i][j] = fmaf(dv[i], pv[j], acc_w[i][j]);
      }
    }
    if (blockIdx.x == 0 && t < rd)
      for (int n = 0; n < SEB_NT; ++n) acc_b += s_d1[n * rdp + t];
  }
  if (w_active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (wr + i < rd && c0 + wc + j < C) dWr[(size_t)(wr + i) * C + c0 + wc + j] = acc_w[i][j];
  }
  if (blockIdx.x == 0 && t < rd) dbr[t] = acc_b;
  if (bn.sums) {
    // combine the (SEB_NT / 2) image-pair threads of every channel quad, then one thread per channel writes the
    // coefficients of dx = a*g + b*x + c and the affine gradients (what bn_bwd_finalize does from fp64 sums)
    __syncthreads();
    double* s_d = reinterpret_cast<double*>(s_mem);               // [2][SEB_NT/2][SEB_CC] doubles = 8 KB <= the tiles
    const int quad = t % (SEB_CC / 4), pair = t / (SEB_CC / 4);
    if (pair < SEB_NT / 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s_d[pair * SEB_CC + quad * 4 + j] = bn_g[j];
        s_d[(SEB_NT / 2 + pair) * SEB_CC + quad * 4 + j] = bn_gx[j];
      }
    }
    __syncthreads();
    if (t < SEB_CC && c0 + t < C) {
      double sg = 0, sgx = 0;
      for (int q = 0; q < SEB_NT / 2; ++q) { sg += s_d[q * SEB_CC + t]; sgx += s_d[(SEB_NT / 2 + q) * SEB_CC + t]; }
      const int c = c0 + t;
      const float mean = bn.rec[2 * C + c], rstd = bn.rec[3 * C + c];
      const double sgxh = (double)rstd * (sgx - (double)mean * sg);   // sum g*xhat
      const float a = bn.gamma[c] * rstd;
      const float m1 = (float)(sg / bn.count), m2 = (float)(sgxh / bn.count);
      bn.coef[c] = a;
      bn.coef[C + c] = -a * rstd * m2;
      bn.coef[2 * C + c] = a * (mean * rstd * m2 - m1);
      bn.dgamma[c] = (float)sgxh;
      bn.dbeta[c] = (float)sg;
    }
  }
}

// ------------------------------------------------------------------------------------------------ activation backward
// g = (dA*gate[n,c] + dmean[n,c]*inv_hw) * silu'(bn(x)) ; bstats += (sum g, sum g*xhat).  dA / gate / dmean may be null.
// act == 0: no SiLU (g = upstream), used for BN layers without activation.
__global__ void __launch_bounds__(TPB) act_bwd_kernel(const uint4* __restrict__ dA, const float* __restrict__ gate,
                                                      const float* __restrict__ dmean, float inv_hw,
                                                      const uint4* __restrict__ x, const float* __restrict__ rec,
                                                      uint4* __restrict__ g_out, double* __restrict__ bstats, int HW, int C,
                                                      int V, int VX, int RY, int act) {
  __shared__ float s_red[16 * TPB];
  const RowMap m = row_map(V, VX, RY, blockIdx.z);
  const int n = blockIdx.y;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (m.active) {
    const f8 sc = ldf8(rec + 8 * m.v), sh = ldf8(rec + C + 8 * m.v);
    f8 gt, dm;
#pragma unroll
    for (int i = 0; i < 8; ++i) { gt.v[i] = 1.f; dm.v[i] = 0.f; }
    if (gate) gt = ldf8(gate + (size_t)n * C + 8 * m.v);
    if (dmean) {
      dm = ldf8(dmean + (size_t)n * C + 8 * m.v);
#pragma unroll
      for (int i = 0; i < 8; ++i) dm.v[i] *= inv_hw;
    }
    const int step = gridDim.x * RY;
    for (int r = blockIdx.x * RY + m.ry; r < HW; r += UNR * step) {
      uint4 da[UNR], xa[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < HW) {
          const size_t idx = ((size_t)n * HW + r + u * step) * V + m.v;
          xa[u] = __ldg(x + idx);
          if (dA) da[u] = __ldg(dA + idx);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r + u * step < HW) {
          const size_t idx = ((size_t)n * HW + r + u * step) * V + m.v;
          const f8 a = unpack8(xa[u]);
          f8 d;
          if (dA) d = unpack8(da[u]);
          f8 o;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float up = (dA ? d.v[i] * gt.v[i] : 0.f) + dm.v[i];
            o.v[i] = act ? up * silu_gradf_(fmaf(a.v[i], sc.v[i], sh.v[i])) : up;
          }
          const uint4 packed = pack8(o);
          g_out[idx] = packed;
          const f8 oq = unpack8(packed);   // statistics over the bf16-rounded values that are stored
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i] += oq.v[i];
            acc[8 + i] = fmaf(oq.v[i], a.v[i], acc[8 + i]);        // sum g*x; turned into sum g*xhat after the loop
          }
        }
      }
    }
    // sum g*xhat = rstd * (sum g*x - mean * sum g): mean / rstd are only live here, not across the streaming loop
    const f8 mu = ldf8(rec + 2 * C + 8 * m.v), rs = ldf8(rec + 3 * C + 8 * m.v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[8 + i] = rs.v[i] * (acc[8 + i] - mu.v[i] * acc[i]);
  }
  block_reduce_rows<16>(acc, s_red, m);
  if (m.ry == 0 && m.active) {
    double* rep = bstats + (size_t)((blockIdx.x + blockIdx.y * gridDim.x) % TRT_STAT_REPLICAS) * 2 * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(rep + 8 * m.v + i, (double)acc[i]);
      atomicAdd(rep + C + 8 * m.v + i, (double)acc[8 + i]);
    }
  }
}

// fp32 [N, K] weight -> bf16 [N, K] and bf16 [K, N] (transposed copy for dgrad)
__global__ void pack_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, __nv_bfloat16* __restrict__ ot,
                              int N, int K) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int n = n0 + j, k = k0 + threadIdx.x;
    float v = 0.f;
    if (n < N && k < K) {
      v = w[(size_t)n * K + k];
      o[(size_t)n * K + k] = __float2bfloat16_rn(v);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  if (ot) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int k = k0 + j, n = n0 + threadIdx.x;
      if (n < N && k < K) ot[(size_t)k * N + n] = __float2bfloat16_rn(tile[threadIdx.x][j]);
    }
  }
}

// all 1x1-conv weights of a model in ONE launch: table[i] = {w, o, ot, N, K, first_tile} (int64 each, device memory)
__global__ void pack_w_batch_kernel(const long long* __restrict__ table, int count) {
  __shared__ float tile[32][33];
  __shared__ int s_entry;
  if (threadIdx.y == 0) {
    // last entry whose first tile is <= blockIdx.x (first tiles ascend, entry 0 starts at tile 0): one warp counts them
    // with independent loads - ONE L2 round trip; the binary search this replaces was a chain of log2(count) dependent
    // loads in front of a block that only moves 8 KB (17 k such blocks per step)
    int n_le = 0;
    for (int base = 0; base < count; base += 32) {
      const int i = base + threadIdx.x;
      const bool le = i < count && __ldg(table + i * 6 + 5) <= (long long)blockIdx.x;
      n_le += __popc(__ballot_sync(0xffffffffu, le));
    }
    if (threadIdx.x == 0) s_entry = n_le - 1;
  }
  __syncthreads();
  const long long* e = table + s_entry * 6;
  const float* w = reinterpret_cast<const float*>(e[0]);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e[1]);
  __nv_bfloat16* ot = reinterpret_cast<__nv_bfloat16*>(e[2]);
  const int N = (int)e[3], K = (int)e[4];
  const int local = blockIdx.x - (int)e[5], tiles_k = (K + 31) / 32;
  const int k0 = (local % tiles_k) * 32, n0 = (local / tiles_k) * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int n = n0 + j, k = k0 + threadIdx.x;
    float v = 0.f;
    if (n < N && k < K) {
      v = w[(size_t)n * K + k];
      o[(size_t)n * K + k] = __float2bfloat16_rn(v);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  if (ot) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int k = k0 + j, n = n0 + threadIdx.x;
      if (n < N && k < K) ot[(size_t)k * N + n] = __float2bfloat16_rn(tile[threadIdx.x][j]);
    }
  }
}

__global__ void scale_f32_kernel(float* __restrict__ x, size_t n, float alpha) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= alpha;
}

}  // namespace

// Block count of the kernels that end in per-channel fp64 atomics.  Same-address atomics serialise at the L2 (~25-50 ns
// each, measured: 4x the blocks = 3x the time on the 7x7 layers), so the grid follows the bytes streamed - about 384 KB per
// block - between half a wave and three blocks per SM.  (env TEETHRT_RED_KB overrides the per-block quantum: bring-up knob)
static int red_blocks(size_t bytes) {
  static int kb = 0;
  if (kb == 0) {
    const char* e = getenv("TEETHRT_RED_KB");
    kb = e ? atoi(e) : 384;
    if (kb < 1) kb = 384;
  }
  long long b = (long long)(bytes / ((size_t)kb << 10));
  const int lo = trt_num_sms() / 2, hi = 3 * trt_num_sms();
  return b < lo ? lo : (b > hi ? hi : (int)b);
}

#define CHECK_C(C) TRT_REQUIRE((C) > 0 && (C) % 8 == 0, "%s: channels must be a positive multiple of 8 (got %d)", __func__, (C))

extern "C" int trt_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, long long* num_batches_tracked, float* rec, int C, double count,
                               float eps, float momentum, cudaStream_t stream) {
  TRT_REQUIRE(stats && gamma && beta && rec && C > 0 && count > 0, "trt_bn_finalize: bad argument");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(stats, gamma, beta, running_mean, running_var,
                                                          num_batches_tracked, rec, C, count, eps, momentum);
  return trt_check_launch("trt_bn_finalize");
}

extern "C" int trt_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                float* rec, int C, float eps, cudaStream_t stream) {
  TRT_REQUIRE(gamma && beta && running_mean && running_var && rec && C > 0, "trt_bn_fold_eval: bad argument");
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, stream>>>(gamma, beta, running_mean, running_var, rec, C, eps);
  return trt_check_launch("trt_bn_fold_eval");
}

extern "C" int trt_bn_bwd_finalize(const double* bstats, const float* rec, const float* gamma, float* coef, float* dgamma,
                                   float* dbeta, int C, double count, cudaStream_t stream) {
  TRT_REQUIRE(bstats && rec && gamma && coef && dgamma && dbeta && C > 0, "trt_bn_bwd_finalize: bad argument");
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(bstats, rec, gamma, coef, dgamma, dbeta, C, count, 0);
  return trt_check_launch("trt_bn_bwd_finalize");
}

static int check_fin(const trt_bn_fin_t* f, const char* who) {
  if (f && !(f->stats && f->gamma && f->beta && f->rec && f->count > 0))
    return trt_set_error(TRT_ERR_INVALID, "%s: incomplete lazy BatchNorm record", who);
  return TRT_OK;
}

// A lazy prologue reads 16 fp64 words (128 bytes, L2 hits) per channel IN EVERY BLOCK.  On the wide, narrow-channel layers
// that is a few MB in total and replaces a 4-5 us launch; on the 14x14 / 7x7 layers (hundreds of blocks x up to 2048
// channels each) it is 100+ MB of redundant L2 reads, i.e. more than the tensor itself (measured: lazy everywhere was
// 0.18 ms/step SLOWER than separate launches).  So the entry point decides: in-prologue while the redundant bytes stay
// under ~24 MB (about 3 us of L2 bandwidth), otherwise it enqueues the finalise kernel itself in front of the consumer -
// same arithmetic, same results either way.  TEETHRT_LAZY_FORCE=1 / 0 pins the choice (tests, A/B runs).
static bool lazy_pays(long long blocks, int ch_per_block) {
  const char* e = getenv("TEETHRT_LAZY_FORCE");
  if (e && *e) return *e != '0';
  return (size_t)blocks * (size_t)ch_per_block * 128u <= ((size_t)24 << 20);
}
static void finalize_now(const trt_bn_fin_t& f, int C, cudaStream_t stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(f.stats, f.gamma, f.beta, f.running_mean, f.running_var,
                                                          f.num_batches_tracked, f.rec, C, f.count, f.eps, f.momentum);
  trt_count_launch(1);
}

extern "C" int trt_bn_apply(const void* x, const float* rec, const void* residual, void* out, const trt_bn_fin_t* fin_host,
                            int rows, int C, int act, cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(x && (rec || fin_host) && out && rows > 0, "trt_bn_apply: bad argument");
  int rc;
  if ((rc = check_fin(fin_host, "trt_bn_apply"))) return rc;
  trt_bn_fin_t fin = {};
  if (fin_host) fin = *fin_host;
  const Launch L = plan(C);
  dim3 grid(row_blocks((rows + UNR - 1) / UNR, L.RY, L.slabs, 6 * trt_num_sms()), L.slabs);
  int lazy = fin_host ? 1 : 0;
  if (lazy && !lazy_pays((long long)grid.x * grid.y, L.VX * 8)) {
    finalize_now(fin, C, stream);
    rec = fin.rec;
    lazy = 0;
  }
  bn_apply_kernel<<<grid, TPB, 0, stream>>>((const uint4*)x, rec, (const uint4*)residual, (uint4*)out, rows, C, L.V, L.VX,
                                            L.RY, act, lazy, fin);
  return trt_check_launch("trt_bn_apply");
}

extern "C" int trt_pool_act(const void* x, const float* rec, float* pooled_sum, int zeroed, const trt_bn_fin_t* fin_host, int N,
                            int HW, int C, int act, cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(x && pooled_sum && N > 0 && HW > 0, "trt_pool_act: bad argument");
  int rc;
  if ((rc = check_fin(fin_host, "trt_pool_act"))) return rc;
  trt_bn_fin_t fin = {};
  if (fin_host) fin = *fin_host;
  const Launch L = plan_img(C, HW, true);
  const bool single = one_pass(L, HW);
  if (!zeroed && !single) TRT_CUDA(cudaMemsetAsync(pooled_sum, 0, (size_t)N * C * sizeof(float), stream));
  int target = env_mult("TEETHRT_POOL_BLOCKS", 4) * trt_num_sms() / (N * L.slabs);     // per-image outputs: no cross-block contention, latency-bound when fewer
  if (target < 1) target = 1;
  dim3 grid(single ? 1 : row_blocks((HW + UNR - 1) / UNR, L.RY, 1, target), N, L.slabs);
  int lazy = fin_host ? 1 : 0;
  if (lazy && !lazy_pays((long long)grid.x * grid.y * grid.z, L.VX * 8)) {
    finalize_now(fin, C, stream);
    rec = fin.rec;
    lazy = 0;
  }
  TRT_CUDA(trt_launch(pool_act_kernel, grid, dim3(TPB), 0, stream, (const uint4*)x, rec, pooled_sum, HW, C, L.V, L.VX, L.RY, act, lazy, fin));
  return trt_check_launch("trt_pool_act");
}

extern "C" int trt_se_fwd(const float* pooled_sum, float inv_hw, const float* Wr, const float* br, const float* We,
                          const float* be, float* s1, float* gate, void* apply_x, int HW, int N, int C, int rd,
                          cudaStream_t stream) {
  TRT_REQUIRE(pooled_sum && Wr && br && We && be && s1 && gate && N > 0 && C > 0 && rd > 0, "trt_se_fwd: bad argument");
  TRT_REQUIRE(!apply_x || (HW > 0 && C % 8 == 0), "trt_se_fwd: apply_x needs HW > 0 and C %% 8 == 0");
  static const int small_off = [] { const char* e = getenv("TEETHRT_SE_SMALL"); return (e && *e == '0') ? 1 : 0; }();   // A/B switch
  if (N <= SE_SMALL_N && !apply_x && !small_off) {
    TRT_CUDA(trt_launch(se_reduce_small_kernel, dim3(rd), dim3(TPB), 0, stream, pooled_sum, inv_hw, Wr, br, s1, N, C, rd));
    const size_t smem_s = ((size_t)SE_SCC * (rd + 1) + (size_t)SE_SMALL_N * rd) * sizeof(float);
    TRT_REQUIRE(smem_s <= 48 * 1024, "trt_se_fwd: rd %d too large for the small-batch expand kernel", rd);
    TRT_CUDA(trt_launch(se_expand_small_kernel, dim3((C + SE_SCC - 1) / SE_SCC), dim3(TPB), smem_s, stream, (const float*)s1, We, be, gate, N, C, rd));
    trt_count_launch(1);
    return trt_check_launch("trt_se_fwd");
  }
  TRT_CUDA(trt_launch(se_reduce_kernel, dim3(rd, (N + 7) / 8), dim3(TPB), (size_t)C * sizeof(float), stream, pooled_sum, inv_hw, Wr, br, s1, N, C, rd));
  int splits = N >= 32 ? 4 : (N >= 8 ? 2 : 1);
  const int npb = (N + splits - 1) / splits;
  splits = (N + npb - 1) / npb;
  const size_t smem = ((size_t)SE_CC * (rd + 1) + (size_t)npb * rd + (apply_x ? (size_t)npb * SE_CC : 0)) * sizeof(float);
  TRT_REQUIRE(smem <= 96 * 1024, "trt_se_fwd: batch %d x rd %d too large for one block", N, rd);
  // the opt-in is per device and cheap: set on every call (a process-wide "done" flag only covered the first device used)
  TRT_CUDA(cudaFuncSetAttribute(se_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  TRT_CUDA(trt_launch(se_expand_kernel, dim3((C + SE_CC - 1) / SE_CC, splits), dim3(TPB), smem, stream, s1, We, be, gate, N, C, rd, npb,
                      reinterpret_cast<uint4*>(apply_x), HW));
  trt_count_launch(1);
  return trt_check_launch("trt_se_fwd");
}

extern "C" int trt_gate_apply(const void* x, const float* rec, const float* gate, void* out, int N, int HW, int C,
                              cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(x && gate && out && N > 0 && HW > 0, "trt_gate_apply: bad argument");
  const Launch L = plan(C);
  int target = env_mult("TEETHRT_GATE_BLOCKS", 4) * trt_num_sms() / (N * L.slabs);
  if (target < 1) target = 1;
  dim3 grid(row_blocks((HW + UNR - 1) / UNR, L.RY, 1, target), N, L.slabs);
  TRT_CUDA(trt_launch(gate_apply_kernel, grid, dim3(TPB), 0, stream, (const uint4*)x, rec, gate, (uint4*)out, HW, C, L.V, L.VX, L.RY));
  return trt_check_launch("trt_gate_apply");
}

extern "C" int trt_bn_bwd_reduce(const void* dy, const void* x, const float* rec, double* bstats, int rows, int C,
                                 cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(dy && x && rec && bstats && rows > 0, "trt_bn_bwd_reduce: bad argument");
  const Launch L = plan(C);
  dim3 grid(row_blocks((rows + UNR - 1) / UNR, L.RY, L.slabs, red_blocks((size_t)rows * C * 4)), L.slabs);
  bn_bwd_reduce_kernel<<<grid, TPB, 0, stream>>>((const uint4*)dy, (const uint4*)x, rec, bstats, rows, C, L.V, L.VX, L.RY);
  return trt_check_launch("trt_bn_bwd_reduce");
}

extern "C" int trt_affine2(const void* dy, const void* x, const float* coef, void* out, const trt_bn_bwd_fin_t* fin_host,
                           int rows, int C, cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(dy && x && (coef || fin_host) && out && rows > 0, "trt_affine2: bad argument");
  TRT_REQUIRE(!fin_host || (fin_host->bstats && fin_host->rec && fin_host->gamma && fin_host->dgamma && fin_host->dbeta &&
                            fin_host->count > 0), "trt_affine2: incomplete lazy BatchNorm-backward record");
  trt_bn_bwd_fin_t fin = {};
  if (fin_host) fin = *fin_host;
  const Launch L = plan(C);
  dim3 grid(row_blocks((rows + UNR - 1) / UNR, L.RY, L.slabs, 6 * trt_num_sms()), L.slabs);
  int lazy = fin_host ? 1 : 0;
  if (lazy && fin.coef && !lazy_pays((long long)grid.x * grid.y, L.VX * 8)) {
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(fin.bstats, fin.rec, fin.gamma, fin.coef, fin.dgamma, fin.dbeta, C, fin.count, fin.raw_x);
    trt_count_launch(1);
    coef = fin.coef;
    lazy = 0;
  }
  affine2_kernel<<<grid, TPB, 0, stream>>>((const uint4*)dy, (const uint4*)x, coef, (uint4*)out, rows, C, L.V, L.VX, L.RY,
                                           lazy, fin);
  return trt_check_launch("trt_affine2");
}

extern "C" int trt_se_bwd_reduce(const void* dA, const void* x, const float* rec, float* sums, int zeroed, int full, int N,
                                 int HW, int C, cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(dA && x && rec && sums && N > 0 && HW > 0, "trt_se_bwd_reduce: bad argument");
  const Launch L = plan_img(C, HW);
  const bool single = one_pass(L, HW);
  const size_t NC = (size_t)N * C;
  if (!zeroed && !single) TRT_CUDA(cudaMemsetAsync(sums, 0, (full ? 5 : 1) * NC * sizeof(float), stream));
  int target = 3 * trt_num_sms() / (N * L.slabs);     // per-image outputs: no cross-block contention, latency-bound when fewer
  if (target < 1) target = 1;
  dim3 grid(single ? 1 : row_blocks((HW + UNR - 1) / UNR, L.RY, 1, target), N, L.slabs);
  const char* v8 = getenv("TEETHRT_SE_REDUCE_V8");          // "1": the 8-channel-per-thread version (A/B switch)
  if (full && !single && !(v8 && *v8 == '1')) {
    // four channels per thread: V4 = C / 4 vectors of 8 bytes
    Launch L4;
    L4.V = C / 4;
    L4.slabs = (L4.V + TPB - 1) / TPB;
    L4.VX = (L4.V + L4.slabs - 1) / L4.slabs;
    L4.RY = TPB / L4.VX;
    // one full wave: the kernel is compiled for 4 resident blocks per SM, and the isolated sweep (tools/gpu_elt_sweep.sh:
    // 771 / 716 / 816 / 997 us per step at 3 / 4 / 6 / 12) has its minimum exactly there - more blocks only add reduction
    // tails and atomics, fewer leave SMs idle
    int tgt = env_mult("TEETHRT_SEBR_BLOCKS", 4) * trt_num_sms() / (N * L4.slabs);
    if (tgt < 1) tgt = 1;
    dim3 g4(row_blocks((HW + UNR - 1) / UNR, L4.RY, 1, tgt), N, L4.slabs);
    se_bwd_reduce5_kernel<<<g4, TPB, 0, stream>>>((const uint2*)dA, (const uint2*)x, rec, sums, HW, C, L4.V, L4.VX, L4.RY, NC);
  } else if (full) se_bwd_reduce_kernel<true><<<grid, TPB, 0, stream>>>((const uint4*)dA, (const uint4*)x, rec, sums, HW, C, L.V, L.VX, L.RY, NC);
  else se_bwd_reduce_kernel<false><<<grid, TPB, 0, stream>>>((const uint4*)dA, (const uint4*)x, rec, sums, HW, C, L.V, L.VX, L.RY, NC);
  return trt_check_launch("trt_se_bwd_reduce");
}

extern "C" int trt_act_bwd_apply(const void* dA, const float* gate, const float* dmean, float inv_hw, const void* x,
                                 const float* rec, const float* coef, void* out, int N, int HW, int C, cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(dA && gate && dmean && x && rec && coef && out && N > 0 && HW > 0, "trt_act_bwd_apply: bad argument");
  const Launch L = plan(C);
  int target = env_mult("TEETHRT_ABA_BLOCKS", 4) * trt_num_sms() / (N * L.slabs);
  if (target < 1) target = 1;
  dim3 grid(row_blocks((HW + UNR - 1) / UNR, L.RY, 1, target), N, L.slabs);
  act_bwd_apply_kernel<<<grid, TPB, 0, stream>>>((const uint4*)dA, gate, dmean, inv_hw, (const uint4*)x, rec, coef, (uint4*)out,
                                                 HW, C, L.V, L.VX, L.RY);
  return trt_check_launch("trt_act_bwd_apply");
}

extern "C" int trt_se_bwd(const float* dgate_pre, const float* gate, const float* s1, const float* pooled_sum, float inv_hw,
                          const float* Wr, const float* We, float* ds2, float* ds1, float* dmean, float* dWr, float* dbr,
                          float* dWe, float* dbe, int ds1_zeroed, const trt_se_bn_t* bn_host, int N, int C, int rd,
                          cudaStream_t stream) {
  TRT_REQUIRE(dgate_pre && gate && s1 && pooled_sum && Wr && We && ds2 && ds1 && dmean && dWr && dbr && dWe && dbe,
              "trt_se_bwd: null pointer");
  SeBn bn = {};
  if (bn_host) {
    TRT_REQUIRE(bn_host->sums && bn_host->rec && bn_host->gamma && bn_host->coef && bn_host->dgamma && bn_host->dbeta && bn_host->count > 0,
                "trt_se_bwd: incomplete BatchNorm-backward record");
    TRT_REQUIRE(rd <= SEB_RD, "trt_se_bwd: the BatchNorm-backward tail needs rd <= %d (got %d)", SEB_RD, rd);
    bn.sums = bn_host->sums; bn.gate = gate; bn.rec = bn_host->rec; bn.gamma = bn_host->gamma; bn.coef = bn_host->coef;
    bn.dgamma = bn_host->dgamma; bn.dbeta = bn_host->dbeta; bn.count = bn_host->count;
  }
  if (rd <= SEB_RD && (SEB_NT / 4) * ((rd + 3) / 4) >= 1) {
    // fused path: ds1 (scratch) accumulates K1's per-chunk partial sums of ds2.We (zeroed here); K2 applies silu' on load
    const int rdp = (rd + 3) & ~3;
    if (!ds1_zeroed) TRT_CUDA(cudaMemsetAsync(ds1, 0, (size_t)N * rd * sizeof(float), stream));
    const int blocks = (C + SEB_CC - 1) / SEB_CC;
    const size_t smem1 = ((size_t)SEB_NT * rdp + (size_t)SEB_NT * SEB_CC + (size_t)SEB_CC * rdp) * sizeof(float);
    size_t smem2 = ((size_t)SEB_NT * rdp + (size_t)rdp * SEB_CC + (size_t)SEB_NT * SEB_CC) * sizeof(float);
    if (bn.sums && smem2 < (size_t)SEB_NT * SEB_CC * sizeof(double)) smem2 = (size_t)SEB_NT * SEB_CC * sizeof(double);   // the tail's fp64 scratch
    TRT_CUDA(cudaFuncSetAttribute(se_bwd_k1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    TRT_CUDA(cudaFuncSetAttribute(se_bwd_k2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    se_bwd_k1_kernel<<<blocks, TPB, smem1, stream>>>(dgate_pre, gate, s1, We, ds2, ds1, dWe, dbe, N, C, rd);
    trt_count_launch(1);
    se_bwd_k2_kernel<<<blocks, TPB, smem2, stream>>>(ds1, s1, pooled_sum, inv_hw, Wr, dmean, dWr, dbr, N, C, rd, bn);
    return trt_check_launch("trt_se_bwd");
  }
  TRT_REQUIRE(!bn.sums, "trt_se_bwd: the BatchNorm-backward tail is only built into the fused path");
  se_bwd_a_kernel<<<dim3(rd, (N + 7) / 8), TPB, (size_t)C * sizeof(float), stream>>>(dgate_pre, gate, s1, We, ds2, ds1, N, C, rd);
  {
    int splits = N >= 32 ? 4 : (N >= 8 ? 2 : 1);
    const int npb = (N + splits - 1) / splits;
    splits = (N + npb - 1) / npb;
    se_bwd_b_kernel<<<dim3((C + TPB - 1) / TPB, splits), TPB, (size_t)npb * rd * sizeof(float), stream>>>(ds1, Wr, dmean, N, C, rd, npb);
  }
  trt_count_launch(1);
  se_bwd_w_kernel<<<(C * rd * 8 + TPB - 1) / TPB, TPB, 0, stream>>>(ds2, ds1, s1, pooled_sum, inv_hw, dWr, dbr, dWe, dbe, N, C, rd);
  trt_count_launch(1);
  return trt_check_launch("trt_se_bwd");
}

extern "C" int trt_act_bwd(const void* dA, const float* gate, const float* dmean, float inv_hw, const void* x,
                           const float* rec, void* g_out, double* bstats, int N, int HW, int C, int act, cudaStream_t stream) {
  CHECK_C(C);
  TRT_REQUIRE(x && rec && g_out && bstats && N > 0 && HW > 0 && (dA || dmean), "trt_act_bwd: bad argument");
  const Launch L = plan(C);
  int target = 3 * trt_num_sms() / (N * L.slabs);     // per-image outputs: no cross-block contention, latency-bound when fewer
  if (target < 1) target = 1;
  dim3 grid(row_blocks((HW + UNR - 1) / UNR, L.RY, 1, target), N, L.slabs);
  act_bwd_kernel<<<grid, TPB, 0, stream>>>((const uint4*)dA, gate, dmean, inv_hw, (const uint4*)x, rec, (uint4*)g_out,
                                           bstats, HW, C, L.V, L.VX, L.RY, act);
  return trt_check_launch("trt_act_bwd");
}

extern "C" int trt_pack_w1x1(const float* w, void* w_bf16, void* wt_bf16, int N, int K, cudaStream_t stream) {
  TRT_REQUIRE(w && w_bf16 && N > 0 && K > 0, "trt_pack_w1x1: bad argument");
  dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
  pack_w_kernel<<<grid, block, 0, stream>>>(w, (__nv_bfloat16*)w_bf16, (__nv_bfloat16*)wt_bf16, N, K);
  return trt_check_launch("trt_pack_w1x1");
}

extern "C" int trt_pack_w1x1_batch(const long long* table_dev, int count, int total_tiles, cudaStream_t stream) {
  TRT_REQUIRE(table_dev && count > 0 && total_tiles > 0, "trt_pack_w1x1_batch: bad argument");
  pack_w_batch_kernel<<<total_tiles, dim3(32, 8), 0, stream>>>(table_dev, count);
  return trt_check_launch("trt_pack_w1x1_batch");
}

extern "C" int trt_scale_f32(float* x, size_t n, float alpha, cudaStream_t stream) {
  TRT_REQUIRE(x && n > 0, "trt_scale_f32: bad argument");
  int grid = (int)((n + 255) / 256);
  if (grid > 4 * trt_num_sms()) grid = 4 * trt_num_sms();
  TRT_CUDA(trt_launch(scale_f32_kernel, dim3(grid), dim3(256), 0, stream, x, n, alpha));
  return trt_check_launch("trt_scale_f32");
}
