// Squeeze-excite MLP of an MBConv block at training batch sizes, forward and backward, ONE launch each.
//
// The two-launch versions (eltwise.cu: se_reduce + se_expand, se_bwd_k1 + se_bwd_k2) cost 16 us / 27 us per block inside the
// captured train step for ~20 MFLOP (tools/knockout.py: 0.53 + 0.88 ms per step, every microsecond of it exposed): the forward
// re-read the pooled means once per reduced channel (77 MB of L2 reads), the backward funnelled 1.2 M fp32 atomics into 7 k
// addresses, and every launch boundary is a dependent round trip.  Here the grid is one block per channel chunk (<= one block
// per SM, so all blocks are co-resident) and the two contractions that reduce over CHANNELS - s1 = mean.Wr^T in the forward,
// ds1 = ds2.We in the backward - are done split-K: every block writes the partial product of its chunk, a grid-wide barrier,
// every block sums a slice of the partials, a second barrier, and the second half of the MLP runs on the finished vector.
// Nothing is atomically accumulated, so the result is bit-reproducible.
//
// Replaces: timm SqueezeExcite.forward inside `self.backbone(x_img)` (experiments/multimodal_v1/train_mm_joint_dualtask.py:154)
// and its autograd backward (:248).
#include "common.cuh"
#include "../../include/teethrt.h"

namespace {

constexpr int TPB = 256;
constexpr int NT = 64;          // images per pass
constexpr int MAX_CC = 32;      // channels per block (multiple of 4)
constexpr int BAR_BYTES = 256;  // workspace header: arrival counter at +0, generation at +128

#ifdef TRT_SE_TIMING
// bring-up instrumentation (never compiled into the shipped library): clock64 deltas of block 0 / thread 0 per phase
__device__ unsigned long long g_se_dbg[16];
#define SE_TICK_INIT long long tick__ = clock64()
#define SE_TICK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long now__ = clock64(); atomicAdd(&g_se_dbg[slot], (unsigned long long)(now__ - tick__)); tick__ = now__; } } while (0)
#else
#define SE_TICK_INIT do {} while (0)
#define SE_TICK(slot) do {} while (0)
#endif

// Sense-reversing grid barrier over co-resident blocks (grid <= SM count, checked by the host).  Thread 0 reads the generation
// BEFORE it arrives (the last arriver bumps it); the other threads wait at the block barrier.  Release / acquire operations at
// GPU scope instead of __threadfence(): that compiles to MEMBAR.SC.GPU + CCTL.IVALL (a sequentially-consistent fence plus an
// L1 flush), three of which per barrier made the barrier cost ~3 us.  Data written by other blocks is read with ld.global.cg.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* count = bar;
    unsigned int* gen = bar + 32;
    unsigned int g, old;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(gen) : "memory");
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(count) : "memory");
    if (old == nblocks - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(count), "r"(0u) : "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gen) : "memory");
    } else {
      unsigned int cur;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(gen) : "memory");
      } while (cur == g);
    }
  }
  __syncthreads();
}

// Sum of the G partial [N][rdp] products for this block's slice of element quads; 16 threads share a quad (each adds every
// 16th partial), the sub-lane-0 thread gets the total.  All threads of the block must call it (full-mask shuffles).
template <typename F>
__device__ __forceinline__ void reduce_partials(const float* __restrict__ part, int G, int total4, int b, F&& finish) {
  const int per = (total4 + G - 1) / G;
  const int q0 = b * per, q1 = min(q0 + per, total4);
  const int sub = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const float4* p4 = reinterpret_cast<const float4*>(part);
  for (int base = q0; base < q1; base += TPB / 16) {
    const int q = base + grp;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < q1) {
#pragma unroll 4
      for (int g = sub; g < G; g += 16) {
        const float4 v = __ldcg(p4 + (size_t)g * total4 + q);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (sub == 0 && q < q1) finish(q, acc);
  }
}

__device__ __forceinline__ void fma44(float (&acc)[4][4], const float4& a, const float4& b) {
  const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
}

// ------------------------------------------------------------------------------------------------ forward
//   s1[n,r]   = br[r] + sum_c (pooled[n,c] / HW) * Wr[r,c]
//   gate[n,c] = sigmoid(be[c] + sum_r silu(s1[n,r]) * We[c,r])
__global__ void __launch_bounds__(TPB) se_fwd_fused_kernel(const float* __restrict__ pooled, float inv_hw,
                                                           const float* __restrict__ Wr, const float* __restrict__ br,
                                                           const float* __restrict__ We, const float* __restrict__ be,
                                                           float* __restrict__ s1, float* __restrict__ gate,
                                                           float* __restrict__ part, unsigned int* bar, int N, int C, int rd, int cc) {
  extern __shared__ __align__(16) float sm[];
  const int rdp = (rd + 3) & ~3, G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
  const int c0 = b * cc, quads = cc >> 2;
  SE_TICK_INIT;
  {
    // ---- phase 1: partial s1 over this block's channels
    float* s_wrT = sm;                 // [cc][rdp]   Wr chunk, transposed, pre-scaled by 1/HW
    float* s_poT = sm + cc * rdp;      // [cc][NT]    pooled sums of the pass, transposed
    for (int i = t; i < rdp * cc; i += TPB) {
      const int r = i / cc, c = i - r * cc;
      s_wrT[c * rdp + r] = (r < rd && c0 + c < C) ? __ldg(Wr + (size_t)r * C + c0 + c) * inv_hw : 0.f;
    }
    for (int n0 = 0; n0 < N; n0 += NT) {
      const int nn = min(NT, N - n0);
      __syncthreads();
      for (int i = t; i < NT * cc; i += TPB) {
        const int n = i / cc, c = i - n * cc;
        s_poT[c * NT + n] = (n < nn && c0 + c < C) ? __ldg(pooled + (size_t)(n0 + n) * C + c0 + c) : 0.f;
      }
      __syncthreads();
      for (int item = t; item < (NT / 4) * (rdp >> 2); item += TPB) {
        const int nb = (item % (NT / 4)) * 4, rb = (item / (NT / 4)) * 4;
        float acc[4][4] = {};
#pragma unroll 4
        for (int c = 0; c < cc; ++c)
          fma44(acc, *reinterpret_cast<const float4*>(s_poT + c * NT + nb), *reinterpret_cast<const float4*>(s_wrT + c * rdp + rb));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (nb + i < nn)
            *reinterpret_cast<float4*>(part + ((size_t)b * N + n0 + nb + i) * rdp + rb) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
  }
  SE_TICK(0);
  grid_barrier(bar, G);
  SE_TICK(1);
  reduce_partials(part, G, N * rdp / 4, b, [&](int q, const float4& v) {
    const int n = (q * 4) / rdp, r0 = (q * 4) - n * rdp;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r0 + j < rd) s1[(size_t)n * rd + r0 + j] = vv[j] + __ldg(br + r0 + j);
  });
  SE_TICK(2);
  grid_barrier(bar, G);
  SE_TICK(3);
  {
    // ---- phase 2: gates of this block's channels
    float* s_a1T = sm;                 // [rdp][NT]   silu(s1) of the pass, transposed
    float* s_weT = sm + rdp * NT;      // [rdp][cc]
    for (int i = t; i < cc * rdp; i += TPB) {
      const int c = i / rdp, r = i - c * rdp;
      s_weT[r * cc + c] = (r < rd && c0 + c < C) ? __ldg(We + (size_t)(c0 + c) * rd + r) : 0.f;
    }
    for (int n0 = 0; n0 < N; n0 += NT) {
      const int nn = min(NT, N - n0);
      __syncthreads();
      for (int i = t; i < NT * rdp; i += TPB) {
        const int n = i / rdp, r = i - n * rdp;
        s_a1T[r * NT + n] = (n < nn && r < rd) ? siluf_(__ldcg(s1 + (size_t)(n0 + n) * rd + r)) : 0.f;
      }
      __syncthreads();
      for (int item = t; item < (NT / 2) * quads; item += TPB) {
        const int c = (item % quads) * 4, n = (item / quads) * 2;
        if (c0 + c >= C) continue;
        const float4 bias = __ldg(reinterpret_cast<const float4*>(be + c0 + c));
        float acc[2][4] = {{bias.x, bias.y, bias.z, bias.w}, {bias.x, bias.y, bias.z, bias.w}};
#pragma unroll 4
        for (int r = 0; r < rd; ++r) {
          const float2 a = *reinterpret_cast<const float2*>(s_a1T + r * NT + n);
          const float4 w = *reinterpret_cast<const float4*>(s_weT + r * cc + c);
          acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
          acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
          acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
          acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (n + i < nn)
            *reinterpret_cast<float4*>(gate + (size_t)(n0 + n + i) * C + c0 + c) =
                make_float4(sigmoidf_(acc[i][0]), sigmoidf_(acc[i][1]), sigmoidf_(acc[i][2]), sigmoidf_(acc[i][3]));
      }
    }
  }
  SE_TICK(4);
}

// ------------------------------------------------------------------------------------------------ backward
//   ds2 = dgate_pre * g * (1 - g);  dWe[c,r] = sum_n ds2[n,c] * silu(s1[n,r]);  dbe[c] = sum_n ds2[n,c]
//   ds1[n,r] = silu'(s1[n,r]) * sum_c ds2[n,c] * We[c,r]                                  (split over the channel chunks)
//   dmean[n,c] = sum_r ds1[n,r] * Wr[r,c];  dWr[r,c] = sum_n ds1[n,r] * pooled[n,c] / HW;  dbr[r] = sum_n ds1[n,r]
//   + (optional) the BatchNorm-backward coefficients of the gated activation from the five per-(image, channel) sums of
//     se_bwd_reduce (what se_bwd_k2_kernel does; same arithmetic)
struct SeBnDev {
  const float* sums;           // [5][N][C]; null = not requested
  const float* rec;            // [4][C]
  const float* gamma;
  float *coef, *dgamma, *dbeta;
  double count;
};

__global__ void __launch_bounds__(TPB) se_bwd_fused_kernel(const float* __restrict__ dgate_pre, const float* __restrict__ gate,
                                                           const float* __restrict__ s1, const float* __restrict__ pooled, float inv_hw,
                                                           const float* __restrict__ Wr, const float* __restrict__ We,
                                                           float* __restrict__ ds2_out, float* __restrict__ ds1, float* __restrict__ dmean,
                                                           float* __restrict__ dWr, float* __restrict__ dbr, float* __restrict__ dWe,
                                                           float* __restrict__ dbe, float* __restrict__ part, unsigned int* bar,
                                                           int N, int C, int rd, int cc, const SeBnDev bn) {
  extern __shared__ __align__(16) float sm[];
  const int rdp = (rd + 3) & ~3, G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
  const int c0 = b * cc, quads = cc >> 2, rgroups = rdp >> 2;
  SE_TICK_INIT;
  {
    // ---- phase 1
    float* s_a1 = sm;                       // [NT][rdp]  silu(s1)
    float* s_d2 = s_a1 + NT * rdp;          // [NT][cc]
    float* s_d2T = s_d2 + NT * cc;          // [cc][NT]
    float* s_we = s_d2T + cc * NT;          // [cc][rdp]
    for (int i = t; i < cc * rdp; i += TPB) {
      const int c = i / rdp, r = i - c * rdp;
      s_we[i] = (r < rd && c0 + c < C) ? __ldg(We + (size_t)(c0 + c) * rd + r) : 0.f;
    }
    const int wc = (t % quads) * 4, wr = (t / quads) * 4;        // dWe tile: 4 channels x 4 reduced channels
    const bool w_active = t / quads < rgroups;
    float acc_w[4][4] = {};
    float acc_b = 0.f;
    for (int n0 = 0; n0 < N; n0 += NT) {
      const int nn = min(NT, N - n0);
      __syncthreads();
      for (int i = t; i < NT * rdp; i += TPB) {
        const int n = i / rdp, r = i - n * rdp;
        s_a1[i] = (n < nn && r < rd) ? siluf_(__ldg(s1 + (size_t)(n0 + n) * rd + r)) : 0.f;
      }
      for (int i = t; i < NT * cc; i += TPB) {
        const int n = i / cc, c = i - n * cc;
        float d = 0.f;
        if (n < nn && c0 + c < C) {
          const size_t idx = (size_t)(n0 + n) * C + c0 + c;
          const float g = __ldg(gate + idx);
          d = __ldg(dgate_pre + idx) * g * (1.f - g);
          if (ds2_out) ds2_out[idx] = d;
        }
        s_d2[i] = d;
        s_d2T[c * NT + n] = d;
      }
      __syncthreads();
      if (w_active) {
#pragma unroll 4
        for (int n = 0; n < NT; ++n)
          fma44(acc_w, *reinterpret_cast<const float4*>(s_d2 + n * cc + wc), *reinterpret_cast<const float4*>(s_a1 + n * rdp + wr));
      }
      if (t < cc)
        for (int n = 0; n < NT; ++n) acc_b += s_d2T[t * NT + n];
      for (int item = t; item < (NT / 4) * rgroups; item += TPB) {
        const int nb = (item % (NT / 4)) * 4, rb = (item / (NT / 4)) * 4;
        float acc[4][4] = {};
#pragma unroll 4
        for (int c = 0; c < cc; ++c)
          fma44(acc, *reinterpret_cast<const float4*>(s_d2T + c * NT + nb), *reinterpret_cast<const float4*>(s_we + c * rdp + rb));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (nb + i < nn)
            *reinterpret_cast<float4*>(part + ((size_t)b * N + n0 + nb + i) * rdp + rb) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
    if (w_active) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + wc + i < C && wr + j < rd) dWe[(size_t)(c0 + wc + i) * rd + wr + j] = acc_w[i][j];
    }
    if (t < cc && c0 + t < C) dbe[c0 + t] = acc_b;
  }
  SE_TICK(8);
  grid_barrier(bar, G);
  SE_TICK(9);
  reduce_partials(part, G, N * rdp / 4, b, [&](int q, const float4& v) {
    const int n = (q * 4) / rdp, r0 = (q * 4) - n * rdp;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r0 + j < rd) ds1[(size_t)n * rd + r0 + j] = vv[j] * silu_gradf_(__ldg(s1 + (size_t)n * rd + r0 + j));
  });
  SE_TICK(10);
  grid_barrier(bar, G);
  SE_TICK(11);
  {
    // ---- phase 2
    float* s_d1 = sm;                       // [NT][rdp]
    float* s_d1T = s_d1 + NT * rdp;         // [rdp][NT]
    float* s_wr = s_d1T + rdp * NT;         // [rdp][cc]
    float* s_po = s_wr + rdp * cc;          // [NT][cc]   pooled / HW
    for (int i = t; i < rdp * cc; i += TPB) {
      const int r = i / cc, c = i - r * cc;
      s_wr[i] = (r < rd && c0 + c < C) ? __ldg(Wr + (size_t)r * C + c0 + c) : 0.f;
    }
    // dmean: thread = (2 images, 4 channels); dWr: thread = (4 reduced channels, 4 channels).  With narrow chunks the two
    // use disjoint halves of the block, so they run side by side
    const int dm_items = (NT / 2) * quads;
    const bool dm_active = t < dm_items;
    const int dquad = t % quads, dpair = t / quads;
    const int tw = (dm_items <= TPB / 2 && quads * rgroups <= TPB / 2) ? ((t + TPB / 2) % TPB) : t;
    const int wq = (tw % quads) * 4, wr = (tw / quads) * 4;
    const bool w_active = tw / quads < rgroups;
    float acc_w[4][4] = {};
    float acc_b = 0.f;
    double bn_g[4] = {0, 0, 0, 0}, bn_gx[4] = {0, 0, 0, 0};
    for (int n0 = 0; n0 < N; n0 += NT) {
      const int nn = min(NT, N - n0);
      __syncthreads();
      for (int i = t; i < NT * rdp; i += TPB) {
        const int n = i / rdp, r = i - n * rdp;
        const float d = (n < nn && r < rd) ? __ldcg(ds1 + (size_t)(n0 + n) * rd + r) : 0.f;
        s_d1[i] = d;
        s_d1T[r * NT + n] = d;
      }
      for (int i = t; i < NT * cc; i += TPB) {
        const int n = i / cc, c = i - n * cc;
        s_po[i] = (n < nn && c0 + c < C) ? __ldg(pooled + (size_t)(n0 + n) * C + c0 + c) * inv_hw : 0.f;
      }
      __syncthreads();
      if (dm_active && c0 + dquad * 4 < C) {
        const int dc = dquad * 4, dn = dpair * 2;
        float acc[2][4] = {};
#pragma unroll 4
        for (int r = 0; r < rd; ++r) {
          const float2 a = *reinterpret_cast<const float2*>(s_d1T + r * NT + dn);
          const float4 w = *reinterpret_cast<const float4*>(s_wr + r * cc + dc);
          acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
          acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
          acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
          acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (dn + i < nn)
            *reinterpret_cast<float4*>(dmean + (size_t)(n0 + dn + i) * C + c0 + dc) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (bn.sums) {
          const size_t NC = (size_t)N * C;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (dn + i < nn) {
              const size_t o = (size_t)(n0 + dn + i) * C + c0 + dc;
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(gate + o));
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
        for (int n = 0; n < NT; ++n)
          fma44(acc_w, *reinterpret_cast<const float4*>(s_d1 + n * rdp + wr), *reinterpret_cast<const float4*>(s_po + n * cc + wq));
      }
      if (b == 0 && t < rd)
        for (int n = 0; n < NT; ++n) acc_b += s_d1T[t * NT + n];
    }
    if (w_active) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (wr + i < rd && c0 + wq + j < C) dWr[(size_t)(wr + i) * C + c0 + wq + j] = acc_w[i][j];
    }
    if (b == 0 && t < rd) dbr[t] = acc_b;
    if (bn.sums) {
      __syncthreads();
      double* s_d = reinterpret_cast<double*>(sm);               // [2][NT/2][cc]
      if (dm_active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s_d[dpair * cc + dquad * 4 + j] = bn_g[j];
          s_d[(NT / 2 + dpair) * cc + dquad * 4 + j] = bn_gx[j];
        }
      }
      __syncthreads();
      if (t < cc && c0 + t < C) {
        double sg = 0, sgx = 0;
        for (int q = 0; q < NT / 2; ++q) { sg += s_d[q * cc + t]; sgx += s_d[(NT / 2 + q) * cc + t]; }
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
  SE_TICK(12);
}

// channels per block: 16 while that needs no more blocks than SMs (all blocks must be co-resident for the grid barrier)
int pick_cc(int C) {
  int cc = 16;
  while ((C + cc - 1) / cc > trt_num_sms() && cc < MAX_CC) cc += 4;
  return cc;
}

size_t part_floats(int N, int C, int rd) {
  const int cc = pick_cc(C), G = (C + cc - 1) / cc, rdp = (rd + 3) & ~3;
  return (size_t)G * N * rdp;
}

}  // namespace

#ifdef TRT_SE_TIMING
extern "C" void trt_se_debug_read(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (out16) cudaMemcpyFromSymbol(out16, g_se_dbg, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_se_dbg, z, sizeof(z)); }
}
#endif

extern "C" size_t trt_se_workspace_bytes(int N, int C, int rd) {
  if (N <= 0 || C <= 0 || rd <= 0) return 0;
  return BAR_BYTES + part_floats(N, C, rd) * sizeof(float);
}

extern "C" int trt_se_fwd_fused(const float* pooled_sum, float inv_hw, const float* Wr, const float* br, const float* We,
                                const float* be, float* s1, float* gate, void* workspace, size_t ws_bytes, int N, int C, int rd,
                                cudaStream_t stream) {
  TRT_REQUIRE(pooled_sum && Wr && br && We && be && s1 && gate && workspace && N > 0 && C > 0 && rd > 0, "trt_se_fwd_fused: bad argument");
  TRT_REQUIRE(C % 4 == 0 && (((uintptr_t)gate | (uintptr_t)be | (uintptr_t)workspace) & 15) == 0, "trt_se_fwd_fused: C %% 4 and 16-byte aligned gate / be / workspace");
  TRT_REQUIRE(ws_bytes >= trt_se_workspace_bytes(N, C, rd), "trt_se_fwd_fused: workspace too small (%zu < %zu)", ws_bytes, trt_se_workspace_bytes(N, C, rd));
  const int cc = pick_cc(C), G = (C + cc - 1) / cc, rdp = (rd + 3) & ~3;
  TRT_REQUIRE(G <= trt_num_sms(), "trt_se_fwd_fused: %d channels need more than one block per SM", C);
  const size_t a = (size_t)cc * rdp + (size_t)cc * NT, bq = (size_t)rdp * NT + (size_t)rdp * cc;
  const size_t smem = (a > bq ? a : bq) * sizeof(float);
  TRT_REQUIRE(smem <= 200 * 1024, "trt_se_fwd_fused: rd %d too large", rd);
  TRT_CUDA(cudaFuncSetAttribute(se_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  unsigned int* bar = reinterpret_cast<unsigned int*>(workspace);
  float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + BAR_BYTES);
  se_fwd_fused_kernel<<<G, TPB, smem, stream>>>(pooled_sum, inv_hw, Wr, br, We, be, s1, gate, part, bar, N, C, rd, cc);
  return trt_check_launch("trt_se_fwd_fused");
}

extern "C" int trt_se_bwd_fused(const float* dgate_pre, const float* gate, const float* s1, const float* pooled_sum, float inv_hw,
                                const float* Wr, const float* We, float* ds2, float* ds1, float* dmean, float* dWr, float* dbr,
                                float* dWe, float* dbe, const trt_se_bn_t* bn_host, void* workspace, size_t ws_bytes, int N, int C,
                                int rd, cudaStream_t stream) {
  TRT_REQUIRE(dgate_pre && gate && s1 && pooled_sum && Wr && We && ds1 && dmean && dWr && dbr && dWe && dbe && workspace,
              "trt_se_bwd_fused: null pointer");
  TRT_REQUIRE(N > 0 && C > 0 && rd > 0 && C % 4 == 0 && (((uintptr_t)gate | (uintptr_t)dmean | (uintptr_t)workspace) & 15) == 0,
              "trt_se_bwd_fused: C %% 4 and 16-byte aligned gate / dmean / workspace");
  TRT_REQUIRE(ws_bytes >= trt_se_workspace_bytes(N, C, rd), "trt_se_bwd_fused: workspace too small (%zu < %zu)", ws_bytes, trt_se_workspace_bytes(N, C, rd));
  SeBnDev bn = {};
  if (bn_host) {
    TRT_REQUIRE(bn_host->sums && bn_host->rec && bn_host->gamma && bn_host->coef && bn_host->dgamma && bn_host->dbeta && bn_host->count > 0,
                "trt_se_bwd_fused: incomplete BatchNorm-backward record");
    TRT_REQUIRE((((uintptr_t)bn_host->sums) & 15) == 0, "trt_se_bwd_fused: sums must be 16-byte aligned");
    bn.sums = bn_host->sums; bn.rec = bn_host->rec; bn.gamma = bn_host->gamma; bn.coef = bn_host->coef;
    bn.dgamma = bn_host->dgamma; bn.dbeta = bn_host->dbeta; bn.count = bn_host->count;
  }
  const int cc = pick_cc(C), G = (C + cc - 1) / cc, rdp = (rd + 3) & ~3;
  TRT_REQUIRE(G <= trt_num_sms(), "trt_se_bwd_fused: %d channels need more than one block per SM", C);
  TRT_REQUIRE(rd <= TPB, "trt_se_bwd_fused: rd %d > %d", rd, TPB);
  const size_t p1 = (size_t)NT * rdp + 2 * (size_t)NT * cc + (size_t)cc * rdp;
  const size_t p2 = 2 * (size_t)NT * rdp + (size_t)rdp * cc + (size_t)NT * cc;
  size_t smem = (p1 > p2 ? p1 : p2) * sizeof(float);
  if (smem < (size_t)NT * cc * sizeof(double)) smem = (size_t)NT * cc * sizeof(double);
  TRT_REQUIRE(smem <= 200 * 1024, "trt_se_bwd_fused: rd %d too large", rd);
  TRT_CUDA(cudaFuncSetAttribute(se_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  unsigned int* bar = reinterpret_cast<unsigned int*>(workspace);
  float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + BAR_BYTES);
  se_bwd_fused_kernel<<<G, TPB, smem, stream>>>(dgate_pre, gate, s1, pooled_sum, inv_hw, Wr, We, ds2, ds1, dmean, dWr, dbr, dWe, dbe,
                                               part, bar, N, C, rd, cc, bn);
  return trt_check_launch("trt_se_bwd_fused");
}
