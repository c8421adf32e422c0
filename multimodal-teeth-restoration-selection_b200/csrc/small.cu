// Small fused kernels of the hot path:
//   * MIL gated-attention pooling forward/backward (one CTA per bag: instance matrix staged once in shared memory with
//     coalesced float4 reads, warp-shuffle reductions for the gate dot products, the softmax over instances and the
//     weighted sum).  Replaces AttentionMIL.forward, experiments/vision_v2/train_mil_attention_v1.py:124-130 and
//     MILAttention.forward, ui/gradio_app/infer_mil.py:62-68.
//   * tabular MLP + late-fusion dual heads + dual BCE loss forward/backward in ONE CTA (the whole problem is a few
//     hundred KFLOP; what matters is launch count).  Replaces `self.tab`, `self.fusion`, `cls_head`, `reg_head`
//     (experiments/multimodal_v1/train_mm_joint_dualtask.py:140-159) and the loss (:176-179, :244-247).
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int KC = 16;   // instances processed per register chunk

// ================================================================================================= MIL attention
// Forward = two launches.  (1) scores: grid (bag, hidden-slice); the bag's instance matrix H[K][D] is staged once in shared
// memory with coalesced float4 reads; a warp owns one hidden unit j and computes BOTH gate rows <V_j, H_k> and <U_j, H_k>
// for 16 instances per pass (each H value read from shared memory feeds two FMAs), finishes them with warp shuffles, applies
// tanh * sigmoid and adds w_j * gate into the bag's K scores (shared-memory atomics, then one global atomic per instance).
// (2) pool: one CTA per bag: softmax over the K scores (warp shuffles) and M = sum_k alpha_k H_k with coalesced reads.
// The hidden slice count adapts to the batch so that small batches (6 training bags) still fill the GPU.
// HU = hidden units per warp: 1 when the batch is small (more blocks), 4 when it is large (each instance value read from
// shared memory then feeds 8 gate rows = 32 FMAs, so the loop is FMA- rather than shared-memory-bound).
template <int HU>
__global__ void __launch_bounds__(TPB) mil_score_kernel(const float* __restrict__ H, const float* __restrict__ Vw,
                                                        const float* __restrict__ Vb, const float* __restrict__ Uw,
                                                        const float* __restrict__ Ub, const float* __restrict__ ww,
                                                        float* __restrict__ score, float* __restrict__ gV,
                                                        float* __restrict__ gU, int K, int D, int hid, int JH) {
  constexpr int KCH = HU == 1 ? 16 : 8;      // instances per register pass
  extern __shared__ __align__(16) float sm[];
  float* s_H = sm;                       // [K][D]
  float* s_sc = s_H + (size_t)K * D;     // [K] partial scores of this hidden slice
  const int b = blockIdx.x, j0 = blockIdx.y * JH, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4* Hg = reinterpret_cast<const float4*>(H + (size_t)b * K * D);
  const int D4 = D / 4;
  for (int i = threadIdx.x; i < K * D4; i += TPB) reinterpret_cast<float4*>(s_H)[i] = __ldg(Hg + i);
  for (int k = threadIdx.x; k < K; k += TPB) s_sc[k] = 0.f;
  __syncthreads();
  const int jend = min(j0 + JH, hid);
  for (int jb = j0 + warp * HU; jb < jend; jb += (TPB / 32) * HU) {
    for (int k0 = 0; k0 < K; k0 += KCH) {
      float av[HU][KCH], au[HU][KCH];
#pragma unroll
      for (int u = 0; u < HU; ++u)
#pragma unroll
        for (int k = 0; k < KCH; ++k) av[u][k] = au[u][k] = 0.f;
      for (int d = lane; d < D4; d += 32) {
        float4 v4[HU], u4[HU];
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          const int j = min(jb + u, hid - 1);            // clamped rows are computed and dropped
          v4[u] = __ldg(reinterpret_cast<const float4*>(Vw + (size_t)j * D) + d);
          u4[u] = __ldg(reinterpret_cast<const float4*>(Uw + (size_t)j * D) + d);
        }
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
          if (k0 + k < K) {
            const float4 h = reinterpret_cast<const float4*>(s_H + (size_t)(k0 + k) * D)[d];
#pragma unroll
            for (int u = 0; u < HU; ++u) {
              av[u][k] = fmaf(v4[u].x, h.x, fmaf(v4[u].y, h.y, fmaf(v4[u].z, h.z, fmaf(v4[u].w, h.w, av[u][k]))));
              au[u][k] = fmaf(u4[u].x, h.x, fmaf(u4[u].y, h.y, fmaf(u4[u].z, h.z, fmaf(u4[u].w, h.w, au[u][k]))));
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < HU; ++u) {
        const int j = jb + u;
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
          if (k0 + k < K && j < jend) {                    // uniform across the warp
            const float sv = warp_sum(av[u][k]) + Vb[j], su = warp_sum(au[u][k]) + Ub[j];
            if (lane == 0) {
              const float tv = tanhf(sv), sg = 1.0f / (1.0f + expf(-su));
              if (gV) { gV[((size_t)b * K + k0 + k) * hid + j] = tv; gU[((size_t)b * K + k0 + k) * hid + j] = sg; }
              atomicAdd(s_sc + k0 + k, ww[j] * tv * sg);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += TPB) atomicAdd(score + (size_t)b * K + k, s_sc[k]);
}

// ---- tensor-core path of the score pass: the [B*K, D] x [D, 2*hid] projection is a dense contraction.
// fp32 accuracy is kept by splitting both operands into bf16 pairs x = hi + lo and taking three of the four products,
//     x.w ~ hi.Whi + lo.Whi + hi.Wlo      (error ~2^-18 relative per product, the dropped lo.Wlo term)
// as ONE tcgen05 GEMM over a 3*D-long reduction: A' = [hi | lo | hi] (stored once as [hi | lo]; the third segment re-reads
// the first through the TMA coordinates), B' = [Whi | Whi | Wlo], weight rows interleaved (V_j, U_j) so the epilogue sees a
// gate's two pre-activations in adjacent accumulator columns and reduces tanh * sigmoid * w to one score per instance.
__global__ void __launch_bounds__(TPB) mil_split_h_kernel(const float4* __restrict__ H, uint2* __restrict__ out, size_t rows, int D4) {
  // out row layout: [hi (D bf16) | lo (D bf16)]; one thread = 4 consecutive features
  for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < rows * D4; i += (size_t)gridDim.x * TPB) {
    const size_t r = i / D4;
    const int d = (int)(i - r * D4);
    const float4 x = __ldg(H + i);
    const float h0 = bf16_round(x.x), h1 = bf16_round(x.y), h2 = bf16_round(x.z), h3 = bf16_round(x.w);
    uint2 hi, lo;
    hi.x = pack_bf16(h0, h1); hi.y = pack_bf16(h2, h3);
    lo.x = pack_bf16(x.x - h0, x.y - h1); lo.y = pack_bf16(x.z - h2, x.w - h3);
    out[r * 2 * D4 + d] = hi;
    out[r * 2 * D4 + D4 + d] = lo;
  }
}
// W' [2*hid][3*D] bf16: row 2j = V_j, row 2j+1 = U_j, columns [Whi | Whi | Wlo]; bias2[2j] = Vb_j, bias2[2j+1] = Ub_j
__global__ void __launch_bounds__(TPB) mil_split_w_kernel(const float* __restrict__ Vw, const float* __restrict__ Uw,
                                                          const float* __restrict__ Vb, const float* __restrict__ Ub,
                                                          __nv_bfloat16* __restrict__ out, float* __restrict__ bias2, int hid, int D) {
  const int n = blockIdx.x, j = n >> 1;
  const float* w = (n & 1) ? Uw + (size_t)j * D : Vw + (size_t)j * D;
  __nv_bfloat16* o = out + (size_t)n * 3 * D;
  for (int d = threadIdx.x; d < D; d += TPB) {
    const float x = __ldg(w + d), h = bf16_round(x);
    o[d] = __float2bfloat16_rn(h);
    o[D + d] = __float2bfloat16_rn(h);
    o[2 * D + d] = __float2bfloat16_rn(x - h);
  }
  if (threadIdx.x == 0) bias2[n] = (n & 1) ? Ub[j] : Vb[j];
}

// A holds the raw scores on entry and the softmax weights on exit
__global__ void __launch_bounds__(TPB) mil_pool_kernel(const float* __restrict__ H, const float* __restrict__ wb,
                                                       float* __restrict__ A, float* __restrict__ M, int K, int D) {
  extern __shared__ __align__(16) float sm[];
  float* s_att = sm;   // [K]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < K; k += TPB) s_att[k] = A[(size_t)b * K + k] + wb[0];
  __syncthreads();
  if (warp == 0) {   // softmax over the K instances (dim=1 of [B,K]; dim=0 of the twin's single bag)
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, s_att[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < K; k += 32) sum += expf(s_att[k] - mx);
    sum = warp_sum(sum);
    for (int k = lane; k < K; k += 32) {
      const float a = expf(s_att[k] - mx) / sum;
      s_att[k] = a;
      A[(size_t)b * K + k] = a;
    }
  }
  __syncthreads();
  const float4* Hg = reinterpret_cast<const float4*>(H + (size_t)b * K * D);
  const int D4 = D / 4;
  for (int d = threadIdx.x; d < D4; d += TPB) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int k = 0; k < K; ++k) {
      const float a = s_att[k];
      const float4 h = __ldg(Hg + (size_t)k * D4 + d);
      acc.x = fmaf(a, h.x, acc.x); acc.y = fmaf(a, h.y, acc.y); acc.z = fmaf(a, h.z, acc.z); acc.w = fmaf(a, h.w, acc.w);
    }
    reinterpret_cast<float4*>(M + (size_t)b * D)[d] = acc;
  }
}

// Backward = three launches:
//   (1) per bag: dalpha_k = <dM, H_k>, softmax backward, gate gradients; gV / gU are OVERWRITTEN with dv / du [B,K,hid]
//   (2) weight gradients as a blocked product over ALL bags: dW[j][d] += sum_{b,k} dvu[b,k,j] H[b,k,d]; each output element
//       is owned by one block (no atomics; the old kernel issued 2*hid*D global atomics per bag)
//   (3) dH[b,k,d] = alpha_k dM[d] + sum_j dv[b,k,j] Vw[j][d] + du[b,k,j] Uw[j][d]; grid (bag, D-slice)
__global__ void __launch_bounds__(TPB) mil_bwd_gate_kernel(const float* __restrict__ dM, const float* __restrict__ H,
                                                           const float* __restrict__ A, float* __restrict__ gV,
                                                           float* __restrict__ gU, const float* __restrict__ ww,
                                                           float* __restrict__ dVb, float* __restrict__ dUb,
                                                           float* __restrict__ dww, float* __restrict__ dwb, int K, int D,
                                                           int hid) {
  extern __shared__ __align__(16) float sm[];
  float* s_da = sm;   // [K]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D4 = D / 4;
  const float4* Hg = reinterpret_cast<const float4*>(H + (size_t)b * K * D);
  const float4* dMg = reinterpret_cast<const float4*>(dM + (size_t)b * D);
  for (int k = warp; k < K; k += TPB / 32) {   // dalpha_k = <dM, H_k>
    float acc = 0.f;
    for (int d = lane; d < D4; d += 32) {
      const float4 m = __ldg(dMg + d), h = __ldg(Hg + (size_t)k * D4 + d);
      acc = fmaf(m.x, h.x, fmaf(m.y, h.y, fmaf(m.z, h.z, fmaf(m.w, h.w, acc))));
    }
    acc = warp_sum(acc);
    if (lane == 0) s_da[k] = acc;
  }
  __syncthreads();
  if (warp == 0) {   // softmax backward
    float dot = 0.f;
    for (int k = lane; k < K; k += 32) dot = fmaf(A[(size_t)b * K + k], s_da[k], dot);
    dot = warp_sum(dot);
    float sb = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float da = A[(size_t)b * K + k] * (s_da[k] - dot);
      s_da[k] = da;
      sb += da;
    }
    sb = warp_sum(sb);
    if (lane == 0) atomicAdd(dwb, sb);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < hid; j += TPB) {
    float aw = 0.f, av = 0.f, au = 0.f;
    const float wj = ww[j];
    for (int k = 0; k < K; ++k) {
      const size_t idx = ((size_t)b * K + k) * hid + j;
      const float tv = gV[idx], su = gU[idx];
      const float da = s_da[k];
      aw = fmaf(da, tv * su, aw);
      const float dg = da * wj;
      const float dv = dg * su * (1.f - tv * tv), du = dg * tv * su * (1.f - su);
      gV[idx] = dv;
      gU[idx] = du;
      av += dv; au += du;
    }
    atomicAdd(dww + j, aw);
    atomicAdd(dVb + j, av);
    atomicAdd(dUb + j, au);
  }
}

constexpr int MW_R = 16;     // gate rows per block
constexpr int MW_C = 256;    // feature columns per block (one per thread)
constexpr int MW_CH = 32;    // (bag, instance) rows per shared-memory chunk
__global__ void __launch_bounds__(TPB) mil_bwd_weight_kernel(const float* __restrict__ H, const float* __restrict__ dv,
                                                             const float* __restrict__ du, float* __restrict__ dVw,
                                                             float* __restrict__ dUw, int BK, int D, int hid) {
  __shared__ __align__(16) float s_h[MW_CH][MW_C];
  __shared__ __align__(16) float s_g[MW_CH][MW_R];
  const int nb = gridDim.x / 2, d0 = blockIdx.y * MW_C, t = threadIdx.x;   // first half of the row blocks = V, second = U
  const bool is_u = (int)blockIdx.x >= nb;
  const float* g = is_u ? du : dv;
  const int j0 = ((int)blockIdx.x - (is_u ? nb : 0)) * MW_R;
  float acc[MW_R];
#pragma unroll
  for (int i = 0; i < MW_R; ++i) acc[i] = 0.f;
  for (int r0 = 0; r0 < BK; r0 += MW_CH) {
    __syncthreads();
    for (int i = t; i < MW_CH * MW_C; i += TPB) {
      const int r = i / MW_C, c = i % MW_C;
      s_h[r][c] = (r0 + r < BK && d0 + c < D) ? __ldg(H + (size_t)(r0 + r) * D + d0 + c) : 0.f;
    }
    for (int i = t; i < MW_CH * MW_R; i += TPB) {
      const int r = i / MW_R, c = i % MW_R;
      s_g[r][c] = (r0 + r < BK && j0 + c < hid) ? __ldg(g + (size_t)(r0 + r) * hid + j0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < MW_CH; ++r) {
      const float h = s_h[r][t];
#pragma unroll
      for (int i4 = 0; i4 < MW_R / 4; ++i4) {
        const float4 gg = *reinterpret_cast<const float4*>(&s_g[r][4 * i4]);
        acc[4 * i4] = fmaf(gg.x, h, acc[4 * i4]); acc[4 * i4 + 1] = fmaf(gg.y, h, acc[4 * i4 + 1]);
        acc[4 * i4 + 2] = fmaf(gg.z, h, acc[4 * i4 + 2]); acc[4 * i4 + 3] = fmaf(gg.w, h, acc[4 * i4 + 3]);
      }
    }
  }
  float* dst = is_u ? dUw : dVw;
  if (d0 + t < D) {
#pragma unroll
    for (int i = 0; i < MW_R; ++i)
      if (j0 + i < hid) dst[(size_t)(j0 + i) * D + d0 + t] += acc[i];      // this block owns these elements
  }
}

constexpr int MH_C = 128;    // feature columns per block of the dH kernel
__global__ void __launch_bounds__(TPB) mil_bwd_dh_kernel(const float* __restrict__ dM, const float* __restrict__ A,
                                                         const float* __restrict__ dv, const float* __restrict__ du,
                                                         const float* __restrict__ Vw, const float* __restrict__ Uw,
                                                         float* __restrict__ dH, int K, int D, int hid) {
  extern __shared__ __align__(16) float sm[];
  float* s_dv = sm;                         // [K][hid]
  float* s_du = s_dv + (size_t)K * hid;     // [K][hid]
  const int b = blockIdx.x, d0 = blockIdx.y * MH_C;
  for (int i = threadIdx.x; i < K * hid; i += TPB) {
    s_dv[i] = dv[(size_t)b * K * hid + i];
    s_du[i] = du[(size_t)b * K * hid + i];
  }
  __syncthreads();
  const int dl = threadIdx.x % MH_C, kh = threadIdx.x / MH_C;      // two instance halves per column
  const int d = d0 + dl;
  if (d >= D) return;
  const float dm = dM[(size_t)b * D + d];
  for (int k0 = kh * 8; k0 < K; k0 += 16) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 2
    for (int j = 0; j < hid; ++j) {
      const float wv = __ldg(Vw + (size_t)j * D + d), wu = __ldg(Uw + (size_t)j * D + d);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k0 + k < K) acc[k] = fmaf(s_dv[(size_t)(k0 + k) * hid + j], wv, fmaf(s_du[(size_t)(k0 + k) * hid + j], wu, acc[k]));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k0 + k < K) dH[((size_t)b * K + k0 + k) * D + d] = fmaf(A[(size_t)b * K + k0 + k], dm, acc[k]);
  }
}

// ================================================================================================= tab MLP + heads + loss
// counter-based uniform in [0,1): splitmix64 of (seed, stream, index)
__device__ __forceinline__ float urand(unsigned long long seed, unsigned long long stream_id, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (stream_id * 0x100000001B3ull + idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float keep_scale(float p, unsigned long long seed, unsigned long long stream_id, unsigned long long idx) {
  if (p <= 0.f) return 1.f;
  return urand(seed, stream_id, idx) >= p ? 1.f / (1.f - p) : 0.f;
}

struct TabParams {
  int B, T, Hd, F;          // batch, tab_in, tab_hidden, image feature dim
  int train;                // train-mode BN1d (batch statistics) + dropout
  float drop_p, bn_eps, bn_momentum, alpha, beta;
  unsigned long long seed;  // dropout seed
  const unsigned long long* step;   // device step counter mixed into the seed (may be null)
  // inputs
  const float* feat;   // [B,F]
  const float* xtab;   // [B,T]
  // parameters
  const float *W0, *b0, *bn_g, *bn_b, *W1, *b1, *cls_w, *cls_b, *reg_w, *reg_b;
  float *bn_rm, *bn_rv;
  long long* bn_nbt;
  // optional loss inputs
  const float *y_hard, *y_soft, *sample_w;
  // outputs
  float *logit, *reg, *loss;        // [B], [B], [1]
  float *dlogit, *dreg;             // [B] gradient of the loss w.r.t. the two head outputs (written when targets given)
  // scratch (global, caller provided): z0 [B,Hd], a1 [B,Hd], ft [B,Hd], bnstat [2*Hd]
  float* scratch;
  int big;                          // 1: [B][Hd] arrays stay in the global scratch (batch too large for shared memory)
};

constexpr int TAB_TPB = 1024;   // threads of the one-block tab MLP kernels: every phase is a short latency chain, so 32 warps
                                // (not 8) keep the LSU busy - 256 threads measured 38 us forward / 74 us backward under ncu
// Tab MLP in ONE block with the whole problem in shared memory (z0 / a1 [B][Hd] in place, W1 padded): the earlier version
// walked global scratch with dependent loads (170 us for ~0.6 MFLOP).  Global scratch still receives z0, a1, ft and the BN
// statistics for the backward pass.
__global__ void __launch_bounds__(TAB_TPB, 1) tab_heads_fwd_kernel(const TabParams p) {
  extern __shared__ __align__(16) float tsm[];
  const int B = p.B, T = p.T, Hd = p.Hd;
  const bool big = p.big != 0;             // batch too large for shared memory: z0 / a1 are read back from the global scratch
  float* s_w1 = tsm + (big ? 0 : (size_t)B * Hd);      // [Hd][Hd+1]
  float* s_part = s_w1 + (size_t)Hd * (Hd + 1);   // [2][TAB_TPB] partial sums of the BatchNorm1d statistics
  float* s_stat = s_part + 2 * TAB_TPB;        // mean[Hd], rstd[Hd]
  float* z0 = p.scratch;
  float* a1 = z0 + (size_t)B * Hd;
  float* ft = a1 + (size_t)B * Hd;
  float* bnstat = ft + (size_t)B * Hd;   // mean[Hd], rstd[Hd]
  float* s_z = big ? z0 : tsm;             // [B][Hd]   z0 ...
  float* s_a = big ? a1 : tsm;             // ... then a1 (in place when in shared memory)
  pdl_launch_dependents();
  pdl_wait();
  const unsigned long long seed = p.seed + (p.step ? *p.step * 0x9E3779B97F4A7C15ull : 0ull);
  const int t = threadIdx.x;
  for (int i = t; i < Hd * Hd; i += TAB_TPB) s_w1[(i / Hd) * (Hd + 1) + (i % Hd)] = __ldg(p.W1 + i);
  for (int i = t; i < B * Hd; i += TAB_TPB) {
    const int b = i / Hd, j = i % Hd;
    float acc = p.b0[j];
    for (int tt = 0; tt < T; ++tt) acc = fmaf(__ldg(p.xtab + b * T + tt), __ldg(p.W0 + j * T + tt), acc);
    s_z[i] = acc;
    if (!big) z0[i] = acc;
  }
  __syncthreads();
  // BatchNorm1d statistics: thread = (feature j, batch slice); two-pass (mean, then centred sum of squares) like torch
  const int parts = TAB_TPB / Hd > 0 ? TAB_TPB / Hd : 1;
  const int j = t % Hd, part = t / Hd;
  if (p.train) {
    float sacc = 0.f;
    if (part < parts) for (int b = part; b < B; b += parts) sacc += s_z[b * Hd + j];
    s_part[t] = sacc;
    __syncthreads();
    if (t < Hd) {
      float m = 0.f;
      for (int q = 0; q < parts; ++q) m += s_part[q * Hd + t];
      s_stat[t] = m / B;
    }
    __syncthreads();
    float vacc = 0.f;
    if (part < parts) {
      const float mean = s_stat[j];
      for (int b = part; b < B; b += parts) { const float d = s_z[b * Hd + j] - mean; vacc = fmaf(d, d, vacc); }
    }
    s_part[TAB_TPB + t] = vacc;
    __syncthreads();
    if (t < Hd) {
      float v = 0.f;
      for (int q = 0; q < parts; ++q) v += s_part[TAB_TPB + q * Hd + t];
      const float mean = s_stat[t], var = v / B;
      s_stat[Hd + t] = rsqrtf(var + p.bn_eps);
      p.bn_rm[t] = (1.f - p.bn_momentum) * p.bn_rm[t] + p.bn_momentum * mean;
      p.bn_rv[t] = (1.f - p.bn_momentum) * p.bn_rv[t] + p.bn_momentum * (B > 1 ? v / (B - 1) : var);
      if (t == 0 && p.bn_nbt) *p.bn_nbt += 1;
    }
  } else if (t < Hd) {
    s_stat[t] = p.bn_rm[t];
    s_stat[Hd + t] = rsqrtf(p.bn_rv[t] + p.bn_eps);
  }
  __syncthreads();
  if (t < 2 * Hd) bnstat[t] = s_stat[t];
  for (int i = t; i < B * Hd; i += TAB_TPB) {
    const int jj = i % Hd;
    float v = (s_z[i] - s_stat[jj]) * s_stat[Hd + jj] * p.bn_g[jj] + p.bn_b[jj];
    v = fmaxf(v, 0.f);
    if (p.train) v *= keep_scale(p.drop_p, seed, 1, i);
    s_a[i] = v;
    if (!big) a1[i] = v;
  }
  __syncthreads();
  for (int i = t; i < B * Hd; i += TAB_TPB) {
    const int b = i / Hd, jj = i % Hd;
    float acc = p.b1[jj];
    const float* arow = s_a + b * Hd;
    const float* wrow = s_w1 + jj * (Hd + 1);
#pragma unroll 8
    for (int tt = 0; tt < Hd; ++tt) acc = fmaf(arow[tt], wrow[tt], acc);
    ft[i] = fmaxf(acc, 0.f);
  }
  if (threadIdx.x == 0 && p.loss) p.loss[0] = 0.f;     // the heads kernel (next launch) accumulates into it
}

// fused vector -> two heads (+ dual BCE): one block per sample, lanes over the F + Hd inputs
__global__ void __launch_bounds__(TPB) heads_fwd_kernel(const TabParams p) {
  const int B = p.B, Hd = p.Hd, F = p.F, b = blockIdx.x;
  const float* ft = p.scratch + (size_t)2 * B * Hd;
  __shared__ float s_c[TPB / 32], s_r[TPB / 32];
  pdl_launch_dependents();
  pdl_wait();
  const unsigned long long seed = p.seed + (p.step ? *p.step * 0x9E3779B97F4A7C15ull : 0ull);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float lc = 0.f, lr = 0.f;
  for (int i = threadIdx.x; i < F + Hd; i += TPB) {
    float f = i < F ? __ldg(p.feat + (size_t)b * F + i) : ft[b * Hd + (i - F)];
    if (p.train) f *= keep_scale(p.drop_p, seed, 2, (unsigned long long)b * (F + Hd) + i);
    lc = fmaf(f, __ldg(p.cls_w + i), lc);
    lr = fmaf(f, __ldg(p.reg_w + i), lr);
  }
  lc = warp_sum(lc);
  lr = warp_sum(lr);
  if (lane == 0) { s_c[warp] = lc; s_r[warp] = lr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    lc = p.cls_b[0]; lr = p.reg_b[0];
    for (int i = 0; i < TPB / 32; ++i) { lc += s_c[i]; lr += s_r[i]; }
    p.logit[b] = lc;
    p.reg[b] = lr;
    if (p.y_hard) {
      const float w = p.sample_w ? p.sample_w[b] : 1.f;
      const float yh = p.y_hard[b], ys = p.y_soft[b];
      const float bh = fmaxf(lc, 0.f) - lc * yh + log1pf(expf(-fabsf(lc)));
      const float bs = fmaxf(lr, 0.f) - lr * ys + log1pf(expf(-fabsf(lr)));
      atomicAdd(p.loss, w * (p.alpha * bh + p.beta * bs) / B);
      p.dlogit[b] = p.alpha * w * (1.f / (1.f + expf(-lc)) - yh) / B;
      p.dreg[b] = p.beta * w * (1.f / (1.f + expf(-lr)) - ys) / B;
    }
  }
}

struct TabBwdParams {
  TabParams f;                 // same forward description (scratch still holds z0, a1, ft, bnstat)
  const float *dlogit, *dreg;  // [B]
  float* dfeat;                // [B,F]
  float *dW0, *db0, *dbn_g, *dbn_b, *dW1, *db1, *dcls_w, *dcls_b, *dreg_w, *dreg_b;   // written (=), not accumulated
  float* scratch2;             // dz1 [B,Hd], da1 [B,Hd]
};

// head weight gradients + gradient into the fused vector: one thread per input of the fused vector, blocks over inputs.
// The batch loop runs 8 samples at a time with all loads issued before the first use (the plain loop was a chain of
// dependent L2 round trips: ~30 us for 64 samples).
constexpr int HB_U = 8;
__global__ void __launch_bounds__(TPB) heads_bwd_kernel(const TabBwdParams q) {
  const TabParams& p = q.f;
  const int B = p.B, Hd = p.Hd, F = p.F;
  const float* __restrict__ ft = p.scratch + (size_t)2 * B * Hd;
  const float* __restrict__ feat = p.feat;
  const float* __restrict__ dlogit = q.dlogit;
  const float* __restrict__ dreg = q.dreg;
  float* __restrict__ dfeat = q.dfeat;
  float* __restrict__ dz1 = q.scratch2;
  const unsigned long long seed = p.seed + (p.step ? *p.step * 0x9E3779B97F4A7C15ull : 0ull);
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i < F + Hd) {
    float gc = 0.f, gr = 0.f;
    const float wc = p.cls_w[i], wr = p.reg_w[i];
    for (int b0 = 0; b0 < B; b0 += HB_U) {
      float f[HB_U], dl[HB_U], dr[HB_U];
#pragma unroll
      for (int u = 0; u < HB_U; ++u) {
        const int b = b0 + u;
        if (b < B) {
          f[u] = i < F ? __ldg(feat + (size_t)b * F + i) : ft[b * Hd + (i - F)];
          dl[u] = __ldg(dlogit + b); dr[u] = __ldg(dreg + b);
        }
      }
#pragma unroll
      for (int u = 0; u < HB_U; ++u) {
        const int b = b0 + u;
        if (b < B) {
          const float ks = p.train ? keep_scale(p.drop_p, seed, 2, (unsigned long long)b * (F + Hd) + i) : 1.f;
          const float fk = f[u] * ks;
          gc = fmaf(dl[u], fk, gc);
          gr = fmaf(dr[u], fk, gr);
          const float df = (dl[u] * wc + dr[u] * wr) * ks;
          if (i < F) dfeat[(size_t)b * F + i] = df;
          else dz1[b * Hd + (i - F)] = f[u] > 0.f ? df : 0.f;   // through the last ReLU
        }
      }
    }
    q.dcls_w[i] = gc;
    q.dreg_w[i] = gr;
  }
  if (i == 0) {
    float sc = 0.f, sr = 0.f;
    for (int b = 0; b < B; ++b) { sc += dlogit[b]; sr += dreg[b]; }
    q.dcls_b[0] = sc;
    q.dreg_b[0] = sr;
  }
}

// Tab MLP backward in ONE block with dz1 / a1 / z0 / da1 ([B][Hd] each) and W1 staged in shared memory, like the forward:
// every loop below used to walk global scratch with dependent loads.
__global__ void __launch_bounds__(TAB_TPB, 1) tab_heads_bwd_kernel(const TabBwdParams q) {
  extern __shared__ __align__(16) float tsm[];
  const TabParams& p = q.f;
  const int B = p.B, T = p.T, Hd = p.Hd;
  const float* z0 = p.scratch;
  const float* a1 = z0 + (size_t)B * Hd;
  const float* ft = a1 + (size_t)B * Hd;
  const float* bnstat = ft + (size_t)B * Hd;
  const float* dz1 = q.scratch2;
  // the four [B][Hd] arrays live in shared memory (row pitch Hd+1: read by column in dW1 / db1) when the batch fits,
  // otherwise they are used where they are in the global scratch (row pitch Hd; da1 = second half of scratch2)
  const bool big = p.big != 0;
  const int ld = big ? Hd : Hd + 1;
  const size_t arr = big ? 0 : (size_t)B * (Hd + 1);
  const float* s_dz1 = big ? dz1 : tsm;
  const float* s_a1 = big ? a1 : tsm + arr;
  const float* s_z0 = big ? z0 : tsm + 2 * arr;
  float* s_da = big ? q.scratch2 + (size_t)B * Hd : tsm + 3 * arr;   // da1, then dz0 in place
  float* s_w1 = tsm + 4 * arr;                      // [Hd][Hd]
  float* s_xt = s_w1 + (size_t)Hd * Hd;             // [B][T]
  float* s_red = s_xt + (size_t)B * T;              // [2][TAB_TPB] partial sums of the BatchNorm1d backward
  float* s_sum = s_red + 2 * TAB_TPB;               // [2][Hd]
  const unsigned long long seed = p.seed + (p.step ? *p.step * 0x9E3779B97F4A7C15ull : 0ull);
  const int t_ = threadIdx.x;
  if (!big) {
    for (int i = t_; i < B * Hd; i += TAB_TPB) {
      const int b = i / Hd, j = i - b * Hd;
      tsm[b * ld + j] = dz1[i];
      tsm[arr + b * ld + j] = a1[i];
      tsm[2 * arr + b * ld + j] = z0[i];
    }
  }
  for (int i = t_; i < Hd * Hd; i += TAB_TPB) s_w1[i] = __ldg(p.W1 + i);
  for (int i = t_; i < B * T; i += TAB_TPB) s_xt[i] = __ldg(p.xtab + i);
  __syncthreads();
  // second linear: dW1[j][t] = sum_b dz1[b][j] a1[b][t]; db1; da1 = dz1 . W1
  for (int i = t_; i < Hd * Hd; i += TAB_TPB) {
    const int j = i / Hd, t = i % Hd;
    float acc = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b) acc = fmaf(s_dz1[b * ld + j], s_a1[b * ld + t], acc);
    q.dW1[i] = acc;
  }
  for (int j = t_; j < Hd; j += TAB_TPB) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += s_dz1[b * ld + j];
    q.db1[j] = acc;
  }
  for (int i = t_; i < B * Hd; i += TAB_TPB) {
    const int b = i / Hd, t = i % Hd;
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < Hd; ++j) acc = fmaf(s_dz1[b * ld + j], s_w1[j * Hd + t], acc);
    // through dropout and ReLU of the first layer (a1 > 0 <=> kept and positive)
    const float ks = p.train ? keep_scale(p.drop_p, seed, 1, i) : 1.f;
    s_da[b * ld + t] = s_a1[b * ld + t] > 0.f ? acc * ks : 0.f;
  }
  __syncthreads();
  // BatchNorm1d backward (batch statistics in train mode, plain scaling in eval) -> overwrite da1 with dz0.
  // thread = (feature j, batch slice): the two sums over the batch are short strided partials joined through shared memory
  {
    const int parts = TAB_TPB / Hd, j = t_ % Hd, part = t_ / Hd;     // Hd <= 256 (checked by the host): parts >= 4
    float s1 = 0.f, s2 = 0.f;
    if (part < parts) {
      const float mean = bnstat[j], rstd = bnstat[Hd + j];
      for (int b = part; b < B; b += parts) {
        const float d = s_da[b * ld + j], xh = (s_z0[b * ld + j] - mean) * rstd;
        s1 += d; s2 = fmaf(d, xh, s2);
      }
    }
    s_red[t_] = s1;
    s_red[TAB_TPB + t_] = s2;
    __syncthreads();
    if (t_ < Hd) {
      float t1 = 0.f, t2 = 0.f;
      for (int k = 0; k < parts; ++k) { t1 += s_red[k * Hd + t_]; t2 += s_red[TAB_TPB + k * Hd + t_]; }
      s_sum[t_] = t1;
      s_sum[Hd + t_] = t2;
      q.dbn_b[t_] = t1;
      q.dbn_g[t_] = t2;
    }
    __syncthreads();
    for (int i = t_; i < B * Hd; i += TAB_TPB) {
      const int b = i / Hd, jj = i - b * Hd;
      const float mean = bnstat[jj], rstd = bnstat[Hd + jj], gma = p.bn_g[jj];
      const float d = s_da[b * ld + jj], xh = (s_z0[b * ld + jj] - mean) * rstd;
      s_da[b * ld + jj] = p.train ? gma * rstd * (d - s_sum[jj] / B - xh * s_sum[Hd + jj] / B) : gma * rstd * d;
    }
  }
  __syncthreads();
  for (int i = t_; i < Hd * T; i += TAB_TPB) {
    const int j = i / T, t = i % T;
    float acc = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b) acc = fmaf(s_da[b * ld + j], s_xt[b * T + t], acc);
    q.dW0[i] = acc;
  }
  for (int j = t_; j < Hd; j += TAB_TPB) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += s_da[b * ld + j];
    q.db0[j] = acc;
  }
}


// ================================================================================================= MIL head + BCE
// logit[b] = <dropout(M[b]), w> + bias   (MILNet.drop + MILNet.head, train_mil_attention_v1.py:146-147)
__global__ void __launch_bounds__(TPB) linear1_fwd_kernel(const float* __restrict__ M, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ logit, int D,
                                                          float drop_p, unsigned long long seed,
                                                          const unsigned long long* __restrict__ step) {
  __shared__ float s[TPB / 32];
  const int b = blockIdx.x;
  const unsigned long long sd = seed + (step ? *step * 0x9E3779B97F4A7C15ull : 0ull);
  float acc = 0.f;
  for (int d = threadIdx.x; d < D; d += TPB)
    acc = fmaf(M[(size_t)b * D + d] * keep_scale(drop_p, sd, 3, (unsigned long long)b * D + d), w[d], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = bias[0];
    for (int i = 0; i < TPB / 32; ++i) t += s[i];
    logit[b] = t;
  }
}

// dM[b][d] = dlogit[b]*w[d]*keep ; dw[d] = sum_b dlogit[b]*M[b][d]*keep ; db = sum_b dlogit[b]   (written, single block)
__global__ void __launch_bounds__(TPB) linear1_bwd_kernel(const float* __restrict__ dlogit, const float* __restrict__ M,
                                                          const float* __restrict__ w, float* __restrict__ dM,
                                                          float* __restrict__ dw, float* __restrict__ db, int B, int D,
                                                          float drop_p, unsigned long long seed,
                                                          const unsigned long long* __restrict__ step) {
  const unsigned long long sd = seed + (step ? *step * 0x9E3779B97F4A7C15ull : 0ull);
  for (int d = blockIdx.x * TPB + threadIdx.x; d < D; d += gridDim.x * TPB) {
    float acc = 0.f;
    const float wd = w[d];
    for (int b = 0; b < B; ++b) {
      const float ks = keep_scale(drop_p, sd, 3, (unsigned long long)b * D + d);
      const float dl = dlogit[b];
      dM[(size_t)b * D + d] = dl * wd * ks;
      acc = fmaf(dl, M[(size_t)b * D + d] * ks, acc);
    }
    dw[d] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += dlogit[b];
    db[0] = t;
  }
}

// loss = mean_b w_b * bce(logit_b, y_b) ; dlogit_b = w_b * (sigmoid(logit_b) - y_b) / B   (train_mil_attention_v1.py:183)
__global__ void __launch_bounds__(TPB) bce_logits_kernel(const float* __restrict__ logit, const float* __restrict__ y,
                                                         const float* __restrict__ sw, float* __restrict__ loss,
                                                         float* __restrict__ dlogit, int B) {
  __shared__ float s[TPB / 32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += TPB) {
    const float x = logit[b], t = y[b], w = sw ? sw[b] : 1.f;
    acc += w * (fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)))) / B;
    if (dlogit) dlogit[b] = w * (1.f / (1.f + expf(-x)) - t) / B;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < TPB / 32; ++i) t += s[i];
    loss[0] = t;
  }
}

}  // namespace

extern "C" size_t trt_mil_attn_smem_bytes(int K, int D, int hid, int backward) {
  if (backward) return (size_t)2 * K * hid * sizeof(float);
  return ((size_t)K * D + K) * sizeof(float);
}

extern "C" int trt_mil_attn_fwd(const float* H, const float* Vw, const float* Vb, const float* Uw, const float* Ub,
                                const float* ww, const float* wb, float* M, float* A, float* gV, float* gU, int B, int K,
                                int D, int hid, cudaStream_t stream) {
  TRT_REQUIRE(H && Vw && Vb && Uw && Ub && ww && wb && M && A, "trt_mil_attn_fwd: null pointer");
  TRT_REQUIRE(B > 0 && K > 0 && D > 0 && D % 4 == 0 && hid > 0, "trt_mil_attn_fwd: bad shape");
  TRT_REQUIRE((gV == nullptr) == (gU == nullptr), "trt_mil_attn_fwd: gV and gU must be given together");
  const size_t smem = trt_mil_attn_smem_bytes(K, D, hid, 0);
  TRT_REQUIRE(smem <= 220 * 1024, "trt_mil_attn_fwd: bag of %d x %d does not fit in shared memory", K, D);
  TRT_CUDA(cudaFuncSetAttribute(mil_score_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per device
  TRT_CUDA(cudaFuncSetAttribute(mil_score_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  // hidden slices: enough blocks for ~2 per SM on small batches, one slice (H staged once per bag) on large ones
  // hidden units per warp: the 4-unit variant (8 gate rows per shared-memory read) was measured SLOWER at B = 1024 (203
  // registers, one block per SM, exposed L2 latency on the weight rows), so the 1-unit variant serves every batch size
  const int hu = 1;
  int JS = (2 * trt_num_sms() + B - 1) / B;
  const int max_js = (hid + 8 * hu - 1) / (8 * hu);           // at least one pass of the block's 8 warps
  if (JS > max_js) JS = max_js;
  if (JS < 1) JS = 1;
  const int JH = ((hid + JS - 1) / JS + hu - 1) / hu * hu;
  JS = (hid + JH - 1) / JH;
  TRT_CUDA(cudaMemsetAsync(A, 0, (size_t)B * K * sizeof(float), stream));      // A accumulates the raw scores first
  if (hu == 4) mil_score_kernel<4><<<dim3(B, JS), TPB, smem, stream>>>(H, Vw, Vb, Uw, Ub, ww, A, gV, gU, K, D, hid, JH);
  else mil_score_kernel<1><<<dim3(B, JS), TPB, smem, stream>>>(H, Vw, Vb, Uw, Ub, ww, A, gV, gU, K, D, hid, JH);
  trt_count_launch(1);
  mil_pool_kernel<<<B, TPB, (size_t)K * sizeof(float), stream>>>(H, wb, A, M, K, D);
  return trt_check_launch("trt_mil_attn_fwd");
}

int trt_gemm_mil_scores(const void* A_split, const void* W_split, int M, int hid, int D, const float* bias2, const float* w,
                        float* score, float* gv, float* gu, cudaStream_t stream);     // gemm_tc.cu

extern "C" size_t trt_mil_attn_tc_workspace_bytes(int B, int K, int D, int hid) {
  const size_t hs = (size_t)B * K * 2 * D * 2, wsz = (size_t)2 * hid * 3 * D * 2, bs = (size_t)2 * hid * 4;
  return ((hs + 255) & ~(size_t)255) + ((wsz + 255) & ~(size_t)255) + bs;
}

extern "C" int trt_mil_attn_fwd_tc(const float* H, const float* Vw, const float* Vb, const float* Uw, const float* Ub,
                                   const float* ww, const float* wb, float* M, float* A, float* gV, float* gU, int B, int K,
                                   int D, int hid, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  TRT_REQUIRE(H && Vw && Vb && Uw && Ub && ww && wb && M && A && workspace, "trt_mil_attn_fwd_tc: null pointer");
  TRT_REQUIRE(B > 0 && K > 0 && D > 0 && D % 64 == 0 && hid > 0 && hid % 8 == 0 && 2 * hid <= 4096, "trt_mil_attn_fwd_tc: bad shape (D %% 64, hid %% 8)");
  TRT_REQUIRE((gV == nullptr) == (gU == nullptr), "trt_mil_attn_fwd_tc: gV and gU must be given together");
  TRT_REQUIRE(workspace_bytes >= trt_mil_attn_tc_workspace_bytes(B, K, D, hid) && (((uintptr_t)workspace) & 255) == 0,
              "trt_mil_attn_fwd_tc: workspace too small or not 256-byte aligned");
  TRT_REQUIRE((size_t)K * sizeof(float) <= 48 * 1024, "trt_mil_attn_fwd_tc: bag too large");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const size_t rows = (size_t)B * K;
  const size_t hs = (rows * 2 * D * 2 + 255) & ~(size_t)255, wsz = ((size_t)2 * hid * 3 * D * 2 + 255) & ~(size_t)255;
  __nv_bfloat16* Hs = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(ws + hs);
  float* bias2 = reinterpret_cast<float*>(ws + hs + wsz);
  TRT_CUDA(cudaMemsetAsync(A, 0, rows * sizeof(float), stream));                // A accumulates the raw scores first
  size_t blocks = (rows * (D / 4) + TPB - 1) / TPB;
  if (blocks > (size_t)8 * trt_num_sms()) blocks = (size_t)8 * trt_num_sms();
  mil_split_h_kernel<<<(int)blocks, TPB, 0, stream>>>(reinterpret_cast<const float4*>(H), reinterpret_cast<uint2*>(Hs), rows, D / 4);
  trt_count_launch(1);
  mil_split_w_kernel<<<2 * hid, TPB, 0, stream>>>(Vw, Uw, Vb, Ub, Ws, bias2, hid, D);
  trt_count_launch(1);
  int rc = trt_gemm_mil_scores(Hs, Ws, (int)rows, hid, D, bias2, ww, A, gV, gU, stream);
  if (rc) return rc;
  mil_pool_kernel<<<B, TPB, (size_t)K * sizeof(float), stream>>>(H, wb, A, M, K, D);
  return trt_check_launch("trt_mil_attn_fwd_tc");
}

extern "C" int trt_mil_attn_bwd(const float* dM, const float* H, const float* A, float* gV, float* gU,
                                const float* Vw, const float* Uw, const float* ww, float* dH, float* dVw, float* dVb,
                                float* dUw, float* dUb, float* dww, float* dwb, int B, int K, int D, int hid,
                                cudaStream_t stream) {
  TRT_REQUIRE(dM && H && A && gV && gU && Vw && Uw && ww && dH && dVw && dVb && dUw && dUb && dww && dwb,
              "trt_mil_attn_bwd: null pointer");
  TRT_REQUIRE(B > 0 && K > 0 && D > 0 && D % 4 == 0 && hid > 0, "trt_mil_attn_bwd: bad shape");
  const size_t smem = trt_mil_attn_smem_bytes(K, D, hid, 1);
  TRT_REQUIRE(smem <= 220 * 1024, "trt_mil_attn_bwd: bag of %d x %d does not fit in shared memory", K, hid);
  TRT_CUDA(cudaFuncSetAttribute(mil_bwd_dh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  mil_bwd_gate_kernel<<<B, TPB, (size_t)K * sizeof(float), stream>>>(dM, H, A, gV, gU, ww, dVb, dUb, dww, dwb, K, D, hid);
  trt_count_launch(1);
  mil_bwd_weight_kernel<<<dim3(2 * ((hid + MW_R - 1) / MW_R), (D + MW_C - 1) / MW_C), TPB, 0, stream>>>(H, gV, gU, dVw, dUw, B * K, D, hid);
  trt_count_launch(1);
  mil_bwd_dh_kernel<<<dim3(B, (D + MH_C - 1) / MH_C), TPB, smem, stream>>>(dM, A, gV, gU, Vw, Uw, dH, K, D, hid);
  return trt_check_launch("trt_mil_attn_bwd");
}

extern "C" size_t trt_tab_heads_scratch_floats(int B, int Hd) { return (size_t)5 * B * Hd + 2 * Hd; }

extern "C" int trt_tab_heads_fwd(const float* feat, const float* xtab, const float* const* params_host, float* bn_rm,
                                 float* bn_rv, long long* bn_nbt, const float* y_hard, const float* y_soft,
                                 const float* sample_w, float* logit, float* reg, float* loss, float* dlogit, float* dreg,
                                 float* scratch, int B, int T, int Hd, int F, int train, float drop_p, float alpha,
                                 float beta, unsigned long long seed, const unsigned long long* step, cudaStream_t stream) {
  TRT_REQUIRE(feat && xtab && params_host && logit && reg && scratch && bn_rm && bn_rv, "trt_tab_heads_fwd: null pointer");
  TRT_REQUIRE(B > 0 && T > 0 && Hd > 0 && F > 0, "trt_tab_heads_fwd: bad shape");
  TRT_REQUIRE(!(train && B < 2), "trt_tab_heads_fwd: Expected more than 1 value per channel when training (BatchNorm1d, batch %d)", B);
  TRT_REQUIRE(!y_hard || (y_soft && loss && dlogit && dreg), "trt_tab_heads_fwd: loss outputs missing");
  TabParams p;
  p.B = B; p.T = T; p.Hd = Hd; p.F = F; p.train = train;
  p.drop_p = drop_p; p.bn_eps = 1e-5f; p.bn_momentum = 0.1f; p.alpha = alpha; p.beta = beta;
  p.seed = seed; p.step = step;
  p.feat = feat; p.xtab = xtab;
  p.W0 = params_host[0]; p.b0 = params_host[1]; p.bn_g = params_host[2]; p.bn_b = params_host[3]; p.W1 = params_host[4]; p.b1 = params_host[5];
  p.cls_w = params_host[6]; p.cls_b = params_host[7]; p.reg_w = params_host[8]; p.reg_b = params_host[9];
  p.bn_rm = bn_rm; p.bn_rv = bn_rv; p.bn_nbt = bn_nbt;
  p.y_hard = y_hard; p.y_soft = y_soft; p.sample_w = sample_w;
  p.logit = logit; p.reg = reg; p.loss = loss; p.dlogit = dlogit; p.dreg = dreg;
  p.scratch = scratch;
  TRT_REQUIRE(Hd <= 256, "trt_tab_heads_fwd: tab_hidden %d > 256 not built", Hd);
  const size_t tfixed = ((size_t)Hd * (Hd + 1) + 2 * TAB_TPB + 2 * Hd) * sizeof(float);
  size_t tsmem = tfixed + (size_t)B * Hd * sizeof(float);
  p.big = tsmem > 200 * 1024;          // the reference has no batch limit: large batches walk the global scratch instead
  if (p.big) tsmem = tfixed;
  TRT_CUDA(cudaFuncSetAttribute(tab_heads_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  TRT_CUDA(trt_launch(tab_heads_fwd_kernel, dim3(1), dim3(TAB_TPB), tsmem, stream, p));
  trt_count_launch(1);
  TRT_CUDA(trt_launch(heads_fwd_kernel, dim3(B), dim3(TPB), 0, stream, p));
  return trt_check_launch("trt_tab_heads_fwd");
}

extern "C" int trt_tab_heads_bwd(const float* feat, const float* xtab, const float* const* params_host, const float* bn_rm,
                                 const float* bn_rv, const float* dlogit, const float* dreg, float* dfeat,
                                 float* const* grads_host, float* scratch, int B, int T, int Hd, int F, int train, float drop_p,
                                 unsigned long long seed, const unsigned long long* step, cudaStream_t stream) {
  TRT_REQUIRE(feat && xtab && params_host && dlogit && dreg && dfeat && grads_host && scratch, "trt_tab_heads_bwd: null pointer");
  TabBwdParams q;
  TabParams& p = q.f;
  p.B = B; p.T = T; p.Hd = Hd; p.F = F; p.train = train;
  p.drop_p = drop_p; p.bn_eps = 1e-5f; p.bn_momentum = 0.1f; p.alpha = 0; p.beta = 0; p.big = 0;
  p.seed = seed; p.step = step;
  p.feat = feat; p.xtab = xtab;
  p.W0 = params_host[0]; p.b0 = params_host[1]; p.bn_g = params_host[2]; p.bn_b = params_host[3]; p.W1 = params_host[4]; p.b1 = params_host[5];
  p.cls_w = params_host[6]; p.cls_b = params_host[7]; p.reg_w = params_host[8]; p.reg_b = params_host[9];
  p.bn_rm = const_cast<float*>(bn_rm); p.bn_rv = const_cast<float*>(bn_rv); p.bn_nbt = nullptr;
  p.y_hard = p.y_soft = p.sample_w = nullptr;
  p.logit = p.reg = p.loss = p.dlogit = p.dreg = nullptr;
  p.scratch = scratch;
  q.dlogit = dlogit; q.dreg = dreg; q.dfeat = dfeat;
  q.dW0 = grads_host[0]; q.db0 = grads_host[1]; q.dbn_g = grads_host[2]; q.dbn_b = grads_host[3]; q.dW1 = grads_host[4]; q.db1 = grads_host[5];
  q.dcls_w = grads_host[6]; q.dcls_b = grads_host[7]; q.dreg_w = grads_host[8]; q.dreg_b = grads_host[9];
  q.scratch2 = scratch + (size_t)3 * B * Hd + 2 * Hd;
  heads_bwd_kernel<<<(F + Hd + TPB - 1) / TPB, TPB, 0, stream>>>(q);
  trt_count_launch(1);
  TRT_REQUIRE(Hd <= 256, "trt_tab_heads_bwd: tab_hidden %d > 256 not built", Hd);
  const size_t bfixed = ((size_t)Hd * Hd + (size_t)B * T + 2 * TAB_TPB + 2 * Hd) * sizeof(float);
  size_t bsmem = bfixed + (size_t)4 * B * (Hd + 1) * sizeof(float);
  p.big = bsmem > 200 * 1024;
  if (p.big) bsmem = bfixed;
  TRT_REQUIRE(bsmem <= 200 * 1024, "trt_tab_heads_bwd: batch %d x tab_in %d does not fit in shared memory", B, T);
  TRT_CUDA(cudaFuncSetAttribute(tab_heads_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tab_heads_bwd_kernel<<<1, TAB_TPB, bsmem, stream>>>(q);
  return trt_check_launch("trt_tab_heads_bwd");
}


extern "C" int trt_linear1_fwd(const float* M, const float* w, const float* bias, float* logit, int B, int D, float drop_p,
                               unsigned long long seed, const unsigned long long* step, cudaStream_t stream) {
  TRT_REQUIRE(M && w && bias && logit && B > 0 && D > 0, "trt_linear1_fwd: bad argument");
  linear1_fwd_kernel<<<B, TPB, 0, stream>>>(M, w, bias, logit, D, drop_p, seed, step);
  return trt_check_launch("trt_linear1_fwd");
}

extern "C" int trt_linear1_bwd(const float* dlogit, const float* M, const float* w, float* dM, float* dw, float* db, int B,
                               int D, float drop_p, unsigned long long seed, const unsigned long long* step,
                               cudaStream_t stream) {
  TRT_REQUIRE(dlogit && M && w && dM && dw && db && B > 0 && D > 0, "trt_linear1_bwd: bad argument");
  linear1_bwd_kernel<<<(D + TPB - 1) / TPB, TPB, 0, stream>>>(dlogit, M, w, dM, dw, db, B, D, drop_p, seed, step);
  return trt_check_launch("trt_linear1_bwd");
}

extern "C" int trt_bce_logits(const float* logit, const float* y, const float* sample_w, float* loss, float* dlogit, int B,
                              cudaStream_t stream) {
  TRT_REQUIRE(logit && y && loss && B > 0, "trt_bce_logits: bad argument");
  bce_logits_kernel<<<1, TPB, 0, stream>>>(logit, y, sample_w, loss, dlogit, B);
  return trt_check_launch("trt_bce_logits");
}
