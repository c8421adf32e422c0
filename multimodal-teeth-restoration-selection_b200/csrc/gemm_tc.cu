// tcgen05 / TMEM / TMA GEMMs for the 1x1 convolutions of the EfficientNet encoder (SURVEY.md K2, 92 % of the MACs).
//
//   trt_gemm_bf16      C[M,N] = epi(A[M,K] . B[N,K]^T)       forward (B = W[Cout,Cin]) and dgrad (B = W^T[Cin,Cout])
//                      both operands K-major, 128B-swizzled TMA tiles; persistent CTAs (1/SM, 448 threads): warp 0 = TMA
//                      producer, warp 1 = single-thread tcgen05.mma issuer, warps 2-13 = up to three independent 4-warp
//                      epilogue groups, each draining its own fp32 accumulator from TMEM.  The weights stay resident in
//                      shared memory for the whole CTA when one n-block fits.  Epilogue: folded-BN scale/shift, SiLU,
//                      residual add, bf16 tile staged in padded shared memory, then ONE pass in which every 16-byte shared
//                      read feeds a coalesced st.global and the per-channel sum / sum^2 (train-mode BN statistics).
//                      (A TMA tensor store of the C tile was measured slower on these 48..576-byte rows: DESIGN.md 5.)
//   trt_gemm_wgrad_bf16  O[p,q] += sum_m P[m,p] * Q[m,q]      weight gradient: both operands MN-major (the reduction
//                      runs over rows), split over m across CTAs, fp32 red.global.add epilogue.
//
// Replaces: the cuDNN/cuBLAS dispatch behind `self.backbone(x_img)` (train_mm_joint_dualtask.py:154) and its autograd
// backward (:248) for every conv_pw / conv_pwl / conv_head of timm's EfficientNet.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/teethrt.h"

namespace {

constexpr int BM = 128;          // UMMA M (rows of the output tile = TMEM lanes)
constexpr int BK = 64;           // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int GEMM_THREADS = 448;        // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..13 = three epilogue groups
constexpr int EPI_THREADS = 128;         // threads of one epilogue group (one warp per TMEM lane quarter)
constexpr int MAX_GROUPS = 3;
constexpr int WGRAD_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB

struct GemmParams {
  int M, N, K;
  int block_n, num_m_blocks, num_n_blocks, num_k_blocks;
  int stages, acc_stride, nacc, tmem_cols, flags;
  int ngroups;    // epilogue groups in use (each owns one staging buffer); tile i of a CTA goes to group i % ngroups
  int b_resident; // 1: the whole B operand (all k-blocks, one n-block) is loaded once per CTA and stays in shared memory
  __nv_bfloat16* C;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  double* stats;  // [TRT_STAT_REPLICAS][2][N]
  const __nv_bfloat16* bn_x;   // TRT_EPI_BNBWD: [M,N] input of the BatchNorm whose backward sums this launch accumulates
  int a_kblocks;  // > 0: A has only this many k-blocks; k-block kb of the product reads A k-block kb % a_kblocks (split operands)
  // TRT_EPI_MILGATE: columns come in (V_j, U_j) pairs; score[m] += sum_j w[j] * tanh(acc[2j] + bias[2j]) * sigmoid(acc[2j+1] + bias[2j+1])
  const float* mil_bias;   // [N] interleaved (Vb_j, Ub_j)
  const float* mil_w;      // [N/2]
  float* mil_score;        // [M] accumulated (atomicAdd): zeroed by the caller
  float* mil_gv;           // [M][N/2] tanh(.) (optional, with mil_gu)
  float* mil_gu;           // [M][N/2] sigmoid(.)
};

#ifdef TRT_GEMM_TIMING
// bring-up instrumentation (never compiled into the shipped library): per-phase clock64 totals of epilogue group 0's first thread
__device__ unsigned long long g_gemm_dbg[16];
// producer / MMA-issuer phases of CTA 0 (slots 8..11): cycles waiting for a free stage, issuing TMA, waiting for data, issuing MMAs
#define TRT_ROLE_TICK(slot) do { if (blockIdx.x == 0) { const long long now__ = clock64(); atomicAdd(&g_gemm_dbg[slot], (unsigned long long)(now__ - rtick__)); rtick__ = now__; } } while (0)
#define TRT_ROLE_TICK_INIT long long rtick__ = clock64()
#define TRT_TICK(slot) do { if (threadIdx.x == 64) { const long long now__ = clock64(); atomicAdd(&g_gemm_dbg[slot], (unsigned long long)(now__ - tick__)); tick__ = now__; } } while (0)
#define TRT_TICK_INIT long long tick__ = clock64()
#else
#define TRT_TICK(slot) do {} while (0)
#define TRT_TICK_INIT do {} while (0)
#define TRT_ROLE_TICK(slot) do {} while (0)
#define TRT_ROLE_TICK_INIT do {} while (0)
#endif

struct SmemLayout {
  uint32_t a_off, b_off, c_off, cpitch, cbuf_bytes, bar_off, total;
};
__host__ __device__ inline SmemLayout make_layout(int block_n, int stages, int ngroups, int b_slots, int nostage = 0) {
  SmemLayout L;
  uint32_t b_stage = (uint32_t)block_n * 128u;
  L.a_off = 0;
  L.b_off = L.a_off + (uint32_t)stages * A_STAGE_BYTES;
  L.c_off = L.b_off + (uint32_t)b_slots * b_stage;      // b_slots = stages (streamed) or k-blocks (resident)
  // staged C tile: row-major [128][block_n] bf16 with the row pitch padded to an ODD number of 16-byte granules, so the
  // drain (32 lanes = 32 consecutive rows, same column granule) hits 8 distinct bank groups per quarter-warp
  L.cpitch = (((uint32_t)block_n >> 3) | 1u) * 16u;
  L.cbuf_bytes = (BM * L.cpitch + 1023u) & ~1023u;
  if (L.cbuf_bytes < 16384u) L.cbuf_bytes = 16384u;       // also the scratch of the statistics flush
  if (nostage) L.cbuf_bytes = 1024u;                      // epilogues that keep everything in registers (gated-attention scores)
  L.bar_off = L.c_off + (uint32_t)ngroups * L.cbuf_bytes;
  L.total = L.bar_off + 256;
  return L;
}

// MODE 0: the standard epilogue; 1: + BatchNorm-backward sums (TRT_EPI_BNBWD); 2: gated-attention scores (TRT_EPI_MILGATE).
// A template parameter, not a flag: the extra live registers of modes 1 / 2 would otherwise spill the standard epilogue
// (128 registers at 448 threads).
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = TRT_ALIGNED_SMEM(smem_raw, 1024);
  const SmemLayout L = make_layout(p.block_n, p.stages, p.ngroups, p.b_resident ? p.num_k_blocks : p.stages, MODE == 2 ? 1 : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;     // [nacc] accumulator complete (MMA -> epilogue group)
  uint64_t* tempty_bar = tfull_bar + 4;             // [nacc] accumulator drained (epilogue group -> MMA)
  uint64_t* bres_bar = tempty_bar + 4;              // resident B landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * 128u;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < p.nacc; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 1);      // one elected arrival per group (after the group's own barrier)
    }
    ptx::mbar_init(bres_bar, 1);
    ptx::fence_barrier_init();
  }
  pdl_launch_dependents();
  pdl_wait();                      // everything above (barriers, descriptor prefetch) overlapped the previous kernel's tail
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {   // elect.sync, not lane == 0: ptxas then issues the uniform-datapath TMA / MMA instructions directly instead of inside a per-lane loop
      int stage = 0;
      uint32_t phase = 0;
      TRT_ROLE_TICK_INIT;
      if (p.b_resident) {
        // these GEMMs are bound by the TMA unit's box-row rate (~7 cycles per <=128-byte row, measured): the weights of the
        // CTA's n-block are the same for every tile it computes, so their rows are fetched once instead of once per tile.
        // With several n-blocks the grid is a multiple of their number, so a CTA's tiles t = blockIdx.x + i * gridDim.x all
        // share the n-block blockIdx.x % num_n_blocks (host: gemm_launch)
        const int n_res = (blockIdx.x % p.num_n_blocks) * p.block_n;
        ptx::mbar_expect_tx(bres_bar, (uint32_t)p.num_k_blocks * b_stage_bytes);
        for (int kb = 0; kb < p.num_k_blocks; ++kb)
          ptx::tma_load_2d(smem + L.b_off + kb * b_stage_bytes, &tmap_b, bres_bar, kb * BK, n_res);
      }
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / p.num_n_blocks) * BM;
        const int n0 = (t % p.num_n_blocks) * p.block_n;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          TRT_ROLE_TICK(9);
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          TRT_ROLE_TICK(8);
          ptx::mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + (p.b_resident ? 0u : b_stage_bytes));
          ptx::tma_load_2d(smem + L.a_off + stage * A_STAGE_BYTES, &tmap_a, &full_bar[stage],
                           (MODE == 2 && p.a_kblocks > 0 ? kb % p.a_kblocks : kb) * BK, m0);
          if (!p.b_resident) ptx::tma_load_2d(smem + L.b_off + stage * b_stage_bytes, &tmap_b, &full_bar[stage], kb * BK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::instr_desc_bf16(BM, p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TRT_ROLE_TICK_INIT;
      if (p.b_resident) ptx::mbar_wait(bres_bar, 0);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          TRT_ROLE_TICK(11);
          ptx::mbar_wait(&full_bar[stage], phase);
          TRT_ROLE_TICK(10);
          ptx::tc_fence_after();
          TRT_ROLE_TICK(12);
          const uint32_t a_addr = ptx::smem_u32(smem + L.a_off + stage * A_STAGE_BYTES);
          const uint32_t b_addr = ptx::smem_u32(smem + L.b_off + (p.b_resident ? kb : stage) * b_stage_bytes);
          const int k_rem = p.K - kb * BK;
          const int nk = k_rem >= BK ? BK / UK : (k_rem + UK - 1) / UK;
#ifdef TRT_GEMM_TIMING
          const int nk_dbg = (p.flags & (1 << 21)) ? 1 : nk;       // ablation: one MMA per k-block (wrong result, timing only)
#else
          const int nk_dbg = nk;
#endif
          for (int k = 0; k < nk_dbg; ++k) {
            const uint64_t da = ptx::smem_desc(a_addr + k * (UK * 2), 0, 1024);
            const uint64_t db = ptx::smem_desc(b_addr + k * (UK * 2), 0, 1024);
            ptx::umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          TRT_ROLE_TICK(13);
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          TRT_ROLE_TICK(14);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
        if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: up to three independent groups of four warps =====================
    // Group g owns the CTA's tiles g, g + G, g + 2G, ... and one staging buffer.  For its tile it (1) drains the TMEM
    // accumulator (one warp per lane quarter, 32 columns per tcgen05.ld): epilogue math -> bf16 -> shared memory; (2) walks
    // the staged tile 16 bytes per access: the same shared-memory read feeds a fully coalesced st.global (the tile's rows
    // are contiguous in HBM) and the BatchNorm column sums (FHFMA on the packed bf16, kept in registers across tiles).
    // Three tiles are in flight per SM, so one group's TMEM round trips and store issue overlap the others' (a single group,
    // a drain/store role split and a TMA tensor store of the 288-byte-pitch rows were all measured slower: profiles/).
    const int g = (warp - 2) >> 2;
    if (g < p.ngroups) {
      const int q = warp & 3;                       // TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31)
      const int row_l = q * 32 + lane;              // row inside the 128-row tile
      const int gt = threadIdx.x - 64 - g * EPI_THREADS;   // 0..127 inside the group
      const bool f_ss = p.flags & TRT_EPI_SCALE_SHIFT, f_silu = p.flags & TRT_EPI_SILU;
      const bool f_res = p.flags & TRT_EPI_RESIDUAL, f_stats = p.flags & TRT_EPI_STATS;
      constexpr bool f_bnbwd = MODE == 1;               // second sum = sum y * bn_x instead of sum y^2 (BatchNorm backward)
      const uint32_t cpitch = L.cpitch;
      uint8_t* cstage = smem + L.c_off + g * L.cbuf_bytes;
      const int n_oct = p.block_n >> 3;             // 8-column octets per tile row
      const int rg_count = EPI_THREADS / n_oct;     // row groups of the store pass
      const int so = gt % n_oct, srg = gt / n_oct;  // this thread's octet / row group
      const bool s_active = srg < rg_count;
      const int bar_id = 1 + g;
      float st_s[8], st_q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) st_s[i] = st_q[i] = 0.f;
      int st_n0 = -1;
      auto flush_stats = [&]() {
        // the group's staging buffer (free between tiles) holds partial[rg][block_n] for sum and sum^2
        float* ps = reinterpret_cast<float*>(cstage);
        float* pq = ps + rg_count * p.block_n;
        if (s_active) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            ps[srg * p.block_n + so * 8 + i] = st_s[i];
            pq[srg * p.block_n + so * 8 + i] = st_q[i];
          }
        }
        ptx::named_bar_sync(bar_id, EPI_THREADS);
        for (int c = gt; c < 2 * p.block_n; c += EPI_THREADS) {
          const int k = c >= p.block_n, cc = c - k * p.block_n;
          const float* src = k ? pq : ps;
          float tot = 0.f;
          for (int r = 0; r < rg_count; ++r) tot += src[r * p.block_n + cc];
          if (st_n0 + cc < p.N)
            atomicAdd(p.stats + ((size_t)(blockIdx.x % TRT_STAT_REPLICAS) * 2 + k) * p.N + st_n0 + cc, (double)tot);
        }
        ptx::named_bar_sync(bar_id, EPI_THREADS);
#pragma unroll
        for (int i = 0; i < 8; ++i) st_s[i] = st_q[i] = 0.f;
      };
      int seq = g;                                  // index of the tile in this CTA's sequence
      TRT_TICK_INIT;
      for (int t = blockIdx.x + g * gridDim.x; t < num_tiles; t += p.ngroups * gridDim.x, seq += p.ngroups) {
        const int m0 = (t / p.num_n_blocks) * BM;
        const int n0 = (t % p.num_n_blocks) * p.block_n;
        const int row = m0 + row_l;
        const int acc = seq % p.nacc;
        const uint32_t acc_par = (uint32_t)(seq / p.nacc) & 1u;
        if (f_stats && st_n0 != n0) {
          if (st_n0 >= 0) flush_stats();
          st_n0 = n0;
        }
        TRT_TICK(0);
        ptx::mbar_wait_sleep(&tfull_bar[acc], acc_par, 32);
        ptx::tc_fence_after();
        TRT_TICK(1);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
        if constexpr (MODE == 2) {
          // gated-attention scores straight from the accumulator: a thread owns one instance row, its (V_j, U_j) pairs are
          // adjacent columns; nothing is staged or stored except the optional gate activations the backward pass needs
          float part = 0.f;
          const int hid = p.N >> 1;
          for (int c0 = 0; c0 < p.block_n; c0 += 32) {
            uint32_t r[32];
            ptx::tmem_ld32(taddr + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = n0 + c0 + 2 * i;
              if (col < p.N) {
                const float2 bb = __ldg(reinterpret_cast<const float2*>(p.mil_bias + col));
                const float v = __uint_as_float(r[2 * i]) + bb.x, u = __uint_as_float(r[2 * i + 1]) + bb.y;
                const float tv = 2.0f * sigmoidf_(2.0f * v) - 1.0f, su = sigmoidf_(u);      // tanh(v) = 2 sigmoid(2v) - 1
                part = fmaf(__ldg(p.mil_w + (col >> 1)) * tv, su, part);
                if (p.mil_gv && row < p.M) {
                  p.mil_gv[(size_t)row * hid + (col >> 1)] = tv;
                  p.mil_gu[(size_t)row * hid + (col >> 1)] = su;
                }
              }
            }
          }
          ptx::tc_fence_before();
          ptx::named_bar_sync(bar_id, EPI_THREADS);          // accumulator drained by all four warps
          if (gt == 0) ptx::mbar_arrive(&tempty_bar[acc]);
          if (row < p.M) atomicAdd(p.mil_score + row, part);
          continue;
        }
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t r[32];
          const bool full = c0 + 32 <= p.block_n;    // the tile may end on a 16-column boundary
          if (full) ptx::tmem_ld32(taddr + c0, r);
          else ptx::tmem_ld16(taddr + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int o8 = 0; o8 < 4; ++o8) {
            if (o8 >= 2 && !full) break;
            const int col = n0 + c0 + o8 * 8;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[o8 * 8 + i]);
            if (f_ss && col < p.N) {
              const f8 sc = ldf8(p.scale + col), sh = ldf8(p.shift + col);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc.v[i], sh.v[i]);
            }
            if (f_silu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = siluf_(v[i]);
            }
            if (f_res && row < p.M && col < p.N) {
              const f8 rr = unpack8(ldg16(p.residual + (size_t)row * p.N + col));
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += rr.v[i];
            }
            uint4 o;
            o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
            o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
            *reinterpret_cast<uint4*>(cstage + row_l * cpitch + ((c0 >> 3) + o8) * 16) = o;
          }
        }
        ptx::tc_fence_before();
        TRT_TICK(2);
        ptx::named_bar_sync(bar_id, EPI_THREADS);            // accumulator drained + tile staged, by all four warps
        if (gt == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        TRT_TICK(3);
        if (s_active && n0 + so * 8 < p.N) {
          // store + statistics over the bf16-rounded values (rows past M are exact zeros: zero-filled A rows)
          const uint8_t* cbase = cstage + so * 16;
          __nv_bfloat16* gbase = p.C + (size_t)m0 * p.N + n0 + so * 8;
          const int rows_here = min(BM, p.M - m0);
          const uint32_t one2 = 0x3f803f80u;                  // bf16x2 {1, 1}
          const __nv_bfloat16* xbase = f_bnbwd ? p.bn_x + (size_t)m0 * p.N + n0 + so * 8 : nullptr;
          for (int r0 = srg; r0 < rows_here; r0 += 4 * rg_count) {
            uint4 w[4], xq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (r0 + u * rg_count < rows_here) {
                w[u] = *reinterpret_cast<const uint4*>(cbase + (r0 + u * rg_count) * cpitch);
                if (f_bnbwd) xq[u] = ldg16(xbase + (size_t)(r0 + u * rg_count) * p.N);
              }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (r0 + u * rg_count < rows_here) *reinterpret_cast<uint4*>(gbase + (size_t)(r0 + u * rg_count) * p.N) = w[u];
            if (f_stats) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (r0 + u * rg_count < rows_here) {
                  const uint32_t ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
                  const uint32_t xs[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const uint16_t lo = (uint16_t)(ws[i] & 0xffffu), hi = (uint16_t)(ws[i] >> 16);
                    const uint16_t qlo = f_bnbwd ? (uint16_t)(xs[i] & 0xffffu) : lo, qhi = f_bnbwd ? (uint16_t)(xs[i] >> 16) : hi;
                    st_s[2 * i] = ptx::fhfma(lo, (uint16_t)(one2 & 0xffffu), st_s[2 * i]);
                    st_s[2 * i + 1] = ptx::fhfma(hi, (uint16_t)(one2 >> 16), st_s[2 * i + 1]);
                    st_q[2 * i] = ptx::fhfma(lo, qlo, st_q[2 * i]);
                    st_q[2 * i + 1] = ptx::fhfma(hi, qhi, st_q[2 * i + 1]);
                  }
                }
              }
            }
          }
        }
        TRT_TICK(4);
        ptx::named_bar_sync(bar_id, EPI_THREADS);            // staging buffer free for the group's next tile
        TRT_TICK(5);
      }
      if (f_stats && st_n0 >= 0) flush_stats();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ wgrad
struct WgradParams {
  int M, Cp, Cq;
  int block_q;              // UMMA N (multiple of 16, <= 256)
  int num_p_blocks, num_q_blocks, splits, mblocks_per_split, num_mblocks;
  int stages, tmem_cols, vec4;
  int q_store;              // columns q >= q_store are computed but not stored (zero-padded operands, e.g. the stem's 27 taps)
  long long so_p, so_q;     // output strides (elements)
  float* out;
  // descriptor knobs (defaults follow the canonical MN-major SW128 layout; overridable by the bring-up test)
  uint32_t lbo, sbo, kstep_bytes;
};

__global__ void __launch_bounds__(WGRAD_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q,
                  const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = TRT_ALIGNED_SMEM(smem_raw, 1024);
  const int q_chunks = (p.block_q + 63) / 64;
  const uint32_t p_stage = 2 * 8192, q_stage = (uint32_t)q_chunks * 8192;
  uint8_t* sp = smem;
  uint8_t* sq = smem + (uint32_t)p.stages * p_stage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sq + (uint32_t)p.stages * q_stage);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* done_bar = empty_bar + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
  const int p0 = (tile / p.num_q_blocks) * BM;
  const int q0 = (tile % p.num_q_blocks) * p.block_q;
  const int mb_begin = split * p.mblocks_per_split;
  const int mb_end = min(mb_begin + p.mblocks_per_split, p.num_mblocks);
  const int nmb = mb_end - mb_begin;   // host guarantees >= 1
  const int p_boxes = min(2, (p.Cp - p0 + 63) / 64);
  const int q_boxes = min(q_chunks, (p.Cq - q0 + 63) / 64);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_p);
    ptx::prefetch_tmap(&tmap_q);
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nmb; ++i) {
        const int r0 = (mb_begin + i) * BK;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full_bar[stage], (uint32_t)(p_boxes + q_boxes) * 8192u);
        for (int j = 0; j < p_boxes; ++j)
          ptx::tma_load_2d(sp + stage * p_stage + j * 8192, &tmap_p, &full_bar[stage], p0 + 64 * j, r0);
        for (int j = 0; j < q_boxes; ++j)
          ptx::tma_load_2d(sq + stage * q_stage + j * 8192, &tmap_q, &full_bar[stage], q0 + 64 * j, r0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::instr_desc_bf16(BM, p.block_q, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nmb; ++i) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(sp + stage * p_stage);
        const uint32_t b_addr = ptx::smem_u32(sq + stage * q_stage);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          const uint64_t da = ptx::smem_desc(a_addr + k * p.kstep_bytes, p.lbo, p.sbo);
          const uint64_t db = ptx::smem_desc(b_addr + k * p.kstep_bytes, p.lbo, p.sbo);
          ptx::umma_f16(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(done_bar);
    }
  } else {
    const int q = warp & 3;
    const int pr = p0 + q * 32 + lane;
    ptx::mbar_wait(done_bar, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < p.block_q; c0 += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(taddr + c0, r);
      ptx::tmem_ld_wait();
      if (pr < p.Cp) {
        float* o = p.out + (long long)pr * p.so_p;
        if (p.vec4) {      // contiguous q, 16-byte aligned rows: one red.global.add.v4.f32 per 4 columns
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int qc = q0 + c0 + i;
            if (qc < p.q_store)
              atomicAdd(reinterpret_cast<float4*>(o + qc), make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                         __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int qc = q0 + c0 + i;
            if (qc < p.q_store) atomicAdd(o + (long long)qc * p.so_q, __uint_as_float(r[i]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// TEETHRT_GEMM_RES_TILED=0: weights stay resident only when the whole N fits one n-block (the round-1 rule; A/B switch)
const int g_res_tiled_n = [] { const char* e = getenv("TEETHRT_GEMM_RES_TILED"); return (e && *e == '0') ? 0 : 1; }();

int pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

int pick_block_n(int N) {
  if (N <= 256) return (N + 15) / 16 * 16;
  int best = 256, best_pad = 1 << 30;
  const int cand[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    int pad = (N + cand[i] - 1) / cand[i] * cand[i];
    if (pad < best_pad) { best_pad = pad; best = cand[i]; }
  }
  return best;
}

// Shared-memory plan of one tile width: is the n-block's whole weight slab resident, how many epilogue groups (one staging
// buffer + one accumulator each) fit beside the pipeline stages, how many stages.  Used by the launcher and by the tile-width
// model, so the model prices exactly what will be launched.
struct TilePlan { int resident, ngroups, stages; };
TilePlan plan_tile(int block_n, int num_k_blocks, int num_n_blocks, int nostage) {
  TilePlan t;
  const int b_stage = block_n * 128, sms = trt_num_sms();
  t.resident = (num_k_blocks * b_stage <= 64 * 1024 && num_n_blocks <= sms / 2 && !nostage &&
                g_res_tiled_n >= (num_n_blocks > 1 ? 1 : 0)) ? 1 : 0;
  const int stage_bytes = A_STAGE_BYTES + (t.resident ? 0 : b_stage);
  const int fixed_bytes = (t.resident ? num_k_blocks * b_stage : 0) + 2048;
  const int cbuf_bytes = (int)make_layout(block_n, 2, 1, 2, nostage).cbuf_bytes;
  // as many epilogue groups as fit beside min(k-blocks + 1, 3) pipeline stages.  Several groups must each be the ONLY consumer
  // of "their" accumulator barrier (mbarrier parity waits alias if a waiter can fall two phases behind), so with G > 1 groups
  // the accumulators are G as well: tile i -> group i % G -> accumulator i % G.
  const int want_stages = num_k_blocks + 1 < 3 ? num_k_blocks + 1 : 3;
  const int acc_stride = (block_n + 31) & ~31, max_acc = 512 / acc_stride;
  t.ngroups = 1;
  for (int gq = MAX_GROUPS < max_acc ? MAX_GROUPS : max_acc; gq >= 1; --gq)
    if ((220 * 1024 - gq * cbuf_bytes - fixed_bytes) / stage_bytes >= (gq == 1 ? 2 : want_stages)) { t.ngroups = gq; break; }
  int stages = (220 * 1024 - t.ngroups * cbuf_bytes - fixed_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  t.stages = stages;
  return t;
}

// Forward/dgrad tile width: the width whose busiest CTA finishes first under a two-term model fitted to an isolated sweep of
// every 1x1-conv shape of the encoder at batch 64 (tools/gemm_probe.py -> profiles/r02_gemm_probe.jsonl; 44 shapes x up to 5
// widths, each timed inside a CUDA graph):
//   load     = 7 cycles per 128-byte A box row (activations, streamed from HBM) + 2 per B row (weights: L2 hits)
//   epilogue = 48 cycles per output column of a tile, divided by the epilogue groups that work in parallel
//   cost     = max(load, epilogue) on the CTA with the most tiles
// The model picks the measured optimum on all but one of those shapes (total regret 0.9 us per step; the round-1 model, which
// charged every tile a constant epilogue and B rows like A rows, lost 106 us per step to its choices).
int pick_block_n_fwd(int M, int N, int K) {
  const int mb = (M + BM - 1) / BM, kb = (K + BK - 1) / BK, sms = trt_num_sms();
  int cand[5], nc = 0;
  if (N <= 256) cand[nc++] = (N + 15) / 16 * 16;
  const int tiled[4] = {256, 192, 128, 64};
  for (int i = 0; i < 4; ++i)
    if (tiled[i] < N) cand[nc++] = tiled[i];
  int best = cand[0];
  double best_cost = 1e30;
  for (int i = 0; i < nc; ++i) {
    const int bn = cand[i], nb = (N + bn - 1) / bn;
    const TilePlan t = plan_tile(bn, kb, nb, 0);
    const long long tiles = (long long)mb * nb;
    long long grid = tiles < sms ? tiles : sms;
    if (t.resident && nb > 1) grid = grid / nb * nb;
    const long long waves = (tiles + grid - 1) / grid;
    const double a_rows = (double)waves * kb * 128;
    const double b_rows = t.resident ? (double)kb * bn : (double)waves * kb * bn;
    const double load = 7.0 * a_rows + 2.0 * b_rows;
    const double epi = 48.0 * (double)waves * bn / t.ngroups;
    const double cost = load > epi ? load : epi;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

struct MilEpi { int a_kblocks; const float* bias; const float* w; float* score; float* gv; float* gu; };

static int gemm_launch(const void* A, const void* B, void* C, int M, int N, int K, int flags, const float* scale,
                       const float* shift, const void* residual, double* stats, int block_n_override, cudaStream_t stream,
                       const MilEpi* mil = nullptr, const void* bn_x = nullptr) {
  TRT_REQUIRE(A && B && (C || mil), "trt_gemm_bf16: null operand");
  TRT_REQUIRE(M > 0 && N > 0 && K > 0 && (N % 8) == 0 && (K % 8) == 0, "trt_gemm_bf16: M,N,K must be >0 and N,K multiples of 8 (got %d %d %d)", M, N, K);
  TRT_REQUIRE(!(flags & TRT_EPI_SCALE_SHIFT) || (scale && shift), "trt_gemm_bf16: scale/shift missing");
  TRT_REQUIRE(!(flags & TRT_EPI_RESIDUAL) || residual, "trt_gemm_bf16: residual missing");
  TRT_REQUIRE(!(flags & TRT_EPI_STATS) || stats, "trt_gemm_bf16: stats missing");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.block_n = block_n_override > 0 ? block_n_override : pick_block_n_fwd(M, N, K);
  TRT_REQUIRE(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "trt_gemm_bf16: bad block_n %d", p.block_n);
  p.num_m_blocks = (M + BM - 1) / BM;
  p.num_n_blocks = (N + p.block_n - 1) / p.block_n;
  TRT_REQUIRE(p.num_n_blocks == 1 || p.block_n % 64 == 0, "trt_gemm_bf16: tiled N needs block_n %% 64 == 0");
  p.num_k_blocks = (K + BK - 1) / BK;
  p.acc_stride = (p.block_n + 31) & ~31;
  p.nacc = 512 / p.acc_stride < 3 ? 512 / p.acc_stride : 3;
  p.tmem_cols = pow2_cols(p.nacc * p.acc_stride);
  p.flags = flags;
  p.scale = scale; p.shift = shift;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.stats = stats;
  p.bn_x = reinterpret_cast<const __nv_bfloat16*>(bn_x);
  TRT_REQUIRE(!(flags & TRT_EPI_BNBWD) || ((flags & TRT_EPI_STATS) && bn_x && (((uintptr_t)bn_x) & 15) == 0),
              "trt_gemm_bf16: the BatchNorm-backward sums need TRT_EPI_STATS and a 16-byte aligned bn_x");
  p.a_kblocks = 0;
  p.mil_bias = p.mil_w = nullptr; p.mil_score = p.mil_gv = p.mil_gu = nullptr;
  if (mil) {
    p.a_kblocks = mil->a_kblocks; p.mil_bias = mil->bias; p.mil_w = mil->w; p.mil_score = mil->score; p.mil_gv = mil->gv; p.mil_gu = mil->gu;
  }
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  const int sms = trt_num_sms();
  const int nostage = (flags & TRT_EPI_MILGATE) ? 1 : 0;
  const TilePlan tp = plan_tile(p.block_n, p.num_k_blocks, p.num_n_blocks, nostage);
  p.b_resident = tp.resident;
  p.ngroups = tp.ngroups;
  p.stages = tp.stages;
  const int max_acc = 512 / p.acc_stride;
  p.nacc = p.ngroups > 1 ? p.ngroups : (max_acc >= 2 ? 2 : 1);
  p.tmem_cols = pow2_cols(p.nacc * p.acc_stride);
  SmemLayout L = make_layout(p.block_n, p.stages, p.ngroups, p.b_resident ? p.num_k_blocks : p.stages, nostage);
  size_t smem_bytes = (size_t)L.total + 1024;      // slack for the manual 1024B alignment
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // > half an SM's smem: exactly one persistent CTA per SM
  CUtensorMap ta, tb;
  int rc;
  const uint64_t a_cols = p.a_kblocks > 0 ? (uint64_t)p.a_kblocks * BK : (uint64_t)K;     // physical width of A
  if ((rc = trt_make_tmap_2d(&ta, A, (uint64_t)M, a_cols, a_cols, BM, BK))) return rc;
  if ((rc = trt_make_tmap_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)p.block_n, BK))) return rc;
  TRT_REQUIRE((((uintptr_t)C) & 15) == 0, "trt_gemm_bf16: C must be 16-byte aligned");
  TRT_REQUIRE(!(flags & TRT_EPI_MILGATE) || (mil && mil->bias && mil->w && mil->score && (N % 2) == 0 && (mil->gv == nullptr) == (mil->gu == nullptr)),
              "trt_gemm_bf16: incomplete gated-attention epilogue");
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  int grid = tiles < sms ? tiles : sms;
  if (p.b_resident && p.num_n_blocks > 1) grid = grid / p.num_n_blocks * p.num_n_blocks;     // every CTA keeps ONE n-block
#define TRT_GEMM_GO(MODE)                                                                                                          \
  do {                                                                                                                             \
    TRT_CUDA(cudaFuncSetAttribute(gemm_kmajor_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); /* per device */ \
    TRT_CUDA(trt_launch(gemm_kmajor_kernel<MODE>, dim3(grid), dim3(GEMM_THREADS), smem_bytes, stream, ta, tb, p));               \
  } while (0)
  if (flags & TRT_EPI_MILGATE) TRT_GEMM_GO(2);
  else if (flags & TRT_EPI_BNBWD) TRT_GEMM_GO(1);
  else TRT_GEMM_GO(0);
#undef TRT_GEMM_GO
  return trt_check_launch("trt_gemm_bf16");
}

extern "C" int trt_gemm_bf16_bnbwd(const void* A, const void* B, void* C, int M, int N, int K, int flags, const void* residual,
                                   const void* bn_x, double* bstats, cudaStream_t stream) {
  return gemm_launch(A, B, C, M, N, K, (flags & TRT_EPI_RESIDUAL) | TRT_EPI_STATS | TRT_EPI_BNBWD, nullptr, nullptr, residual, bstats,
                     0, stream, nullptr, bn_x);
}

extern "C" int trt_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int flags, const float* scale,
                             const float* shift, const void* residual, double* stats, int block_n_override,
                             cudaStream_t stream) {
  return gemm_launch(A, B, C, M, N, K, flags, scale, shift, residual, stats, block_n_override, stream);
}

#ifdef TRT_GEMM_TIMING
extern "C" int trt_debug_gemm_timing(unsigned long long* out8, int reset) {
  if (out8) cudaMemcpyFromSymbol(out8, g_gemm_dbg, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_gemm_dbg, z, sizeof(z)); }
  return 0;
}
#endif

extern "C" int trt_gemm_wgrad_bf16(const void* P, const void* Q, float* out, int M, int Cp, int Cq, long long so_p,
                                   long long so_q, int q_store, int lbo, int sbo, int kstep_bytes, cudaStream_t stream) {
  TRT_REQUIRE(P && Q && out, "trt_gemm_wgrad_bf16: null operand");
  TRT_REQUIRE(M > 0 && Cp > 0 && Cq > 0 && (Cp % 8) == 0 && (Cq % 8) == 0, "trt_gemm_wgrad_bf16: bad shape %d %d %d", M, Cp, Cq);
  {
    // out[p,q] = sum_m P[m,p] Q[m,q] is symmetric in the roles of P and Q: put the operand on the 128-row UMMA M side that
    // gives fewer output tiles, so each operand column block is streamed from HBM fewer times (e.g. 144 x 24: one tile
    // instead of two; measured 422 MB of DRAM reads for 270 MB of operands before)
    auto tiles_of = [](int cp, int cq) {
      const int bq = cq <= 256 ? cq : pick_block_n(cq);
      return ((cp + BM - 1) / BM) * ((cq + bq - 1) / bq);
    };
    if (q_store <= 0 && tiles_of(Cq, Cp) < tiles_of(Cp, Cq)) {
      const void* t = P; P = Q; Q = t;
      int c = Cp; Cp = Cq; Cq = c;
      long long so = so_p; so_p = so_q; so_q = so;
    }
  }
  WgradParams p;
  p.M = M; p.Cp = Cp; p.Cq = Cq;
  p.block_q = Cq <= 256 ? (Cq + 15) / 16 * 16 : pick_block_n(Cq);
  p.num_p_blocks = (Cp + BM - 1) / BM;
  p.num_q_blocks = (Cq + p.block_q - 1) / p.block_q;
  p.num_mblocks = (M + BK - 1) / BK;
  const int tiles = p.num_p_blocks * p.num_q_blocks;
  // split the row reduction across CTAs, but keep >= 4 k-blocks per CTA: every extra split costs a full tile of atomics.
  // One wave (1 x SMs) measured best inside the train step, where these kernels run on the low-priority side stream
  // beside the data-gradient chain (2 x: +0.07 ms/step, 4 x: +0.16, 8 x: +0.23; tools/gpu_prio_ab.sh)
  static const int split_mult = getenv("TEETHRT_WGRAD_SPLIT_MULT") ? atoi(getenv("TEETHRT_WGRAD_SPLIT_MULT")) : 1;
  int splits = (split_mult * trt_num_sms() + tiles - 1) / tiles;
  if (splits > p.num_mblocks / 4) splits = p.num_mblocks / 4;
  if (splits < 1) splits = 1;
  p.mblocks_per_split = (p.num_mblocks + splits - 1) / splits;
  p.splits = (p.num_mblocks + p.mblocks_per_split - 1) / p.mblocks_per_split;   // no empty split
  p.tmem_cols = pow2_cols(p.block_q);
  p.so_p = so_p; p.so_q = so_q; p.out = out;
  p.q_store = (q_store > 0 && q_store < Cq) ? q_store : Cq;
  p.vec4 = (so_q == 1 && (so_p % 4) == 0 && (p.q_store % 4) == 0 && (((uintptr_t)out) & 15) == 0) ? 1 : 0;
  p.lbo = lbo > 0 ? (uint32_t)lbo : 8192u;
  p.sbo = sbo > 0 ? (uint32_t)sbo : 1024u;
  p.kstep_bytes = kstep_bytes > 0 ? (uint32_t)kstep_bytes : 2048u;
  const int q_chunks = (p.block_q + 63) / 64;
  const int stage_bytes = (2 + q_chunks) * 8192;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > p.mblocks_per_split + 1) stages = p.mblocks_per_split + 1;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + 256 + 1024;
  CUtensorMap tp, tq;
  int rc;
  if ((rc = trt_make_tmap_2d(&tp, P, (uint64_t)M, (uint64_t)Cp, (uint64_t)Cp, BK, 64))) return rc;
  if ((rc = trt_make_tmap_2d(&tq, Q, (uint64_t)M, (uint64_t)Cq, (uint64_t)Cq, BK, 64))) return rc;
  TRT_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));    // per device
  gemm_wgrad_kernel<<<tiles * p.splits, WGRAD_THREADS, smem_bytes, stream>>>(tp, tq, p);
  return trt_check_launch("trt_gemm_wgrad_bf16");
}

// internal (small.cu): A' = [hi | lo | hi] (physically [M, 2*D] bf16), B' = [Whi | Whi | Wlo] ([2*hid, 3*D] bf16, rows
// interleaved V_j, U_j) -> gated-attention scores; see trt_mil_attn_fwd_tc
int trt_gemm_mil_scores(const void* A_split, const void* W_split, int M, int hid, int D, const float* bias2, const float* w,
                        float* score, float* gv, float* gu, cudaStream_t stream) {
  MilEpi mil = {2 * D / BK, bias2, w, score, gv, gu};
  // few instance rows (6 training bags = one 128-row tile): narrow n-blocks put more SMs on the weight stream; partial
  // scores of the n-blocks meet in the atomicAdd
  const int m_tiles = (M + BM - 1) / BM;
  const int bn = (2 * hid) % 64 == 0 && m_tiles * ((2 * hid + 255) / 256) < 32 ? 64 : 0;
  return gemm_launch(A_split, W_split, nullptr, M, 2 * hid, 3 * D, TRT_EPI_MILGATE, nullptr, nullptr, nullptr, nullptr, bn, stream, &mil);
}
