// Shared device/host helpers for libteethrt (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/teethrt.h"

#if defined(__CUDA_ARCH__) && !(__CUDA_ARCH__ >= 1000)
#error "libteethrt is written for sm_100a (B200) only; compile with -gencode arch=compute_100a,code=sm_100a"
#endif

// ---------------------------------------------------------------- status / error plumbing (abi.cu)
int trt_set_error(int status, const char* fmt, ...);
int trt_check_launch(const char* what);   // also counts one kernel launch
void trt_count_launch(int n);             // extra launches of entry points that enqueue more than one kernel
#define TRT_REQUIRE(cond, ...) do { if (!(cond)) return trt_set_error(TRT_ERR_INVALID, __VA_ARGS__); } while (0)
#define TRT_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return trt_set_error(TRT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); } while (0)

int trt_num_sms();

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The batch-1 forward is ~200 dependent kernels of a few microseconds each: launch latency, not work.  A kernel launched
// through trt_launch() with TEETHRT_PDL != 0 carries cudaLaunchAttributeProgrammaticStreamSerialization: it may start while
// its predecessor in the stream is still draining, runs its prologue (barrier init, descriptor prefetch, index math) and
// blocks in pdl_wait() until the predecessor has completed and its writes are visible.  Every kernel launched this way calls
// pdl_launch_dependents() first (so ITS successor can be staged early) and pdl_wait() before touching global memory; both are
// no-ops in a launch without the attribute.  Works inside stream capture (programmatic graph edges).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool trt_pdl_enabled();   // abi.cu: TEETHRT_PDL (default on)
template <typename... KArgs, typename... Args>
inline cudaError_t trt_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = trt_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- small device utilities
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct f8 { float v[8]; };
__device__ __forceinline__ f8 unpack8(const uint4& q) {
  f8 r;
  r.v[0] = bf16_lo(q.x); r.v[1] = bf16_hi(q.x); r.v[2] = bf16_lo(q.y); r.v[3] = bf16_hi(q.y);
  r.v[4] = bf16_lo(q.z); r.v[5] = bf16_hi(q.z); r.v[6] = bf16_lo(q.w); r.v[7] = bf16_hi(q.w);
  return r;
}
__device__ __forceinline__ uint4 pack8(const f8& r) {
  uint4 q;
  q.x = pack_bf16(r.v[0], r.v[1]); q.y = pack_bf16(r.v[2], r.v[3]);
  q.z = pack_bf16(r.v[4], r.v[5]); q.w = pack_bf16(r.v[6], r.v[7]);
  return q;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ f8 ldf8(const float* p) {  // 8 consecutive fp32 (32 B aligned)
  f8 r;
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

__device__ __forceinline__ f8 lds8(const float* p) {  // 8 consecutive fp32 in shared memory (32 B aligned)
  f8 r;
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// sigmoid = rcp(1 + 2^(-x log2 e)): two MUFU ops, no range fix-ups needed (2^+big = inf -> rcp = 0; 2^-big = 0 -> rcp(1) = 1)
__device__ __forceinline__ float sigmoidf_(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// silu(x*sc + sh) with the exponent argument as its own FFMA: sc2 = -log2(e)*sc, sh2 = -log2(e)*sh (one FMUL less per element)
__device__ __forceinline__ float silu_affine_(float x, float sc, float sh, float sc2, float sh2) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(x, sc2, sh2)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(x, sc, sh) * r;
}
// d/dx silu(x) = s * (1 + x * (1 - s))
__device__ __forceinline__ float silu_gradf_(float x) { float s = sigmoidf_(x); return s * (1.0f + x * (1.0f - s)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Align the dynamic shared-memory base WITHOUT laundering it through an integer: an offset added to the __shared__ symbol
// keeps the pointer in the shared address space (LDS/STS, no aliasing with global stores); a uintptr_t round trip turns
// every access into a generic LD.E/ST.E.
#define TRT_ALIGNED_SMEM(raw, align) \
  ((raw) + ((((uint32_t)__cvta_generic_to_shared(raw) + (uint32_t)((align) - 1)) & ~(uint32_t)((align) - 1)) - \
            (uint32_t)__cvta_generic_to_shared(raw)))

// ---------------------------------------------------------------- lazy BatchNorm records (consumer-side finalisation)
// total of a replicated statistics entry: stats is [TRT_STAT_REPLICAS][2][C]; idx in [0, 2C)
__device__ __forceinline__ double stat_total(const double* stats, int C, int idx) {
  double t = 0;
#pragma unroll
  for (int r = 0; r < TRT_STAT_REPLICAS; ++r) t += __ldcg(stats + (size_t)r * 2 * C + idx);
  return t;
}
// One channel of a train-mode BatchNorm: fp64 {sum, sum^2} -> {scale, shift, mean, rstd}.  The ONE definition of that
// arithmetic: bn_finalize_kernel and every lazy consumer call it, so a consumer that derives scale/shift from the
// statistics gets bit-identical values to a later kernel that reads the published record.
struct BnChannel { float sc, sh, mean, rstd; double var; };
__device__ __forceinline__ BnChannel bn_channel(const double* stats, const float* gamma, const float* beta, int C, int c,
                                                double count, float eps) {
  BnChannel r;
  const double mean = stat_total(stats, C, c) / count;
  double var = stat_total(stats, C, C + c) / count - mean * mean;
  if (var < 0) var = 0;
  r.var = var;
  r.rstd = (float)(1.0 / sqrt(var + (double)eps));
  r.mean = (float)mean;
  r.sc = gamma[c] * r.rstd;
  r.sh = beta[c] - r.mean * r.sc;
  return r;
}
__device__ __forceinline__ void bn_publish(const BnChannel& b, float* rec, float* running_mean, float* running_var, int C,
                                           int c, double count, float momentum) {
  rec[c] = b.sc;
  rec[C + c] = b.sh;
  rec[2 * C + c] = b.mean;
  rec[3 * C + c] = b.rstd;
  if (running_mean) {
    const double unb = count > 1 ? b.var * count / (count - 1) : b.var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * b.mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}
// Called by ALL threads of a block: scale/shift of channels [c0, c0 + nch) into shared memory, straight from the producer's
// statistics (16 independent L2 loads per channel, ~1 us, instead of a 4-5 us finalise launch between producer and
// consumer).  The publishing block also writes the record and the running statistics for every later consumer.
__device__ __forceinline__ void bn_lazy_block(const trt_bn_fin_t& f, int C, int c0, int nch, float* s_sc, float* s_sh,
                                              bool publish) {
  for (int i = threadIdx.x; i < nch; i += blockDim.x) {
    const int c = c0 + i;
    if (c < C) {
      const BnChannel b = bn_channel(f.stats, f.gamma, f.beta, C, c, f.count, f.eps);
      s_sc[i] = b.sc;
      s_sh[i] = b.sh;
      if (publish) bn_publish(b, f.rec, f.running_mean, f.running_var, C, c, f.count, f.momentum);
    }
  }
  if (publish && c0 == 0 && threadIdx.x == 0 && f.num_batches_tracked) *f.num_batches_tracked += 1;
  __syncthreads();
}
// backward: bstats {sum dy, sum dy*xhat} -> dx = a*dy + b*x + c for one channel (what bn_bwd_finalize_kernel publishes)
struct BnBwdChannel { float a, b, c, dgamma, dbeta; };
__device__ __forceinline__ BnBwdChannel bn_bwd_channel(const double* bstats, const float* rec, const float* gamma, int C, int c,
                                                       double count, int raw_x = 0) {
  BnBwdChannel r;
  const float mean = rec[2 * C + c], rstd = rec[3 * C + c];
  const double sdy = stat_total(bstats, C, c);
  double sdyx = stat_total(bstats, C, C + c);
  if (raw_x) sdyx = (double)rstd * (sdyx - (double)mean * sdy);      // the producer summed dy * x: -> sum dy * xhat
  const float a = gamma[c] * rstd;
  const float m1 = (float)(sdy / count), m2 = (float)(sdyx / count);
  r.a = a;
  r.b = -a * rstd * m2;
  r.c = a * (mean * rstd * m2 - m1);
  r.dgamma = (float)sdyx;
  r.dbeta = (float)sdy;
  return r;
}

// ---------------------------------------------------------------- PTX: mbarrier / TMA / tcgen05
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// waiting with back-off: a warp that spins on try_wait keeps taking issue slots from the warps that share its scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns = 64) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled TMA load: c0 = coordinate along the contiguous dim, c1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 4D tiled TMA load (NHWC activations: c0 = channel, c1 = x, c2 = y, c3 = image); coordinates may be negative or run past
// the tensor: out-of-bounds elements arrive as zeros, which is exactly the convolution's zero padding.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// fp32 acc += bf16 * bf16 (FHFMA.BF16), operands straight from 16-bit register halves
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}

// --- tensor memory
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets TMEM lane (taddr.lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// --- descriptors (bit layout: cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS; PTX ISA tcgen05 "matrix descriptor")
constexpr uint32_t LAYOUT_SW128 = 2;
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;              // descriptor version 1 (sm_100)
  d |= (uint64_t)LAYOUT_SW128 << 61;   // 128-byte swizzle
  return d;
}
// bf16 x bf16 -> fp32, dense.  a_mn / b_mn: 1 = operand is MN-major in shared memory, 0 = K-major.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
}  // namespace ptx

// ---------------------------------------------------------------- host: TMA descriptor encode (driver entry point, no -lcuda)
// 2D row-major bf16 tensor [rows, cols] (cols contiguous, row pitch = ld elements), box [box_rows, box_cols], 128B swizzle.
int trt_make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_rows, uint32_t box_cols);
// 4D NHWC bf16 tensor [N,H,W,C] (C % 8 == 0), box [1, box_h, box_w, box_c], no swizzle (smem tile = [box_h][box_w][box_c]).
int trt_make_tmap_nhwc(CUtensorMap* out, const void* base, int N, int H, int W, int C, int box_c, int box_w, int box_h);
