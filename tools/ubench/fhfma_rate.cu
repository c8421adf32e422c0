// Issue-rate microbenchmark: FFMA vs FHFMA.BF16 (fma.rn.f32.bf16) vs MUFU.TANH vs MUFU.EX2+RCP on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma_rate fhfma_rate.cu && ./fhfma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float fhfma(uint32_t a, uint32_t b, float c, bool hi) {
  float d;
  uint16_t ah = hi ? (uint16_t)(a >> 16) : (uint16_t)(a & 0xffff), bh = hi ? (uint16_t)(b >> 16) : (uint16_t)(b & 0xffff);
  asm volatile("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(ah), "h"(bh), "f"(c));
  return d;
}
template <int MODE>
__global__ void k(float* out, const uint32_t* in, int iters) {
  float acc[16];
  uint32_t a = in[threadIdx.x], b = in[threadIdx.x + 1024];
  float fa = __uint_as_float(a), fb = __uint_as_float(b);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(fa), "f"(fb));
      if (MODE == 1) acc[i] = fhfma(a, b, acc[i], i & 1);
      if (MODE == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(acc[i]));
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(acc[i])); }
      if (MODE == 4) { uint32_t h = __float_as_uint(acc[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h)); acc[i] = __uint_as_float(h); }
      if (MODE == 5) { uint32_t h = __float_as_uint(acc[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h)); acc[i] = __uint_as_float(h); }
      if (MODE == 6) { uint32_t h = __float_as_uint(acc[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h)); acc[i] = __uint_as_float(h); }
      if (MODE == 7) {   // the candidate packed sigmoid: 2 x f32 -> f16x2, tanh, affine, back to 2 x f32 (per PAIR of elements)
        uint32_t h; float lo = acc[i], hi = acc[(i + 1) & 15];
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h) : "r"(0x38003800u));
        float a, b;
        asm volatile("{.reg .f16 l, u; mov.b32 {l, u}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, u;}" : "=f"(a), "=f"(b) : "r"(h));
        acc[i] = a + b;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int ops_per) {
  float* out; uint32_t* in;
  cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&in, 8192); cudaMemset(in, 0x3f, 8192);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 1024>>>(out, in, 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(out, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = 148.0 * 2 * 1024 * iters * 16.0 * ops_per;
  printf("%-12s %8.3f ms  %8.2f Tinst/s (thread-level)  = %.1f lanes/clk/SM @1.9GHz\n", name, ms, inst / ms / 1e9, inst / ms / 1e3 / 148 / 1.9e6);
}
int main() {
  run<0>("FFMA", 1); run<1>("FHFMA.BF16", 1); run<2>("MUFU.TANH", 1); run<3>("EX2+RCP", 2);
  run<4>("TANH.F16x2", 1); run<5>("EX2.F16x2", 1); run<6>("TANH.BF16x2", 1); run<7>("sigmoid pair", 1);
  return 0;
}
