// How long does a "last block done" counter cost?  G blocks each do __threadfence + one atomicAdd (with return) on ONE word.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_counter(unsigned* c, unsigned* out) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { unsigned t = atomicAdd(c, 1u); if (t == gridDim.x - 1) *out = t; }
}
__global__ void k_empty(unsigned* c, unsigned* out) { if (threadIdx.x == 9999) *out = *c; }
__global__ void k_f64(double* d, int n) {  // every block adds to the same n doubles
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(d + i, 1.0);
}
int main() {
  unsigned *c, *o; double* d;
  cudaMalloc(&c, 4); cudaMalloc(&o, 4); cudaMalloc(&d, 8 * 4096); cudaMemset(c, 0, 4); cudaMemset(d, 0, 8 * 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int G : {148, 296, 444, 1776, 7104}) {
    float ms_e, ms_c, ms_f;
    k_empty<<<G, 256>>>(c, o); cudaEventRecord(e0); for (int i = 0; i < 20; ++i) k_empty<<<G, 256>>>(c, o); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_e, e0, e1);
    k_counter<<<G, 256>>>(c, o); cudaEventRecord(e0); for (int i = 0; i < 20; ++i) k_counter<<<G, 256>>>(c, o); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c, e0, e1);
    k_f64<<<G, 256>>>(d, 2048); cudaEventRecord(e0); for (int i = 0; i < 20; ++i) k_f64<<<G, 256>>>(d, 2048); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_f, e0, e1);
    printf("G=%5d: empty %.2f us  counter %.2f us  (+%.2f us)   f64 atomics on 2048 shared addresses %.2f us\n", G, ms_e * 50, ms_c * 50, (ms_c - ms_e) * 50, ms_f * 50);
  }
  return 0;
}
