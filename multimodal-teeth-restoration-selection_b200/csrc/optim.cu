// Fused optimiser step over FLAT fp32 parameter / gradient / moment buffers: global grad-norm (clip_grad_norm_),
// AdamW with decoupled weight decay, per-iteration cosine learning rate — the step counter and the schedule live on the
// device so the whole train step replays as one CUDA graph.  HBM-bound: 28 B per parameter (read p,g,m,v; write p,m,v).
//
// Replaces: GradScaler.unscale_ + clip_grad_norm_(…, 1.0) + AdamW.step + CosineAnnealingLR.step,
// experiments/multimodal_v1/train_mm_joint_dualtask.py:217-220,249-254 (same recipe at
// experiments/vision_v2/train_mil_attention_v1.py:170-188).
#include "common.cuh"

namespace {

struct OptState {               // layout mirrored by teethrt/optim.py (8 x 8 bytes)
  unsigned long long step;      // number of optimiser steps taken
  double lr0, t_max, beta1, beta2;
  float lr, bc1, bc2, pad;      // values for the CURRENT step (written by optim_advance)
  double skipped;               // optimiser steps skipped because the gradient norm was not finite
};

__global__ void optim_advance_kernel(OptState* s) {
  const unsigned long long t = ++s->step;
  double lr = s->lr0;
  if (s->t_max > 0) lr = s->lr0 * (1.0 + cos(3.14159265358979323846 * (double)(t - 1) / s->t_max)) * 0.5;
  s->lr = (float)lr;
  s->bc1 = (float)(1.0 - pow(s->beta1, (double)t));
  s->bc2 = (float)(1.0 - pow(s->beta2, (double)t));
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ g, size_t n4, const float* __restrict__ tail,
                                                    int ntail, double* __restrict__ out) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g + i);
    acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) acc = fmaf(tail[threadIdx.x], tail[threadIdx.x], acc);
  __shared__ float s[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += (double)s[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, size_t n, OptState* __restrict__ st,
                                                    const double* __restrict__ normsq, float* __restrict__ norm_out,
                                                    float gscale, float max_norm, float eps, float wd) {
  const float lr = st->lr, bc1 = st->bc1, bc2 = st->bc2;
  const float b1 = (float)st->beta1, b2 = (float)st->beta2;
  float coef = gscale;
  if (normsq) {
    const float total = sqrtf((float)normsq[0]) * gscale;     // clip_grad_norm_: norm of the (averaged) gradient
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) norm_out[0] = total;
    // one inf/NaN gradient (bf16 overflow) would poison every parameter through coef: leave p/m/v untouched and count the
    // skip, which is what the reference's GradScaler.step does under --amp (train_mm_joint_dualtask.py:249-253)
    if (!isfinite(total)) {
      if (blockIdx.x == 0 && threadIdx.x == 0) st->skipped += 1.0;
      return;
    }
    if (max_norm > 0.f) coef *= fminf(1.0f, max_norm / (total + 1e-6f));
  }
  const float decay = 1.0f - lr * wd, step_size = lr / bc1, rsq_bc2 = rsqrtf(bc2);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] * decay - step_size * mi / (sqrtf(vi) * rsq_bc2 + eps);
  }
}

}  // namespace

extern "C" size_t trt_optim_state_bytes(void) { return sizeof(OptState); }

extern "C" int trt_optim_advance(void* state, cudaStream_t stream) {
  TRT_REQUIRE(state, "trt_optim_advance: null state");
  optim_advance_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<OptState*>(state));
  return trt_check_launch("trt_optim_advance");
}

extern "C" int trt_grad_sumsq(const float* g, size_t n, double* out, cudaStream_t stream) {
  TRT_REQUIRE(g && out && n > 0 && (((uintptr_t)g) & 15) == 0, "trt_grad_sumsq: bad argument (16-byte aligned buffer needed)");
  TRT_CUDA(cudaMemsetAsync(out, 0, sizeof(double), stream));
  const size_t n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  const int cap = 8 * trt_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  sumsq_kernel<<<grid, 256, 0, stream>>>((const float4*)g, n4, g + n4 * 4, (int)(n - n4 * 4), out);
  return trt_check_launch("trt_grad_sumsq");
}

extern "C" int trt_adamw_step(float* p, const float* g, float* m, float* v, size_t n, const void* state,
                              const double* normsq, float* norm_out, float grad_scale, float max_norm, float eps,
                              float weight_decay, cudaStream_t stream) {
  TRT_REQUIRE(p && g && m && v && state && n > 0, "trt_adamw_step: bad argument");
  int grid = (int)((n + 255) / 256);
  const int cap = 16 * trt_num_sms();
  if (grid > cap) grid = cap;
  adamw_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n, reinterpret_cast<OptState*>(const_cast<void*>(state)), normsq, norm_out,
                                         grad_scale, max_norm, eps, weight_decay);
  return trt_check_launch("trt_adamw_step");
}
