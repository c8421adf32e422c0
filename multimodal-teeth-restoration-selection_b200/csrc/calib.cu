// Post-epoch calibration of a fold: temperature-scaled BCE and its derivative for the LBFGS fit,
// calibrated probabilities, and the confusion counts of the whole threshold sweep plus the rank statistic behind ROC-AUC
// in one pass.  Replaces reference train_mm_joint_dualtask.py:162-186 (TemperatureScaler, compute_metrics) and :271-295
// (the fit and the 61-point sweep).  All three are latency-sized kernels (a validation fold is 10^2..10^4 samples): one
// launch each, exact integer counts, fp64 accumulation, deterministic reduction order.
#include "common.cuh"

namespace {

constexpr int CAL_TPB = 512;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0;
  if (w == 0) {
    t = l < (int)(blockDim.x >> 5) ? red[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}
__device__ __forceinline__ long long block_sum_ll(long long v, long long* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  long long t = 0;
  if (w == 0) {
    t = l < (int)(blockDim.x >> 5) ? red[l] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;
}

// loss = mean_i bce(l_i / T, y_i), T = exp(log_T);  d loss / d log_T = mean_i (sigmoid(z_i) - y_i) * (-z_i),  z_i = l_i / T.
// The quotient is rounded to fp32 first, as the reference's `logits / T` tensor is (train_mm_joint_dualtask.py:168-170).
__global__ void __launch_bounds__(CAL_TPB) temperature_nll_kernel(const float* __restrict__ logits, const float* __restrict__ y,
                                                                  const float* __restrict__ log_T, float* __restrict__ out, int n) {
  __shared__ double red[CAL_TPB / 32];
  const float T = expf(log_T[0]);
  double loss = 0, grad = 0;
  for (int i = threadIdx.x; i < n; i += CAL_TPB) {
    const float z = logits[i] / T, t = y[i];
    const float az = fabsf(z);
    loss += (double)(fmaxf(z, 0.f) - z * t + log1pf(expf(-az)));
    const float s = 1.0f / (1.0f + expf(-z));
    grad += (double)((s - t) * -z);
  }
  loss = block_sum(loss, red);
  grad = block_sum(grad, red);
  if (threadIdx.x == 0) { out[0] = (float)(loss / n); out[1] = (float)(grad / n); }
}

__global__ void __launch_bounds__(CAL_TPB) scaled_sigmoid_kernel(const float* __restrict__ logits, float T, float* __restrict__ prob, int n) {
  const int i = blockIdx.x * CAL_TPB + threadIdx.x;
  if (i < n) prob[i] = 1.0f / (1.0f + expf(-(logits[i] / T)));
}

// blocks [0, nthr): confusion counts at thr[b] -> counts[b] = {tp, fp, fn, tn};   prediction is (double)prob >= thr
// blocks [nthr, gridDim): a slice of the positives against every negative -> auc[0] += 2*[p_i > p_j] + [p_i == p_j]
//                         block nthr also writes the class sizes and the number of labels outside {0,1}
__global__ void __launch_bounds__(CAL_TPB) binary_metrics_kernel(const float* __restrict__ prob, const float* __restrict__ y, int n,
                                                                 const double* __restrict__ thr, int nthr,
                                                                 long long* __restrict__ counts, unsigned long long* __restrict__ auc) {
  __shared__ long long red[CAL_TPB / 32];
  const int b = blockIdx.x;
  if (b < nthr) {
    const double t = thr[b];
    long long tp = 0, fp = 0, fn = 0, tn = 0;
    for (int i = threadIdx.x; i < n; i += CAL_TPB) {
      const bool pos = y[i] == 1.0f, pred = (double)prob[i] >= t;
      tp += pos && pred; fp += !pos && pred; fn += pos && !pred; tn += !pos && !pred;
    }
    tp = block_sum_ll(tp, red); fp = block_sum_ll(fp, red); fn = block_sum_ll(fn, red); tn = block_sum_ll(tn, red);
    if (threadIdx.x == 0) { counts[b * 4 + 0] = tp; counts[b * 4 + 1] = fp; counts[b * 4 + 2] = fn; counts[b * 4 + 3] = tn; }
    return;
  }
  const int slice = b - nthr, nslices = gridDim.x - nthr;
  if (slice == 0) {
    long long np = 0, nn = 0, bad = 0;
    for (int i = threadIdx.x; i < n; i += CAL_TPB) {
      const float t = y[i];
      np += t == 1.0f; nn += t == 0.0f; bad += !(t == 1.0f || t == 0.0f);
    }
    np = block_sum_ll(np, red); nn = block_sum_ll(nn, red); bad = block_sum_ll(bad, red);
    if (threadIdx.x == 0) { auc[1] = np; auc[2] = nn; auc[3] = bad; }
  }
  long long wins2 = 0;
  for (int i = slice; i < n; i += nslices) {           // block-uniform: every thread walks the negatives for positive i
    if (y[i] != 1.0f) continue;
    const float pi = prob[i];
    for (int j = threadIdx.x; j < n; j += CAL_TPB) {
      if (y[j] != 0.0f) continue;
      const float pj = prob[j];
      wins2 += pi > pj ? 2 : (pi == pj ? 1 : 0);
    }
  }
  wins2 = block_sum_ll(wins2, red);
  if (threadIdx.x == 0 && wins2) atomicAdd(auc, (unsigned long long)wins2);
}

}  // namespace

extern "C" int trt_temperature_nll(const float* logits, const float* targets, const float* log_T, float* loss_grad, int n,
                                   cudaStream_t stream) {
  TRT_REQUIRE(logits && targets && log_T && loss_grad, "trt_temperature_nll: null pointer");
  TRT_REQUIRE(n > 0, "trt_temperature_nll: empty validation set");
  temperature_nll_kernel<<<1, CAL_TPB, 0, stream>>>(logits, targets, log_T, loss_grad, n);
  return trt_check_launch("trt_temperature_nll");
}

extern "C" int trt_scaled_sigmoid(const float* logits, float T, float* prob, int n, cudaStream_t stream) {
  TRT_REQUIRE(logits && prob, "trt_scaled_sigmoid: null pointer");
  TRT_REQUIRE(n >= 0, "trt_scaled_sigmoid: bad length");
  TRT_REQUIRE(T > 0.f, "trt_scaled_sigmoid: temperature must be positive, got %g", (double)T);
  if (n == 0) return TRT_OK;
  scaled_sigmoid_kernel<<<(n + CAL_TPB - 1) / CAL_TPB, CAL_TPB, 0, stream>>>(logits, T, prob, n);
  return trt_check_launch("trt_scaled_sigmoid");
}

extern "C" int trt_binary_metrics(const float* prob, const float* y, int n, const double* thr, int nthr, long long* counts,
                                  long long* auc, cudaStream_t stream) {
  TRT_REQUIRE(prob && y && auc, "trt_binary_metrics: null pointer");
  TRT_REQUIRE(n > 0, "trt_binary_metrics: empty input");
  TRT_REQUIRE(nthr >= 0 && (nthr == 0 || (thr && counts)), "trt_binary_metrics: threshold buffers missing");
  TRT_CUDA(cudaMemsetAsync(auc, 0, 4 * sizeof(long long), stream));
  const int slices = n < 2 * trt_num_sms() ? n : 2 * trt_num_sms();
  binary_metrics_kernel<<<nthr + slices, CAL_TPB, 0, stream>>>(prob, y, n, thr, nthr, counts,
                                                                 reinterpret_cast<unsigned long long*>(auc));
  return trt_check_launch("trt_binary_metrics");
}
