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
template <typename P>
__global__ void __launch_bounds__(CAL_TPB) binary_metrics_kernel(const P* __restrict__ prob, const float* __restrict__ y, int n,
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
    const P pi = prob[i];
    for (int j = threadIdx.x; j < n; j += CAL_TPB) {
      if (y[j] != 0.0f) continue;
      const P pj = prob[j];
      wins2 += pi > pj ? 2 : (pi == pj ? 1 : 0);
    }
  }
  wins2 = block_sum_ll(wins2, red);
  if (threadIdx.x == 0 && wins2) atomicAdd(auc, (unsigned long long)wins2);
}

// ---------------------------------------------------------------------------------------------------------------------
// Meta-learner of the late-fusion stacker: L2-regularised logistic regression over <= 4 stream
// probabilities, the model sklearn's LogisticRegression(max_iter=1000) fits at experiments/fusion_v1/stack_blend.py:245-247
// and ui/gradio_app/stack_meta.py:55-57:   min_w,b  0.5 |w|^2 + C * sum_i log(1 + exp(-s_i (w.x_i + b))).
// One block runs damped Newton to the optimum (the problem is strictly convex; 5-8 iterations): every pass reduces the
// loss, gradient and Hessian in fp64 over all rows, thread 0 checks the Armijo condition (halving the step on failure),
// solves the (D+1)x(D+1) system by Cholesky and broadcasts the next iterate.
// ---------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(CAL_TPB) logreg_newton_kernel(const double* __restrict__ X, const float* __restrict__ y, int n,
                                                                double Creg, int max_iter, double tol, double* __restrict__ coef,
                                                                double* __restrict__ info) {
  constexpr int D1 = D + 1, NH = D1 * (D1 + 1) / 2, NR = 1 + D1 + NH;
  __shared__ double red[CAL_TPB / 32];
  __shared__ double tot[NR];
  __shared__ double w[D1], w_prev[D1], g_prev[D1];
  __shared__ double f_prev;
  __shared__ int state;          // 0 = keep iterating, 1 = converged
  if (threadIdx.x == 0) { for (int j = 0; j < D1; ++j) { w[j] = 0; w_prev[j] = 0; g_prev[j] = 0; } f_prev = INFINITY; state = 0; }
  __syncthreads();
  int it = 0;
  double gmax = INFINITY;
  for (; it < max_iter; ++it) {
    double acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0;
    for (int i = threadIdx.x; i < n; i += CAL_TPB) {
      double x[D1];
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = X[(size_t)i * D + j];
      x[D] = 1.0;
      double z = 0;
#pragma unroll
      for (int j = 0; j < D1; ++j) z += w[j] * x[j];
      const double t = y[i];
      const double e = exp(-fabs(z));
      acc[0] += fmax(z, 0.0) - t * z + log1p(e);                 // log(1 + e^z) - y z
      const double p = z >= 0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
      const double r = p - t, sw = p * (1.0 - p);
      int h = 1 + D1;
#pragma unroll
      for (int j = 0; j < D1; ++j) {
        acc[1 + j] += r * x[j];
#pragma unroll
        for (int k = j; k < D1; ++k) acc[h++] += sw * x[j] * x[k];
      }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      const double v = block_sum(acc[j], red);
      if (threadIdx.x == 0) tot[j] = v;
    }
    if (threadIdx.x == 0) {
      double f = Creg * tot[0], g[D1], H[D1][D1];
      for (int j = 0; j < D; ++j) f += 0.5 * w[j] * w[j];
      int h = 1 + D1;
      for (int j = 0; j < D1; ++j) {
        g[j] = Creg * tot[1 + j] + (j < D ? w[j] : 0.0);
        for (int k = j; k < D1; ++k) { H[j][k] = H[k][j] = Creg * tot[h++] + ((j == k && j < D) ? 1.0 : 0.0); }
      }
      double slope = 0;
      for (int j = 0; j < D1; ++j) slope += g_prev[j] * (w[j] - w_prev[j]);
      if (f > f_prev + 1e-4 * slope && it > 0) {                 // Armijo failed: halve the step from the accepted point
        for (int j = 0; j < D1; ++j) w[j] = w_prev[j] + 0.5 * (w[j] - w_prev[j]);
      } else {
        gmax = 0;
        for (int j = 0; j < D1; ++j) gmax = fmax(gmax, fabs(g[j]));
        for (int j = 0; j < D1; ++j) { w_prev[j] = w[j]; g_prev[j] = g[j]; }
        f_prev = f;
        if (gmax <= tol) state = 1;
        else {
          // Cholesky H = L L^T, solve H d = -g
          double L[D1][D1];
          bool ok = true;
          for (int j = 0; j < D1 && ok; ++j) {
            double d = H[j][j];
            for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
            if (!(d > 0)) { ok = false; break; }
            L[j][j] = sqrt(d);
            for (int i2 = j + 1; i2 < D1; ++i2) {
              double v = H[i2][j];
              for (int k = 0; k < j; ++k) v -= L[i2][k] * L[j][k];
              L[i2][j] = v / L[j][j];
            }
          }
          double dl[D1];
          if (ok) {
            for (int j = 0; j < D1; ++j) { double v = -g[j]; for (int k = 0; k < j; ++k) v -= L[j][k] * dl[k]; dl[j] = v / L[j][j]; }
            for (int j = D1 - 1; j >= 0; --j) { double v = dl[j]; for (int k = j + 1; k < D1; ++k) v -= L[k][j] * dl[k]; dl[j] = v / L[j][j]; }
          } else {
            for (int j = 0; j < D1; ++j) dl[j] = -g[j] / (fabs(H[j][j]) + 1.0);   // not reached for C > 0, n > 0: H is PD
          }
          for (int j = 0; j < D1; ++j) w[j] = w_prev[j] + dl[j];
        }
      }
    }
    __syncthreads();
    if (state) break;
  }
  if (threadIdx.x == 0) {
    for (int j = 0; j < D1; ++j) coef[j] = w_prev[j];
    info[0] = (double)(it + (state ? 1 : 0)); info[1] = gmax; info[2] = f_prev;
  }
}

__global__ void __launch_bounds__(CAL_TPB) logreg_predict_kernel(const double* __restrict__ X, int n, int d, const double* __restrict__ coef,
                                                                 double* __restrict__ p) {
  const int i = blockIdx.x * CAL_TPB + threadIdx.x;
  if (i >= n) return;
  double z = coef[d];
  for (int j = 0; j < d; ++j) z += coef[j] * X[(size_t)i * d + j];
  const double e = exp(-fabs(z));
  p[i] = z >= 0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
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

extern "C" int trt_binary_metrics(const void* prob, int prob_is_f64, const float* y, int n, const double* thr, int nthr,
                                  long long* counts, long long* auc, cudaStream_t stream) {
  TRT_REQUIRE(prob && y && auc, "trt_binary_metrics: null pointer");
  TRT_REQUIRE(n > 0, "trt_binary_metrics: empty input");
  TRT_REQUIRE(nthr >= 0 && (nthr == 0 || (thr && counts)), "trt_binary_metrics: threshold buffers missing");
  TRT_CUDA(cudaMemsetAsync(auc, 0, 4 * sizeof(long long), stream));
  const int slices = n < 2 * trt_num_sms() ? n : 2 * trt_num_sms();
  if (prob_is_f64)
    binary_metrics_kernel<double><<<nthr + slices, CAL_TPB, 0, stream>>>(static_cast<const double*>(prob), y, n, thr, nthr, counts,
                                                                           reinterpret_cast<unsigned long long*>(auc));
  else
    binary_metrics_kernel<float><<<nthr + slices, CAL_TPB, 0, stream>>>(static_cast<const float*>(prob), y, n, thr, nthr, counts,
                                                                          reinterpret_cast<unsigned long long*>(auc));
  return trt_check_launch("trt_binary_metrics");
}

extern "C" int trt_logreg_fit(const double* X, const float* y, int n, int d, double C, int max_iter, double tol, double* coef,
                              double* info, cudaStream_t stream) {
  TRT_REQUIRE(X && y && coef && info, "trt_logreg_fit: null pointer");
  TRT_REQUIRE(n > 0 && C > 0 && max_iter > 0, "trt_logreg_fit: bad argument n=%d C=%g max_iter=%d", n, C, max_iter);
  TRT_REQUIRE(d >= 1 && d <= 4, "trt_logreg_fit: %d features not built (1..4 stream probabilities)", d);
  switch (d) {
    case 1: logreg_newton_kernel<1><<<1, CAL_TPB, 0, stream>>>(X, y, n, C, max_iter, tol, coef, info); break;
    case 2: logreg_newton_kernel<2><<<1, CAL_TPB, 0, stream>>>(X, y, n, C, max_iter, tol, coef, info); break;
    case 3: logreg_newton_kernel<3><<<1, CAL_TPB, 0, stream>>>(X, y, n, C, max_iter, tol, coef, info); break;
    default: logreg_newton_kernel<4><<<1, CAL_TPB, 0, stream>>>(X, y, n, C, max_iter, tol, coef, info); break;
  }
  return trt_check_launch("trt_logreg_fit");
}

extern "C" int trt_logreg_predict(const double* X, int n, int d, const double* coef, double* prob, cudaStream_t stream) {
  TRT_REQUIRE(X && coef && prob, "trt_logreg_predict: null pointer");
  TRT_REQUIRE(n > 0 && d >= 1, "trt_logreg_predict: bad shape n=%d d=%d", n, d);
  logreg_predict_kernel<<<(n + CAL_TPB - 1) / CAL_TPB, CAL_TPB, 0, stream>>>(X, n, d, coef, prob);
  return trt_check_launch("trt_logreg_predict");
}
