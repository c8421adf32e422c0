// The 3x3 stride-2 stem convolution (K = 27 is too thin for the tensor cores), forward and weight gradient,
// TF-'same' padding, NCHW fp32/bf16 in -> NHWC bf16 out, BN statistics (train) or folded BN + SiLU (eval) in the epilogue.
// (The depthwise convolutions live in dwconv.cu.)
//
// Replaces cuDNN/ATen depthwise + stem conv launches inside `self.backbone(x_img)`
// (experiments/multimodal_v1/train_mm_joint_dualtask.py:154) and their autograd backward (:248).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ stem 3x3 s2 conv (3 -> CS)
// x: NCHW (fp32 or bf16) ; out: NHWC bf16.  One thread = one output pixel, all CS channels in registers.
template <int CS, typename InT>
__global__ void __launch_bounds__(128) stem_fwd_kernel(const InT* __restrict__ x, const float* __restrict__ w,
                                                       __nv_bfloat16* __restrict__ out, const float* __restrict__ out_rec,
                                                       double* __restrict__ stats, int N, int H, int W, int OH, int OW,
                                                       int pad_t, int pad_l) {
  __shared__ __align__(16) float s_w[27 * CS];
  __shared__ __align__(16) __nv_bfloat16 s_o[128 * CS];
  for (int i = threadIdx.x; i < 27 * CS; i += 128) {
    const int tap = i / CS, co = i % CS;
    s_w[i] = __ldg(w + co * 27 + tap);
  }
  __syncthreads();
  const long long total = (long long)N * OH * OW;
  const long long pix = (long long)blockIdx.x * 128 + threadIdx.x;
  const bool valid = pix < total;
  float acc[CS];
#pragma unroll
  for (int i = 0; i < CS; ++i) acc[i] = 0.f;
  if (valid) {
    const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iy = oy * 2 - pad_t + kh, ix = ox * 2 - pad_l + kw;
          float v = 0.f;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = (float)x[((size_t)(n * 3 + ci) * H + iy) * W + ix];
          v = bf16_round(v);   // the tensor-core path of the other layers consumes bf16 activations: keep the stem consistent
          const float* wr = s_w + (ci * 9 + kh * 3 + kw) * CS;
#pragma unroll
          for (int co = 0; co < CS; co += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + co);
            acc[co] = fmaf(v, wv.x, acc[co]); acc[co + 1] = fmaf(v, wv.y, acc[co + 1]);
            acc[co + 2] = fmaf(v, wv.z, acc[co + 2]); acc[co + 3] = fmaf(v, wv.w, acc[co + 3]);
          }
        }
    if (out_rec) {
#pragma unroll
      for (int co = 0; co < CS; ++co) acc[co] = siluf_(fmaf(acc[co], out_rec[co], out_rec[CS + co]));
    }
  }
  uint4* orow = reinterpret_cast<uint4*>(s_o + threadIdx.x * CS);
#pragma unroll
  for (int j = 0; j < CS / 8; ++j) {
    uint4 q;
    q.x = pack_bf16(acc[8 * j], acc[8 * j + 1]); q.y = pack_bf16(acc[8 * j + 2], acc[8 * j + 3]);
    q.z = pack_bf16(acc[8 * j + 4], acc[8 * j + 5]); q.w = pack_bf16(acc[8 * j + 6], acc[8 * j + 7]);
    orow[j] = q;
    if (valid) reinterpret_cast<uint4*>(out + (size_t)pix * CS)[j] = q;
  }
  if (stats) {
    __syncthreads();
    if (threadIdx.x < CS) {
      float s = 0.f, q = 0.f;
      const long long rem = total - (long long)blockIdx.x * 128;
      const int cnt = rem < 128 ? (int)rem : 128;
      for (int r = 0; r < cnt; ++r) {
        const float v = __bfloat162float(s_o[r * CS + threadIdx.x]);
        s += v; q = fmaf(v, v, q);
      }
      double* rep = stats + (size_t)(blockIdx.x % TRT_STAT_REPLICAS) * 2 * CS;
      atomicAdd(rep + threadIdx.x, (double)s);
      atomicAdd(rep + CS + threadIdx.x, (double)q);
    }
  }
}

// dW[co][tap] += sum_pix dS[pix][co] * patch[pix][tap]
template <int CS, typename InT>
__global__ void __launch_bounds__(128) stem_wgrad_kernel(const InT* __restrict__ x, const __nv_bfloat16* __restrict__ ds,
                                                         float* __restrict__ dw, int N, int H, int W, int OH, int OW,
                                                         int pad_t, int pad_l, int chunks_per_block) {
  constexpr int PX = 64;
  __shared__ __align__(16) float s_ds[PX * CS];
  __shared__ float s_patch[PX * 28];
  const long long total = (long long)N * OH * OW;
  const int cog = threadIdx.x % (CS / 4), tg = threadIdx.x / (CS / 4);   // 4 output channels x 3 taps per thread
  const bool worker = tg < 9;
  float acc[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) acc[a][b] = 0.f;
  for (int ch = 0; ch < chunks_per_block; ++ch) {
    const long long p0 = ((long long)blockIdx.x * chunks_per_block + ch) * PX;
    if (p0 >= total) break;
    __syncthreads();
    for (int i = threadIdx.x; i < PX * CS; i += 128) {
      const long long pix = p0 + i / CS;
      s_ds[i] = pix < total ? __bfloat162float(ds[(size_t)pix * CS + i % CS]) : 0.f;
    }
    for (int i = threadIdx.x; i < PX * 27; i += 128) {
      const int p = i / 27, tap = i % 27;
      const long long pix = p0 + p;
      float v = 0.f;
      if (pix < total) {
        const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
        const int ci = tap / 9, kh = (tap % 9) / 3, kw = tap % 3;
        const int iy = oy * 2 - pad_t + kh, ix = ox * 2 - pad_l + kw;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = bf16_round((float)x[((size_t)(n * 3 + ci) * H + iy) * W + ix]);
      }
      s_patch[p * 28 + tap] = v;
    }
    __syncthreads();
    if (worker) {
#pragma unroll 4
      for (int p = 0; p < PX; ++p) {
        const float4 d = *reinterpret_cast<const float4*>(s_ds + p * CS + cog * 4);
        const float t0 = s_patch[p * 28 + tg * 3], t1 = s_patch[p * 28 + tg * 3 + 1], t2 = s_patch[p * 28 + tg * 3 + 2];
        acc[0][0] = fmaf(d.x, t0, acc[0][0]); acc[0][1] = fmaf(d.x, t1, acc[0][1]); acc[0][2] = fmaf(d.x, t2, acc[0][2]);
        acc[1][0] = fmaf(d.y, t0, acc[1][0]); acc[1][1] = fmaf(d.y, t1, acc[1][1]); acc[1][2] = fmaf(d.y, t2, acc[1][2]);
        acc[2][0] = fmaf(d.z, t0, acc[2][0]); acc[2][1] = fmaf(d.z, t1, acc[2][1]); acc[2][2] = fmaf(d.z, t2, acc[2][2]);
        acc[3][0] = fmaf(d.w, t0, acc[3][0]); acc[3][1] = fmaf(d.w, t1, acc[3][1]); acc[3][2] = fmaf(d.w, t2, acc[3][2]);
      }
    }
  }
  if (worker) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) atomicAdd(dw + (cog * 4 + a) * 27 + tg * 3 + b, acc[a][b]);
  }
}

// im2col of the stem: NCHW image -> patches [N*OH*OW, 32] bf16, column = ci*9 + kh*3 + kw (27 taps, 5 zero columns), so the
// train-mode stem runs as tcgen05 GEMMs: forward = patches . W^T (BN statistics in the epilogue), weight gradient =
// dS^T . patches.  One thread per output pixel; the 64-byte row is written as four 16-byte stores.
template <typename InT>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const InT* __restrict__ x, uint4* __restrict__ patches, int N, int H,
                                                          int W, int OH, int OW, int pad_t, int pad_l) {
  const long long total = (long long)N * OH * OW;
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= total) return;
  const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
  float v[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * 2 - pad_t + kh;
      const InT* row = x + ((size_t)(n * 3 + ci) * H + (iy >= 0 && iy < H ? iy : 0)) * W;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * 2 - pad_l + kw;
        v[ci * 9 + kh * 3 + kw] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? (float)row[ix] : 0.f;
      }
    }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_bf16(v[8 * j], v[8 * j + 1]); q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
    q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
    patches[(size_t)pix * 4 + j] = q;
  }
}

// stem weight [CS,3,3,3] fp32 -> bf16 [CS,32] (27 taps + 5 zero columns)
__global__ void stem_pack_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int CS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CS * 32) return;
  const int co = i / 32, t = i % 32;
  o[i] = __float2bfloat16_rn(t < 27 ? w[co * 27 + t] : 0.f);
}

inline void same_pad(int i, int k, int s, int& out, int& pad_before) {
  out = (i + s - 1) / s;
  int total = (out - 1) * s + k - i;
  if (total < 0) total = 0;
  pad_before = total / 2;
}

}  // namespace

extern "C" int trt_stem_fwd(const void* x, int x_is_bf16, const float* w, void* out, const float* out_rec, double* stats,
                            int N, int H, int W, int CS, cudaStream_t stream) {
  TRT_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0, "trt_stem_fwd: bad argument");
  TRT_REQUIRE(CS == 32 || CS == 48, "trt_stem_fwd: stem width %d not built (32 = B0, 48 = B4)", CS);
  int OH, OW, pt, pl;
  same_pad(H, 3, 2, OH, pt);
  same_pad(W, 3, 2, OW, pl);
  const long long total = (long long)N * OH * OW;
  const int grid = (int)((total + 127) / 128);
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  if (CS == 32) {
    if (x_is_bf16) stem_fwd_kernel<32, __nv_bfloat16><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)x, w, o, out_rec, stats, N, H, W, OH, OW, pt, pl);
    else stem_fwd_kernel<32, float><<<grid, 128, 0, stream>>>((const float*)x, w, o, out_rec, stats, N, H, W, OH, OW, pt, pl);
  } else {
    if (x_is_bf16) stem_fwd_kernel<48, __nv_bfloat16><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)x, w, o, out_rec, stats, N, H, W, OH, OW, pt, pl);
    else stem_fwd_kernel<48, float><<<grid, 128, 0, stream>>>((const float*)x, w, o, out_rec, stats, N, H, W, OH, OW, pt, pl);
  }
  return trt_check_launch("trt_stem_fwd");
}

extern "C" int trt_stem_im2col(const void* x, int x_is_bf16, void* patches, int N, int H, int W, cudaStream_t stream) {
  TRT_REQUIRE(x && patches && N > 0 && H > 0 && W > 0, "trt_stem_im2col: bad argument");
  int OH, OW, pt, pl;
  same_pad(H, 3, 2, OH, pt);
  same_pad(W, 3, 2, OW, pl);
  const long long total = (long long)N * OH * OW;
  const int grid = (int)((total + 255) / 256);
  if (x_is_bf16) stem_im2col_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (uint4*)patches, N, H, W, OH, OW, pt, pl);
  else stem_im2col_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (uint4*)patches, N, H, W, OH, OW, pt, pl);
  return trt_check_launch("trt_stem_im2col");
}

extern "C" int trt_stem_pack_w(const float* w, void* w_bf16, int CS, cudaStream_t stream) {
  TRT_REQUIRE(w && w_bf16 && CS > 0, "trt_stem_pack_w: bad argument");
  stem_pack_w_kernel<<<(CS * 32 + 127) / 128, 128, 0, stream>>>(w, (__nv_bfloat16*)w_bf16, CS);
  return trt_check_launch("trt_stem_pack_w");
}

extern "C" int trt_stem_wgrad(const void* x, int x_is_bf16, const void* ds, float* dw, int N, int H, int W, int CS,
                              cudaStream_t stream) {
  TRT_REQUIRE(x && ds && dw && N > 0, "trt_stem_wgrad: bad argument");
  TRT_REQUIRE(CS == 32 || CS == 48, "trt_stem_wgrad: stem width %d not built", CS);
  int OH, OW, pt, pl;
  same_pad(H, 3, 2, OH, pt);
  same_pad(W, 3, 2, OW, pl);
  const long long total = (long long)N * OH * OW;
  const long long chunks = (total + 63) / 64;
  int cpb = (int)((chunks + 4 * trt_num_sms() - 1) / (4 * trt_num_sms()));
  if (cpb < 1) cpb = 1;
  const int grid = (int)((chunks + cpb - 1) / cpb);
  const __nv_bfloat16* d = (const __nv_bfloat16*)ds;
  if (CS == 32) {
    if (x_is_bf16) stem_wgrad_kernel<32, __nv_bfloat16><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)x, d, dw, N, H, W, OH, OW, pt, pl, cpb);
    else stem_wgrad_kernel<32, float><<<grid, 128, 0, stream>>>((const float*)x, d, dw, N, H, W, OH, OW, pt, pl, cpb);
  } else {
    if (x_is_bf16) stem_wgrad_kernel<48, __nv_bfloat16><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)x, d, dw, N, H, W, OH, OW, pt, pl, cpb);
    else stem_wgrad_kernel<48, float><<<grid, 128, 0, stream>>>((const float*)x, d, dw, N, H, W, OH, OW, pt, pl, cpb);
  }
  return trt_check_launch("trt_stem_wgrad");
}
