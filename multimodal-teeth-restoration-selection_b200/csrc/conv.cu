// Spatial convolutions of the EfficientNet encoder that are NOT dense contractions: the 3x3 stride-2 stem (K = 27) and
// the depthwise k3/k5 stride-1/2 convolutions, forward and backward, TF-'same' padding, NHWC bf16.
// Memory-bound: input tiles (with halo) are staged once in shared memory with the producer's BatchNorm + SiLU applied on
// load, outputs carry the BatchNorm batch statistics (train) or the folded BN + SiLU + squeeze-excite pooling (eval) in the
// epilogue, so no activation tensor is materialised between a conv and its norm.
//
// Replaces cuDNN/ATen depthwise + stem conv launches inside `self.backbone(x_img)`
// (experiments/multimodal_v1/train_mm_joint_dualtask.py:154) and their autograd backward (:248).
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int CL = 8;     // channel lanes per block (8 lanes x 8 channels = 64 channels)
constexpr int TOH = 8;    // output tile height

struct DwGeom {
  int N, H, W, C, OH, OW, S, pad_t, pad_l, tiles_x, tiles_y;
};

__device__ __forceinline__ uint4 zero4() { return make_uint4(0, 0, 0, 0); }

// per-block reduction of per-thread channel partials over the 32 pixel-threads that share a channel lane;
// s_red must hold 32*64*NV floats; result for channel (lane*8+i) is returned to threads with pt == 0.
template <int NV>
__device__ __forceinline__ void reduce_over_pt(float (&acc)[NV][8], float* s_red, int lane, int pt) {
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[(k * 32 + pt) * 64 + lane * 8 + i] = acc[k][i];
  __syncthreads();
  if (pt == 0) {
    for (int q = 1; q < 32; ++q)
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] += s_red[(k * 32 + q) * 64 + lane * 8 + i];
  }
}

// ------------------------------------------------------------------------------------------------ depthwise forward
// in_rec : producer BN record (scale, shift, mean, rstd)[C] -> input = silu(x*scale+shift); null = input used as is
// out_rec: eval-mode folded BN of THIS conv's norm -> out = silu(acc*scale+shift), pooled[n,c] += sum(out); null = raw out
// stats  : train-mode: stats[0][c] += sum(out), stats[1][c] += sum(out^2) over the bf16-rounded raw outputs
template <int K, int S>
__global__ void __launch_bounds__(TPB) dwconv_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ in_rec,
                                                         const float* __restrict__ w, uint4* __restrict__ out,
                                                         const float* __restrict__ out_rec, float* __restrict__ pooled,
                                                         double* __restrict__ stats, const DwGeom g) {
  constexpr int TOW = S == 1 ? 16 : 8;
  constexpr int P = TOW / 4;
  constexpr int IH = (TOH - 1) * S + K, IW = (TOW - 1) * S + K;
  constexpr int ROWV = (P - 1) * S + K;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* s_in = reinterpret_cast<uint4*>(smem);                               // [IH][IW][CL]
  float* s_w = reinterpret_cast<float*>(smem + (size_t)IH * IW * CL * 16);    // [K*K][64]
  float* s_red = reinterpret_cast<float*>(smem);                              // reused after compute

  const int V = g.C / 8;
  const int tile = blockIdx.x, cb = blockIdx.y, n = blockIdx.z;
  const int oy0 = (tile / g.tiles_x) * TOH, ox0 = (tile % g.tiles_x) * TOW;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;

  for (int i = threadIdx.x; i < K * K * 64; i += TPB) {
    const int tap = i / 64, c = cb * 64 + (i % 64);
    s_w[i] = c < g.C ? __ldg(w + (size_t)c * K * K + tap) : 0.f;
  }
  {
    f8 sc, sh;
    if (in_rec && cvalid) { sc = ldf8(in_rec + 8 * cv); sh = ldf8(in_rec + g.C + 8 * cv); }
    const int gy0 = oy0 * S - g.pad_t, gx0 = ox0 * S - g.pad_l;
    for (int i = pt; i < IH * IW; i += TPB / CL) {
      const int iy = i / IW, ix = i - iy * IW;
      const int gy = gy0 + iy, gx = gx0 + ix;
      uint4 v = zero4();
      if (cvalid && gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) {
        v = __ldg(x + ((size_t)(n * g.H + gy) * g.W + gx) * V + cv);
        if (in_rec) {
          f8 a = unpack8(v);
#pragma unroll
          for (int k = 0; k < 8; ++k) a.v[k] = siluf_(fmaf(a.v[k], sc.v[k], sh.v[k]));
          v = pack8(a);
        }
      }
      s_in[i * CL + lane] = v;
    }
  }
  __syncthreads();

  const int oy = pt / 4, oxb = (pt % 4) * P;
  float acc[P][8];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
#pragma unroll
  for (int kh = 0; kh < K; ++kh) {
    f8 row[ROWV];
#pragma unroll
    for (int j = 0; j < ROWV; ++j) row[j] = unpack8(s_in[((oy * S + kh) * IW + oxb * S + j) * CL + lane]);
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (kh * K + kw) * 64 + lane * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (kh * K + kw) * 64 + lane * 8 + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(row[p * S + kw].v[i], wv[i], acc[p][i]);
    }
  }

  float red[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
  f8 osc, osh;
  if (out_rec && cvalid) { osc = ldf8(out_rec + 8 * cv); osh = ldf8(out_rec + g.C + 8 * cv); }
  const int gy = oy0 + oy;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int gx = ox0 + oxb + p;
    if (cvalid && gy < g.OH && gx < g.OW) {
      f8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = out_rec ? siluf_(fmaf(acc[p][i], osc.v[i], osh.v[i])) : acc[p][i];
      const uint4 q = pack8(o);
      out[((size_t)(n * g.OH + gy) * g.OW + gx) * V + cv] = q;
      const f8 r = unpack8(q);
#pragma unroll
      for (int i = 0; i < 8; ++i) { red[0][i] += r.v[i]; red[1][i] = fmaf(r.v[i], r.v[i], red[1][i]); }
    }
  }
  if (stats || pooled) {
    reduce_over_pt<2>(red, s_red, lane, pt);
    if (pt == 0 && cvalid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (stats) {
          atomicAdd(stats + 8 * cv + i, (double)red[0][i]);
          atomicAdd(stats + g.C + 8 * cv + i, (double)red[1][i]);
        }
        if (pooled) atomicAdd(pooled + (size_t)n * g.C + 8 * cv + i, red[0][i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ depthwise backward (data)
// dD (upstream of this conv) = coef ? a*gy + b*y_raw + c : gy     [BN backward of the conv's own norm fused on load]
// dIn = convT(dD, w);  if x_rec: g = dIn * silu'(x_raw*scale+shift), bstats += (sum g, sum g*xhat); else g = dIn
template <int K>
__global__ void __launch_bounds__(TPB) dwconv_bwd_data_kernel(const uint4* __restrict__ gy_, const uint4* __restrict__ y_raw,
                                                              const float* __restrict__ coef, const float* __restrict__ w,
                                                              const uint4* __restrict__ x_raw, const float* __restrict__ x_rec,
                                                              uint4* __restrict__ g_out, double* __restrict__ bstats,
                                                              const DwGeom g, int DH, int DW) {
  constexpr int TIW = 16, P = 4;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* s_d = reinterpret_cast<uint4*>(smem);                              // [DH][DW][CL]
  float* s_w = reinterpret_cast<float*>(smem + (size_t)DH * DW * CL * 16);  // [K*K][64]
  float* s_red = reinterpret_cast<float*>(smem);
  const int V = g.C / 8, S = g.S;
  const int tile = blockIdx.x, cb = blockIdx.y, n = blockIdx.z;
  const int iy0 = (tile / g.tiles_x) * TOH, ix0 = (tile % g.tiles_x) * TIW;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  // first output row/col that can touch this input tile (floor division for possibly negative numerators)
  const int ny = iy0 + g.pad_t - (K - 1), nx = ix0 + g.pad_l - (K - 1);
  const int oyb = ny <= 0 ? 0 : (ny + S - 1) / S, oxb = nx <= 0 ? 0 : (nx + S - 1) / S;

  for (int i = threadIdx.x; i < K * K * 64; i += TPB) {
    const int tap = i / 64, c = cb * 64 + (i % 64);
    s_w[i] = c < g.C ? __ldg(w + (size_t)c * K * K + tap) : 0.f;
  }
  {
    f8 ca, cbv, cc;
    if (coef && cvalid) { ca = ldf8(coef + 8 * cv); cbv = ldf8(coef + g.C + 8 * cv); cc = ldf8(coef + 2 * g.C + 8 * cv); }
    for (int i = pt; i < DH * DW; i += TPB / CL) {
      const int dy = i / DW, dx = i - dy * DW;
      const int oy = oyb + dy, ox = oxb + dx;
      uint4 v = zero4();
      if (cvalid && oy < g.OH && ox < g.OW) {
        const size_t idx = ((size_t)(n * g.OH + oy) * g.OW + ox) * V + cv;
        v = __ldg(gy_ + idx);
        if (coef) {
          f8 a = unpack8(v);
          const f8 yr = unpack8(__ldg(y_raw + idx));
#pragma unroll
          for (int k = 0; k < 8; ++k) a.v[k] = fmaf(ca.v[k], a.v[k], fmaf(cbv.v[k], yr.v[k], cc.v[k]));
          v = pack8(a);
        }
      }
      s_d[i * CL + lane] = v;
    }
  }
  __syncthreads();

  const int iy = iy0 + pt / 4, ixb = ix0 + (pt % 4) * P;
  float acc[P][8];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
  for (int kh = 0; kh < K; ++kh) {
    const int ty = iy + g.pad_t - kh;
    if (ty < 0 || (ty % S) != 0) continue;
    const int oy = ty / S;
    if (oy >= g.OH) continue;
    for (int kw = 0; kw < K; ++kw) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (kh * K + kw) * 64 + lane * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (kh * K + kw) * 64 + lane * 8 + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const int tx = ixb + p + g.pad_l - kw;
        if (tx < 0 || (tx % S) != 0) continue;
        const int ox = tx / S;
        if (ox >= g.OW) continue;
        const f8 d = unpack8(s_d[((oy - oyb) * DW + (ox - oxb)) * CL + lane]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(d.v[i], wv[i], acc[p][i]);
      }
    }
  }

  float red[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][i] = red[1][i] = 0.f;
  f8 sc, sh, mu, rs;
  if (x_rec && cvalid) {
    sc = ldf8(x_rec + 8 * cv); sh = ldf8(x_rec + g.C + 8 * cv);
    mu = ldf8(x_rec + 2 * g.C + 8 * cv); rs = ldf8(x_rec + 3 * g.C + 8 * cv);
  }
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int ix = ixb + p;
    if (cvalid && iy < g.H && ix < g.W) {
      const size_t idx = ((size_t)(n * g.H + iy) * g.W + ix) * V + cv;
      f8 o;
      if (x_rec) {
        const f8 xr = unpack8(__ldg(x_raw + idx));
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = acc[p][i] * silu_gradf_(fmaf(xr.v[i], sc.v[i], sh.v[i]));
        const uint4 q = pack8(o);
        g_out[idx] = q;
        const f8 r = unpack8(q);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          red[0][i] += r.v[i];
          red[1][i] = fmaf(r.v[i], (xr.v[i] - mu.v[i]) * rs.v[i], red[1][i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = acc[p][i];
        g_out[idx] = pack8(o);
      }
    }
  }
  if (x_rec && bstats) {
    reduce_over_pt<2>(red, s_red, lane, pt);
    if (pt == 0 && cvalid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(bstats + 8 * cv + i, (double)red[0][i]);
        atomicAdd(bstats + g.C + 8 * cv + i, (double)red[1][i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ depthwise backward (weight)
// dW[c][kh][kw] += sum_{n,oy,ox} dD[n,oy,ox,c] * act(x)[n, oy*S-pad_t+kh, ox*S-pad_l+kw, c]
template <int K>
__global__ void __launch_bounds__(TPB) dwconv_bwd_weight_kernel(const uint4* __restrict__ gy_, const uint4* __restrict__ y_raw,
                                                                const float* __restrict__ coef, const uint4* __restrict__ x,
                                                                const float* __restrict__ in_rec, float* __restrict__ dw,
                                                                const DwGeom g, int IH, int IW) {
  constexpr int TOW = 8, NTAP = K * K, GROUPS = 32 / NTAP;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* s_in = reinterpret_cast<uint4*>(smem);                       // [IH][IW][CL]
  uint4* s_d = s_in + (size_t)IH * IW * CL;                           // [TOH][TOW][CL]
  const int V = g.C / 8, S = g.S;
  const int tile = blockIdx.x, cb = blockIdx.y;
  const int oy0 = (tile / g.tiles_x) * TOH, ox0 = (tile % g.tiles_x) * TOW;
  const int lane = threadIdx.x % CL, pt = threadIdx.x / CL;
  const int cv = cb * CL + lane;
  const bool cvalid = cv < V;
  const int tap = pt % NTAP, grp = pt / NTAP;
  const bool worker = grp < GROUPS;
  const int kh = tap / K, kw = tap % K;
  f8 sc, sh, ca, cbv, cc;
  if (in_rec && cvalid) { sc = ldf8(in_rec + 8 * cv); sh = ldf8(in_rec + g.C + 8 * cv); }
  if (coef && cvalid) { ca = ldf8(coef + 8 * cv); cbv = ldf8(coef + g.C + 8 * cv); cc = ldf8(coef + 2 * g.C + 8 * cv); }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int gy0 = oy0 * S - g.pad_t, gx0 = ox0 * S - g.pad_l;
  for (int n = blockIdx.z; n < g.N; n += gridDim.z) {
    __syncthreads();
    for (int i = pt; i < IH * IW; i += TPB / CL) {
      const int iy = i / IW, ix = i - iy * IW;
      const int gy = gy0 + iy, gx = gx0 + ix;
      uint4 v = zero4();
      if (cvalid && gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) {
        v = __ldg(x + ((size_t)(n * g.H + gy) * g.W + gx) * V + cv);
        if (in_rec) {
          f8 a = unpack8(v);
#pragma unroll
          for (int k = 0; k < 8; ++k) a.v[k] = siluf_(fmaf(a.v[k], sc.v[k], sh.v[k]));
          v = pack8(a);
        }
      }
      s_in[i * CL + lane] = v;
    }
    for (int i = pt; i < TOH * TOW; i += TPB / CL) {
      const int oy = oy0 + i / TOW, ox = ox0 + i % TOW;
      uint4 v = zero4();
      if (cvalid && oy < g.OH && ox < g.OW) {
        const size_t idx = ((size_t)(n * g.OH + oy) * g.OW + ox) * V + cv;
        v = __ldg(gy_ + idx);
        if (coef) {
          f8 a = unpack8(v);
          const f8 yr = unpack8(__ldg(y_raw + idx));
#pragma unroll
          for (int k = 0; k < 8; ++k) a.v[k] = fmaf(ca.v[k], a.v[k], fmaf(cbv.v[k], yr.v[k], cc.v[k]));
          v = pack8(a);
        }
      }
      s_d[i * CL + lane] = v;
    }
    __syncthreads();
    if (worker) {
      for (int o = grp; o < TOH * TOW; o += GROUPS) {
        const int oy = o / TOW, ox = o % TOW;
        const f8 d = unpack8(s_d[o * CL + lane]);
        const f8 a = unpack8(s_in[((oy * S + kh) * IW + ox * S + kw) * CL + lane]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(d.v[i], a.v[i], acc[i]);
      }
    }
  }
  if (worker && cvalid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dw + (size_t)(8 * cv + i) * NTAP + tap, acc[i]);
  }
}

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
      atomicAdd(stats + threadIdx.x, (double)s);
      atomicAdd(stats + CS + threadIdx.x, (double)q);
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

inline void same_pad(int i, int k, int s, int& out, int& pad_before) {
  out = (i + s - 1) / s;
  int total = (out - 1) * s + k - i;
  if (total < 0) total = 0;
  pad_before = total / 2;
}

}  // namespace

extern "C" int trt_dwconv_fwd(const void* x, const float* in_rec, const float* w, void* out, const float* out_rec,
                              float* pooled_sum, double* stats, int N, int H, int W, int C, int k, int s,
                              cudaStream_t stream) {
  TRT_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "trt_dwconv_fwd: bad argument");
  TRT_REQUIRE((k == 3 || k == 5) && (s == 1 || s == 2), "trt_dwconv_fwd: only k in {3,5}, s in {1,2}");
  DwGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.S = s;
  same_pad(H, k, s, g.OH, g.pad_t);
  same_pad(W, k, s, g.OW, g.pad_l);
  const int tow = s == 1 ? 16 : 8;
  g.tiles_x = (g.OW + tow - 1) / tow;
  g.tiles_y = (g.OH + TOH - 1) / TOH;
  if (pooled_sum) TRT_CUDA(cudaMemsetAsync(pooled_sum, 0, (size_t)N * C * sizeof(float), stream));
  const int ih = (TOH - 1) * s + k, iw = (tow - 1) * s + k;
  size_t smem = (size_t)ih * iw * CL * 16 + (size_t)k * k * 64 * 4;
  const size_t red_bytes = 2 * 32 * 64 * 4;
  if (smem < red_bytes) smem = red_bytes;
  dim3 grid(g.tiles_x * g.tiles_y, (C / 8 + CL - 1) / CL, N);
#define LAUNCH_DW(KK, SS)                                                                                          \
  do {                                                                                                             \
    static bool attr = false;                                                                                      \
    if (!attr) { TRT_CUDA(cudaFuncSetAttribute(dwconv_fwd_kernel<KK, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; } \
    dwconv_fwd_kernel<KK, SS><<<grid, TPB, smem, stream>>>((const uint4*)x, in_rec, w, (uint4*)out, out_rec, pooled_sum, stats, g); \
  } while (0)
  if (k == 3 && s == 1) LAUNCH_DW(3, 1);
  else if (k == 3 && s == 2) LAUNCH_DW(3, 2);
  else if (k == 5 && s == 1) LAUNCH_DW(5, 1);
  else LAUNCH_DW(5, 2);
#undef LAUNCH_DW
  return trt_check_launch("trt_dwconv_fwd");
}

extern "C" int trt_dwconv_bwd(const void* gy, const void* y_raw, const float* coef, const float* w, const void* x_raw,
                              const float* x_rec, void* g_out, double* bstats, float* dw, int N, int H, int W, int C, int k,
                              int s, cudaStream_t stream) {
  TRT_REQUIRE(gy && w && x_raw && dw && N > 0 && C > 0 && C % 8 == 0, "trt_dwconv_bwd: bad argument");
  TRT_REQUIRE(!coef || y_raw, "trt_dwconv_bwd: coef needs y_raw");
  TRT_REQUIRE((k == 3 || k == 5) && (s == 1 || s == 2), "trt_dwconv_bwd: only k in {3,5}, s in {1,2}");
  DwGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.S = s;
  same_pad(H, k, s, g.OH, g.pad_t);
  same_pad(W, k, s, g.OW, g.pad_l);
  const int cblocks = (C / 8 + CL - 1) / CL;
  if (g_out) {   // data gradient (skipped for the first layer, whose input needs no gradient)
    g.tiles_x = (W + 15) / 16;
    g.tiles_y = (H + TOH - 1) / TOH;
    const int DH = (TOH + k - 2) / s + 2, DW = (16 + k - 2) / s + 2;
    size_t smem = (size_t)DH * DW * CL * 16 + (size_t)k * k * 64 * 4;
    const size_t red_bytes = 2 * 32 * 64 * 4;
    if (smem < red_bytes) smem = red_bytes;
    dim3 grid(g.tiles_x * g.tiles_y, cblocks, N);
    if (k == 3) {
      static bool attr = false;
      if (!attr) { TRT_CUDA(cudaFuncSetAttribute(dwconv_bwd_data_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
      dwconv_bwd_data_kernel<3><<<grid, TPB, smem, stream>>>((const uint4*)gy, (const uint4*)y_raw, coef, w, (const uint4*)x_raw, x_rec, (uint4*)g_out, bstats, g, DH, DW);
    } else {
      static bool attr = false;
      if (!attr) { TRT_CUDA(cudaFuncSetAttribute(dwconv_bwd_data_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
      dwconv_bwd_data_kernel<5><<<grid, TPB, smem, stream>>>((const uint4*)gy, (const uint4*)y_raw, coef, w, (const uint4*)x_raw, x_rec, (uint4*)g_out, bstats, g, DH, DW);
    }
    int rc = trt_check_launch("trt_dwconv_bwd(data)");
    if (rc) return rc;
  }
  {
    g.tiles_x = (g.OW + 7) / 8;
    g.tiles_y = (g.OH + TOH - 1) / TOH;
    const int IH = (TOH - 1) * s + k, IW = 7 * s + k;
    const size_t smem = ((size_t)IH * IW + TOH * 8) * CL * 16;
    int zsplit = (4 * trt_num_sms()) / (g.tiles_x * g.tiles_y * cblocks);
    if (zsplit < 1) zsplit = 1;
    if (zsplit > N) zsplit = N;
    dim3 grid(g.tiles_x * g.tiles_y, cblocks, zsplit);
    if (k == 3) {
      static bool attr = false;
      if (!attr) { TRT_CUDA(cudaFuncSetAttribute(dwconv_bwd_weight_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
      dwconv_bwd_weight_kernel<3><<<grid, TPB, smem, stream>>>((const uint4*)gy, (const uint4*)y_raw, coef, (const uint4*)x_raw, x_rec, dw, g, IH, IW);
    } else {
      static bool attr = false;
      if (!attr) { TRT_CUDA(cudaFuncSetAttribute(dwconv_bwd_weight_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
      dwconv_bwd_weight_kernel<5><<<grid, TPB, smem, stream>>>((const uint4*)gy, (const uint4*)y_raw, coef, (const uint4*)x_raw, x_rec, dw, g, IH, IW);
    }
  }
  return trt_check_launch("trt_dwconv_bwd(weight)");
}

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
