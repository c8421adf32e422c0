// Orientation normalisation of the input stage: the device half of src/preprocessing/normalise.py:19-57
// `deskew` — cv2.cvtColor(BGR2GRAY) -> cv2.Canny(50, 150) -> first/second moments of the edge coordinates (the host turns
// them into the PCA angle) -> cv2.warpAffine(INTER_LINEAR, BORDER_REPLICATE).  Every stage reproduces OpenCV 4.x's
// integer arithmetic bit for bit (numpy restatements in oracle/ref_preproc.py are the spec):
//   gray   = (3735 B + 19235 G + 9798 R + 2^14) >> 15
//   Sobel  = 3x3, replicated border; |dx| + |dy| magnitude; magnitudes outside the image are 0
//   NMS    = OpenCV's TG22 fixed-point sector test (shift 15), asymmetric > / >= comparisons
//   edges  = 8-connected components of the NMS survivors above `low` that contain a pixel above `high`
//   warp   = 10-bit fixed-point source coordinates (double products rounded half-to-even, no FMA contraction), 5-bit
//            sub-pixel position, 15-bit bilinear weights (32767/1 at the integer position, as OpenCV's saturated table)
// All kernels are HBM/L2-streaming over uint8; nothing here is GEMM-shaped.
#include "common.cuh"

namespace {

constexpr int CT = 32;                 // output tile edge
constexpr int GT = CT + 4;             // gray tile (halo 2)
constexpr int MT = CT + 2;             // magnitude tile (halo 1)

__global__ void __launch_bounds__(CT * 8) canny_nms_kernel(const uint8_t* __restrict__ bgr, int H, int W, int low, int high,
                                                           uint8_t* __restrict__ map) {
  __shared__ int16_t gray[GT][GT + 2];
  __shared__ int16_t sdx[MT][MT + 2], sdy[MT][MT + 2];
  __shared__ int smag[MT][MT + 1];
  const int x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
  const int tid = threadIdx.y * CT + threadIdx.x;
  for (int i = tid; i < GT * GT; i += CT * 8) {
    const int ty = i / GT, tx = i - ty * GT;
    const int y = min(max(y0 + ty - 2, 0), H - 1), x = min(max(x0 + tx - 2, 0), W - 1);     // BORDER_REPLICATE
    const uint8_t* p = bgr + ((size_t)y * W + x) * 3;
    gray[ty][tx] = (int16_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15);
  }
  __syncthreads();
  for (int i = tid; i < MT * MT; i += CT * 8) {
    const int ty = i / MT, tx = i - ty * MT;
    const int y = y0 + ty - 1, x = x0 + tx - 1;
    int dx = 0, dy = 0, m = 0;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const int a = gray[ty][tx], b = gray[ty][tx + 1], c = gray[ty][tx + 2];
      const int d = gray[ty + 1][tx], f = gray[ty + 1][tx + 2];
      const int g = gray[ty + 2][tx], h = gray[ty + 2][tx + 1], k = gray[ty + 2][tx + 2];
      dx = (c + 2 * f + k) - (a + 2 * d + g);
      dy = (g + 2 * h + k) - (a + 2 * b + c);
      m = abs(dx) + abs(dy);
    }
    sdx[ty][tx] = (int16_t)dx; sdy[ty][tx] = (int16_t)dy; smag[ty][tx] = m;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < CT; r += 8) {
    const int x = x0 + threadIdx.x, y = y0 + r;
    if (x >= W || y >= H) continue;
    const int ty = r + 1, tx = threadIdx.x + 1;
    const int m = smag[ty][tx];
    uint8_t out = 0;
    if (m > low) {
      const int xs = sdx[ty][tx], ys = sdy[ty][tx];
      const long long ax = abs(xs), ay = (long long)abs(ys) << 15;
      const long long tg22x = ax * 13573;
      bool keep;
      if (ay < tg22x) keep = m > smag[ty][tx - 1] && m >= smag[ty][tx + 1];
      else {
        const long long tg67x = tg22x + (ax << 16);
        if (ay > tg67x) keep = m > smag[ty - 1][tx] && m >= smag[ty + 1][tx];
        else {
          const int s = (xs ^ ys) < 0 ? -1 : 1;
          keep = m > smag[ty - 1][tx - s] && m > smag[ty + 1][tx + s];
        }
      }
      if (keep) out = m > high ? 2 : 1;
    }
    map[(size_t)y * W + x] = out;
  }
}

// One propagation pass: inside each tile the strong label floods through candidates until the tile is stable (shared
// memory), across tiles it advances one tile per pass; *changed is set when any pixel was promoted.
__global__ void __launch_bounds__(CT * 8) canny_hysteresis_kernel(uint8_t* __restrict__ map, int H, int W, int* __restrict__ changed) {
  __shared__ uint8_t t[MT][MT + 2];
  __shared__ int again;
  const int x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
  const int tid = threadIdx.y * CT + threadIdx.x;
  for (int i = tid; i < MT * MT; i += CT * 8) {
    const int ty = i / MT, tx = i - ty * MT;
    const int y = y0 + ty - 1, x = x0 + tx - 1;
    t[ty][tx] = (y >= 0 && y < H && x >= 0 && x < W) ? map[(size_t)y * W + x] : 0;
  }
  bool promoted_any = false;
  for (;;) {
    __syncthreads();
    if (tid == 0) again = 0;
    __syncthreads();
    bool promoted = false;
    for (int r = threadIdx.y; r < CT; r += 8) {
      const int ty = r + 1, tx = threadIdx.x + 1;
      if (t[ty][tx] == 1) {
        const bool nb = t[ty - 1][tx - 1] == 2 || t[ty - 1][tx] == 2 || t[ty - 1][tx + 1] == 2 || t[ty][tx - 1] == 2 ||
                        t[ty][tx + 1] == 2 || t[ty + 1][tx - 1] == 2 || t[ty + 1][tx] == 2 || t[ty + 1][tx + 1] == 2;
        if (nb) { t[ty][tx] = 2; promoted = true; }       // monotone 1 -> 2: a racing read sees it this sweep or the next
      }
    }
    if (promoted) { again = 1; promoted_any = true; }
    __syncthreads();
    if (!again) break;
  }
  for (int r = threadIdx.y; r < CT; r += 8) {
    const int x = x0 + threadIdx.x, y = y0 + r;
    if (x < W && y < H && t[r + 1][threadIdx.x + 1] == 2) map[(size_t)y * W + x] = 2;
  }
  if (promoted_any) *changed = 1;
}

// edges (u8, 255/0) from the label map and the six coordinate moments {N, Sy, Sx, Syy, Sxy, Sxx} of the edge pixels
__global__ void __launch_bounds__(256) canny_finish_kernel(const uint8_t* __restrict__ map, int H, int W, uint8_t* __restrict__ edges,
                                                           unsigned long long* __restrict__ mom) {
  __shared__ unsigned long long red[6][8];
  unsigned long long a[6] = {0, 0, 0, 0, 0, 0};
  const size_t n = (size_t)H * W;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const bool e = map[i] == 2;
    if (edges) edges[i] = e ? 255 : 0;
    if (e) {
      const unsigned long long y = i / W, x = i - y * W;
      a[0] += 1; a[1] += y; a[2] += x; a[3] += y * y; a[4] += x * y; a[5] += x * x;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    unsigned long long s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s) atomicAdd(mom + threadIdx.x, s);
  }
}

struct Affine { double m[6]; };

template <int CH>
__global__ void __launch_bounds__(256) warp_affine_kernel(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst,
                                                          int DH, int DW, const Affine A) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= DW) return;
  constexpr double AB_SCALE = 1024.0;
  // saturate_cast<int>(double) = round half to even; explicit _rn intrinsics keep nvcc from fusing the multiply-adds
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(A.m[0], (double)x), AB_SCALE));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(A.m[3], (double)x), AB_SCALE));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m[1], (double)y), A.m[2]), AB_SCALE)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(A.m[4], (double)y), A.m[5]), AB_SCALE)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  int sx = X >> 5, sy = Y >> 5;
  sx = max(min(sx, 32767), -32768); sy = max(min(sy, 32767), -32768);          // saturate_cast<short>
  const int fx = X & 31, fy = Y & 31;
  int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }     // OpenCV's short table saturates 32768 and repairs the sum on the far tap
  const int xa = min(max(sx, 0), W - 1), xb = min(max(sx + 1, 0), W - 1);
  const int ya = min(max(sy, 0), H - 1), yb = min(max(sy + 1, 0), H - 1);
  const uint8_t* p00 = src + ((size_t)ya * W + xa) * CH;
  const uint8_t* p01 = src + ((size_t)ya * W + xb) * CH;
  const uint8_t* p10 = src + ((size_t)yb * W + xa) * CH;
  const uint8_t* p11 = src + ((size_t)yb * W + xb) * CH;
  uint8_t* q = dst + ((size_t)y * DW + x) * CH;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int v = (p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11 + (1 << 14)) >> 15;
    q[c] = (uint8_t)min(max(v, 0), 255);
  }
}

}  // namespace

extern "C" int trt_canny_nms_bgr_u8(const uint8_t* bgr, int h, int w, int low, int high, uint8_t* map, cudaStream_t stream) {
  TRT_REQUIRE(bgr && map, "trt_canny_nms_bgr_u8: null pointer");
  TRT_REQUIRE(h > 0 && w > 0 && low >= 0 && high >= low, "trt_canny_nms_bgr_u8: bad argument h=%d w=%d low=%d high=%d", h, w, low, high);
  canny_nms_kernel<<<dim3((w + CT - 1) / CT, (h + CT - 1) / CT), dim3(CT, 8), 0, stream>>>(bgr, h, w, low, high, map);
  return trt_check_launch("trt_canny_nms_bgr_u8");
}

extern "C" int trt_canny_hysteresis_pass(uint8_t* map, int h, int w, int* changed, cudaStream_t stream) {
  TRT_REQUIRE(map && changed && h > 0 && w > 0, "trt_canny_hysteresis_pass: bad argument");
  canny_hysteresis_kernel<<<dim3((w + CT - 1) / CT, (h + CT - 1) / CT), dim3(CT, 8), 0, stream>>>(map, h, w, changed);
  return trt_check_launch("trt_canny_hysteresis_pass");
}

extern "C" int trt_canny_finish(const uint8_t* map, int h, int w, uint8_t* edges, long long* moments, cudaStream_t stream) {
  TRT_REQUIRE(map && moments && h > 0 && w > 0, "trt_canny_finish: bad argument");
  TRT_CUDA(cudaMemsetAsync(moments, 0, 6 * sizeof(long long), stream));
  const size_t n = (size_t)h * w;
  const int blocks = (int)((n + 256 * 16 - 1) / (256 * 16) < (size_t)(4 * trt_num_sms()) ? (n + 256 * 16 - 1) / (256 * 16) : 4 * trt_num_sms());
  canny_finish_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(map, h, w, edges, reinterpret_cast<unsigned long long*>(moments));
  return trt_check_launch("trt_canny_finish");
}

extern "C" int trt_warp_affine_linear_u8(const uint8_t* src, int h, int w, int channels, uint8_t* dst, int dh, int dw,
                                         const double* inverse_map_host, cudaStream_t stream) {
  TRT_REQUIRE(src && dst && inverse_map_host, "trt_warp_affine_linear_u8: null pointer");
  TRT_REQUIRE(h > 0 && w > 0 && dh > 0 && dw > 0 && h < 32768 && w < 32768, "trt_warp_affine_linear_u8: bad shape");
  TRT_REQUIRE(channels == 1 || channels == 3, "trt_warp_affine_linear_u8: %d channels not built (1 or 3)", channels);
  Affine A;
  for (int i = 0; i < 6; ++i) A.m[i] = inverse_map_host[i];
  dim3 grid((dw + 255) / 256, dh);
  if (channels == 3) warp_affine_kernel<3><<<grid, 256, 0, stream>>>(src, h, w, dst, dh, dw, A);
  else warp_affine_kernel<1><<<grid, 256, 0, stream>>>(src, h, w, dst, dh, dw, A);
  return trt_check_launch("trt_warp_affine_linear_u8");
}
