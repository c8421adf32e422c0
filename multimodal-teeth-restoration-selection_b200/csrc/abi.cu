// libteethrt: status/error plumbing, device init and the TMA descriptor encoder shared by all entry points.
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/teethrt.h"

static thread_local char g_err[512] = "";

int trt_set_error(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

static unsigned long long g_launches = 0;
void trt_count_launch(int n) { g_launches += (unsigned long long)n; }
extern "C" unsigned long long trt_launch_count(void) { return g_launches; }

int trt_check_launch(const char* what) {
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return trt_set_error(TRT_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return TRT_OK;
}

extern "C" const char* trt_last_error_string(void) { return g_err; }
extern "C" int trt_version(void) { return TEETHRT_VERSION; }
extern "C" int trt_stat_replicas(void) { return TRT_STAT_REPLICAS; }

// Programmatic dependent launch is a win for the inference chain (batch-1 forward: ~200 dependent kernels of a few
// microseconds; measured 6.68 -> 6.28 ms for the 5-fold x 3-TTA ensemble) and a small loss inside the train step (11.66 ->
// 11.98 ms: early-launched successors compete with the side-stream weight-gradient kernels for SM slots), so it is a mode
// the host turns on around the eval forward (trt_set_pdl) rather than a process-wide default.  TEETHRT_PDL=0 forbids it, =2 forces it everywhere (A/B switch).
static int g_pdl = 0;
bool trt_pdl_enabled() {
  static const int mode = [] { const char* e = getenv("TEETHRT_PDL"); return (e && *e) ? (*e - '0') : 1; }();    // 0 never, 1 when the host asks, 2 always (A/B)
  return mode == 2 || (mode == 1 && g_pdl);
}
extern "C" int trt_set_pdl(int on) {
  const int prev = g_pdl;
  g_pdl = on ? 1 : 0;
  return prev;
}

static int g_num_sms = 0;
int trt_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn_t g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode = (encode_fn_t)fn;
}

int trt_make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_rows, uint32_t box_cols) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return trt_set_error(TRT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
  if (((uintptr_t)base & 15) || ((ld_elems * 2) & 15))
    return trt_set_error(TRT_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  if (box_cols * 2 > 128 || box_rows > 256)
    return trt_set_error(TRT_ERR_INVALID, "TMA box %u x %u too large for 128B swizzle", box_rows, box_cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return trt_set_error(TRT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                         (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
  return TRT_OK;
}

int trt_make_tmap_nhwc(CUtensorMap* out, const void* base, int N, int H, int W, int C, int box_c, int box_w, int box_h) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return trt_set_error(TRT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
  if (((uintptr_t)base & 15) || (C % 8) != 0)
    return trt_set_error(TRT_ERR_INVALID, "TMA NHWC operand must be 16-byte aligned with C %% 8 == 0");
  if (box_c > 256 || box_w > 256 || box_h > 256 || (box_c * 2) % 16 != 0)
    return trt_set_error(TRT_ERR_INVALID, "TMA NHWC box %d x %d x %d not encodable", box_h, box_w, box_c);
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return trt_set_error(TRT_ERR_CUDA, "cuTensorMapEncodeTiled(NHWC) failed (%d) N=%d H=%d W=%d C=%d box=%dx%dx%d", (int)r, N, H, W,
                         C, box_h, box_w, box_c);
  return TRT_OK;
}

extern "C" int trt_init(int device) {
  // checks the device without making it current: the caller's current device (torch's, in the Python host) is not ours to move
  int major = 0, minor = 0;
  TRT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  TRT_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10)
    return trt_set_error(TRT_ERR_UNSUPPORTED, "libteethrt needs an sm_100a device (B200); found sm_%d%d — no fallback path exists",
                         major, minor);
  int n = 0;
  TRT_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  if (n > 0) g_num_sms = n;          // one box holds identical GPUs; grids are sized for this count
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return trt_set_error(TRT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  return TRT_OK;
}
