"""Bring-up aid: per-phase clock64 totals of the GEMM epilogue (library built with -DTRT_GEMM_TIMING)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, teethrt
from teethrt import ops
from teethrt._lib import lib
teethrt.init()
lib.trt_debug_gemm_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 8)()
E = ops.EPI_STATS
SHAPES = [(64 * 112 * 112, 64, 144, E), (64 * 112 * 112, 64, 144, E | (1 << 16)), (64 * 112 * 112, 64, 144, E | (1 << 17)), (64 * 112 * 112, 64, 144, E | (1 << 18)),
          (64 * 112 * 112, 64, 144, E | (1 << 16) | (1 << 17)), (64 * 112 * 112, 64, 144, E | (1 << 16) | (1 << 17) | (1 << 18)), (64 * 112 * 112, 64, 144, 0)]
SHAPES_OLD = [(64 * 112 * 112, 64, 144, ops.EPI_STATS), (64 * 112 * 112, 32, 144, ops.EPI_STATS), (64 * 112 * 112, 16, 144, ops.EPI_STATS),
          (64 * 112 * 112, 24, 64, ops.EPI_STATS), (64 * 112 * 112, 24, 256, ops.EPI_STATS), (64 * 112 * 112, 128, 144, ops.EPI_STATS)]
for (M, K, N, flags) in [(64 * 56 * 56, 32, 192, E), (64 * 56 * 56, 192, 32, 0), (64 * 28 * 28, 56, 336, E), (64 * 14 * 14, 112, 672, E),
                         (64 * 14 * 14, 960, 160, E), (64 * 7 * 7, 272, 1632, E), (64 * 7 * 7, 1632, 272, E)] + 0 * [(64 * 112 * 112, 24, 144, ops.EPI_STATS), (64 * 112 * 112, 24, 24, ops.EPI_STATS), (64 * 112 * 112, 144, 24, 0),
                         (64 * 56 * 56, 32, 192, ops.EPI_STATS), (64 * 14 * 14, 160, 960, ops.EPI_STATS), (64 * 7 * 7, 1632, 272, ops.EPI_STATS)]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); st = ops.new_stats(N, "cuda")
    for _ in range(2): ops.gemm(A, W, flags, stats=st if flags else None, out=C)
    torch.cuda.synchronize(); lib.trt_debug_gemm_timing(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(A, W, flags, stats=st if flags else None, out=C); e1.record(); torch.cuda.synchronize()
    lib.trt_debug_gemm_timing(buf, 0)
    tiles = ((M + 127) // 128) * max(1, (N + 255) // 256)
    import math
    # thread 64 = group 0: it sees every third tile of its CTA (every tile when only one group is active)
    v = [x / max(1, tiles / 3) for x in buf[:8]]
    print(f"M={M} K={K} N={N} flags={flags:#x}: {e0.elapsed_time(e1)*1e3:.1f} us, group-0 per-tile cycles (if 3 groups): top {v[0]:.0f} wait_tfull {v[1]:.0f} "
          f"drain {v[2]:.0f} barrier {v[3]:.0f} store+stats {v[4]:.0f} closing barrier {v[5]:.0f}")
