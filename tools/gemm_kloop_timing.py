"""Bring-up aid (library built with -DTRT_GEMM_TIMING): where CTA 0's producer and MMA threads spend their cycles per k-block.
slots: 8 producer waits for a free stage, 9 producer issues TMA (+ loop), 10 MMA thread waits for data, 11 MMA thread issues."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, teethrt
from teethrt import ops
from teethrt._lib import lib
teethrt.init()
raw = lib._cdll
raw.trt_debug_gemm_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
# 1 << 21: one MMA per k-block instead of four (timing ablation, wrong result)
CASES = [(49, 1632, 272, 0), (49, 1632, 272, 1 << 21), (3136, 272, 1632, 0), (3136, 272, 1632, 1 << 21), (12544, 960, 160, 0), (12544, 960, 160, 1 << 21)]
for (M, K, N, dbg) in CASES:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(A, W, dbg, out=C)
    torch.cuda.synchronize(); raw.trt_debug_gemm_timing(None, 1)
    R = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(R): ops.gemm(A, W, dbg, out=C)
    e1.record(); torch.cuda.synchronize()
    raw.trt_debug_gemm_timing(buf, 0)
    kb = (K + 63) // 64
    v = [buf[i] / R for i in range(16)]
    print(f"M={M} K={K} N={N} kb={kb} dbg={dbg:#x} {e0.elapsed_time(e1) * 1e3 / R:.1f} us/launch (eager): CTA0 totals per launch (cycles): producer wait_empty {v[8]:.0f} issue {v[9]:.0f} | mma wait_full {v[10]:.0f} fence {v[12]:.0f} mma {v[13]:.0f} commit {v[14]:.0f} loop {v[11]:.0f}", flush=True)
