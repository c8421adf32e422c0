#!/usr/bin/env python
"""Fixed cost of a launch inside a CUDA graph: a tiny tcgen05 GEMM, a tiny depthwise convolution and a tiny elementwise
kernel, each alone (R back-to-back launches in one graph) and interleaved.  If a pair costs more than the sum of its parts
the kernels do not overlap each other's tail / the SM is reconfigured between them (shared-memory carve-out flips).
The carve-out / shared-memory-floor switches this probe was first run with (device-wide cudaFuncCachePreferShared / L1, the
GEMM without its 120 KB floor) changed nothing and were removed from the library; PROBE_PDL=1 turns programmatic dependent
launch on."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import teethrt
from teethrt import ops

teethrt.init()
if os.environ.get("PROBE_PDL") == "1":
    from teethrt._lib import lib as _lib
    _lib.trt_set_pdl(1)
dev = "cuda"
R = 200


def graph_time(fn, reps=R):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        ts.sort()
    return ts[len(ts) // 2]


def main():
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("TEETHRT_") or k.startswith("PROBE_")}}
    for (M, K, N) in [(128, 64, 64), (3136, 272, 1632), (49, 2688, 448), (12544, 48, 24)]:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        W = torch.randn(N, K, device=dev).to(torch.bfloat16)
        C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        out[f"gemm_{M}x{K}x{N}"] = round(graph_time(lambda: ops.gemm(A, W, 0, out=C)), 2)
    v = torch.ones(4096, device=dev)
    out["scale_f32"] = round(graph_time(lambda: ops.scale_f32(v, 1.0)), 2)
    # tiny depthwise convolution (7x7, 64 channels, one image)
    x = torch.randn(1, 7, 7, 64, device=dev).to(torch.bfloat16)
    w = torch.randn(64, 1, 3, 3, device=dev)
    y = torch.empty_like(x)
    out["dwconv"] = round(graph_time(lambda: ops.dwconv_fwd(x, None, w, y, 1, 7, 7, 3, 1)), 2)
    A = torch.randn(128, 64, device=dev).to(torch.bfloat16)
    W = torch.randn(64, 64, device=dev).to(torch.bfloat16)
    C = torch.empty(128, 64, device=dev, dtype=torch.bfloat16)

    def pair_ge():
        ops.gemm(A, W, 0, out=C)
        ops.scale_f32(v, 1.0)

    def pair_gd():
        ops.gemm(A, W, 0, out=C)
        ops.dwconv_fwd(x, None, w, y, 1, 7, 7, 3, 1)

    def pair_de():
        ops.dwconv_fwd(x, None, w, y, 1, 7, 7, 3, 1)
        ops.scale_f32(v, 1.0)

    out["pair_gemm_scale"] = round(graph_time(pair_ge, R // 2), 2)
    out["pair_gemm_dwconv"] = round(graph_time(pair_gd, R // 2), 2)
    out["pair_dwconv_scale"] = round(graph_time(pair_de, R // 2), 2)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
