#!/usr/bin/env python
"""Marginal cost of each kernel class inside the captured train step: the class's entry point is replaced by a no-op (the
numbers the step computes are then garbage - only the TIMING is read), a fresh trainer is captured and timed, and the
difference to the untouched step is what the class really costs once overlap between the main and the side stream is
accounted for.  A per-launch ncu table cannot tell that: it serialises the two streams.

    python tools/knockout.py [--steps 40] [--only se_bwd,gemm_wgrad]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import teethrt
from teethrt._lib import lib
from teethrt.modules import MMJointDualHead
from teethrt.train import DualTaskTrainer
import bench

NOOP = lambda *a: 0  # noqa: E731


def dw_bwd_filter(which, real):
    def call(*a):
        g_out, dw = a[4], a[6]
        if which == "weight":
            if g_out is None:
                return 0
            a = list(a)
            a[6] = None
            return real(*a)
        if dw is None:
            return 0
        a = list(a)
        a[4] = None
        a[5] = None
        return real(*a)
    return call


CLASSES = {
    "se_bwd_mlp": ["trt_se_bwd"],
    "se_bwd_reduce": ["trt_se_bwd_reduce"],
    "act_bwd_apply": ["trt_act_bwd_apply"],
    "se_fwd_mlp": ["trt_se_fwd"],
    "pool_act": ["trt_pool_act"],
    "gate_apply": ["trt_gate_apply"],
    "affine2": ["trt_affine2"],
    "bn_apply": ["trt_bn_apply"],
    "gemm_wgrad": ["trt_gemm_wgrad_bf16"],
    "gemm_fwd_dgrad": ["trt_gemm_bf16", "trt_gemm_bf16_bnbwd"],
    "dwconv_fwd": ["trt_dwconv_fwd"],
    "dwconv_bwd_weight": ["trt_dwconv_bwd:weight"],
    "dwconv_bwd_data": ["trt_dwconv_bwd:data"],
    "bn_finalize": ["trt_bn_finalize", "trt_bn_bwd_finalize"],
    "adamw": ["trt_adamw_step"],
}


def run(dev, steps, B=64):
    torch.manual_seed(0)
    model = MMJointDualHead('tf_efficientnet_b4_ns', tab_in=bench.TAB, tab_hidden=64, drop=0.2).to(dev)
    tr = DualTaskTrainer(model, lr=3e-4, weight_decay=1e-4, t_max=100000, alpha=1.0, beta=0.3, grad_clip=1.0, graph=True, seed=1234)
    batches = [bench.synth_batch(B, 1000 + i, device=dev) for i in range(2)]
    for i in range(5 + tr.graph_warmup):
        tr.step(*batches[i % 2])
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            tr.step(*batches[i % 2])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    del tr, model
    torch.cuda.empty_cache()
    return best


INFER_CLASSES = {
    "gemm": ["trt_gemm_bf16"],
    "dwconv_fwd": ["trt_dwconv_fwd"],
    "se_fwd_mlp": ["trt_se_fwd"],
    "gate_apply": ["trt_gate_apply"],
    "pool_act": ["trt_pool_act"],
    "stem": ["trt_stem_fwd"],
    "tab_heads": ["trt_tab_heads_fwd"],
}


def run_infer(dev, steps, batch):
    """p50 of the graph-replayed MMNet eval forward at `batch` images (microseconds -> returned in ms)."""
    from teethrt.modules import MMNet
    from teethrt.infer import _GraphedForward
    torch.manual_seed(0)
    m = MMNet().to(dev).eval()
    x, t = torch.randn(batch, 3, bench.IMG, bench.IMG, device=dev), torch.randn(batch, bench.TAB, device=dev)
    g = _GraphedForward(lambda a, b: m(a, b)[0], [x, t])
    for _ in range(30):
        g(x, t)
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps * 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g(x, t)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--infer", type=int, default=0, help="N > 0: the batch-N eval forward instead of the train step")
    args = ap.parse_args()
    teethrt.init(0)
    dev = torch.device("cuda", 0)
    global run, CLASSES
    if args.infer:
        CLASSES = INFER_CLASSES
        run = lambda dev_, steps_: run_infer(dev_, steps_, args.infer)  # noqa: E731
    names = [n for n in args.only.split(",") if n] or list(CLASSES)
    base = run(dev, args.steps)
    lines = [{"knockout": None, "ms_per_step": round(base, 3)}]
    print(json.dumps(lines[-1]), flush=True)
    for n in names:
        saved = {}
        for sym in CLASSES[n]:
            s, _, which = sym.partition(":")
            real = getattr(lib, s)
            saved[s] = real
            setattr(lib, s, dw_bwd_filter(which, real) if which else NOOP)
        try:
            ms = run(dev, args.steps)
        finally:
            for s, real in saved.items():
                setattr(lib, s, real)
        lines.append({"knockout": n, "ms_per_step": round(ms, 3), "marginal_ms": round(base - ms, 3)})
        print(json.dumps(lines[-1]), flush=True)
    again = run(dev, args.steps)
    lines.append({"knockout": None, "ms_per_step": round(again, 3), "note": "baseline again (drift check)"})
    print(json.dumps(lines[-1]), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
