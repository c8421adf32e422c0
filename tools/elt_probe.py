#!/usr/bin/env python
"""Per-layer cost of the per-image streaming kernels of an MBConv block at batch 64 (tf_efficientnet_b4 shapes), timed as
back-to-back calls inside one CUDA graph over ROTATING tensor copies (> L2 in total, so reads come from HBM as in the step).
Prints one JSON line per (op, shape) with the full-width and the narrow-slab geometry side by side."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, teethrt
from teethrt import ops
sys.path.insert(0, os.path.join(ROOT, "tools"))
from launch_probe import graph_time
teethrt.init()
N = 64
SHAPES = [(112 * 112, 48, 1), (56 * 56, 144, 1), (56 * 56, 192, 3), (28 * 28, 192, 1), (28 * 28, 336, 3), (14 * 14, 336, 1), (14 * 14, 672, 6),
          (14 * 14, 960, 5), (7 * 7, 960, 1), (7 * 7, 1632, 8), (7 * 7, 2688, 1)]      # (HW, C, blocks)
only = os.environ.get("ELT_ONLY", "")
tot = {}
for HW, C, cnt in SHAPES:
    nbytes = N * HW * C * 2
    K = max(2, min(16, (300 << 20) // nbytes + 1))
    g = torch.Generator(device="cuda").manual_seed(C)
    xs = [torch.randn(N * HW, C, device="cuda", generator=g).to(torch.bfloat16) for _ in range(K)]
    ds = [torch.randn(N * HW, C, device="cuda", generator=g).to(torch.bfloat16) for _ in range(K)]
    outs = [torch.empty_like(xs[0]) for _ in range(K)]
    rec = torch.stack([torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1, torch.rand(C, device="cuda") + 0.5])
    coef = torch.randn(3, C, device="cuda") * 0.1
    gate, dmean = torch.rand(N, C, device="cuda"), torch.randn(N, C, device="cuda")
    pooled, sums = torch.zeros(N, C, device="cuda"), torch.zeros(5, N, C, device="cuda")
    i = [0]

    def nxt():
        i[0] = (i[0] + 1) % K
        return i[0]
    cases = {
        "pool_act": lambda: ops.pool_act(xs[nxt()], rec, pooled, N, HW, act=1, zeroed=True),
        "gate_apply": lambda: ops.gate_apply(xs[nxt()], rec, gate, outs[i[0]], N, HW),
        "se_bwd_reduce": lambda: ops.se_bwd_reduce(ds[nxt()], xs[i[0]], rec, sums, N, HW, zeroed=True, full=True),
        "act_bwd_apply": lambda: ops.act_bwd_apply(ds[nxt()], gate, dmean, 1.0 / HW, xs[i[0]], rec, coef, ds[i[0]], N, HW),
        "affine2": lambda: ops.affine2(ds[nxt()], xs[i[0]], coef, outs[i[0]]),
        "bn_apply": lambda: ops.bn_apply(xs[nxt()], rec, outs[i[0]], act=0),
    }
    for name, fn in cases.items():
        if only and name not in only.split(","):
            continue
        r = {"op": name, "HW": HW, "C": C, "MB": round(nbytes / 1e6, 1)}
        for narrow in ("0", "1"):
            os.environ["TEETHRT_NARROW_SLABS"] = narrow
            r["narrow" + narrow] = round(graph_time(fn, 4 * K), 2)
        print(json.dumps(r), flush=True)
        for narrow in ("0", "1"):
            tot[(name, narrow)] = tot.get((name, narrow), 0.0) + cnt * r["narrow" + narrow]
    del xs, ds, outs
    torch.cuda.empty_cache()
print(json.dumps({"per_step_us": {f"{k[0]}/narrow{k[1]}": round(v, 1) for k, v in tot.items()}}))
