#!/usr/bin/env python
"""SE MLP forward / backward per layer shape of tf_efficientnet_b4 at batch 64: two-launch kernels vs the one-launch (grid
barrier) kernels, each timed as 100 back-to-back calls inside one CUDA graph (no other stream active)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, teethrt
from teethrt import ops
sys.path.insert(0, os.path.join(ROOT, "tools"))
from launch_probe import graph_time
teethrt.init()
N = int(os.environ.get("SE_N", "64"))
SHAPES = [(48, 12, 2), (144, 6, 4), (192, 8, 4), (336, 14, 4), (672, 28, 6), (960, 40, 6), (1632, 68, 8), (2688, 112, 2)]   # (C, rd, blocks)
tot = {"fwd2": 0.0, "fwd1": 0.0, "bwd2": 0.0, "bwd1": 0.0}
for C, rd, cnt in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(C)
    R = lambda *s: torch.randn(*s, device="cuda", generator=g)
    pooled, Wr, br, We, be = R(N, C).abs() * 49, R(rd, C) * C ** -0.5, R(rd) * 0.1, R(C, rd) * rd ** -0.5, R(C) * 0.1
    s1, gate = torch.empty(N, rd, device="cuda"), torch.empty(N, C, device="cuda")
    ws = ops.se_workspace(N, C, rd, "cuda")
    sums, rec, gamma = R(5, N, C), torch.stack([R(C) * 0.1 + 1, R(C) * 0.1, R(C) * 0.2, R(C).abs() + 0.5]), R(C) * 0.1 + 1
    o = dict(ds2=torch.empty(N, C, device="cuda"), ds1=torch.zeros(N, rd, device="cuda"), dmean=torch.empty(N, C, device="cuda"),
             dWr=torch.empty_like(Wr), dbr=torch.empty_like(br), dWe=torch.empty_like(We), dbe=torch.empty_like(be),
             coef=torch.empty(3, C, device="cuda"), dgamma=torch.empty(C, device="cuda"), dbeta=torch.empty(C, device="cuda"))
    bn = ops.se_bn(sums, rec, gamma, o["coef"], o["dgamma"], o["dbeta"], N * 49)

    def fwd(w):
        ops.se_fwd(pooled, 1 / 49, Wr, br, We, be, s1, gate, ws=w)

    def bwd(w):
        ops.se_bwd(sums[0], gate, s1, pooled, 1 / 49, Wr, We, o["ds2"], o["ds1"], o["dmean"], o["dWr"], o["dbr"], o["dWe"], o["dbe"],
                   ds1_zeroed=False, bn=bn, ws=w)
    r = {"C": C, "rd": rd, "fwd2": graph_time(lambda: fwd(None), 100), "fwd1": graph_time(lambda: fwd(ws), 100),
         "bwd2": graph_time(lambda: bwd(None), 100), "bwd1": graph_time(lambda: bwd(ws), 100)}
    for k in tot:
        tot[k] += cnt * r[k]
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
print(json.dumps({"per_step_us": {k: round(v, 1) for k, v in tot.items()}, "N": N}))
