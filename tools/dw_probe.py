#!/usr/bin/env python
"""Depthwise kernels (forward, data gradient, weight gradient) on every distinct tf_efficientnet_b4 layer at batch 64, each
timed as back-to-back launches in one CUDA graph over rotating tensor copies (HBM-cold).  Prints microseconds, the HBM floor of
the bytes each pass must move, and the per-step totals weighted by how often the shape occurs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, teethrt
from teethrt import ops
sys.path.insert(0, os.path.join(ROOT, "tools"))
from launch_probe import graph_time
teethrt.init()
bf16 = torch.bfloat16
N = 64
PK = 6527.1
# (H, C, k, s, count in B4)
SHAPES = [(112, 48, 3, 1, 1), (112, 24, 3, 1, 1), (112, 144, 3, 2, 1), (56, 192, 3, 1, 3), (56, 192, 5, 2, 1), (28, 336, 5, 1, 3),
          (28, 336, 3, 2, 1), (14, 672, 3, 1, 5), (14, 672, 5, 1, 1), (14, 960, 5, 1, 5), (14, 960, 5, 2, 1), (7, 1632, 5, 1, 7),
          (7, 1632, 3, 1, 1), (7, 2688, 3, 1, 1)]
tot = {"fwd": 0.0, "data": 0.0, "weight": 0.0, "fwd_floor": 0.0, "data_floor": 0.0, "weight_floor": 0.0}
for H, C, k, s, cnt in SHAPES:
    OH = ops.same_out(H, s)
    in_b, out_b = N * H * H * C * 2, N * OH * OH * C * 2
    R = max(2, min(8, (300 << 20) // (in_b + out_b) + 1))
    xs = [(torch.randn(N, H, H, C, device="cuda") + 0.2).to(bf16) for _ in range(R)]
    ys = [torch.empty(N, OH, OH, C, device="cuda", dtype=bf16) for _ in range(R)]
    gys = [torch.randn(N, OH, OH, C, device="cuda").to(bf16) for _ in range(R)]
    gxs = [torch.empty(N, H, H, C, device="cuda", dtype=bf16) for _ in range(R)]
    rec = torch.stack([torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1,
                       torch.rand(C, device="cuda") + 0.5]).contiguous()
    w = torch.randn(C, 1, k, k, device="cuda") / k
    stats, bst, dw = ops.new_stats(C, "cuda"), ops.new_stats(C, "cuda"), torch.zeros_like(w)
    i = [0]

    def nxt():
        i[0] = (i[0] + 1) % R
        return i[0]
    t_f = graph_time(lambda: ops.dwconv_fwd(xs[nxt()], rec, w, ys[i[0]], N, H, H, k, s, stats=stats), 4 * R)
    t_d = graph_time(lambda: ops.dwconv_bwd(gys[nxt()], w, xs[i[0]], rec, gxs[i[0]], bst, None, N, H, H, k, s), 4 * R)
    t_w = graph_time(lambda: ops.dwconv_bwd(gys[nxt()], w, xs[i[0]], rec, None, None, dw, N, H, H, k, s), 4 * R)
    fl_f, fl_d, fl_w = (in_b + out_b) / PK / 1e3, (2 * in_b + out_b) / PK / 1e3, (in_b + out_b) / PK / 1e3
    print(json.dumps({"H": H, "C": C, "k": k, "s": s, "count": cnt, "fwd_us": round(t_f, 1), "fwd_floor": round(fl_f, 1),
                      "data_us": round(t_d, 1), "data_floor": round(fl_d, 1), "weight_us": round(t_w, 1), "weight_floor": round(fl_w, 1)}), flush=True)
    for kname, v in (("fwd", t_f), ("data", t_d), ("weight", t_w), ("fwd_floor", fl_f), ("data_floor", fl_d), ("weight_floor", fl_w)):
        tot[kname] += cnt * v
    del xs, ys, gys, gxs
    torch.cuda.empty_cache()
print(json.dumps({k_: round(v, 1) for k_, v in tot.items()}))
