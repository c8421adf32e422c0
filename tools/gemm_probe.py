#!/usr/bin/env python
"""Tile-width sweep of the forward / data-gradient 1x1-conv GEMMs on the tf_efficientnet_b4 shapes at batch 64, each timed as
back-to-back launches inside a CUDA graph over rotating operand copies (no host gaps, reads from HBM).  One JSON line per
(layer, direction): microseconds for the library's own choice (block_n 0) and for every legal override."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, teethrt
from teethrt import ops
sys.path.insert(0, os.path.join(ROOT, "tools"))
from launch_probe import graph_time
teethrt.init()
# (HW, Cin, Cout, count): expand / project pairs of every stage + head
LAYERS = [(112, 48, 24, 1), (112, 24, 24, 1), (112, 24, 144, 1), (56, 144, 32, 1), (56, 32, 192, 3), (56, 192, 32, 3), (28, 192, 56, 1),
          (28, 56, 336, 3), (28, 336, 56, 3), (14, 336, 112, 1), (14, 112, 672, 6), (14, 672, 112, 5), (14, 672, 160, 1),
          (14, 160, 960, 6), (14, 960, 160, 5), (7, 960, 272, 1), (7, 272, 1632, 8), (7, 1632, 272, 7), (7, 1632, 448, 1),
          (7, 448, 2688, 1), (7, 2688, 448, 1), (7, 448, 1792, 1)]
B = int(os.environ.get("GEMM_B", "64"))
tot = {"fwd_auto": 0.0, "fwd_best": 0.0, "dgrad_auto": 0.0, "dgrad_best": 0.0}


def cands(N):
    c = [0]
    if N <= 256:
        c.append((N + 15) // 16 * 16)
    c += [b for b in (64, 128, 192, 256) if b < N]
    return c


def sweep(M, K, N, stats):
    nbytes = (M * K + M * N) * 2
    R = max(2, min(12, (300 << 20) // nbytes + 1))
    As = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(R)]
    Cs = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(R)]
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    st = ops.new_stats(N, "cuda") if stats else None
    i = [0]
    res = {}
    for bn in cands(N):
        def fn():
            i[0] = (i[0] + 1) % R
            ops.gemm(As[i[0]], W, ops.EPI_STATS if stats else 0, stats=st, out=Cs[i[0]], block_n=bn)
        try:
            res[str(bn)] = round(graph_time(fn, 4 * R), 2)
        except Exception as e:  # an override the kernel rejects
            res[str(bn)] = None
    return res


for hw, ci, co, cnt in LAYERS:
    M = B * hw * hw
    f = sweep(M, ci, co, True)
    d = sweep(M, co, ci, False)
    bf = min(v for v in f.values() if v)
    bd = min(v for v in d.values() if v)
    tot["fwd_auto"] += cnt * f["0"]; tot["fwd_best"] += cnt * bf
    tot["dgrad_auto"] += cnt * d["0"]; tot["dgrad_best"] += cnt * bd
    print(json.dumps({"hw": hw, "cin": ci, "cout": co, "count": cnt, "fwd": f, "dgrad": d}), flush=True)
print(json.dumps({k: round(v, 1) for k, v in tot.items()}))
