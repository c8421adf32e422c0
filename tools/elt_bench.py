"""Streaming / reducing channel-wise kernels on the largest tf_efficientnet_b4_ns activation shapes at batch 64."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, teethrt
from teethrt import ops
teethrt.init()
bf16 = torch.bfloat16
PK = 6527.1
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
SHAPES = [(112, 144, 1), (56, 192, 3), (28, 336, 3), (14, 672, 6), (14, 960, 6), (7, 1632, 8), (7, 2688, 1)]


def timed(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


N = 64
tot = {}
for hw, C, cnt in SHAPES:
    HW = hw * hw
    x = torch.randn(N * HW, C, device="cuda").to(bf16); dA = torch.randn(N * HW, C, device="cuda").to(bf16)
    rec = torch.stack([torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1, torch.rand(C, device="cuda") + 0.5]).contiguous()
    gate = torch.rand(N, C, device="cuda"); dmean = torch.randn(N, C, device="cuda"); coef = torch.randn(3, C, device="cuda")
    out = torch.empty_like(x); bst = ops.new_stats(C, "cuda"); pooled = torch.empty(N, C, device="cuda")
    B = x.numel() * 2
    res = {
        "act_bwd": (timed(lambda: ops.act_bwd(dA, gate, dmean, 1.0 / HW, x, rec, out, bst, N, HW, act=1)), 3 * B),
        "se_bwd_reduce": (timed(lambda: ops.se_bwd_reduce(dA, x, rec, pooled, N, HW)), 2 * B),
        "pool_act": (timed(lambda: ops.pool_act(x, rec, pooled, N, HW, act=1)), B),
        "bn_bwd_reduce": (timed(lambda: ops.bn_bwd_reduce(dA, x, rec, bst)), 2 * B),
        "gate_apply": (timed(lambda: ops.gate_apply(x, rec, gate, out, N, HW)), 2 * B),
        "affine2": (timed(lambda: ops.affine2(dA, x, coef, out)), 3 * B),
        "bn_apply": (timed(lambda: ops.bn_apply(x, rec, out, act=1)), 2 * B),
    }
    print(f"HW={hw:3d} C={C:4d} x{cnt}: " + "  ".join(f"{k} {t:6.1f}us {b / t / 1e3:5.0f}GB/s" for k, (t, b) in res.items()), flush=True)
    for k, (t, b) in res.items():
        tot[k] = tot.get(k, 0) + cnt * t
print(json.dumps({k: round(v, 1) for k, v in tot.items()}))
