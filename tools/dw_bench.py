"""Depthwise-conv kernels on every distinct tf_efficientnet_b4_ns layer shape at batch 64 (L2 flushed between launches).
Prints us, bytes of the tensors touched and the GB/s / fraction of the measured HBM peak per shape and pass, plus the sum
weighted by how often each shape occurs in the network.  `--only i` restricts to one shape (for ncu --set full)."""
import argparse, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import teethrt
from teethrt import ops
teethrt.init()
bf16 = torch.bfloat16
ap = argparse.ArgumentParser()
ap.add_argument("--only", type=int, default=-1)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--noflush", action="store_true")
args = ap.parse_args()
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6527.1
# (H, C, k, s, count in B4)
SHAPES = [(112, 48, 3, 1, 1), (112, 24, 3, 1, 1), (112, 144, 3, 2, 1), (56, 192, 3, 1, 3), (56, 192, 5, 2, 1), (28, 336, 5, 1, 3),
          (28, 336, 3, 2, 1), (14, 672, 3, 1, 5), (14, 672, 5, 1, 1), (14, 960, 5, 1, 5), (14, 960, 5, 2, 1), (7, 1632, 5, 1, 7),
          (7, 1632, 3, 1, 1), (7, 2688, 3, 1, 1)]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timed(fn):
    fn()
    ts = []
    for _ in range(args.reps):
        if not args.noflush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


N = args.batch
tot = {"fwd": 0.0, "bwd": 0.0, "fwd_ideal": 0.0, "bwd_ideal": 0.0}
for i, (H, C, k, s, cnt) in enumerate(SHAPES):
    if args.only >= 0 and i != args.only:
        continue
    torch.manual_seed(i)
    OH = ops.same_out(H, s)
    x = (torch.randn(N, H, H, C, device="cuda") + 0.2).to(bf16)
    rec = torch.stack([torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1,
                       torch.rand(C, device="cuda") + 0.5]).contiguous()
    w = torch.randn(C, 1, k, k, device="cuda") / k
    y = torch.empty(N, OH, OH, C, device="cuda", dtype=bf16)
    stats = ops.new_stats(C, "cuda")
    gy = torch.randn(N, OH, OH, C, device="cuda").to(bf16)
    coef = torch.stack([torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.05, torch.randn(C, device="cuda") * 0.05]).contiguous()
    g_out = torch.empty_like(x)
    bst = ops.new_stats(C, "cuda")
    dw = torch.zeros_like(w)
    tf = timed(lambda: ops.dwconv_fwd(x, rec, w, y, N, H, H, k, s, stats=stats))
    dD = torch.empty_like(gy)

    def bwd():
        ops.affine2(gy, y, coef, dD)
        ops.dwconv_bwd(dD, w, x, rec, g_out, bst, dw, N, H, H, k, s)
    tb = timed(bwd)
    bf = (x.numel() + y.numel()) * 2
    bb = (2 * y.numel() + 2 * x.numel()) * 2          # gy, y_raw, x_raw read once, g_out written (layer-fused lower bound)
    print(f"[{i:2d}] H={H:3d} C={C:4d} k{k} s{s} x{cnt}: fwd {tf:7.1f} us {bf / tf / 1e3:6.0f} GB/s ({100 * bf / tf / 1e3 / PK:4.1f}%)   "
          f"bwd {tb:7.1f} us {bb / tb / 1e3:6.0f} GB/s ({100 * bb / tb / 1e3 / PK:4.1f}%)", flush=True)
    tot["fwd"] += cnt * tf; tot["bwd"] += cnt * tb
    tot["fwd_ideal"] += cnt * bf / PK / 1e3; tot["bwd_ideal"] += cnt * bb / PK / 1e3
print(json.dumps({k: round(v, 1) for k, v in tot.items()}))
