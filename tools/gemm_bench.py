"""1x1-conv GEMM shapes of tf_efficientnet_b4_ns at batch 64 (forward with BN-stats epilogue, dgrad, wgrad), L2 flushed.
Prints us and the fraction of two floors: HBM copy peak on all bytes, and the measured pure-write cap on the bytes written."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, teethrt
from teethrt import ops
teethrt.init()
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6527.1
WR = 3880.0      # GB/s, pure-write cap measured with tools/ubench/bw_probe.py
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
# (HW, Cin, Cout, count): expand / project pairs of every stage + head
LAYERS = [(112, 48, 24, 1), (112, 24, 24, 1), (112, 24, 144, 1), (56, 144, 32, 1), (56, 32, 192, 3), (56, 192, 32, 3), (28, 192, 56, 1),
          (28, 56, 336, 3), (28, 336, 56, 3), (14, 336, 112, 1), (14, 112, 672, 6), (14, 672, 112, 5), (14, 672, 160, 1),
          (14, 160, 960, 6), (14, 960, 160, 5), (7, 960, 272, 1), (7, 272, 1632, 8), (7, 1632, 272, 7), (7, 1632, 448, 1),
          (7, 448, 2688, 1), (7, 2688, 448, 1), (7, 448, 1792, 1)]


def timed(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


tot = [0.0, 0.0, 0.0, 0.0]
for hw, ci, co, cnt in LAYERS:
    M = 64 * hw * hw
    A = torch.randn(M, ci, device="cuda").to(torch.bfloat16)
    W = torch.randn(co, ci, device="cuda").to(torch.bfloat16)
    Wt = W.t().contiguous()
    C = torch.empty(M, co, device="cuda", dtype=torch.bfloat16)
    dA = torch.empty(M, ci, device="cuda", dtype=torch.bfloat16)
    st = ops.new_stats(co, "cuda")
    dW = torch.zeros(co, ci, device="cuda")
    tf = timed(lambda: ops.gemm(A, W, ops.EPI_STATS, stats=st, out=C))
    td = timed(lambda: ops.gemm(C, Wt, 0, out=dA))
    tw = timed(lambda: ops.gemm_wgrad(C, A, dW))
    bytes_all = (M * ci + M * co) * 2
    floor_f = max(bytes_all / PK, M * co * 2 / WR) / 1e3
    floor_d = max(bytes_all / PK, M * ci * 2 / WR) / 1e3
    floor_w = bytes_all / PK / 1e3
    print(f"HW={hw:3d} {ci:4d}->{co:4d} x{cnt}: fwd {tf:6.1f} us ({100 * floor_f / tf:3.0f}% of floor {floor_f:5.1f})  dgrad {td:6.1f} ({100 * floor_d / td:3.0f}%)  "
          f"wgrad {tw:6.1f} ({100 * floor_w / tw:3.0f}%)", flush=True)
    tot[0] += cnt * tf; tot[1] += cnt * td; tot[2] += cnt * tw; tot[3] += cnt * (floor_f + floor_d + floor_w)
print(json.dumps({"fwd_us": round(tot[0], 1), "dgrad_us": round(tot[1], 1), "wgrad_us": round(tot[2], 1), "floor_us": round(tot[3], 1)}))
