"""Driver for the `ncu --set full` capture of the kernels round 1 left without ncu evidence (VERDICT r01, row d): CLAHE
(histogram / LUT / apply on 64 x 1024^2 radiograph-like images), MIL gated-attention pooling forward + backward (6 and 1024
bags x 16 x 1280), the tab MLP + heads + dual BCE forward / backward at batch 64, and the optimiser (grad norm + AdamW over
the 17.56 M flat parameters).  One launch of each after warm-up, bracketed by cudaProfilerStart/Stop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import teethrt
from teethrt import ops, preproc
import ref_preproc as P
teethrt.init()
dev = "cuda"
imgs = torch.from_numpy(np.stack([P.image_set("radiograph", 1024, 1024, seed=i % 4) for i in range(64)])).cuda()
D, K, hid = 1280, 16, 128
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s, scale=1.0: torch.randn(*s, device=dev, generator=g) * scale
Vw, Vb, Uw, Ub, ww, wb = R(hid, D, scale=D ** -0.5), R(hid), R(hid, D, scale=D ** -0.5), R(hid), R(hid, scale=hid ** -0.5), R(1)
Hs = {B: R(B, K, D) for B in (6, 1024)}
B, F_, T, Hd = 64, 1792, 9, 64
feat, xtab = R(B, F_), R(B, T)
params = [R(Hd, T, scale=0.3), R(Hd), R(Hd) * 0.1 + 1, R(Hd) * 0.1, R(Hd, Hd, scale=0.12), R(Hd), R(F_ + Hd, scale=0.02), R(1), R(F_ + Hd, scale=0.02), R(1)]
rm, rv, nbt = torch.zeros(Hd, device=dev), torch.ones(Hd, device=dev), torch.zeros((), device=dev, dtype=torch.int64)
yh, ys = (torch.rand(B, device=dev) < 0.6).float(), torch.rand(B, device=dev)
scratch = ops.tab_heads_scratch(B, Hd, dev)
grads = [torch.empty_like(p) for p in params]
dfeat = torch.empty_like(feat)
n = 17_557_344
p_, g_, m_, v_ = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
v_.abs_()
st = ops.OptimState(torch.device(dev), 3e-4, t_max=1000)
nsq, gn = torch.zeros(1, device=dev, dtype=torch.float64), torch.zeros(1, device=dev)


def run():
    preproc.apply_clahe(imgs)
    for Bb, H in Hs.items():
        M, A, gV, gU = ops.mil_attn_fwd(H, Vw, Vb, Uw, Ub, ww, wb, save=True)
        gs = [torch.zeros_like(t) for t in (Vw, Vb, Uw, Ub, ww, wb)]
        ops.mil_attn_bwd(torch.ones_like(M), H, A, gV, gU, Vw, Uw, ww, *gs)
    out = ops.tab_heads_fwd(feat, xtab, params, rm, rv, nbt, scratch, True, 0.2, targets=(yh, ys, None), seed=1)
    ops.tab_heads_bwd(feat, xtab, params, rm, rv, out["dlogit"], out["dreg"], dfeat, grads, scratch, True, 0.2, seed=1)
    st.advance()
    ops.grad_sumsq(g_, nsq)
    ops.adamw_step(p_, g_, m_, v_, st, nsq, gn, 1.0, 1.0)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
