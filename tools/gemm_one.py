import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, teethrt
from teethrt import ops
teethrt.init()
hw, ci, co = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
M = 64 * hw * hw
A = torch.randn(M, ci, device="cuda").to(torch.bfloat16); W = torch.randn(co, ci, device="cuda").to(torch.bfloat16); Wt = W.t().contiguous()
C = torch.empty(M, co, device="cuda", dtype=torch.bfloat16); dA = torch.empty(M, ci, device="cuda", dtype=torch.bfloat16)
st = ops.new_stats(co, "cuda")
print("fwd...", flush=True); ops.gemm(A, W, ops.EPI_STATS, stats=st, out=C); torch.cuda.synchronize(); print("fwd ok", flush=True)
ref = (A.float() @ W.float().t())
print("fwd err", float((C.float() - ref).abs().max() / ref.abs().max()), flush=True)
print("dgrad...", flush=True); ops.gemm(C, Wt, 0, out=dA); torch.cuda.synchronize(); print("dgrad ok", flush=True)
dW = torch.zeros(co, ci, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for name, fn in (("fwd", lambda: ops.gemm(A, W, ops.EPI_STATS, stats=st, out=C)), ("dgrad", lambda: ops.gemm(C, Wt, 0, out=dA)), ("wgrad", lambda: ops.gemm_wgrad(C, A, dW))):
    for i in range(8):
        flush.zero_(); fn(); torch.cuda.synchronize(); print(name, i, "ok", flush=True)
