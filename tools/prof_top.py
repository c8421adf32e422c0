"""Driver for `ncu --set full` captures of the top kernels of the train step on their largest real shapes (one launch each
after warm-up): 1x1-conv GEMM forward (24->144 @112x112, BN-stats epilogue), its dgrad (144->24), wgrad, depthwise forward /
data gradient / weight gradient (192 ch k3 @56x56; 144 ch k3 s2 @112x112), the two SE / BatchNorm backward passes and the SE
squeeze on 192 ch @56x56."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, teethrt
from teethrt import ops
teethrt.init()
bf16 = torch.bfloat16
N = 64
M, K, Nn = N * 112 * 112, 24, 144
A = torch.randn(M, K, device="cuda").to(bf16); W = torch.randn(Nn, K, device="cuda").to(bf16); Wt = W.t().contiguous()
C = torch.empty(M, Nn, device="cuda", dtype=bf16); dA = torch.empty(M, K, device="cuda", dtype=bf16)
st = ops.new_stats(Nn, "cuda"); dW = torch.zeros(Nn, K, device="cuda")
def dw_case(H, Cc, k, s):
    OH = ops.same_out(H, s)
    x = (torch.randn(N, H, H, Cc, device="cuda") + 0.2).to(bf16)
    rec = torch.stack([torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1, torch.randn(Cc, device="cuda") * 0.1, torch.rand(Cc, device="cuda") + 0.5]).contiguous()
    w = torch.randn(Cc, 1, k, k, device="cuda") / k
    y = torch.empty(N, OH, OH, Cc, device="cuda", dtype=bf16); dD = torch.randn(N, OH, OH, Cc, device="cuda").to(bf16)
    return dict(x=x, rec=rec, w=w, y=y, dD=dD, g=torch.empty_like(x), st=ops.new_stats(Cc, "cuda"), bst=ops.new_stats(Cc, "cuda"), dw=torch.zeros_like(w), H=H, k=k, s=s,
                act=torch.empty_like(x) if s == 1 else None)
c1, c2 = dw_case(56, 192, 3, 1), dw_case(112, 144, 3, 2)
gate = torch.rand(N, 192, device="cuda"); dmean = torch.randn(N, 192, device="cuda")
sums = torch.zeros(5, N, 192, device="cuda"); coef = torch.randn(3, 192, device="cuda") * 0.1; pooled = torch.zeros(N, 192, device="cuda")
def run():
    ops.gemm(A, W, ops.EPI_STATS, stats=st, out=C)
    ops.gemm(C, Wt, 0, out=dA)
    ops.gemm_wgrad(C, A, dW)
    for c in (c1, c2):
        # the shipped step: the stride-1 forward also stores its activated input (act_out); data gradient and weight gradient
        # are separate launches, the stride-1 weight gradient reads the saved activation (no silu(bn(x)) recomputation)
        ops.dwconv_fwd(c["x"], c["rec"], c["w"], c["y"], N, c["H"], c["H"], c["k"], c["s"], stats=c["st"], act_out=c["act"])
        ops.dwconv_bwd(c["dD"], c["w"], c["x"], c["rec"], c["g"], c["bst"], None, N, c["H"], c["H"], c["k"], c["s"])
        if c["act"] is not None:
            ops.dwconv_bwd(c["dD"], c["w"], c["act"], None, None, None, c["dw"], N, c["H"], c["H"], c["k"], c["s"])
        # "before" row (TEETHRT_DW_SAVE_ACT=0; still the stride-2 path): the weight gradient recomputes silu(bn(x)) per tile
        ops.dwconv_bwd(c["dD"], c["w"], c["x"], c["rec"], None, None, c["dw"], N, c["H"], c["H"], c["k"], c["s"])
    # SE / BatchNorm backward of the depthwise output, merged path: five-sum pass, then the apply pass
    ops.se_bwd_reduce(c1["dD"].view(-1, 192), c1["y"].view(-1, 192), c1["rec"], sums, N, 3136, zeroed=False, full=True)
    ops.act_bwd_apply(c1["dD"].view(-1, 192), gate, dmean, 1.0 / 3136, c1["y"].view(-1, 192), c1["rec"], coef, c1["g"].view(-1, 192), N, 3136)
    ops.pool_act(c1["y"].view(-1, 192), c1["rec"], pooled, N, 3136, act=1)
for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
