"""Short driver for `ncu --set full`: the largest 1x1-conv GEMM launch of the train step (blocks.1.0.conv_pw forward,
A[64*112*112, 24] x W[144, 24]^T, BN-statistics epilogue) and the largest weight-gradient launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import teethrt
from teethrt import ops
teethrt.init()
M, K, N = 64 * 112 * 112, 24, 144
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
stats = ops.new_stats(N, "cuda")
dW = torch.zeros(N, K, device="cuda")
for _ in range(6):
    ops.gemm(A, W, ops.EPI_STATS, stats=stats, out=C)
    ops.gemm_wgrad(C, A, dW)
torch.cuda.synchronize()
print("ok", float(C.float().abs().mean()))
