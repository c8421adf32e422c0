"""Bring-up aid: sweep the MN-major shared-memory descriptor knobs of trt_gemm_wgrad_bf16 on a small problem."""
import itertools
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import teethrt
from teethrt import ops

teethrt.init()
torch.manual_seed(0)
M, Cp, Cq = 256, 128, 64
P = torch.randn(M, Cp, device="cuda").to(torch.bfloat16)
Q = torch.randn(M, Cq, device="cuda").to(torch.bfloat16)
ref = P.float().t() @ Q.float()
for lbo, sbo, kstep in itertools.product([8192, 1024, 128, 16384], [1024, 8192, 128, 2048], [2048, 32, 256, 4096]):
    out = torch.zeros(Cp, Cq, device="cuda")
    try:
        ops.gemm_wgrad(P, Q, out, lbo=lbo, sbo=sbo, kstep=kstep)
        torch.cuda.synchronize()
        err = float((out - ref).abs().max() / ref.abs().max())
    except Exception as e:  # noqa
        err = f"ERR {e}"
    print(f"lbo={lbo} sbo={sbo} kstep={kstep} rel_err={err}", flush=True)
