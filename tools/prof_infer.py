"""One eager batch-1 MMNet forward (B4 @224 + tab) bracketed by cudaProfilerStart/Stop, for the ncu launch list of the
inference path (`ncu --metrics gpu__time_duration.sum --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, teethrt
from teethrt.modules import MMNet
teethrt.init()
torch.manual_seed(0)
m = MMNet().cuda().eval()
x, t = torch.randn(1, 3, 224, 224, device="cuda"), torch.randn(1, 9, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x, t)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x, t)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
