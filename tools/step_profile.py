"""Per-launch roofline of ONE train step (configs[1]).

Run under   ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off
            --csv --log-file gpurun_out/step_launches.csv python tools/step_profile.py --log gpurun_out/step_ops.json
The script runs warm-up steps without graphs, then brackets one eager step with cudaProfilerStart/Stop while logging every
C-ABI call (op name, tensor shapes, bytes of the tensors it touches, kernels it launched).  tools/join_profile.py joins the
two files into a per-launch table (time, algorithmic GB/s, fraction of the measured HBM peak).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import teethrt  # noqa: E402
from teethrt import ops  # noqa: E402
from teethrt._lib import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log", default="gpurun_out/step_ops.json")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--img", type=int, default=224)
ap.add_argument("--mil", action="store_true")
args = ap.parse_args()

LOG = []
ENABLED = [False]


def wrap(name, fn):
    def w(*a, **k):
        if not ENABLED[0]:
            return fn(*a, **k)
        c0 = lib.trt_launch_count()
        r = fn(*a, **k)
        c1 = lib.trt_launch_count()
        shapes, nbytes = [], 0
        flat = list(a) + list(k.values())
        for t in list(flat):
            if isinstance(t, (list, tuple, dict)):
                flat.extend(t.values() if isinstance(t, dict) else t)
        ins = [t for t in flat if isinstance(t, torch.Tensor) and t.is_cuda]
        have = {t.data_ptr() for t in ins}
        outs = r if isinstance(r, (list, tuple)) else ([*r.values()] if isinstance(r, dict) else [r])
        outs = [t for t in outs if isinstance(t, torch.Tensor) and t.is_cuda and t.data_ptr() not in have]
        for t in ins + outs:      # a tensor passed twice (in-place) counts twice: read + write
            shapes.append(list(t.shape))
            nbytes += t.numel() * t.element_size()
        LOG.append(dict(op=name, shapes=shapes, bytes=nbytes, launches=int(c1 - c0)))
        return r
    return w


for n in dir(ops):
    f = getattr(ops, n)
    if callable(f) and getattr(f, "__module__", None) == ops.__name__ and not n.startswith("_") and n not in (
            "same_out", "same_pad", "tab_heads_scratch", "OptimState", "bn_fin", "bn_bwd_fin", "se_bn", "new_stats", "stats_total"):
        setattr(ops, n, wrap(n, f))

ops.OptimState.advance = wrap("optim_advance", ops.OptimState.advance)     # a method, not a module-level op
teethrt.init(0)
torch.manual_seed(0)
dev = torch.device("cuda", 0)
if args.mil:
    from teethrt.modules import MILNet
    from teethrt.train import MILTrainer
    model = MILNet("tf_efficientnet_b0_ns", 0.2).to(dev)
    tr = MILTrainer(model, graph=False)
    batch = [torch.randn(6, 16, 3, args.img, args.img, device=dev), (torch.rand(6, device=dev) < 0.6).float()]
else:
    from teethrt.modules import MMJointDualHead
    from teethrt.train import DualTaskTrainer
    model = MMJointDualHead("tf_efficientnet_b4_ns", tab_in=9, tab_hidden=64, drop=0.2).to(dev)
    tr = DualTaskTrainer(model, t_max=100, graph=False)
    B = args.batch
    batch = [torch.randn(B, 3, args.img, args.img, device=dev), torch.randn(B, 9, device=dev),
             (torch.rand(B, device=dev) < 0.6).float(), torch.rand(B, device=dev)]
for _ in range(3):
    tr.step(*batch)
torch.cuda.synchronize()
ENABLED[0] = True
torch.cuda.profiler.start()
tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ENABLED[0] = False
os.makedirs(os.path.dirname(args.log) or ".", exist_ok=True)
json.dump(LOG, open(args.log, "w"))
print("ops", len(LOG), "launches", sum(e["launches"] for e in LOG), "loss", float(tr.loss))
