"""Secondary BASELINE configs on one B200: MIL pooling (config 2), MIL train step, input stage (config 3), with the
reference's CPU path timed beside them.  Writes JSON lines (one per measurement) to stdout."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import teethrt
from teethrt import ops, preproc
from teethrt.modules import MILNet
from teethrt.train import MILTrainer
teethrt.init()
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def gpu_time(fn, n=20, warm=3, l2flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        if l2flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts)


out = []
# ---- config 2: MIL gated-attention pooling alone
D, K, hid = 1280, 16, 128
Vw, Vb = torch.randn(hid, D, device="cuda") * D ** -0.5, torch.zeros(hid, device="cuda")
Uw, Ub = torch.randn(hid, D, device="cuda") * D ** -0.5, torch.zeros(hid, device="cuda")
ww, wb = torch.randn(hid, device="cuda") * hid ** -0.5, torch.zeros(1, device="cuda")
for B in (6, 64, 1024):
    H = torch.randn(B, K, D, device="cuda")
    t = gpu_time(lambda: ops.mil_attn_fwd(H, Vw, Vb, Uw, Ub, ww, wb), l2flush=False)
    alg = B * (K * D * 4 + D * 4 + K * 4) + 2 * hid * D * 4
    out.append({"what": "mil_attn_fwd", "B": B, "K": K, "D": D, "hid": hid, "us": t * 1e6, "bags_per_s": B / t,
                "algorithmic_GBps": alg / t / 1e9, "hbm_frac": alg / t / 1e9 / PK["hbm_gbs"], "gflops": B * 10.5e6 / t / 1e9})
# ---- MIL train step (B=6 bags x 16 instances @224, B0 encoder)
torch.manual_seed(0)
m = MILNet().cuda()
tr = MILTrainer(m, lr=2e-4, t_max=1000, graph=True)
bags = torch.randn(6, 16, 3, 224, 224, device="cuda")
y = (torch.rand(6, device="cuda") < 0.6).float()
for _ in range(6):
    tr.step(bags, y)
torch.cuda.synchronize()
t = gpu_time(lambda: tr.step(bags, y), n=20, warm=2, l2flush=False)
out.append({"what": "mil_train_step", "bags": 6, "instances": 16, "img": 224, "ms": t * 1e3, "crops_per_s": 96 / t,
            "tflops": 96 * 3 * 0.769e9 / t / 1e12, "launches_per_step": tr.launches_per_step})
# ---- config 3: input stage on 1024^2 radiographs
import ref_preproc as P
import cv2
for n in (1, 64):
    imgs = torch.from_numpy(np.stack([P.image_set("radiograph", 1024, 1024, seed=i % 4) for i in range(n)])).cuda()
    dst = torch.empty_like(imgs)
    ws = torch.empty(teethrt.lib.trt_clahe_workspace_bytes(n), device="cuda", dtype=torch.uint8)
    tab = preproc.device_tables("cuda")
    from teethrt._lib import check, ptr, stream
    fn = lambda: check(teethrt.lib.trt_clahe_bgr_u8(ptr(imgs), ptr(dst), n, 1024, 1024, 3.0, ptr(tab), ptr(ws), ws.numel(), stream()))
    t = gpu_time(fn, l2flush=(n == 1))
    px = n * 1024 * 1024
    out.append({"what": "clahe_bgr_u8", "n": n, "us": t * 1e6, "Mpx_per_s": px / t / 1e6, "algorithmic_GBps": 6 * px / t / 1e9,
                "hbm_frac_6Bpx": 6 * px / t / 1e9 / PK["hbm_gbs"], "practical_GBps_9Bpx": 9 * px / t / 1e9})
    st = preproc.InputStage(n, 1024, 1024, size=224, dtype=torch.bfloat16)
    t = gpu_time(lambda: st(imgs, 0), l2flush=(n == 1))
    out.append({"what": "input_stage_clahe_resize_normalize", "n": n, "us": t * 1e6, "images_per_s": n / t,
                "algorithmic_GBps": n * 3.45e6 / t / 1e9})
img = P.image_set("radiograph", 1024, 1024)
for thr in (1, os.cpu_count()):
    cv2.setNumThreads(thr)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); P.apply_clahe_cv2(img); ts.append(time.perf_counter() - t0)
    out.append({"what": "cpu_reference_apply_clahe", "threads": thr, "ms": statistics.median(ts) * 1e3,
                "Mpx_per_s": 1024 * 1024 / statistics.median(ts) / 1e6})
for o in out:
    print(json.dumps(o))
