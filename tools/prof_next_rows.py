"""Driver for `ncu --set full` captures of the "next"-row kernels on realistic sizes, one launch each after
warm-up: Canny NMS / hysteresis / finish and warpAffine on a 1024x1024 image, the two Pillow resample passes of
1024^2 -> 256 -> 224, temperature NLL and the 61-threshold metrics kernel on a 10,007-sample fold, the Newton logistic fit
on 20,000 x 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import teethrt
from teethrt import ops, preproc, calib
import ref_preproc as P
import ref_calib as RC
teethrt.init()
img = torch.from_numpy(P.tooth_image(1024, 1024, 3, 33.0)).cuda()
z, y, _ = RC.calib_cases()["large"]
zd, yd = torch.tensor(z).cuda(), torch.tensor(y).cuda()
probs = ops.scaled_sigmoid(zd, 1.0)
thr = torch.as_tensor(calib.SWEEP, device="cuda")
rng = np.random.RandomState(0)
yl = (rng.rand(20000) < 0.5).astype(np.float32)
X = torch.tensor(1 / (1 + np.exp(-(rng.randn(20000, 2) + (2 * yl[:, None] - 1))))).cuda()
yl = torch.tensor(yl).cuda()
M = preproc.rotation_matrix_2d((512, 512), 33.0)


def run():
    preproc.canny(img)
    preproc.warp_affine(img, M, (1024, 1024))
    preproc.resize_center_crop(img, 256, 224)
    ops.temperature_nll(zd, yd, torch.zeros(1, device="cuda"))
    ops.binary_metrics(probs, yd, thr)
    ops.logreg_fit(X, yl)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
