"""Mint tests/golden/calib_golden.json from the REFERENCE's own code (run in the build container only).

    python tests/golden/make_calib_golden.py

For every seeded validation fold of oracle.ref_calib.calib_cases() this drives the reference's TemperatureScaler and
compute_metrics (imported unchanged from experiments/multimodal_v1/train_mm_joint_dualtask.py on top of the oracle's timm
shim) through the same statements run_fold executes at :271-296 — that block is inline in run_fold, so it is the only part
spelled out here — and stores T, the chosen threshold, the metrics at it and at 0.5, and a checksum of the probabilities.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

import ref_calib as RC  # noqa: E402
from conftest import load_reference_module  # noqa: E402


def reference_epoch_block(ref, va_logits, va_y):
    scaler_T = ref.TemperatureScaler()
    opt_T = torch.optim.LBFGS(scaler_T.parameters(), lr=0.1, max_iter=50)
    logits_tensor = torch.tensor(va_logits, dtype=torch.float32)
    targets_tensor = torch.tensor(va_y, dtype=torch.float32)

    def _closure():
        opt_T.zero_grad()
        loss_T = F.binary_cross_entropy_with_logits(scaler_T(logits_tensor), targets_tensor)
        loss_T.backward()
        return loss_T
    try:
        opt_T.step(_closure)
    except Exception:
        pass
    with torch.no_grad():
        adj = scaler_T(logits_tensor).cpu().numpy()
        va_probs = 1 / (1 + np.exp(-adj))
    best_thr, best_f1 = 0.5, -1.0
    for t in np.linspace(0.2, 0.8, 61):
        m = ref.compute_metrics(va_y, va_probs, thr=t)
        if m['f1'] > best_f1:
            best_f1 = m['f1']; best_thr = float(t)
    return scaler_T.temperature(), va_probs, best_thr, ref.compute_metrics(va_y, va_probs, thr=best_thr)


def main():
    ref = load_reference_module("experiments/multimodal_v1/train_mm_joint_dualtask.py", "ref_mm")
    out = {"torch": torch.__version__, "cases": {}}
    for name, (z, y, _names) in RC.calib_cases().items():
        T, probs, thr, m = reference_epoch_block(ref, z, y)
        out["cases"][name] = dict(n=int(len(z)), T=T, thr=thr, metrics=m, metrics_at_half=ref.compute_metrics(y, probs, thr=0.5),
                                  prob_sum=float(np.sum(probs, dtype=np.float64)), prob_first=[float(p) for p in probs[:4]],
                                  f1_curve=[ref.compute_metrics(y, probs, thr=t)['f1'] for t in np.linspace(0.2, 0.8, 61)])
        print(name, out["cases"][name]["T"], thr, m)
    with open(os.path.join(HERE, "calib_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
