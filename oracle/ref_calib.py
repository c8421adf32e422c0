"""ORACLE (test infrastructure — only tests/, __graft_entry__.smoke() and bench.py's CPU arms import it).

CPU restatement of the reference's per-fold calibration and result files, each function citing the
reference file:line it follows.  It uses the same libraries the reference calls (torch.optim.LBFGS, sklearn.metrics,
numpy, pandas), so outputs are comparable value for value.  Parity status: the reference's tests hold no vector for this
path; the restatement is pinned in this container against the functions imported from the reference
(tests/test_oracle_calib.py) and frozen as tests/golden/calib_golden.json by tests/golden/make_golden.py.
"""
import json
from pathlib import Path

import numpy as np
import pandas as pd
import torch
import torch.nn.functional as F
from sklearn.metrics import accuracy_score, precision_recall_fscore_support, roc_auc_score


def fast_round(x, n=4):
    """train_mm_joint_dualtask.py:69-70"""
    return float(np.round(x, n))


def compute_metrics(y_true, y_prob, thr=0.5):
    """train_mm_joint_dualtask.py:181-186"""
    y_pred = (y_prob >= thr).astype(int)
    auc = roc_auc_score(y_true, y_prob) if len(np.unique(y_true)) > 1 else float('nan')
    acc = accuracy_score(y_true, y_pred)
    prec, rec, f1, _ = precision_recall_fscore_support(y_true, y_pred, average='binary', zero_division=0)
    return {'auc': fast_round(auc), 'acc': fast_round(acc), 'prec': fast_round(prec), 'rec': fast_round(rec),
            'f1': fast_round(f1)}


def fit_temperature(va_logits, va_y):
    """train_mm_joint_dualtask.py:162-174 (TemperatureScaler) + :271-287 (LBFGS lr 0.1, 50 iterations, log_T from 0).
    -> (T, calibrated probabilities computed the way :285-287 does, in numpy)."""
    log_T = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.LBFGS([log_T], lr=0.1, max_iter=50)
    logits = torch.tensor(va_logits, dtype=torch.float32)
    targets = torch.tensor(va_y, dtype=torch.float32)

    def closure():
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(logits / log_T.exp(), targets)
        loss.backward()
        return loss
    try:
        opt.step(closure)
    except Exception:
        pass
    with torch.no_grad():
        adj = (logits / log_T.exp()).numpy()
        probs = 1 / (1 + np.exp(-adj))
        return float(log_T.exp().item()), probs


def best_threshold(va_y, va_probs):
    """train_mm_joint_dualtask.py:289-295: first maximum of the ROUNDED F1 over 61 thresholds in [0.2, 0.8]."""
    best_thr, best_f1 = 0.5, -1.0
    for t in np.linspace(0.2, 0.8, 61):
        m = compute_metrics(va_y, va_probs, thr=t)
        if m['f1'] > best_f1:
            best_f1 = m['f1']
            best_thr = float(t)
    return best_thr, compute_metrics(va_y, va_probs, thr=best_thr)


def calibrate_epoch(va_logits, va_y):
    """The post-epoch block :270-296 as one call -> dict(T, thr, metrics, probs)."""
    T, probs = fit_temperature(va_logits, va_y)
    thr, m = best_threshold(va_y, probs)
    return dict(T=T, thr=thr, metrics=m, probs=probs)


def tta_probs(logits_3, T):
    """:326-336: mean of the three TTA logits, then sigmoid(logit / T) in torch fp32."""
    logit = torch.as_tensor(logits_3, dtype=torch.float32).mean(0)
    return logit.numpy(), torch.sigmoid(logit / T).numpy()


def fold_result(fold, thr, T, va, te):
    """:347-360; va / te = (names, y, probs)."""
    return {'fold': fold, 'thr': thr, 'T': T,
            'val_metrics': compute_metrics(va[1], va[2], thr=thr), 'test_metrics': compute_metrics(te[1], te[2], thr=thr),
            'val_oof': pd.DataFrame({'image_name': va[0], 'y': va[1], 'prob': va[2]}),
            'test_pred': pd.DataFrame({'image_name': te[0], 'y': te[1], 'prob': te[2]})}


def write_outputs(outdir, results):
    """:402-434: oof_val.csv, pred_test.csv, summary.json."""
    rows = [{'fold': r['fold'], **r['val_metrics'], **{f'test_{k}': v for k, v in r['test_metrics'].items()}} for r in results]
    oof_all = pd.concat([r['val_oof'] for r in results], axis=0).reset_index(drop=True)
    test_all = pd.concat([r['test_pred'] for r in results], axis=0).reset_index(drop=True)
    keys = ['auc', 'acc', 'prec', 'rec', 'f1']
    val_mean = {k: fast_round(np.mean([r[k] for r in rows])) for k in keys}
    test_mean = {k: fast_round(np.mean([r[f'test_{k}'] for r in rows])) for k in keys}
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    oof_all.to_csv(outdir / "oof_val.csv", index=False)
    test_all.to_csv(outdir / "pred_test.csv", index=False)
    summary = {'val_mean': val_mean, 'test_mean': test_mean, 'fold_details': rows}
    with open(outdir / "summary.json", "w") as f:
        json.dump(summary, f, indent=2)
    return summary


def calib_cases():
    """Seeded validation folds shared by the golden script and the tests: name -> (logits fp32, hard labels fp32, names)."""
    rng = np.random.RandomState(1234)
    cases = {}

    def add(name, n, sep, scale, quant=None, flip=0.0):
        y = (rng.rand(n) < 0.6).astype(np.float32)
        z = (rng.randn(n) + sep * (2 * y - 1)) * scale
        if flip:
            idx = rng.rand(n) < flip
            z[idx] = -z[idx]
        if quant:
            z = np.round(z / quant) * quant
        cases[name] = (z.astype(np.float32), y, [f"img_{name}_{i:05d}.jpg" for i in range(n)])
    add("typical", 613, 1.0, 2.5)                 # over-confident model: T > 1
    add("underconf", 400, 1.5, 0.3)               # T < 1
    add("small", 37, 0.7, 1.0)
    add("ties", 500, 0.8, 1.5, quant=0.5)         # many equal scores: AUC tie handling, thresholds landing on scores
    add("noisy", 2048, 0.2, 4.0, flip=0.1)
    add("large", 10007, 1.0, 2.0)
    z, y, nm = cases["small"]
    cases["one_class"] = (z.copy(), np.ones_like(y), nm)          # AUC is NaN (:183)
    z = np.linspace(-6, 6, 101).astype(np.float32)
    cases["separable"] = (z, (z > 0).astype(np.float32), [f"img_sep_{i}.jpg" for i in range(101)])   # T wants to go to 0
    return cases
