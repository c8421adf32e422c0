"""ORACLE (test infrastructure — only tests/, __graft_entry__.smoke() and bench.py's CPU arms import it).

CPU restatement of the reference's late-fusion stacker on the libraries the reference itself calls
(sklearn LogisticRegression / roc_curve / precision_recall_fscore_support, pandas), each function citing the reference
file:line it follows.  Parity status: no vector in the reference's tests; pinned in this container against the functions
imported from the reference (tests/test_oracle_stack.py) and frozen as tests/golden/stack_golden.json
(tests/golden/make_stack_golden.py).
"""
import json
from pathlib import Path

import numpy as np
import pandas as pd
from sklearn.linear_model import LogisticRegression
from sklearn.metrics import accuracy_score, precision_recall_fscore_support, roc_auc_score, roc_curve

MODES = ['max_f1', 'max_acc', 'youden', 'target_prec', 'target_rec']


def metrics(y, p, thr=0.5):
    """experiments/fusion_v1/stack_blend.py:41-47"""
    yhat = (p >= thr).astype(int)
    auc = roc_auc_score(y, p) if len(np.unique(y)) > 1 else float('nan')
    prec, rec, f1, _ = precision_recall_fscore_support(y, yhat, average='binary', zero_division=0)
    r = lambda z: float(np.round(z, 4))  # noqa: E731
    return dict(auc=r(auc), acc=r(accuracy_score(y, yhat)), prec=r(prec), rec=r(rec), f1=r(f1))


def choose_threshold(y, p, mode='max_f1', target=0.80):
    """stack_blend.py:49-86 (twin: ui/gradio_app/stack_meta.py:66-98)"""
    ts = np.linspace(0.01, 0.99, 199)
    prf = lambda t: precision_recall_fscore_support(y, (p >= t).astype(int), average='binary', zero_division=0)  # noqa: E731
    if mode in ('max_f1', 'max_acc'):
        best_t, best = 0.5, -1
        for t in ts:
            s = prf(t)[2] if mode == 'max_f1' else accuracy_score(y, (p >= t).astype(int))
            if s > best:
                best, best_t = float(s), float(t)
        return best_t
    if mode == 'youden':
        fpr, tpr, thr = roc_curve(y, p)
        return float(thr[np.argmax(tpr - fpr)])
    if mode == 'target_prec':
        ok = [t for t in ts if prf(t)[0] >= target]
        return float(ok[0]) if ok else 0.5
    if mode == 'target_rec':
        ok = [t for t in ts if prf(t)[1] >= target]
        return float(ok[-1]) if ok else 0.5
    return 0.5


def fit_meta(X, y):
    """stack_blend.py:245-247 / stack_meta.py:55-57"""
    return LogisticRegression(max_iter=1000).fit(X, y)


def stack_blend(tab_oof, tab_test, mm_oof, mm_te, outdir, mil_oof=None, mil_te=None, thr_mode='youden', thr_target=0.80):
    """stack_blend.py:224-288 on in-memory frames."""
    tab_oof = tab_oof.rename(columns={'prob': 'prob_tab'}); tab_test = tab_test.rename(columns={'prob': 'prob_tab'})
    oof = tab_oof.merge(mm_oof.rename(columns={'prob': 'prob_mm'}), on=['image_name', 'y'], how='inner')
    test = tab_test.merge(mm_te.rename(columns={'prob': 'prob_mm'}), on=['image_name', 'y'], how='inner')
    use_mil = mil_oof is not None
    if use_mil:
        oof = oof.merge(mil_oof.rename(columns={'prob': 'prob_mil'}), on=['image_name', 'y'], how='inner')
        test = test.merge(mil_te.rename(columns={'prob': 'prob_mil'}), on=['image_name', 'y'], how='inner')
    feat_cols = ['prob_tab', 'prob_mm'] + (['prob_mil'] if use_mil else [])
    X_oof, y_oof = oof[feat_cols].values, oof['y'].values
    meta = fit_meta(X_oof, y_oof)
    p_oof = meta.predict_proba(X_oof)[:, 1]
    thr = choose_threshold(y_oof, p_oof, mode=thr_mode, target=thr_target)
    p_te = meta.predict_proba(test[feat_cols].values)[:, 1]
    summary = {'oof': metrics(y_oof, p_oof, thr), 'test': metrics(test['y'].values, p_te, thr), 'thr': float(np.round(thr, 4)),
               'thr_mode': thr_mode, 'thr_target': thr_target, 'features': feat_cols}
    Path(outdir).mkdir(parents=True, exist_ok=True)
    oof_out = oof[['image_name', 'y']].copy(); oof_out['prob'] = p_oof
    te_out = test[['image_name', 'y']].copy(); te_out['prob'] = p_te
    oof_out.to_csv(Path(outdir) / 'stack_oof.csv', index=False)
    te_out.to_csv(Path(outdir) / 'stack_test.csv', index=False)
    with open(Path(outdir) / 'summary.json', 'w') as f:
        json.dump(summary, f, indent=2)
    return dict(summary, coef=meta.coef_[0].tolist(), intercept=float(meta.intercept_[0]), p_oof=p_oof, p_te=p_te)


def stream_frames(seed=7, n=1500, n_test=400):
    """Seeded OOF / test frames of three correlated streams (tab, mm, mil) -> dict of DataFrames (image_name, y, prob)."""
    rng = np.random.RandomState(seed)
    out = {}
    for part, m in (("oof", n), ("test", n_test)):
        y = (rng.rand(m) < 0.55).astype(np.int64)
        latent = rng.randn(m) + 1.1 * (2 * y - 1)
        names = [f"{part}_{i:05d}.jpg" for i in range(m)]
        for k, (noise, gain) in {"tab": (1.2, 0.8), "mm": (0.6, 1.4), "mil": (0.9, 1.0)}.items():
            z = gain * latent + noise * rng.randn(m)
            p = 1 / (1 + np.exp(-z))
            if k == "mil":
                p = np.round(p, 2)                       # coarse scores: ties in the meta probabilities
            order = rng.permutation(m) if k != "tab" else np.arange(m)      # the merge has to realign rows
            df = pd.DataFrame({'image_name': names, 'y': y, 'prob': p}).iloc[order].reset_index(drop=True)
            if k == "mil":
                df = df.iloc[: m - 37].reset_index(drop=True)                 # inner join drops images a stream lacks
            out[f"{k}_{part}"] = df
    return out
