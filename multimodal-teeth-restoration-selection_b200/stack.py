"""Late-fusion stacker over the stream probabilities, same names, arguments and files as

    experiments/fusion_v1/stack_blend.py   _metrics :41-47, choose_threshold :49-86, main :193-288 (-> stack_blend())
    ui/gradio_app/stack_meta.py            Stacker :7-127

It consumes the oof_val.csv / pred_test.csv files teethrt.calib.write_outputs produces.  The meta-learner (L2 logistic
regression, sklearn's default C = 1) is fitted by Newton on the device (csrc/calib.cu), its probabilities and the confusion
counts behind every threshold mode are device kernels; pandas only reads, merges and writes the frames.  The tabular OOF
builder (LightGBM, stack_blend.py:91-190) is outside this row: pass its two frames / CSVs in.
"""
import json
from pathlib import Path

import numpy as np
import pandas as pd
import torch

from . import ops
from ._lib import init
from .calib import metrics_sweep

THRESHOLDS = np.linspace(0.01, 0.99, 199)          # stack_blend.py:50


def _device():
    return torch.device("cuda", init())


def _metrics(y, p, thr=0.5):
    """stack_blend.py:41-47 (same numbers as calib.compute_metrics; scores compare in fp64)."""
    return metrics_sweep(np.asarray(y, dtype=np.float32), np.asarray(p, dtype=np.float64), [thr])[0]


def _counts(y, p, thresholds):
    dev = _device()
    yt = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32), device=dev)
    pt = p if isinstance(p, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(p, dtype=np.float64), device=dev)
    counts, stat = ops.binary_metrics(pt.to(torch.float64).contiguous(), yt, torch.as_tensor(np.asarray(thresholds, np.float64), device=dev))
    if int(stat[3]):
        raise ValueError("binary labels expected (0 / 1)")
    return counts.cpu().numpy().astype(np.float64)


def choose_threshold(y, p, mode='max_f1', target=0.80):
    """stack_blend.py:49-86 / stack_meta.py:66-98: one kernel launch per call instead of 199 sklearn walks."""
    y = np.asarray(y)
    ts = THRESHOLDS
    if mode in ('max_f1', 'max_acc', 'target_prec', 'target_rec'):
        c = _counts(y, p, ts)
        tp, fp, fn, tn = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
        with np.errstate(divide='ignore', invalid='ignore'):
            if mode == 'max_f1':
                score = np.where(2 * tp + fp + fn > 0, 2 * tp / ((tp + fn) + (tp + fp)), 0.0)
            elif mode == 'max_acc':
                score = (tp + tn) / len(y)
            elif mode == 'target_prec':
                score = np.where(tp + fp > 0, tp / (tp + fp), 0.0)
            else:
                score = np.where(tp + fn > 0, tp / (tp + fn), 0.0)
        if mode in ('max_f1', 'max_acc'):
            best_t, best = 0.5, -1
            for t, s in zip(ts, score):
                if s > best:
                    best, best_t = float(s), float(t)
            return best_t
        ok = [t for t, s in zip(ts, score) if s >= target]
        if not ok:
            return 0.5
        return float(ok[0]) if mode == 'target_prec' else float(ok[-1])
    if mode == 'youden':
        # sklearn.metrics.roc_curve(y, p) (drop_intermediate=True): thresholds are the distinct scores, descending, with
        # (tps, fps) = cumulative counts; points whose second differences vanish are dropped; a leading (0, 0, inf) is added
        pt = torch.as_tensor(np.ascontiguousarray(p, dtype=np.float64), device=_device())
        thr = torch.flip(torch.unique(pt), dims=[0]).cpu().numpy()
        c = _counts(y, pt, thr)
        tps, fps = c[:, 0], c[:, 1]
        if len(fps) > 2:
            keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
            fps, tps, thr = fps[keep], tps[keep], thr[keep]
        tps, fps, thr = np.r_[0, tps], np.r_[0, fps], np.r_[np.inf, thr]
        with np.errstate(divide='ignore', invalid='ignore'):
            fpr = fps / fps[-1] if fps[-1] > 0 else np.repeat(np.nan, fps.shape)
            tpr = tps / tps[-1] if tps[-1] > 0 else np.repeat(np.nan, tps.shape)
        return float(thr[np.argmax(tpr - fpr)])
    return 0.5


class LogisticMeta:
    """The slice of sklearn.linear_model.LogisticRegression the stacker uses: fit(X, y), predict_proba(X), coef_, intercept_."""

    def __init__(self, C=1.0, max_iter=1000):
        self.C, self.max_iter = float(C), int(max_iter)
        self.coef_ = self.intercept_ = self.n_iter_ = None

    def fit(self, X, y):
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.asarray(y)
        classes = np.unique(y)
        if len(classes) != 2:
            raise ValueError(f"This solver needs samples of at least 2 classes in the data, but the data contains only {len(classes)}")
        self.classes_ = classes
        dev = _device()
        coef, info = ops.logreg_fit(torch.as_tensor(X, device=dev), torch.as_tensor((y == classes[1]).astype(np.float32), device=dev),
                                    C=self.C, max_iter=self.max_iter)
        coef, info = coef.cpu().numpy(), info.cpu().numpy()
        self.coef_, self.intercept_, self.n_iter_ = coef[None, :-1].copy(), coef[-1:].copy(), np.array([int(info[0])])
        self._coef_dev = torch.as_tensor(coef, device=dev)
        return self

    def predict_proba(self, X):
        X = np.asarray(X)
        if X.shape[0] <= 4:                 # single-case UI call (stack_meta.py:114-116): three flops, no launch
            z = X.astype(np.float64) @ self.coef_[0] + self.intercept_[0]
            p1 = 1.0 / (1.0 + np.exp(-z))
        else:
            p1 = ops.logreg_predict(torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64), device=self._coef_dev.device),
                                    self._coef_dev).cpu().numpy()
        return np.stack([1.0 - p1, p1], axis=1)


def _frame(src):
    return src.copy() if isinstance(src, pd.DataFrame) else pd.read_csv(src)


def stack_blend(tab_oof, tab_test, oof_mm, pred_mm, outdir, oof_mil='', pred_mil='', thr_mode='youden', thr_target=0.80, log=print):
    """main() of stack_blend.py from step 2 on (:224-288).  tab_oof / tab_test: frames or CSVs with image_name, y, prob
    (what fit_tab_oof returns).  Writes stack_oof.csv, stack_test.csv, summary.json; returns the summary dict."""
    Path(outdir).mkdir(parents=True, exist_ok=True)
    tab_oof = _frame(tab_oof).rename(columns={'prob': 'prob_tab'})
    tab_test = _frame(tab_test).rename(columns={'prob': 'prob_tab'})
    mm_oof = _frame(oof_mm).rename(columns={'prob': 'prob_mm'})
    mm_te = _frame(pred_mm).rename(columns={'prob': 'prob_mm'})
    use_mil = not isinstance(oof_mil, str) or (bool(oof_mil.strip()) and bool(str(pred_mil).strip()))
    oof = tab_oof.merge(mm_oof, on=['image_name', 'y'], how='inner')
    test = tab_test.merge(mm_te, on=['image_name', 'y'], how='inner')
    if use_mil:
        oof = oof.merge(_frame(oof_mil).rename(columns={'prob': 'prob_mil'}), on=['image_name', 'y'], how='inner')
        test = test.merge(_frame(pred_mil).rename(columns={'prob': 'prob_mil'}), on=['image_name', 'y'], how='inner')
    feat_cols = ['prob_tab', 'prob_mm'] + (['prob_mil'] if use_mil else [])
    X_oof, y_oof = oof[feat_cols].values, oof['y'].values
    meta = LogisticMeta(max_iter=1000).fit(X_oof, y_oof)
    p_oof = meta.predict_proba(X_oof)[:, 1]
    thr = choose_threshold(y_oof, p_oof, mode=thr_mode, target=thr_target)
    p_te = meta.predict_proba(test[feat_cols].values)[:, 1]
    m_oof, m_te = _metrics(y_oof, p_oof, thr), _metrics(test['y'].values, p_te, thr)
    oof_out = oof[['image_name', 'y']].copy(); oof_out['prob'] = p_oof
    te_out = test[['image_name', 'y']].copy(); te_out['prob'] = p_te
    oof_out.to_csv(Path(outdir) / 'stack_oof.csv', index=False)
    te_out.to_csv(Path(outdir) / 'stack_test.csv', index=False)
    summary = {'oof': m_oof, 'test': m_te, 'thr': float(np.round(thr, 4)), 'thr_mode': thr_mode, 'thr_target': thr_target,
               'features': feat_cols}
    with open(Path(outdir) / 'summary.json', 'w') as f:
        json.dump(summary, f, indent=2)
    log("Features used:", feat_cols)
    log(f"Threshold mode: {thr_mode} | target: {thr_target} | chosen thr: {thr:.3f}")
    log("=== OOF ===", m_oof)
    log("=== TEST ===", m_te)
    return dict(summary, meta=meta)


class Stacker:
    """ui/gradio_app/stack_meta.py:7-127: image-only meta (MM + MIL) fitted on the OOF files at construction, an equal-weight
    blend with the tabular probability when one is supplied, threshold re-selection by mode."""

    def __init__(self, xlsx_tab, oof_mm, pred_mm, oof_mil, pred_mil, folds=5, thr_mode='max_acc', thr_target=0.80):
        self.xlsx_tab = Path(xlsx_tab)
        self.oof_mm, self.pred_mm = Path(oof_mm), Path(pred_mm)
        self.oof_mil, self.pred_mil = Path(oof_mil), Path(pred_mil)
        self.folds, self.thr_mode, self.thr_target = folds, thr_mode, thr_target
        self.meta_full, self.thr_full = None, 0.5
        self.meta_img, self.thr_img = None, 0.5
        self._fit_metas()

    def _oof_probs(self, meta, frame, cols):
        return meta.predict_proba(frame[cols].values)[:, 1]

    def set_threshold_mode(self, mode, target=0.80):
        self.thr_mode, self.thr_target = mode, target
        if self.meta_full is not None:
            p = self._oof_probs(self.meta_full, self.oof_full, ['prob_tab', 'prob_mm', 'prob_mil'])
            self.thr_full = self._choose_threshold(self.oof_full['y'].values, p, mode, target)
        if self.meta_img is not None:
            p = self._oof_probs(self.meta_img, self.oof_img, ['prob_mm', 'prob_mil'])
            self.thr_img = self._choose_threshold(self.oof_img['y'].values, p, mode, target)

    def _load_oof_tab(self):
        return None                                        # stack_meta.py:33-39: the UI never rebuilds the tabular OOF

    def _fit_metas(self):
        mm_oof = pd.read_csv(self.oof_mm).rename(columns={'prob': 'prob_mm'})
        mil_oof = pd.read_csv(self.oof_mil).rename(columns={'prob': 'prob_mil'})
        self.oof_img = mm_oof.merge(mil_oof, on=['image_name', 'y'], how='inner')
        self.meta_img = LogisticMeta(max_iter=1000).fit(self.oof_img[['prob_mm', 'prob_mil']].values, self.oof_img['y'].values)
        p_oof = self._oof_probs(self.meta_img, self.oof_img, ['prob_mm', 'prob_mil'])
        self.thr_img = self._choose_threshold(self.oof_img['y'].values, p_oof, self.thr_mode, self.thr_target)

    def _choose_threshold(self, y, p, mode='max_acc', target=0.80):
        return choose_threshold(y, p, mode, target)

    def _train_meta_full_if_needed(self, prob_tab_available, tab_oof_df):
        if self.meta_full is not None or not prob_tab_available or tab_oof_df is None:
            return
        self.oof_full = tab_oof_df.rename(columns={'prob': 'prob_tab'}).merge(self.oof_img, on=['image_name', 'y'], how='inner')
        cols = ['prob_tab', 'prob_mm', 'prob_mil']
        self.meta_full = LogisticMeta(max_iter=1000).fit(self.oof_full[cols].values, self.oof_full['y'].values)
        self.thr_full = self._choose_threshold(self.oof_full['y'].values, self._oof_probs(self.meta_full, self.oof_full, cols),
                                               self.thr_mode, self.thr_target)

    def predict_single(self, prob_mm, prob_mil, prob_tab):
        """-> (final probability, threshold, detail dict)  (stack_meta.py:110-127)."""
        X = np.array([[prob_mm, prob_mil]], dtype=np.float32)
        p_img = self.meta_img.predict_proba(X)[:, 1][0]
        if prob_tab is None:
            return float(p_img), float(self.thr_img), {"mode": "img-only"}
        p = 0.5 * p_img + 0.5 * float(prob_tab)
        thr = self._choose_threshold(self.oof_img['y'].values, self._oof_probs(self.meta_img, self.oof_img, ['prob_mm', 'prob_mil']),
                                     self.thr_mode, self.thr_target)
        return float(p), float(thr), {"mode": "hybrid(0.5*img_meta + 0.5*tab)"}
