"""Per-fold calibration, fold loop and result files of the joint dual-task trainer, same names,
arguments and file formats as experiments/multimodal_v1/train_mm_joint_dualtask.py:

    TemperatureScaler                      :162-174   (state-dict key 'log_T', .temperature())
    compute_metrics(y_true, y_prob, thr)   :181-186   {'auc','acc','prec','rec','f1'} rounded to 4 places
    calibrate_epoch(va_logits, va_y)       :270-296   LBFGS temperature fit + 61-point F1 sweep
    run_fold(fold, loaders, args)          :188-360   train / calibrate / checkpoint / TTA predict
    write_outputs(outdir, results)         :402-434   oof_val.csv, pred_test.csv, summary.json

Validation logits stay on the device: the temperature loss/derivative, the calibrated probabilities, the confusion counts
of all 61 thresholds and the AUC rank statistic are libteethrt kernels (csrc/calib.cu); the host reads back 61x4 + 4
integers per epoch instead of walking sklearn 62 times.  The LBFGS driver is torch.optim.LBFGS with the reference's
settings, fed by the kernel, so the fitted temperature follows the same iteration.
"""
import json
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import init, TeethRTError

SWEEP = np.linspace(0.2, 0.8, 61)          # :291


def fast_round(x, n=4):
    return float(np.round(x, n))


def _dev(a, device, keep64=False):
    """-> flat contiguous device tensor, fp32 (fp64 scores stay fp64 when keep64: thresholds compare in double either way)."""
    if isinstance(a, torch.Tensor):
        dt = torch.float64 if keep64 and a.dtype == torch.float64 else torch.float32
        return a.detach().to(device=device, dtype=dt).contiguous().view(-1)
    a = np.asarray(a)
    dt = np.float64 if keep64 and a.dtype == np.float64 else np.float32
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=device).view(-1)


def _device_of(*arrays):
    for a in arrays:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device
    return torch.device("cuda", init())


class TemperatureScaler(nn.Module):
    """Single-temperature Platt scaling; T = exp(log_T).  forward(logits) -> logits / T as in the reference."""

    def __init__(self):
        super().__init__()
        self.log_T = nn.Parameter(torch.zeros(1))

    def forward(self, logits):
        return logits / self.log_T.exp()

    @torch.no_grad()
    def temperature(self):
        return float(self.log_T.exp().item())

    @torch.no_grad()
    def probs(self, logits):
        """sigmoid(logits / T) on the device (one kernel)."""
        return ops.scaled_sigmoid(_dev(logits, self.log_T.device), self.temperature())

    def fit(self, logits, targets, lr=0.1, max_iter=50):
        """The reference's on-the-fly fit (:272-284): LBFGS(lr=0.1, max_iter=50) on mean BCE(logits / T, targets); errors
        inside the optimiser are swallowed like the reference's try/except and leave the last iterate in place."""
        dev = self.log_T.device
        if dev.type != "cuda":
            raise RuntimeError("TemperatureScaler.fit runs on libteethrt kernels: move the scaler to a CUDA device first")
        logits, targets = _dev(logits, dev), _dev(targets, dev)
        if logits.numel() != targets.numel():
            raise ValueError(f"logits ({logits.numel()}) and targets ({targets.numel()}) differ in length")
        opt = torch.optim.LBFGS([self.log_T], lr=lr, max_iter=max_iter)
        out = torch.empty(2, device=dev)

        def closure():
            ops.temperature_nll(logits, targets, self.log_T.data, out)
            self.log_T.grad = out[1:2].clone()
            return out[0].clone()
        try:
            opt.step(closure)
        except TeethRTError:
            raise
        except Exception:
            pass
        self.log_T.grad = None
        return self


def _scores(counts, n):
    tp, fp, fn, tn = (int(c) for c in counts)
    acc = (tp + tn) / n
    prec = tp / (tp + fp) if tp + fp else 0.0                       # zero_division=0
    rec = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2.0 * tp / ((tp + fn) + (tp + fp)) if (tp + fn) + (tp + fp) else 0.0
    return acc, prec, rec, f1


def _auc(stat):
    wins2, npos, nneg, bad = (int(v) for v in stat)
    if bad:
        raise ValueError(f"binary labels expected: {bad} value(s) of y_true are neither 0 nor 1")
    return wins2 / (2.0 * npos * nneg) if npos and nneg else float('nan')


def metrics_sweep(y_true, y_prob, thresholds, weak_thresholds=False):
    """compute_metrics at every threshold with ONE kernel launch and one read-back -> list of dicts."""
    dev = _device_of(y_prob, y_true)
    y, p = _dev(y_true, dev), _dev(y_prob, dev, keep64=True)
    if y.numel() != p.numel() or y.numel() == 0:
        raise ValueError(f"y_true ({y.numel()}) and y_prob ({p.numel()}) must be non-empty and equal in length")
    thr64 = np.atleast_1d(np.asarray(thresholds, dtype=np.float64))
    if weak_thresholds and p.dtype == torch.float32:
        # `(y_prob >= thr)` with a float32 array and a PYTHON float compares in float32 (the scalar is "weak" under NumPy 2's
        # promotion, and value-cast under NumPy 1): (double)p >= (double)(float)thr decides exactly like p >= (float)thr.
        # np.float64 thresholds (the linspace sweep at :291) stay fp64 comparisons, as NumPy 2 evaluates them.
        thr64 = thr64.astype(np.float32).astype(np.float64)
    thr = torch.as_tensor(thr64, device=dev)
    counts, stat = ops.binary_metrics(p, y, thr)
    counts, stat = counts.cpu().numpy(), stat.cpu().numpy()
    auc = _auc(stat)
    out = []
    for c in counts:
        acc, prec, rec, f1 = _scores(c, y.numel())
        out.append({'auc': fast_round(auc), 'acc': fast_round(acc), 'prec': fast_round(prec), 'rec': fast_round(rec),
                    'f1': fast_round(f1)})
    return out


def compute_metrics(y_true, y_prob, thr=0.5):
    weak = isinstance(thr, float) and not isinstance(thr, np.floating)       # np.float64 subclasses float: it is NOT weak
    return metrics_sweep(y_true, y_prob, [thr], weak_thresholds=weak)[0]


def best_threshold(y_true, y_prob):
    """First maximum of the rounded F1 over np.linspace(0.2, 0.8, 61) (:289-295) -> (thr, metrics at thr)."""
    ms = metrics_sweep(y_true, y_prob, SWEEP)
    best_thr, best_f1, best_m = 0.5, -1.0, None
    for t, m in zip(SWEEP, ms):
        if m['f1'] > best_f1:
            best_f1, best_thr, best_m = m['f1'], float(t), m
    if best_m is None:                      # every F1 was NaN-free by construction; kept for symmetry with thr=0.5 default
        best_m = compute_metrics(y_true, y_prob, best_thr)
    return best_thr, best_m


def calibrate_epoch(va_logits, va_y):
    """Post-epoch block (:270-296): -> dict(T, thr, metrics, probs (device fp32), scaler)."""
    dev = _device_of(va_logits, va_y)
    scaler = TemperatureScaler().to(dev).fit(va_logits, va_y)
    probs = scaler.probs(va_logits)
    thr, m = best_threshold(_dev(va_y, dev), probs)
    return dict(T=scaler.temperature(), thr=thr, metrics=m, probs=probs, scaler=scaler)


# ------------------------------------------------------------------------------------------------ fold loop
@torch.no_grad()
def collect_logits(model, loader, device):
    """Eval pass (:258-269) -> (logits, y_h) device fp32; no per-batch host read."""
    model.eval()
    logits, ys = [], []
    for (x_img, x_tab, y_h, _y_s, _w, _names) in loader:
        logit, _ = model(x_img.to(device, non_blocking=True), x_tab.to(device, non_blocking=True))
        logits.append(logit.float().view(-1))
        ys.append(y_h.to(device, non_blocking=True).float().view(-1))
    return torch.cat(logits), torch.cat(ys)


@torch.no_grad()
def predict_tta(model, loader, T, device):
    """_predict (:321-345): mean logit over {identity, W-flip, H-flip}, prob = sigmoid(logit / T)
    -> (logits, probs, y, names) as numpy / list, one device->host read per call."""
    from .infer import tta_logit
    model.eval()
    logits, ys, names_all = [], [], []
    for (x_img, x_tab, y_h, _y_s, _w, names) in loader:
        logits.append(tta_logit(model, x_img.to(device, non_blocking=True), x_tab.to(device, non_blocking=True)).float().view(-1))
        ys.append(y_h.float().view(-1))
        names_all.extend(list(names))
    logits = torch.cat(logits).contiguous()
    probs = ops.scaled_sigmoid(logits, T)
    return logits.cpu().numpy(), probs.cpu().numpy(), torch.cat(ys).numpy(), names_all


def run_fold(fold, loaders, args, model=None, scaler_stats=(None, None), log=print, reload_best=False):
    """run_fold (:188-360) on ready-made loaders = (dl_tr, dl_va, dl_te) that yield the reference's batches
    (x_img, x_tab, y_h, y_s, w, names); `args` carries the reference's argparse names.  Saves
    `<outdir>/mm_dualtask_fold{fold}.pt` with the reference's checkpoint keys whenever validation AUC improves.

    Reference quirk kept by default: `best_state['model'] = model.state_dict()` (:302-303) holds the LIVE tensors, so the
    reload at :318 is a no-op and the OOF / test predictions come from the last epoch's weights with the best epoch's T
    and threshold (the .pt file on disk does hold the best epoch).  reload_best=True predicts with the saved weights."""
    from .modules import MMJointDualHead
    from .infer import TAB_FEATURES
    from .train import DualTaskTrainer
    import pandas as pd
    device = torch.device("cuda", init())
    dl_tr, dl_va, dl_te = loaders
    if model is None:
        model = MMJointDualHead(args.backbone, tab_in=len(TAB_FEATURES), tab_hidden=args.tab_hidden, drop=args.dropout).to(device)
    iters_per_epoch = max(1, len(dl_tr))
    trainer = DualTaskTrainer(model, lr=args.lr, weight_decay=args.weight_decay, t_max=args.epochs * iters_per_epoch,
                              alpha=args.alpha, beta=args.beta, grad_clip=args.grad_clip,
                              use_sample_weights=bool(args.use_sample_weights), graph=getattr(args, "graph", True))
    best_auc, best_state = -1.0, None
    history = {'epoch': [], 'tr_loss': [], 'va_auc': [], 'va_f1': []}
    for epoch in range(1, args.epochs + 1):
        model.train()
        losses = []
        for (x_img, x_tab, y_h, y_s, w, _names) in dl_tr:
            losses.append(trainer.step(x_img, x_tab, y_h, y_s, w).clone())
        tr_loss = float(torch.cat(losses).mean()) if losses else float('nan')      # the epoch's only loss read-back
        va_logits, va_y = collect_logits(model, dl_va, device)
        cal = calibrate_epoch(va_logits, va_y)
        m_va = cal['metrics']
        log(f"[Fold {fold}][Epoch {epoch}] tr_loss={tr_loss:.4f}  val_auc={m_va['auc']:.4f} f1={m_va['f1']:.3f} "
            f"thr*={cal['thr']:.3f}  T={cal['T']:.3f}")
        history['epoch'].append(epoch); history['tr_loss'].append(tr_loss)
        history['va_auc'].append(m_va['auc']); history['va_f1'].append(m_va['f1'])
        if m_va['auc'] > best_auc:
            best_auc = m_va['auc']
            best_state = {'model': model.state_dict(), 'scaler_mean': scaler_stats[0], 'scaler_scale': scaler_stats[1],
                          'thr': cal['thr'], 'T': cal['T'], 'args': dict(vars(args)), 'epoch': epoch}
            outdir = Path(args.outdir); outdir.mkdir(parents=True, exist_ok=True)
            torch.save(best_state, outdir / f"mm_dualtask_fold{fold}.pt")
    if best_state is None:
        raise RuntimeError("validation AUC was NaN in every epoch (single-class validation fold): no checkpoint to reload")
    if reload_best:
        model.load_state_dict(torch.load(Path(args.outdir) / f"mm_dualtask_fold{fold}.pt", map_location=device)['model'])
    T, thr = best_state['T'], best_state['thr']
    _, va_probs, va_y, va_names = predict_tta(model, dl_va, T, device)
    _, te_probs, te_y, te_names = predict_tta(model, dl_te, T, device)
    return {'fold': fold, 'thr': thr, 'T': T, 'history': history,
            'val_metrics': compute_metrics(va_y, va_probs, thr=thr), 'test_metrics': compute_metrics(te_y, te_probs, thr=thr),
            'val_oof': pd.DataFrame({'image_name': va_names, 'y': va_y, 'prob': va_probs}),
            'test_pred': pd.DataFrame({'image_name': te_names, 'y': te_y, 'prob': te_probs})}


def write_outputs(outdir, results):
    """main()'s tail (:402-434): concatenated OOF / test predictions and the summary of per-fold + mean metrics."""
    import pandas as pd
    rows = [{'fold': r['fold'], **r['val_metrics'], **{f'test_{k}': v for k, v in r['test_metrics'].items()}} for r in results]
    keys = ['auc', 'acc', 'prec', 'rec', 'f1']
    summary = {'val_mean': {k: fast_round(np.mean([r[k] for r in rows])) for k in keys},
               'test_mean': {k: fast_round(np.mean([r[f'test_{k}'] for r in rows])) for k in keys},
               'fold_details': rows}
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    pd.concat([r['val_oof'] for r in results], axis=0).reset_index(drop=True).to_csv(outdir / "oof_val.csv", index=False)
    pd.concat([r['test_pred'] for r in results], axis=0).reset_index(drop=True).to_csv(outdir / "pred_test.csv", index=False)
    with open(outdir / "summary.json", "w") as f:
        json.dump(summary, f, indent=2)
    return summary


def finalize_from_ckpts(ckpt_dir, fold_loaders, outdir, folds=5, log=print):
    """experiments/multimodal_v1/finalize_mm_dualtask_from_ckpts.py:110-184: reload every `mm_dualtask_fold{k}.pt`, re-run the
    TTA inference on that fold's validation and the test loader, and write `oof_val.csv`, `pred_test.csv` and the script's
    `summary.json` ({'val_mean', 'test_mean', 'folds': [{'fold','thr','T','val','test'}]}).  `fold_loaders(fold, ckpt)` returns
    (dl_va, dl_te) built from the checkpoint's args / scaler statistics as the script does (:122-139); batches are
    (x_img, x_tab, y, ..., names).  Missing folds are skipped with the script's warning; no fold at all raises SystemExit."""
    import pandas as pd
    from .modules import MMNet
    from .infer import TAB_FEATURES
    device = torch.device("cuda", init())
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    oof_list, test_list, fold_summ = [], [], []
    for fold in range(folds):
        ckpt_path = Path(ckpt_dir) / f"mm_dualtask_fold{fold}.pt"
        if not ckpt_path.exists():
            log(f"[WARN] missing {ckpt_path}, skipping fold {fold}")
            continue
        ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=False)
        a, thr, T = ckpt['args'], ckpt['thr'], ckpt['T']
        model = MMNet(backbone=a['backbone'], tab_in=len(TAB_FEATURES), tab_hidden=a['tab_hidden'], drop=a['dropout']).to(device)
        model.load_state_dict(ckpt['model'])
        model.eval()
        dl_va, dl_te = fold_loaders(fold, ckpt)
        # predict_tta reads (x_img, x_tab, y_h, y_s, w, names); the finalize script's dataset yields (x_img, x_tab, y, names)
        six = lambda dl: ((b[0], b[1], b[2], None, None, b[-1]) for b in dl)  # noqa: E731
        _, pv, yv, nv = predict_tta(model, six(dl_va), T, device)
        _, pt, yt, nt = predict_tta(model, six(dl_te), T, device)
        oof_list.append(pd.DataFrame({'image_name': nv, 'y': yv, 'prob': pv}))
        test_list.append(pd.DataFrame({'image_name': nt, 'y': yt, 'prob': pt}))
        fold_summ.append({'fold': fold, 'thr': thr, 'T': T, 'val': compute_metrics(yv, pv, thr), 'test': compute_metrics(yt, pt, thr)})
        log(f"[Fold {fold}] VAL {fold_summ[-1]['val']} | TEST {fold_summ[-1]['test']}")
    if not oof_list:
        raise SystemExit("No folds finalized. Check ckpt-dir path.")
    pd.concat(oof_list).reset_index(drop=True).to_csv(outdir / "oof_val.csv", index=False)
    pd.concat(test_list).reset_index(drop=True).to_csv(outdir / "pred_test.csv", index=False)
    keys = ['auc', 'acc', 'prec', 'rec', 'f1']
    summary = {'val_mean': {k: fast_round(np.mean([f['val'][k] for f in fold_summ])) for k in keys},
               'test_mean': {k: fast_round(np.mean([f['test'][k] for f in fold_summ])) for k in keys}, 'folds': fold_summ}
    with open(outdir / "summary.json", "w") as f:
        json.dump(summary, f, indent=2)
    log("=== VAL (mean) ===", summary['val_mean'])
    log("=== TEST (mean) ===", summary['test_mean'])
    return summary
