"""GPU parity for the fold calibration path: teethrt.calib (csrc/calib.cu underneath) against the
oracle restatement (oracle/ref_calib.py) and the golden fixture minted from the reference (tests/golden/calib_golden.json).

Bars: confusion counts / thresholds / rounded metrics are integer work -> identical; temperature loss and derivative are
fp32 with fp64 accumulation -> 1e-5 relative; the fitted T follows the reference's LBFGS iteration -> 2e-3 relative on the
well-conditioned folds, 5 % on the folds where the reference's line-search-free LBFGS itself overshoots by orders of
magnitude (T = 2e-5 on 'underconf')."""
import json
import math
import os
import types
import warnings

import numpy as np
import pytest
import torch

import ref_calib as RC   # oracle (checker only)
import ref_models as R

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "calib_golden.json")))
CASES = RC.calib_cases()
WILD = {"underconf", "one_class", "separable"}
warnings.filterwarnings("ignore", message="overflow encountered in exp")


@pytest.fixture(scope="module")
def cal():
    import teethrt
    teethrt.init()
    from teethrt import calib
    return calib


def same_metrics(a, b):
    return all((math.isnan(a[k]) and math.isnan(b[k])) or a[k] == b[k] for k in ('auc', 'acc', 'prec', 'rec', 'f1'))


@pytest.mark.parametrize("name", ["typical", "small", "large", "separable"])
def test_temperature_loss_and_derivative(cal, name):
    from teethrt import ops
    z, y, _ = CASES[name]
    zt, yt = torch.tensor(z), torch.tensor(y)
    for lt in (0.0, 0.3, -1.2, 2.5, -6.0):
        log_T = torch.tensor([lt], requires_grad=True)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(zt.double() / log_T.double().exp(), yt.double())
        loss.backward()
        out = ops.temperature_nll(zt.cuda(), yt.cuda(), torch.tensor([lt], device="cuda")).cpu()
        assert float(out[0]) == pytest.approx(float(loss), rel=1e-5, abs=1e-7), (name, lt)
        assert float(out[1]) == pytest.approx(float(log_T.grad), rel=2e-5, abs=1e-6), (name, lt)


@pytest.mark.parametrize("name", sorted(CASES))
def test_fitted_temperature_follows_reference(cal, name):
    z, y, _ = CASES[name]
    s = cal.TemperatureScaler().cuda().fit(z, y)
    assert list(s.state_dict()) == ["log_T"]
    assert s.temperature() == pytest.approx(GOLD["cases"][name]["T"], rel=5e-2 if name in WILD else 2e-3)
    p = s.probs(z).cpu().numpy()
    with np.errstate(over="ignore"):
        want = 1 / (1 + np.exp(-(z / np.float32(s.temperature()))))
    assert np.abs(p - want).max() < 1e-6


def test_python_float_threshold_compares_in_float32(cal):
    """float32(0.7) < 0.7: the reference's `(y_prob >= thr)` with a float32 array and a Python float is a float32 comparison,
    so a score equal to float32(0.7) is a positive prediction (train_mm_joint_dualtask.py:182); an np.float64 threshold (the
    linspace sweep, :291) compares in fp64 under NumPy 2 and the same score is negative."""
    probs = np.array([np.float32(0.7), 0.1, 0.9, np.float32(0.7), 0.3], np.float32)
    y = np.array([1, 0, 1, 0, 1], np.float32)
    for thr in (0.7, np.float64(0.7), 0.3, np.float64(0.3)):
        assert same_metrics(cal.compute_metrics(y, probs, thr), RC.compute_metrics(y, probs, thr)), thr
    assert cal.compute_metrics(y, probs, 0.7) != cal.compute_metrics(y, probs, np.float64(0.7)) or np.lib.NumpyVersion(np.__version__) < "2.0.0"


@pytest.mark.parametrize("name", sorted(CASES))
def test_metrics_identical_to_oracle_at_every_threshold(cal, name):
    z, y, _ = CASES[name]
    _, probs = RC.fit_temperature(z, y)
    probs = probs.astype(np.float32)
    # the sweep grid plus thresholds that land exactly on scores (>= must include them) and outside [0, 1]
    thr = list(np.linspace(0.2, 0.8, 61)) + [float(probs[0]), float(np.median(probs)), float(np.float32(0.5)), 0.0, 1.0, 1.5]
    got = cal.metrics_sweep(y, probs, thr)
    for t, g in zip(thr, got):
        assert same_metrics(g, RC.compute_metrics(y, probs, t)), (name, t, g, RC.compute_metrics(y, probs, t))
    assert same_metrics(cal.compute_metrics(torch.tensor(y).cuda(), torch.tensor(probs).cuda(), 0.37), RC.compute_metrics(y, probs, 0.37))
    bt, bm = cal.best_threshold(y, probs)
    ot, om = RC.best_threshold(y, probs)
    assert bt == ot and same_metrics(bm, om)


@pytest.mark.parametrize("name", sorted(CASES))
def test_calibrate_epoch_vs_golden(cal, name):
    z, y, _ = CASES[name]
    g = GOLD["cases"][name]
    c = cal.calibrate_epoch(torch.tensor(z).cuda(), torch.tensor(y).cuda())
    probs = c["probs"].cpu().numpy()
    ot, om = RC.best_threshold(y, probs)                      # exact, given the product's own probabilities
    assert c["thr"] == ot and same_metrics(c["metrics"], om)
    if name not in WILD:                                      # and the reference's numbers end to end
        assert abs(c["thr"] - g["thr"]) <= 0.0100001
        assert all(abs(c["metrics"][k] - g["metrics"][k]) <= 5e-3 for k in ("auc", "acc", "prec", "rec", "f1"))
        assert float(np.sum(probs, dtype=np.float64)) == pytest.approx(g["prob_sum"], rel=1e-3)


def test_label_and_shape_errors(cal):
    with pytest.raises(ValueError):
        cal.compute_metrics(np.array([0, 1, 0.5], np.float32), np.array([0.1, 0.9, 0.4], np.float32))      # soft label
    with pytest.raises(ValueError):
        cal.compute_metrics(np.array([0, 1], np.float32), np.array([0.1, 0.9, 0.4], np.float32))
    with pytest.raises(ValueError):
        cal.compute_metrics(np.zeros(0, np.float32), np.zeros(0, np.float32))
    with pytest.raises(ValueError):
        cal.TemperatureScaler().cuda().fit(np.zeros(3, np.float32), np.zeros(4, np.float32))
    with pytest.raises(RuntimeError):
        cal.TemperatureScaler().fit(np.zeros(3, np.float32), np.zeros(3, np.float32))                        # no CPU path


def test_result_files_byte_identical(cal, tmp_path):
    mine, theirs = [], []
    for fold, name in enumerate(("typical", "ties", "small")):
        z, y, names = CASES[name]
        c = cal.calibrate_epoch(z, y)
        p = c["probs"].cpu().numpy()
        va, te = (names, y, p), (names[:25], y[:25], p[:25])
        theirs.append(RC.fold_result(fold, c["thr"], c["T"], va, te))
        import pandas as pd
        mine.append({'fold': fold, 'thr': c["thr"], 'T': c["T"], 'val_metrics': cal.compute_metrics(va[1], va[2], c["thr"]),
                     'test_metrics': cal.compute_metrics(te[1], te[2], c["thr"]),
                     'val_oof': pd.DataFrame({'image_name': va[0], 'y': va[1], 'prob': va[2]}),
                     'test_pred': pd.DataFrame({'image_name': te[0], 'y': te[1], 'prob': te[2]})})
    sa = cal.write_outputs(tmp_path / "a", mine)
    sb = RC.write_outputs(tmp_path / "b", theirs)
    assert sa == sb
    for f in ("oof_val.csv", "pred_test.csv", "summary.json"):
        assert open(tmp_path / "a" / f, "rb").read() == open(tmp_path / "b" / f, "rb").read(), f


def _loader(nb, B, img, seed, last=None):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(nb):
        b = last if (last and i == nb - 1) else B
        yh = (torch.rand(b, generator=g) < 0.6).float()
        x = torch.randn(b, 3, img, img, generator=g) + 0.5 * (2 * yh - 1).view(-1, 1, 1, 1)       # learnable signal
        out.append((x, torch.randn(b, 9, generator=g), yh, (yh * 0.8 + 0.2 * torch.rand(b, generator=g)).clamp(0, 1),
                    torch.ones(b), [f"s{seed}_{i}_{j}.jpg" for j in range(b)]))
    return out


def test_run_fold_end_to_end(cal, tmp_path):
    """Two epochs of the fold loop on a B0 backbone with a ragged last batch: checkpoint contract, result frames, the
    calibration of the final predictions, and the reference's live-state-dict quirk (predictions use the LAST weights)."""
    from teethrt.modules import MMJointDualHead
    torch.manual_seed(0)
    ora = R.seeded_model("mm", seed=5, warm=1, img=64, backbone="tf_efficientnet_b0_ns", drop=0.0)
    model = MMJointDualHead("tf_efficientnet_b0_ns", tab_in=9, tab_hidden=64, drop=0.0).cuda()
    model.load_state_dict(ora.state_dict(), strict=True)
    args = types.SimpleNamespace(backbone="tf_efficientnet_b0_ns", tab_hidden=64, dropout=0.0, lr=3e-4, weight_decay=1e-4,
                                 epochs=2, alpha=1.0, beta=0.3, grad_clip=1.0, use_sample_weights=0, outdir=str(tmp_path),
                                 batch_size=8, img_size=64)
    dl_tr, dl_va, dl_te = _loader(5, 8, 64, 1, last=5), _loader(3, 16, 64, 2, last=7), _loader(2, 16, 64, 3, last=3)
    lines = []
    res = cal.run_fold(0, (dl_tr, dl_va, dl_te), args, model=model, scaler_stats=(np.zeros(9), np.ones(9)), log=lines.append)
    assert len(lines) == 2 and lines[0].startswith("[Fold 0][Epoch 1] tr_loss=")
    ck = torch.load(tmp_path / "mm_dualtask_fold0.pt", weights_only=False)
    assert list(ck) == ['model', 'scaler_mean', 'scaler_scale', 'thr', 'T', 'args', 'epoch']
    assert list(ck['model']) == list(ora.state_dict()) and ck['T'] == res['T'] and ck['thr'] == res['thr']
    assert len(res['val_oof']) == 16 + 16 + 7 and len(res['test_pred']) == 16 + 3
    assert list(res['val_oof'].columns) == ['image_name', 'y', 'prob'] and res['val_oof']['prob'].dtype == np.float32
    assert res['val_oof']['image_name'][0] == "s2_0_0.jpg"
    # final predictions = TTA of the LIVE (last-epoch) weights at the checkpointed T: check against the oracle model
    ora.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ora.eval()
    with torch.no_grad():
        ls = []
        for (x, xt, *_r) in dl_va:
            ls.append(torch.stack([ora(x, xt)[0], ora(torch.flip(x, dims=[3]), xt)[0], ora(torch.flip(x, dims=[2]), xt)[0]]).mean(0))
    want = torch.sigmoid(torch.cat(ls) / res['T']).numpy()
    assert np.abs(res['val_oof']['prob'].to_numpy() - want).max() < 2e-2
    assert same_metrics(res['val_metrics'], RC.compute_metrics(res['val_oof']['y'].to_numpy(), res['val_oof']['prob'].to_numpy(), res['thr']))
    s = cal.write_outputs(tmp_path, [res])
    assert s['fold_details'][0]['fold'] == 0 and (tmp_path / "oof_val.csv").exists()
    # --- finalize_mm_dualtask_from_ckpts.py: reload the checkpoint written above and regenerate the prediction files
    four = lambda dl: [(b[0], b[1], b[2], b[5]) for b in dl]        # that script's dataset yields (x_img, x_tab, y, names)
    seen = []
    fin = cal.finalize_from_ckpts(tmp_path, lambda fold, ck: (four(dl_va), four(dl_te)), tmp_path / "finalized", folds=3, log=lambda *a: seen.append(a))
    assert list(fin) == ['val_mean', 'test_mean', 'folds'] and [f['fold'] for f in fin['folds']] == [0]
    assert sum("[WARN] missing" in str(a[0]) for a in seen) == 2                 # folds 1 and 2 have no checkpoint
    assert fin['folds'][0]['T'] == res['T'] and fin['folds'][0]['thr'] == res['thr']
    import pandas as pd
    oof = pd.read_csv(tmp_path / "finalized" / "oof_val.csv")
    assert list(oof.columns) == ['image_name', 'y', 'prob'] and len(oof) == 39
    # the checkpoint holds the BEST epoch, run_fold predicted with the LAST one (the reference's quirk): same weights only if
    # the last epoch was the best; either way the file must equal the oracle model loaded from that checkpoint
    ck = torch.load(tmp_path / "mm_dualtask_fold0.pt", weights_only=False)
    ora.load_state_dict({k: v.cpu() for k, v in ck['model'].items()})
    with torch.no_grad():
        ls = [torch.stack([ora(x, xt)[0], ora(torch.flip(x, dims=[3]), xt)[0], ora(torch.flip(x, dims=[2]), xt)[0]]).mean(0) for (x, xt, *_r) in dl_va]
    assert np.abs(oof['prob'].to_numpy() - torch.sigmoid(torch.cat(ls) / ck['T']).numpy()).max() < 2e-2
    assert same_metrics(fin['folds'][0]['val'], RC.compute_metrics(oof['y'].to_numpy(), oof['prob'].to_numpy().astype(np.float32), ck['thr']))
    with pytest.raises(SystemExit):
        cal.finalize_from_ckpts(tmp_path / "nowhere", lambda fold, ck: ([], []), tmp_path / "x", log=lambda *a: None)


def test_trainer_handles_ragged_batches(cal):
    """Graph-replayed steps with alternating batch sizes equal the eager path (per-shape plans in _FusedTrainer)."""
    from teethrt.modules import MMJointDualHead
    from teethrt.train import DualTaskTrainer
    sd = R.seeded_model("mm", seed=6, warm=1, img=64, backbone="tf_efficientnet_b0_ns", drop=0.0).state_dict()
    batches = _loader(4, 8, 64, 11) + _loader(1, 5, 64, 12) + _loader(3, 8, 64, 13) + _loader(3, 5, 64, 14) + _loader(2, 8, 64, 15)
    runs = []
    for graph in (False, True):
        m = MMJointDualHead("tf_efficientnet_b0_ns", tab_in=9, tab_hidden=64, drop=0.0).cuda()
        m.load_state_dict(sd, strict=True)
        tr = DualTaskTrainer(m, t_max=50, graph=graph)
        losses, tickets = [], []
        for b in batches:
            losses.append(float(tr.step(*b[:5])))
            tickets.append(tr.loss_async())
        assert [tr.loss_value(t) for t in tickets] == losses          # pinned-ring read-back returns each step's own loss
        runs.append((losses, torch.cat([p.detach().flatten() for p in m.parameters()]).cpu()))
    (l0, p0), (l1, p1) = runs
    # same trajectory: tight while the runs are still close, loose once atomics-order noise has been amplified by 13
    # tiny-batch BatchNorm steps.  (Steps 0 and 1 are eager in BOTH runs and already differ by up to 3e-3 - two eager runs of
    # a B0 at 64 px, whose last stages normalise 32 values per channel, are that far apart (fp32 atomics order -> a bf16 rounding flips -> 32-sample batch statistics amplify it); up to
    # 1.2e-2 measured at step 2.)
    assert abs(l0[0] - l1[0]) < 1e-2, (l0[0], l1[0])       # same weights, same batch, both eager: 3e-4 .. 3e-3 observed
    assert max(abs(a - b) for a, b in zip(l0[:5], l1[:5])) < 2.5e-2 and max(abs(a - b) for a, b in zip(l0, l1)) < 6e-2, (l0, l1)
    # AdamW moves a weight by up to lr per step whatever the gradient's size, so single weights whose gradient is pure
    # atomics-order noise differ by up to 2*lr*steps; the bulk must agree far better than that
    assert (p0 - p1).abs().mean() < 5e-4 and (p0 - p1).abs().max() <= 2 * 3e-4 * len(batches)
    with pytest.raises(ValueError):
        tr.step(*[t[:1] for t in batches[0][:5]])
