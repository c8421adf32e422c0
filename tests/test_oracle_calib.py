"""CPU: pins the calibration oracle (oracle/ref_calib.py) to
  (1) the committed golden fixture minted from the reference's own TemperatureScaler / compute_metrics
      (tests/golden/make_calib_golden.py),
  (2) the reference functions themselves when /root/reference is present."""
import json
import math
import os
import warnings

import numpy as np
import pytest

import ref_calib as RC
from conftest import load_reference_module

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "calib_golden.json")))
CASES = RC.calib_cases()
warnings.filterwarnings("ignore", message="overflow encountered in exp")


def same_metrics(a, b):
    return all((math.isnan(a[k]) and math.isnan(b[k])) or a[k] == b[k] for k in ('auc', 'acc', 'prec', 'rec', 'f1'))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    z, y, _ = CASES[name]
    g = GOLD["cases"][name]
    cal = RC.calibrate_epoch(z, y)
    assert cal["T"] == pytest.approx(g["T"], rel=1e-6)
    assert cal["thr"] == g["thr"]
    assert same_metrics(cal["metrics"], g["metrics"])
    assert same_metrics(RC.compute_metrics(y, cal["probs"], 0.5), g["metrics_at_half"])
    assert float(np.sum(cal["probs"], dtype=np.float64)) == pytest.approx(g["prob_sum"], rel=1e-6)
    assert [RC.compute_metrics(y, cal["probs"], t)['f1'] for t in np.linspace(0.2, 0.8, 61)] == g["f1_curve"]


@pytest.mark.reference
def test_oracle_matches_reference_functions():
    ref = load_reference_module("experiments/multimodal_v1/train_mm_joint_dualtask.py", "ref_mm_calib")
    assert ref.fast_round(0.123456) == RC.fast_round(0.123456)
    for name, (z, y, _) in CASES.items():
        p = 1 / (1 + np.exp(-z))
        for thr in (0.2, 0.37, 0.5, 0.8):
            assert same_metrics(ref.compute_metrics(y, p, thr), RC.compute_metrics(y, p, thr)), (name, thr)
    s = ref.TemperatureScaler()
    assert list(s.state_dict()) == ["log_T"] and s.temperature() == 1.0


def test_first_maximum_of_rounded_f1_wins():
    # two thresholds with F1 equal after rounding: the earlier one is kept (strict '>' at :294)
    y = np.array([1, 1, 0, 0, 1, 0], np.float32)
    p = np.array([0.9, 0.85, 0.1, 0.15, 0.7, 0.75], np.float32)
    thr, m = RC.best_threshold(y, p)
    f1 = [RC.compute_metrics(y, p, t)['f1'] for t in np.linspace(0.2, 0.8, 61)]
    assert thr == float(np.linspace(0.2, 0.8, 61)[int(np.argmax(f1))]) and m['f1'] == max(f1)


def test_output_files(tmp_path):
    res = []
    for fold, name in enumerate(("typical", "small")):
        z, y, names = CASES[name]
        cal = RC.calibrate_epoch(z, y)
        res.append(RC.fold_result(fold, cal["thr"], cal["T"], (names, y, cal["probs"]), (names[:20], y[:20], cal["probs"][:20])))
    summary = RC.write_outputs(tmp_path, res)
    oof = open(tmp_path / "oof_val.csv").read().splitlines()
    assert oof[0] == "image_name,y,prob" and len(oof) == 1 + 613 + 37
    assert len(open(tmp_path / "pred_test.csv").read().splitlines()) == 41
    on_disk = json.load(open(tmp_path / "summary.json"))
    assert on_disk == summary and list(on_disk) == ["val_mean", "test_mean", "fold_details"]
    assert list(on_disk["fold_details"][0]) == ["fold", "auc", "acc", "prec", "rec", "f1", "test_auc", "test_acc", "test_prec",
                                                 "test_rec", "test_f1"]
