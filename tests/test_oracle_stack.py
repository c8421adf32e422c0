"""CPU: pins the stacker oracle (oracle/ref_stack.py) to the golden fixture minted from the reference's
stack_blend.py / stack_meta.py (tests/golden/make_stack_golden.py) and to the reference functions when present."""
import json
import math
import os

import numpy as np
import pytest

import ref_stack as RS
from conftest import load_reference_module

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stack_golden.json")))
FR = RS.stream_frames()


def same(a, b):
    return all((math.isnan(a[k]) and math.isnan(b[k])) or a[k] == b[k] for k in a)


@pytest.mark.parametrize("use_mil", [False, True])
def test_oracle_blend_matches_golden(use_mil, tmp_path):
    g = GOLD["blend"]["mil" if use_mil else "no_mil"]
    for key, rec in g["modes"].items():
        mode, target = key.split("@")
        r = RS.stack_blend(FR["tab_oof"], FR["tab_test"], FR["mm_oof"], FR["mm_test"], tmp_path,
                           FR["mil_oof"] if use_mil else None, FR["mil_test"] if use_mil else None, mode, float(target))
        assert np.allclose(r["coef"], g["coef"], atol=1e-9) and abs(r["intercept"] - g["intercept"]) < 1e-9
        assert len(r["p_oof"]) == g["n_oof"] and len(r["p_te"]) == g["n_test"]
        assert r["thr"] == float(np.round(rec["thr"], 4)) and same(r["oof"], rec["oof"]) and same(r["test"], rec["test"])
    s = json.load(open(tmp_path / "summary.json"))
    assert list(s) == ["oof", "test", "thr", "thr_mode", "thr_target", "features"]
    assert open(tmp_path / "stack_oof.csv").readline().strip() == "image_name,y,prob"


@pytest.mark.reference
def test_oracle_matches_reference_functions():
    sb = load_reference_module("experiments/fusion_v1/stack_blend.py", "ref_stack_blend_t")
    rng = np.random.RandomState(3)
    y = (rng.rand(700) < 0.5).astype(int)
    p = np.clip(0.5 + 0.25 * (2 * y - 1) + 0.3 * rng.randn(700), 0, 1)
    p[::7] = np.round(p[::7], 1)
    for mode in RS.MODES + ["nonsense"]:
        for target in (0.6, 0.8, 0.999):
            assert sb.choose_threshold(y, p, mode, target) == RS.choose_threshold(y, p, mode, target), (mode, target)
    for thr in (0.2, 0.5, 0.77):
        assert same(sb._metrics(y, p, thr), RS.metrics(y, p, thr))
