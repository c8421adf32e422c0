"""Mint tests/golden/stack_golden.json from the REFERENCE's own code (run in the build container only).

    python tests/golden/make_stack_golden.py

Drives experiments/fusion_v1/stack_blend.py's choose_threshold / _metrics and ui/gradio_app/stack_meta.py's Stacker
(imported unchanged) on the seeded stream frames of oracle.ref_stack.stream_frames(); main()'s merge -> LogisticRegression ->
threshold -> metrics steps (:224-262) are spelled out because they are inline in main().
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_stack as RS  # noqa: E402
from conftest import load_reference_module  # noqa: E402


def main():
    sb = load_reference_module("experiments/fusion_v1/stack_blend.py", "ref_stack_blend")
    sm = load_reference_module("ui/gradio_app/stack_meta.py", "ref_stack_meta")
    fr = RS.stream_frames()
    out = {"blend": {}, "stacker": {}}
    for use_mil in (False, True):
        oof = fr["tab_oof"].rename(columns={'prob': 'prob_tab'}).merge(fr["mm_oof"].rename(columns={'prob': 'prob_mm'}), on=['image_name', 'y'], how='inner')
        test = fr["tab_test"].rename(columns={'prob': 'prob_tab'}).merge(fr["mm_test"].rename(columns={'prob': 'prob_mm'}), on=['image_name', 'y'], how='inner')
        if use_mil:
            oof = oof.merge(fr["mil_oof"].rename(columns={'prob': 'prob_mil'}), on=['image_name', 'y'], how='inner')
            test = test.merge(fr["mil_test"].rename(columns={'prob': 'prob_mil'}), on=['image_name', 'y'], how='inner')
        cols = ['prob_tab', 'prob_mm'] + (['prob_mil'] if use_mil else [])
        meta = sb.LogisticRegression(max_iter=1000)
        meta.fit(oof[cols].values, oof['y'].values)
        p_oof = meta.predict_proba(oof[cols].values)[:, 1]
        p_te = meta.predict_proba(test[cols].values)[:, 1]
        rec = {"n_oof": len(oof), "n_test": len(test), "coef": meta.coef_[0].tolist(), "intercept": float(meta.intercept_[0]),
               "p_oof_sum": float(p_oof.sum()), "p_te_sum": float(p_te.sum()), "modes": {}}
        for mode in RS.MODES:
            for target in ((0.8, 0.9) if mode.startswith("target") else (0.8,)):
                thr = sb.choose_threshold(oof['y'].values, p_oof, mode=mode, target=target)
                rec["modes"][f"{mode}@{target}"] = {"thr": thr, "oof": sb._metrics(oof['y'].values, p_oof, thr),
                                                    "test": sb._metrics(test['y'].values, p_te, thr)}
        out["blend"]["mil" if use_mil else "no_mil"] = rec
    with tempfile.TemporaryDirectory() as d:
        paths = {}
        for k in ("mm_oof", "mm_test", "mil_oof", "mil_test"):
            paths[k] = os.path.join(d, k + ".csv")
            fr[k].to_csv(paths[k], index=False)
        for mode in RS.MODES:
            st = sm.Stacker(os.path.join(d, "tab.xlsx"), paths["mm_oof"], paths["mm_test"], paths["mil_oof"], paths["mil_test"],
                            thr_mode=mode, thr_target=0.8)
            a = st.predict_single(0.71, 0.64, None)
            b = st.predict_single(0.31, 0.44, 0.9)
            out["stacker"][mode] = {"thr_img": st.thr_img, "img_only": [a[0], a[1], a[2]], "hybrid": [b[0], b[1], b[2]],
                                    "coef": st.meta_img.coef_[0].tolist(), "intercept": float(st.meta_img.intercept_[0])}
        st.set_threshold_mode('max_f1', 0.8)
        out["stacker"]["switched_to_max_f1"] = st.thr_img
    with open(os.path.join(HERE, "stack_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["stacker"], indent=1)[:1500])


if __name__ == "__main__":
    main()
