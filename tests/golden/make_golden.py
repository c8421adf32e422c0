"""Mint the golden fixtures from the REFERENCE's own code (run in the build container only).

    python tests/golden/make_golden.py

* preproc_golden.json — sha1 of `src.preprocessing.normalise.apply_clahe` and
  `src.preprocessing.pipeline.centre_crop_resize` outputs (imported from /root/reference) on the
  seeded SURVEY.md §8d image set.
* models_golden.pt — outputs of the reference's model classes (train_mm_joint_dualtask.MMJointDualHead,
  infer_mm.MMNet, train_mil_attention_v1.MILNet, infer_mil.MILNet; imported unchanged on top of the
  oracle's timm shim) on seeded weights/inputs.  Weights are rebuilt from the seed by
  `oracle.ref_models.seeded_model`, so only inputs' seeds and the outputs are stored.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

import ref_models as R  # noqa: E402
import ref_preproc as P  # noqa: E402
from conftest import load_reference_module  # noqa: E402

PREPROC_CASES = [("noise", 1024, 1024), ("smooth", 1024, 1024), ("radiograph", 1024, 1024), ("const0", 512, 512),
                 ("const128", 512, 512), ("const255", 512, 512), ("ramp", 1024, 1024), ("noise", 1000, 1003),
                 ("radiograph", 480, 640), ("smooth", 777, 1024)]
RESIZE_SIZES = [224, 512]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_preproc():
    from src.preprocessing.normalise import apply_clahe
    from src.preprocessing.pipeline import centre_crop_resize
    out = {"cv2": __import__("cv2").__version__, "cases": []}
    for name, h, w in PREPROC_CASES:
        img = P.image_set(name, h, w)
        cl = apply_clahe(img)
        rec = {"name": name, "h": h, "w": w, "input": sha(img), "clahe": sha(cl)}
        for s in RESIZE_SIZES:
            rec[f"clahe_resize{s}"] = sha(centre_crop_resize(cl, s))
            rec[f"resize{s}"] = sha(centre_crop_resize(img, s))
        out["cases"].append(rec)
    out["tables"] = {k: P.table_sha(v) for k, v in P.tables().items()}
    return out


def mm_inputs(B, img, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, img, img, generator=g)
    xt = torch.randn(B, 9, generator=g)
    yh = (torch.rand(B, generator=g) < 0.6).float()
    ys = (yh * 0.8 + 0.2 * torch.rand(B, generator=g)).clamp(0, 1)
    return x, xt, yh, ys


def golden_models():
    ref_mm = load_reference_module("experiments/multimodal_v1/train_mm_joint_dualtask.py", "ref_mm")
    ref_imm = load_reference_module("ui/gradio_app/infer_mm.py", "ref_imm")
    ref_mil = load_reference_module("experiments/vision_v2/train_mil_attention_v1.py", "ref_mil")
    ref_imil = load_reference_module("ui/gradio_app/infer_mil.py", "ref_imil")
    torch.set_num_threads(os.cpu_count())
    out = {}

    # --- MM, B4, config 0: fwd batch 8 @224 eval (BASELINE.json configs[0]) + TTA
    sd = R.seeded_model("mm", seed=0, warm=2, img=64).state_dict()
    m = ref_mm.MMJointDualHead().eval()
    m.load_state_dict(sd, strict=True)
    x, xt, yh, ys = mm_inputs(8, 224, 100)
    with torch.no_grad():
        logit, reg = m(x, xt)
    out["mm_b4_fwd224"] = dict(logit=logit, reg=reg)
    twin = ref_imm.MMNet().eval()
    twin.load_state_dict(sd, strict=True)
    with torch.no_grad():
        tl = []
        for dims in (None, [3], [2]):
            xi = x[:2].clone() if dims is None else torch.flip(x[:2], dims=dims)
            tl.append(twin(xi, xt[:2])[0])
        tta = torch.stack(tl, 0).mean(0)
    out["mm_b4_tta224"] = dict(logit=tta, prob_T2p5=torch.sigmoid(tta / 2.5))

    # --- MM, B0 backbone, small image: 3 train steps of the reference loop (:241-256) in fp32, dropout 0
    def train_case(backbone, B, img, steps, key):
        sd0 = R.seeded_model("mm", seed=1, warm=1, img=img, backbone=backbone, drop=0.0).state_dict()
        tm = ref_mm.MMJointDualHead(backbone=backbone, drop=0.0)
        tm.load_state_dict(sd0, strict=True)
        tm.train()
        opt = torch.optim.AdamW(tm.parameters(), lr=3e-4, weight_decay=1e-4)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
        losses, gns = [], []
        for s in range(steps):
            xb, xtb, yhb, ysb = mm_inputs(B, img, 200 + s)
            opt.zero_grad(set_to_none=True)
            lg, rg = tm(xb, xtb)
            loss = 1.0 * ref_mm.bce_logits_with_soft_targets(lg, yhb) + 0.3 * ref_mm.bce_logits_with_soft_targets(rg, ysb)
            loss.backward()
            gn = torch.nn.utils.clip_grad_norm_(tm.parameters(), 1.0)
            opt.step()
            sched.step()
            losses.append(float(loss))
            gns.append(float(gn))
        tm.eval()
        xb, xtb, _, _ = mm_inputs(B, img, 299)
        with torch.no_grad():
            lg, rg = tm(xb, xtb)
        out[key] = dict(losses=torch.tensor(losses), grad_norms=torch.tensor(gns), logit_after=lg, reg_after=rg,
                        bn1_running_mean=tm.backbone.bn1.running_mean.clone(),
                        tab_bn_running_var=tm.tab[1].running_var.clone())
    train_case("tf_efficientnet_b0_ns", 8, 64, 3, "mm_b0_train64")
    train_case("tf_efficientnet_b4_ns", 4, 96, 2, "mm_b4_train96")

    # --- MIL train-side module, K=16 instances
    sdm = R.seeded_model("mil", seed=2, warm=1, img=64).state_dict()
    mm_ = ref_mil.MILNet(drop=0.0).eval()
    mm_.load_state_dict(sdm, strict=True)
    g = torch.Generator().manual_seed(300)
    bags = torch.randn(2, 16, 3, 96, 96, generator=g)
    with torch.no_grad():
        lg, A = mm_(bags)
    out["mil_b0_fwd96"] = dict(logit=lg, A=A)
    # pooling alone (a6)
    H = torch.randn(6, 16, 1280, generator=g)
    with torch.no_grad():
        M, A = mm_.mil(H)
    out["mil_pool"] = dict(M=M, A=A)
    # --- MIL inference twin (hid 256), one bag
    sdt = R.seeded_model("mil_twin", seed=3, warm=1, img=64).state_dict()
    tw = ref_imil.MILNet().eval()
    tw.load_state_dict(sdt, strict=True)
    with torch.no_grad():
        out["mil_twin_fwd96"] = dict(logit=tw(bags[0]))
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "preproc_golden.json"), "w") as f:
        json.dump(golden_preproc(), f, indent=1)
    torch.save(golden_models(), os.path.join(HERE, "models_golden.pt"))
    print("golden fixtures written")
