"""ORACLE (test infrastructure — only tests/, __graft_entry__.smoke() and bench.py's CPU arms import it).

CPU fp32 restatement of the reference's hot-path modules, each citing the reference file:line it
follows.  Parity status: the reference's own tests hold no golden vector for this path
(tests/test_pipeline.py:1-11 only) -> "parity unpinned by reference tests"; instead these
restatements are pinned in this container against the *imported* reference classes
(tests/test_oracle_models.py, which loads /root/reference with the timm shim on sys.path) and the
outputs are frozen as fixtures under tests/golden/ by tests/golden/make_golden.py.
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import timm  # noqa: E402  (the shim in oracle/timm)

TAB_FEATURES = ['depth', 'width', 'enamel_cracks', 'occlusal_load', 'carious_lesion',
                'opposing_type', 'adjacent_teeth', 'age_range', 'cervical_lesion']


class MMJointDualHead(nn.Module):
    """experiments/multimodal_v1/train_mm_joint_dualtask.py:135-160 (twins: ui/gradio_app/infer_mm.py:19-39,
    finalize_mm_dualtask_from_ckpts.py:48-61 — identical state-dict keys, the Dropout-only fusion has no params)."""

    def __init__(self, backbone='tf_efficientnet_b4_ns', tab_in=9, tab_hidden=64, drop=0.2):
        super().__init__()
        self.backbone = timm.create_model(backbone, pretrained=False, num_classes=0, global_pool='avg')
        d = self.backbone.num_features
        self.tab = nn.Sequential(
            nn.Linear(tab_in, tab_hidden), nn.BatchNorm1d(tab_hidden), nn.ReLU(inplace=True),
            nn.Dropout(p=drop), nn.Linear(tab_hidden, tab_hidden), nn.ReLU(inplace=True))
        self.fusion = nn.Sequential(nn.Dropout(p=drop))
        self.cls_head = nn.Linear(d + tab_hidden, 1)
        self.reg_head = nn.Linear(d + tab_hidden, 1)

    def forward(self, x_img, x_tab):
        f = torch.cat([self.backbone(x_img), self.tab(x_tab)], dim=1)
        f = self.fusion(f)
        return self.cls_head(f).squeeze(1), self.reg_head(f).squeeze(1)


def dual_bce_loss(logit, reg, y_hard, y_soft, alpha=1.0, beta=0.3, weight=None):
    """train_mm_joint_dualtask.py:176-179 + :244-247 (pos_weight is parsed at :224-226 but never applied)."""
    lh = F.binary_cross_entropy_with_logits(logit, y_hard, weight=weight, reduction='mean')
    ls = F.binary_cross_entropy_with_logits(reg, y_soft, weight=weight, reduction='mean')
    return alpha * lh + beta * ls


def mm_train_step(model, opt, sched, x_img, x_tab, y_h, y_s, w=None, alpha=1.0, beta=0.3, grad_clip=1.0):
    """train_mm_joint_dualtask.py:241-256 in fp32 (the GradScaler/autocast wrapper is the identity in fp32).
    Returns (loss, pre-clip total grad norm)."""
    opt.zero_grad(set_to_none=True)
    logit, reg = model(x_img, x_tab)
    loss = dual_bce_loss(logit, reg, y_h, y_s, alpha, beta, w)
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip) if grad_clip > 0 else torch.zeros(())
    opt.step()
    if sched is not None:
        sched.step()
    return float(loss.item()), float(gn)


def make_optimizer(model, lr=3e-4, weight_decay=1e-4, t_max=None):
    """train_mm_joint_dualtask.py:217-220 (one param group; cosine stepped per iteration)."""
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=t_max) if t_max else None
    return opt, sched


@torch.no_grad()
def mm_tta_logit(model, x_img, x_tab):
    """train_mm_joint_dualtask.py:326-335 / infer_mm.py:99-105: mean logit over {identity, W-flip, H-flip}."""
    outs = []
    for dims in (None, [3], [2]):
        xi = x_img.clone() if dims is None else torch.flip(x_img, dims=dims)
        outs.append(model(xi, x_tab)[0])
    return torch.stack(outs, 0).mean(0)


def prep_tab(tab_dict, mean, scale):
    """infer_mm.py:75-83: z-score with scale==0 -> 1; None -> the mean itself (standardised zeros)."""
    mean = np.asarray(mean, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    x = mean.copy() if tab_dict is None else np.array([float(tab_dict[k]) for k in TAB_FEATURES], dtype=np.float32)
    z = (x - mean) / np.where(scale == 0, 1.0, scale)
    return torch.tensor(z, dtype=torch.float32).unsqueeze(0)


@torch.no_grad()
def mm_ensemble_prob(models_T, scales, x_img, tab_dict):
    """infer_mm.py:93-108: per fold sigmoid(TTA-mean logit / T_fold), then the mean over folds."""
    probs = []
    for f, (model, T) in enumerate(models_T):
        xt = prep_tab(tab_dict, *scales[f])
        logit = mm_tta_logit(model, x_img, xt)
        probs.append(torch.sigmoid(logit / T).item())
    return float(np.mean(probs)), probs


class AttentionMIL(nn.Module):
    """experiments/vision_v2/train_mil_attention_v1.py:117-130 (gated attention, Ilse et al. 2018)."""

    def __init__(self, in_dim, hid=128):
        super().__init__()
        self.attention_V = nn.Linear(in_dim, hid)
        self.attention_U = nn.Linear(in_dim, hid)
        self.attention_w = nn.Linear(hid, 1)

    def forward(self, H):
        g = torch.tanh(self.attention_V(H)) * torch.sigmoid(self.attention_U(H))
        a = torch.softmax(self.attention_w(g).squeeze(-1), dim=1)
        return torch.einsum('bkd,bk->bd', H, a), a


class MILNet(nn.Module):
    """experiments/vision_v2/train_mil_attention_v1.py:132-148."""

    def __init__(self, backbone='tf_efficientnet_b0_ns', drop=0.2, hid=128):
        super().__init__()
        self.encoder = timm.create_model(backbone, pretrained=False, num_classes=0, global_pool='avg')
        d = self.encoder.num_features
        self.mil = AttentionMIL(d, hid=hid)
        self.drop = nn.Dropout(p=drop)
        self.head = nn.Linear(d, 1)

    def forward(self, x):
        B, K = x.shape[:2]
        feats = self.encoder(x.view(B * K, *x.shape[2:])).view(B, K, -1)
        bag, A = self.mil(feats)
        return self.head(self.drop(bag)).squeeze(1), A


class MILAttentionTwin(nn.Module):
    """ui/gradio_app/infer_mil.py:54-68 (single bag [N,D], softmax over dim 0, hid 256)."""

    def __init__(self, in_dim, hid_dim=256):
        super().__init__()
        self.U = nn.Linear(in_dim, hid_dim)
        self.V = nn.Linear(in_dim, hid_dim)
        self.w = nn.Linear(hid_dim, 1)

    def forward(self, H):
        g = torch.tanh(self.V(H)) * torch.sigmoid(self.U(H))
        alpha = torch.softmax(self.w(g).squeeze(-1), dim=0)
        return (alpha.unsqueeze(-1) * H).sum(0), alpha


class MILNetTwin(nn.Module):
    """ui/gradio_app/infer_mil.py:71-96 (global_pool='' + explicit GAP; one bag -> scalar logit)."""

    def __init__(self, backbone='tf_efficientnet_b0_ns', hid_dim=256):
        super().__init__()
        self.enc = timm.create_model(backbone, pretrained=False, num_classes=0, global_pool='')
        d = self.enc.num_features
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.mil = MILAttentionTwin(d, hid_dim)
        self.head = nn.Linear(d, 1)

    def forward(self, x):
        feats = self.enc(x)
        if feats.ndim == 4:
            feats = self.gap(feats).flatten(1)
        M, _ = self.mil(feats)
        return self.head(M).squeeze(-1)


def remap_mil_keys(sd):
    """ui/gradio_app/infer_mil.py:17-34: encoder.->enc., mil.attention_X.->mil.X."""
    out = {}
    for k, v in sd.items():
        if k.startswith("encoder."):
            k = "enc." + k[len("encoder."):]
        for n in "VUw":
            k = k.replace(f"mil.attention_{n}.", f"mil.{n}.")
        out[k] = v
    return out


def mil_train_step(model, opt, sched, bags, y, grad_clip=1.0):
    """experiments/vision_v2/train_mil_attention_v1.py:177-189 in fp32."""
    opt.zero_grad(set_to_none=True)
    logit, _ = model(bags)
    loss = F.binary_cross_entropy_with_logits(logit, y)
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip) if grad_clip > 0 else torch.zeros(())
    opt.step()
    if sched is not None:
        sched.step()
    return float(loss.item()), float(gn)


def seeded_model(kind, seed=0, warm=2, img=64, **kw):
    """Deterministic weights for fixtures: seed -> init -> randomised BN affine -> `warm` train-mode forwards
    on seeded noise so the BN running statistics are non-trivial (SURVEY.md §8c golden recipe).
    CPU torch RNG is reproducible for a fixed torch build, so the GPU box rebuilds identical weights."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = {"mm": MMJointDualHead, "mil": MILNet, "mil_twin": MILNetTwin}[kind](**kw)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            if isinstance(m, nn.Conv2d) and m.bias is not None:
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    model.train()
    with torch.no_grad():
        for _ in range(warm):
            if kind == "mm":
                for p in (m for m in model.modules() if isinstance(m, nn.Dropout)):
                    p.p_saved, p.p = p.p, 0.0
                model(torch.randn(4, 3, img, img, generator=g), torch.randn(4, 9, generator=g))
                for p in (m for m in model.modules() if isinstance(m, nn.Dropout)):
                    p.p = p.p_saved
            elif kind == "mil":
                model.encoder(torch.randn(4, 3, img, img, generator=g))
            else:
                model.enc(torch.randn(4, 3, img, img, generator=g))
    return model.eval()
