"""Host-side mirror of the reference's nn.Module interface for the hot path (same class names, constructor arguments,
forward signatures, state-dict keys and error behaviour), every forward/backward running on libteethrt kernels.

    MMJointDualHead   experiments/multimodal_v1/train_mm_joint_dualtask.py:135-160
    MMNet             ui/gradio_app/infer_mm.py:19-39                    (same keys; Dropout-only `fuse`)
    AttentionMIL      experiments/vision_v2/train_mil_attention_v1.py:117-130
    MILNet            experiments/vision_v2/train_mil_attention_v1.py:132-148
    MILAttention      ui/gradio_app/infer_mil.py:54-68
    MILNetTwin        ui/gradio_app/infer_mil.py:71-96  (the twin's `MILNet`)

The nn.Linear / nn.BatchNorm1d / nn.Dropout children only HOLD parameters under the reference's key names; their own
forward() is never called.
"""
import itertools

import torch
import torch.nn as nn

from . import ops
from .backbone import create_model

_seed_counter = itertools.count(1)

TAB_PARAM_KEYS = ["tab.0.weight", "tab.0.bias", "tab.1.weight", "tab.1.bias", "tab.4.weight", "tab.4.bias",
                  "cls_head.weight", "cls_head.bias", "reg_head.weight", "reg_head.bias"]


class _TabHeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, feat, xtab, *params):
        train = mod.training
        B = feat.shape[0]
        if train and B < 2:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(xtab.shape[:1]) + (mod.tab_hidden,)}")
        bn = mod.tab[1]
        feat, xtab = feat.contiguous().float(), xtab.contiguous().float()
        plist = [p.detach() for p in params]
        scratch = ops.tab_heads_scratch(B, mod.tab_hidden, feat.device)
        seed = next(_seed_counter) if (train and mod.drop_p > 0) else 0
        out = ops.tab_heads_fwd(feat, xtab, plist, bn.running_mean, bn.running_var, bn.num_batches_tracked if train else None,
                                scratch, train, mod.drop_p if train else 0.0, seed=seed)
        ctx.mod, ctx.saved = mod, (feat, xtab, plist, scratch, train, seed)
        return out["logit"], out["reg"]

    @staticmethod
    def backward(ctx, dlogit, dreg):
        mod = ctx.mod
        feat, xtab, plist, scratch, train, seed = ctx.saved
        bn = mod.tab[1]
        grads = [torch.empty_like(p) for p in plist]
        dfeat = torch.empty_like(feat)
        dl = torch.zeros_like(feat[:, 0]) if dlogit is None else dlogit.contiguous().float()
        dr = torch.zeros_like(feat[:, 0]) if dreg is None else dreg.contiguous().float()
        ops.tab_heads_bwd(feat, xtab, plist, bn.running_mean, bn.running_var, dl, dr, dfeat, grads, scratch, train,
                          mod.drop_p if train else 0.0, seed=seed)
        return (None, dfeat, None) + tuple(grads)


class MMJointDualHead(nn.Module):
    def __init__(self, backbone='tf_efficientnet_b4_ns', tab_in=9, tab_hidden=64, drop=0.2):
        super().__init__()
        self.backbone = create_model(backbone, pretrained=True, num_classes=0, global_pool='avg')
        feat_dim = self.backbone.num_features
        self.tab = nn.Sequential(nn.Linear(tab_in, tab_hidden), nn.BatchNorm1d(tab_hidden), nn.ReLU(inplace=True),
                                 nn.Dropout(p=drop), nn.Linear(tab_hidden, tab_hidden), nn.ReLU(inplace=True))
        self.fusion = nn.Sequential(nn.Dropout(p=drop))
        self.cls_head = nn.Linear(feat_dim + tab_hidden, 1)
        self.reg_head = nn.Linear(feat_dim + tab_hidden, 1)
        self.tab_in, self.tab_hidden, self.drop_p = tab_in, tab_hidden, float(drop)

    def tab_head_params(self):
        named = dict(self.named_parameters())
        return [named[k] for k in TAB_PARAM_KEYS]

    def forward(self, x_img, x_tab):
        f_img = self.backbone(x_img)                                   # (B, F) fp32
        return _TabHeadsFn.apply(self, f_img, x_tab, *self.tab_head_params())


class MMNet(MMJointDualHead):
    """Inference twin (ui/gradio_app/infer_mm.py:19-39): identical parameters; the Dropout-only child is called `fuse`."""

    def __init__(self, backbone='tf_efficientnet_b4_ns', tab_in=9, tab_hidden=64, drop=0.2):
        super().__init__(backbone, tab_in, tab_hidden, drop)
        self.fuse = self.fusion[0]
        del self.fusion


# ================================================================================================= MIL
class _MILAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, Vw, Vb, Uw, Ub, ww, wb):
        Hc = H.contiguous().float()
        ps = [t.detach().contiguous() for t in (Vw, Vb, Uw, Ub, ww, wb)]
        M, A, gV, gU = ops.mil_attn_fwd(Hc, ps[0], ps[1], ps[2], ps[3], ps[4].view(-1), ps[5], save=True)
        ctx.saved = (Hc, A, gV, gU, ps)
        ctx.mark_non_differentiable(A)
        return M, A

    @staticmethod
    def backward(ctx, dM, _dA):
        Hc, A, gV, gU, ps = ctx.saved
        grads = [torch.zeros_like(p) for p in ps]
        dH = ops.mil_attn_bwd(dM.contiguous().float(), Hc, A, gV, gU, ps[0], ps[2], ps[4].view(-1), *grads)
        return (dH,) + tuple(grads)


class _Linear1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, M, w, b, drop_p, seed):
        Mc, wc, bc = M.contiguous().float(), w.detach().contiguous(), b.detach().contiguous()
        ctx.saved = (Mc, wc, drop_p, seed)
        return ops.linear1_fwd(Mc, wc.view(-1), bc, drop_p, seed)

    @staticmethod
    def backward(ctx, dlogit):
        Mc, wc, drop_p, seed = ctx.saved
        dw, db = torch.empty_like(wc), torch.empty(1, device=Mc.device)
        dM = ops.linear1_bwd(dlogit.contiguous().float(), Mc, wc.view(-1), dw, db, drop_p, seed)
        return dM, dw, db, None, None


def _attn(H, V, U, w):
    return _MILAttnFn.apply(H, V.weight, V.bias, U.weight, U.bias, w.weight, w.bias)


class AttentionMIL(nn.Module):
    """Gated attention pooling over K instances: returns (M [B,D], A [B,K])."""

    def __init__(self, in_dim, hid=128):
        super().__init__()
        self.attention_V = nn.Linear(in_dim, hid)
        self.attention_U = nn.Linear(in_dim, hid)
        self.attention_w = nn.Linear(hid, 1)

    def forward(self, H):
        return _attn(H, self.attention_V, self.attention_U, self.attention_w)


class MILNet(nn.Module):
    def __init__(self, backbone='tf_efficientnet_b0_ns', drop=0.2, hid=128):
        super().__init__()
        self.encoder = create_model(backbone, pretrained=True, num_classes=0, global_pool='avg')
        d = self.encoder.num_features
        self.mil = AttentionMIL(d, hid=hid)
        self.drop = nn.Dropout(p=drop)
        self.head = nn.Linear(d, 1)
        self.drop_p = float(drop)

    def forward(self, x):                       # x: (B,K,C,H,W)
        B, K, C, H, W = x.shape
        feats = self.encoder(x.reshape(B * K, C, H, W)).view(B, K, -1)
        bag, A = self.mil(feats)
        p = self.drop_p if self.training else 0.0
        seed = next(_seed_counter) if p > 0 else 0
        logit = _Linear1Fn.apply(bag, self.head.weight, self.head.bias, p, seed)
        return logit, A


class MILAttention(nn.Module):
    """ui/gradio_app/infer_mil.py:54-68: one bag H [N,D] -> (M [D], alpha [N]); softmax over the instance axis."""

    def __init__(self, in_dim, hid_dim=256):
        super().__init__()
        self.U = nn.Linear(in_dim, hid_dim)
        self.V = nn.Linear(in_dim, hid_dim)
        self.w = nn.Linear(hid_dim, 1)

    def forward(self, H):
        M, A = _attn(H.unsqueeze(0), self.V, self.U, self.w)
        return M[0], A[0]


class MILNetTwin(nn.Module):
    """The inference twin's MILNet: forward(x [N,3,H,W]) -> scalar bag logit.  `hid_dim` is inferred from the checkpoint by
    MILEnsemble (the reference hard-codes 256 and then cannot load the trainer's 128-wide weights)."""

    def __init__(self, backbone="tf_efficientnet_b0_ns", pretrained=False, hid_dim=256):
        super().__init__()
        self.enc = create_model(backbone, pretrained=pretrained, num_classes=0, global_pool="")
        d = self.enc.num_features
        self.gap = nn.AdaptiveAvgPool2d(1)          # parameter-free; the pooling itself is fused into the encoder's tail
        self.mil = MILAttention(d, hid_dim=hid_dim)
        self.head = nn.Linear(d, 1)

    @torch.no_grad()
    def forward_instances(self, x):
        return self.enc.forward_pooled(x)           # [N, D]

    def forward(self, x):
        H = self.forward_instances(x)
        M, _ = self.mil(H)
        return _Linear1Fn.apply(M.unsqueeze(0), self.head.weight, self.head.bias, 0.0, 0)[0]
