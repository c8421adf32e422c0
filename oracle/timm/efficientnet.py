"""ORACLE: pure-PyTorch restatement of timm's `tf_efficientnet_b{0,4}_ns`.

State-dict keys follow timm (`conv_stem`, `bn1`, `blocks.{s}.{i}.{conv_pw,bn1,conv_dw,bn2,se.conv_reduce,
se.conv_expand,conv_pwl,bn3}`, `conv_head`, `bn2`); the reference keys on exactly these names at
src/vision/predict_vision.py:9-14,24,33-35.  TF-'same' padding, BN eps 1e-3 / momentum 0.1, SiLU,
SE reduction = round(0.25 * block input channels).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# (type, repeats, kernel, stride, expand, out_channels) of EfficientNet-B0; scaled per variant.
_B0_STAGES = [
    ("ds", 1, 3, 1, 1, 16),
    ("ir", 2, 3, 2, 6, 24),
    ("ir", 2, 5, 2, 6, 40),
    ("ir", 3, 3, 2, 6, 80),
    ("ir", 3, 5, 1, 6, 112),
    ("ir", 4, 5, 2, 6, 192),
    ("ir", 1, 3, 1, 6, 320),
]

# name -> (width multiplier, depth multiplier)
ARCHS = {
    "tf_efficientnet_b0_ns": (1.0, 1.0),
    "tf_efficientnet_b0": (1.0, 1.0),
    "tf_efficientnet_b4_ns": (1.4, 1.8),
    "tf_efficientnet_b4": (1.4, 1.8),
    "tf_efficientnet_b0.ns_jft_in1k": (1.0, 1.0),
    "tf_efficientnet_b4.ns_jft_in1k": (1.4, 1.8),
}

BN_EPS = 1e-3
BN_MOMENTUM = 0.1


def round_channels(c, mult, divisor=8):
    c = c * mult
    new_c = max(divisor, int(c + divisor / 2) // divisor * divisor)
    if new_c < 0.9 * c:
        new_c += divisor
    return new_c


def arch_spec(name):
    """Returns (stem, [(type, k, stride, in, out, mid, se_rd) per block grouped per stage], head_features)."""
    wm, dm = ARCHS[name]
    stem = round_channels(32, wm)
    stages = []
    cin = stem
    for (typ, r, k, s, e, c) in _B0_STAGES:
        cout = round_channels(c, wm)
        reps = int(math.ceil(r * dm))
        blocks = []
        for i in range(reps):
            stride = s if i == 0 else 1
            mid = cin * e
            rd = int(round(cin * 0.25))
            blocks.append(dict(type=typ, k=k, stride=stride, cin=cin, cout=cout, mid=mid, rd=rd))
            cin = cout
        stages.append(blocks)
    feat = round_channels(1280, wm)
    return stem, stages, feat


def same_pad(i, k, s):
    total = max((math.ceil(i / s) - 1) * s + k - i, 0)
    return total // 2, total - total // 2


class Conv2dSame(nn.Conv2d):
    """TF 'same' padding: symmetric for stride 1, dynamic asymmetric (extra pixel bottom/right) otherwise."""

    def forward(self, x):
        k, s = self.kernel_size[0], self.stride[0]
        if s == 1:
            return F.conv2d(x, self.weight, self.bias, 1, (k - 1) // 2, 1, self.groups)
        pt, pb = same_pad(x.shape[-2], k, s)
        pl, pr = same_pad(x.shape[-1], k, s)
        x = F.pad(x, (pl, pr, pt, pb))
        return F.conv2d(x, self.weight, self.bias, s, 0, 1, self.groups)


class BatchNormAct2d(nn.BatchNorm2d):
    def __init__(self, c, act=True):
        super().__init__(c, eps=BN_EPS, momentum=BN_MOMENTUM)
        self.apply_act = act

    def forward(self, x):
        x = super().forward(x)
        return F.silu(x) if self.apply_act else x


class SqueezeExcite(nn.Module):
    def __init__(self, c, rd):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, rd, 1, bias=True)
        self.conv_expand = nn.Conv2d(rd, c, 1, bias=True)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(F.silu(self.conv_reduce(s)))
        return x * torch.sigmoid(s)


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.has_skip = b["stride"] == 1 and b["cin"] == b["cout"]
        self.conv_dw = Conv2dSame(b["cin"], b["cin"], b["k"], b["stride"], groups=b["cin"], bias=False)
        self.bn1 = BatchNormAct2d(b["cin"], act=True)
        self.se = SqueezeExcite(b["cin"], b["rd"])
        self.conv_pw = nn.Conv2d(b["cin"], b["cout"], 1, bias=False)
        self.bn2 = BatchNormAct2d(b["cout"], act=False)

    def forward(self, x):
        y = self.bn1(self.conv_dw(x))
        y = self.se(y)
        y = self.bn2(self.conv_pw(y))
        return y + x if self.has_skip else y


class InvertedResidual(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.has_skip = b["stride"] == 1 and b["cin"] == b["cout"]
        self.conv_pw = nn.Conv2d(b["cin"], b["mid"], 1, bias=False)
        self.bn1 = BatchNormAct2d(b["mid"], act=True)
        self.conv_dw = Conv2dSame(b["mid"], b["mid"], b["k"], b["stride"], groups=b["mid"], bias=False)
        self.bn2 = BatchNormAct2d(b["mid"], act=True)
        self.se = SqueezeExcite(b["mid"], b["rd"])
        self.conv_pwl = nn.Conv2d(b["mid"], b["cout"], 1, bias=False)
        self.bn3 = BatchNormAct2d(b["cout"], act=False)

    def forward(self, x):
        y = self.bn1(self.conv_pw(x))
        y = self.bn2(self.conv_dw(y))
        y = self.se(y)
        y = self.bn3(self.conv_pwl(y))
        return y + x if self.has_skip else y


class EfficientNet(nn.Module):
    def __init__(self, name, global_pool="avg"):
        super().__init__()
        stem, stages, feat = arch_spec(name)
        self.num_features = feat
        self.global_pool_type = global_pool
        self.conv_stem = Conv2dSame(3, stem, 3, 2, bias=False)
        self.bn1 = BatchNormAct2d(stem, act=True)
        self.blocks = nn.Sequential(*[
            nn.Sequential(*[(DepthwiseSeparableConv if b["type"] == "ds" else InvertedResidual)(b) for b in st])
            for st in stages
        ])
        self.conv_head = nn.Conv2d(stages[-1][-1]["cout"], feat, 1, bias=False)
        self.bn2 = BatchNormAct2d(feat, act=True)
        self.classifier = nn.Identity()
        self._init_weights()

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan_out))
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward_features(self, x):
        x = self.bn1(self.conv_stem(x))
        x = self.blocks(x)
        return self.bn2(self.conv_head(x))

    def forward(self, x):
        x = self.forward_features(x)
        if self.global_pool_type == "avg":
            x = x.mean((2, 3))
        return self.classifier(x)


def create_model(model_name, pretrained=False, num_classes=0, global_pool="avg", **kwargs):
    """`pretrained` is ignored (no network): weights are whatever the current torch RNG gives."""
    if model_name not in ARCHS:
        raise RuntimeError(f"oracle timm shim: unknown model {model_name!r}")
    if num_classes != 0:
        raise RuntimeError("oracle timm shim: only num_classes=0 (feature extractor) is restated")
    return EfficientNet(model_name, global_pool=global_pool)
