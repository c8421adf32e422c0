"""The train-time image transform of the joint dual-task trainer on the device (stochastic half):
what `timm.data.create_transform(input_size, is_training=True, auto_augment='rand-m9-mstd0.5-inc1', interpolation='bicubic',
re_prob=0.2, re_mode='pixel', re_count=1, mean, std)` builds at experiments/multimodal_v1/train_mm_joint_dualtask.py:75-84:

    RandomResizedCropAndInterpolation(scale (0.08, 1), ratio (3/4, 4/3), bicubic) -> RandomHorizontalFlip(0.5)
    -> RandAugment(2 ops of the 15 'increasing' ops, magnitude N(9, 0.5) clipped to [0, 10], p = 0.5 each, bicubic,
       fill = round(255 * mean)) -> ToTensor -> Normalize -> RandomErasing(p 0.2, per-pixel normal noise, one box)

Two layers with different parity status:
  * the IMAGE OPERATIONS (functions below with timm's names) are libteethrt kernels that reproduce Pillow bit for bit — the
    same per-pixel code is pinned to Pillow on the CPU (tests/test_augment_cpu.py) and on the GPU (tests/test_augment_gpu.py);
  * the SAMPLING (which ops, magnitudes, signs, crop boxes, erase boxes) restates timm's policy from its published
    algorithm; timm itself is absent from this image, so draw-for-draw parity with timm's use of `random` / `numpy.random` is
    UNPINNED — the distributions are the documented ones, the streams may differ.
Images are uint8 RGB [H,W,3] CUDA tensors (numpy arrays are uploaded); everything stays on the device."""
import math
import random

import numpy as np
import torch

from . import ops
from .preproc import _as_u8_image, normalize_flip, resized_crop

IMG_MEAN_FILL = tuple(min(255, round(255 * x)) for x in (0.485, 0.456, 0.406))      # aa_params['img_mean'] = (124, 116, 104)
_LEVEL_DENOM = 10.0


def _img(img):
    t = _as_u8_image(img)
    if t.shape[2] != 3:
        raise ValueError("RGB image [H,W,3] expected")
    return t


def _lut(img, table):
    lut = torch.as_tensor(np.tile(np.asarray(table, dtype=np.uint8), 3), device=img.device)
    return ops.lut_apply_u8(img, lut)


# ------------------------------------------------------------------------------------------------ the 15 operations
def auto_contrast(img, **_):
    img = _img(img)
    return ops.lut_apply_u8(img, ops.hist_lut_u8(img, 0))


def equalize(img, **_):
    img = _img(img)
    return ops.lut_apply_u8(img, ops.hist_lut_u8(img, 1))


def point_table(name, arg=None, thresh=128):
    """256-entry lookup table of the point operations (host side; the same table Pillow's ImageOps builds)."""
    i = np.arange(256)
    if name == "invert":
        return (255 - i).astype(np.uint8)
    if name == "posterize":
        return (i & (~(2 ** (8 - arg) - 1) & 0xff)).astype(np.uint8)
    if name == "solarize":
        return np.where(i < arg, i, 255 - i).astype(np.uint8)
    if name == "solarize_add":
        return np.where(i < thresh, np.minimum(255, i + arg), i).astype(np.uint8)
    raise KeyError(name)


def invert(img, **_):
    return _lut(_img(img), point_table("invert"))


def posterize(img, bits_to_keep, **_):
    img = _img(img)
    if bits_to_keep >= 8:
        return img
    return _lut(img, point_table("posterize", bits_to_keep))


def solarize(img, thresh, **_):
    return _lut(_img(img), point_table("solarize", thresh))


def solarize_add(img, add, thresh=128, **_):
    return _lut(_img(img), point_table("solarize_add", add, thresh))


def brightness(img, factor, **_):
    return ops.enhance_rgb_u8(_img(img), 0, factor)


def color(img, factor, **_):
    return ops.enhance_rgb_u8(_img(img), 1, factor)


def contrast(img, factor, **_):
    return ops.enhance_rgb_u8(_img(img), 2, factor)


def sharpness(img, factor, **_):
    return ops.enhance_rgb_u8(_img(img), 3, factor)


def _affine(img, matrix, resample="bicubic", fillcolor=IMG_MEAN_FILL):
    return ops.affine_pil_u8(_img(img), matrix, resample == "bicubic", fillcolor)


def shear_x(img, factor, **kw):
    return _affine(img, (1, factor, 0, 0, 1, 0), **kw)


def shear_y(img, factor, **kw):
    return _affine(img, (1, 0, 0, factor, 1, 0), **kw)


def translate_x_rel(img, pct, **kw):
    return _affine(img, (1, 0, pct * _img(img).shape[1], 0, 1, 0), **kw)


def translate_y_rel(img, pct, **kw):
    return _affine(img, (1, 0, 0, 0, 1, pct * _img(img).shape[0]), **kw)


def rotation_matrix(w, h, degrees):
    """PIL.Image.rotate's inverse matrix (centre = (w/2, h/2), no expand), or None for the 0 / 180 / square-90 fast paths."""
    angle = degrees % 360.0
    if angle in (0, 180) or (angle in (90, 270) and w == h):
        return None
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2, h / 2
    m[2], m[5] = m[0] * -cx + m[1] * -cy + m[2], m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def rotate(img, degrees, **kw):
    img = _img(img)
    m = rotation_matrix(img.shape[1], img.shape[0], degrees)
    if m is None:                                  # Pillow's transpose fast paths: exact pixel moves
        k = int(round((degrees % 360.0) / 90.0))
        return img if k == 0 else torch.rot90(img, k, dims=(0, 1)).contiguous()
    return _affine(img, m, **kw)


# ------------------------------------------------------------------------------------------------ magnitude -> argument
def _negate(v, rng):
    return -v if rng.random() > 0.5 else v


def _level_args(name, level, rng):
    if name == "Rotate":
        return (_negate((level / _LEVEL_DENOM) * 30.0, rng),)
    if name in ("ColorIncreasing", "ContrastIncreasing", "BrightnessIncreasing", "SharpnessIncreasing"):
        return (max(0.1, 1.0 + _negate((level / _LEVEL_DENOM) * 0.9, rng)),)
    if name in ("ShearX", "ShearY"):
        return (_negate((level / _LEVEL_DENOM) * 0.3, rng),)
    if name in ("TranslateXRel", "TranslateYRel"):
        return (_negate((level / _LEVEL_DENOM) * 0.45, rng),)
    if name == "PosterizeIncreasing":
        return (4 - int((level / _LEVEL_DENOM) * 4),)
    if name == "SolarizeIncreasing":
        return (256 - int((level / _LEVEL_DENOM) * 256),)
    if name == "SolarizeAdd":
        return (min(128, int((level / _LEVEL_DENOM) * 110)),)
    return ()


RAND_INCREASING_TRANSFORMS = ["AutoContrast", "Equalize", "Invert", "Rotate", "PosterizeIncreasing", "SolarizeIncreasing",
                              "SolarizeAdd", "ColorIncreasing", "ContrastIncreasing", "BrightnessIncreasing",
                              "SharpnessIncreasing", "ShearX", "ShearY", "TranslateXRel", "TranslateYRel"]
_OP_FN = {"AutoContrast": auto_contrast, "Equalize": equalize, "Invert": invert, "Rotate": rotate,
          "PosterizeIncreasing": posterize, "SolarizeIncreasing": solarize, "SolarizeAdd": solarize_add, "ColorIncreasing": color,
          "ContrastIncreasing": contrast, "BrightnessIncreasing": brightness, "SharpnessIncreasing": sharpness, "ShearX": shear_x,
          "ShearY": shear_y, "TranslateXRel": translate_x_rel, "TranslateYRel": translate_y_rel}


class RandAugment:
    """'rand-m9-mstd0.5-inc1': num_layers ops drawn with replacement, each applied with probability 0.5 at a magnitude drawn
    from N(9, 0.5) clipped to [0, 10]."""

    def __init__(self, magnitude=9.0, magnitude_std=0.5, num_layers=2, prob=0.5, resample="bicubic", fillcolor=IMG_MEAN_FILL, seed=None):
        self.magnitude, self.magnitude_std, self.num_layers, self.prob = magnitude, magnitude_std, num_layers, prob
        self.kw = dict(resample=resample, fillcolor=fillcolor)
        self.rng = random.Random(seed)
        self.np_rng = np.random.RandomState(seed)
        self.last = []                                          # (op, args) actually applied to the last image

    def __call__(self, img):
        img = _img(img)
        self.last = []
        for name in self.np_rng.choice(RAND_INCREASING_TRANSFORMS, self.num_layers, replace=True):
            if self.prob < 1.0 and self.rng.random() > self.prob:
                continue
            m = self.magnitude
            if self.magnitude_std > 0:
                m = self.rng.gauss(m, self.magnitude_std)
            m = max(0.0, min(m, _LEVEL_DENOM))
            args = _level_args(str(name), m, self.rng)
            self.last.append((str(name), args))
            img = _OP_FN[str(name)](img, *args, **self.kw)
        return img


def random_resized_crop_box(h, w, rng, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """RandomResizedCropAndInterpolation.get_params: 10 attempts at a log-uniform aspect ratio, then the central fallback."""
    area = h * w
    for _ in range(10):
        target_area = rng.uniform(*scale) * area
        aspect = math.exp(rng.uniform(math.log(ratio[0]), math.log(ratio[1])))
        cw, ch = int(round(math.sqrt(target_area * aspect))), int(round(math.sqrt(target_area / aspect)))
        if cw <= w and ch <= h:
            return rng.randint(0, h - ch), rng.randint(0, w - cw), ch, cw
    in_ratio = w / h
    if in_ratio < min(ratio):
        cw, ch = w, int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        ch, cw = h, int(round(h * max(ratio)))
    else:
        cw, ch = w, h
    return (h - ch) // 2, (w - cw) // 2, ch, cw


class TrainTransform:
    """The whole training transform: uint8 RGB image (any size) -> normalised fp32 / bf16 [3, S, S] CUDA tensor."""

    def __init__(self, img_size, re_prob=0.2, dtype=torch.float32, seed=None):
        self.size, self.re_prob, self.dtype = int(img_size), re_prob, dtype
        self.rng = random.Random(seed)
        self.aug = RandAugment(seed=None if seed is None else seed + 1)
        self.gen = None
        self.seed = seed

    def __call__(self, img):
        img = _img(img)
        top, left, ch, cw = random_resized_crop_box(img.shape[0], img.shape[1], self.rng)
        x = resized_crop(img, top, left, ch, cw, self.size, "bicubic")
        if self.rng.random() < 0.5:                                   # RandomHorizontalFlip precedes RandAugment
            x = torch.flip(x, dims=[1]).contiguous()
        x = self.aug(x)
        # ToTensor + Normalize: the kernel takes BGR HWC and writes RGB CHW, so hand it the channel-reversed view
        out = normalize_flip(torch.flip(x, dims=[2]).contiguous(), 0, dtype=self.dtype)
        return self._erase(out)

    def _erase(self, x):
        """RandomErasing(probability, mode='pixel', min_area 0.02, max_area 1/3, min_aspect 0.3, one box, 10 attempts)."""
        if self.rng.random() > self.re_prob:
            return x
        c, h, w = x.shape
        area = h * w
        for _ in range(10):
            target = self.rng.uniform(0.02, 1 / 3) * area
            aspect = math.exp(self.rng.uniform(math.log(0.3), math.log(1 / 0.3)))
            eh, ew = int(round(math.sqrt(target * aspect))), int(round(math.sqrt(target / aspect)))
            if ew < w and eh < h:
                top, left = self.rng.randint(0, h - eh), self.rng.randint(0, w - ew)
                if self.gen is None:
                    self.gen = torch.Generator(device=x.device)
                    if self.seed is not None:
                        self.gen.manual_seed(self.seed + 2)
                x[:, top:top + eh, left:left + ew] = torch.empty((c, eh, ew), device=x.device, dtype=torch.float32).normal_(generator=self.gen).to(x.dtype)
                break
        return x
