"""ORACLE (test infrastructure — only tests/ import it).

The image operations of timm's RandAugment as timm defines them: thin wrappers over Pillow (timm/data/auto_augment.py in
timm 0.9/1.0; timm is absent from this image and un-pinned by the reference, train_mm_joint_dualtask.py:48-51, so the
wrappers are restated from its published source — they are one Pillow call each).  Pillow itself is present here and on
the GPU box, so it is the checker for the pixel arithmetic.  `*_np` are numpy restatements of Pillow's C arithmetic, the
spec multimodal-teeth-restoration-selection_b200/csrc/augment_core.h is written from; tests/test_augment_cpu.py pins both
to Pillow.  The SAMPLING side of the policy (random / numpy.random draws) has no checker here: parity unpinned.
"""
import math

import numpy as np
from PIL import Image, ImageEnhance, ImageOps

FILL = (124, 116, 104)          # tuple(min(255, round(255 * m)) for m in IMAGENET mean): aa_params['img_mean']
RESAMPLE = {"bicubic": Image.BICUBIC, "bilinear": Image.BILINEAR}


def _kw(resample="bicubic", fillcolor=FILL):
    return dict(resample=RESAMPLE[resample], fillcolor=tuple(fillcolor))


# ---- timm's op functions on PIL images
def shear_x(img, factor, **kw): return img.transform(img.size, Image.AFFINE, (1, factor, 0, 0, 1, 0), **_kw(**kw))  # noqa: E704
def shear_y(img, factor, **kw): return img.transform(img.size, Image.AFFINE, (1, 0, 0, factor, 1, 0), **_kw(**kw))  # noqa: E704
def translate_x_rel(img, pct, **kw): return img.transform(img.size, Image.AFFINE, (1, 0, pct * img.size[0], 0, 1, 0), **_kw(**kw))  # noqa: E704
def translate_y_rel(img, pct, **kw): return img.transform(img.size, Image.AFFINE, (1, 0, 0, 0, 1, pct * img.size[1]), **_kw(**kw))  # noqa: E704
def rotate(img, degrees, **kw): return img.rotate(degrees, **_kw(**kw))  # noqa: E704
def auto_contrast(img, **_): return ImageOps.autocontrast(img)  # noqa: E704
def invert(img, **_): return ImageOps.invert(img)  # noqa: E704
def equalize(img, **_): return ImageOps.equalize(img)  # noqa: E704
def solarize(img, thresh, **_): return ImageOps.solarize(img, thresh)  # noqa: E704
def posterize(img, bits, **_): return img if bits >= 8 else ImageOps.posterize(img, bits)  # noqa: E704
def contrast(img, f, **_): return ImageEnhance.Contrast(img).enhance(f)  # noqa: E704
def color(img, f, **_): return ImageEnhance.Color(img).enhance(f)  # noqa: E704
def brightness(img, f, **_): return ImageEnhance.Brightness(img).enhance(f)  # noqa: E704
def sharpness(img, f, **_): return ImageEnhance.Sharpness(img).enhance(f)  # noqa: E704


def solarize_add(img, add, thresh=128, **_):
    lut = [min(255, i + add) if i < thresh else i for i in range(256)]
    return img.point(lut + lut + lut)


OPS = {"AutoContrast": auto_contrast, "Equalize": equalize, "Invert": invert, "Rotate": rotate, "PosterizeIncreasing": posterize,
       "SolarizeIncreasing": solarize, "SolarizeAdd": solarize_add, "ColorIncreasing": color, "ContrastIncreasing": contrast,
       "BrightnessIncreasing": brightness, "SharpnessIncreasing": sharpness, "ShearX": shear_x, "ShearY": shear_y,
       "TranslateXRel": translate_x_rel, "TranslateYRel": translate_y_rel}


# ---- numpy restatements of Pillow's arithmetic
def luma_np(img):
    r, g, b = [img[..., i].astype(np.int64) for i in range(3)]
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend_np(deg, img, factor):
    d, x = deg.astype(np.float32), img.astype(np.float32)
    t = d + np.float32(factor) * (x - d)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int64))).astype(np.uint8)


def smooth_np(img):
    k = np.array([1, 1, 1, 1, 5, 1, 1, 1, 1], np.float32) / np.float32(13)
    x = img.astype(np.float32)
    H, W = img.shape[:2]
    row = lambda r, kk: (x[1 + r:H - 1 + r, 0:W - 2] * kk[0] + x[1 + r:H - 1 + r, 1:W - 1] * kk[1]) + x[1 + r:H - 1 + r, 2:W] * kk[2]  # noqa: E731
    ss = np.float32(0.5) + np.zeros((H - 2, W - 2, img.shape[2]), np.float32)
    ss = ss + row(1, k[0:3])
    ss = ss + row(0, k[3:6])
    ss = ss + row(-1, k[6:9])
    out = img.copy()
    out[1:-1, 1:-1] = np.where(ss <= 0, 0, np.where(ss >= 255, 255, ss.astype(np.int64))).astype(np.uint8)
    return out


def enhance_np(img, mode, factor):
    if mode == 0:
        deg = np.zeros_like(img)
    elif mode == 1:
        deg = np.repeat(luma_np(img)[..., None], 3, 2)
    elif mode == 2:
        L = luma_np(img)
        deg = np.full_like(img, int(L.astype(np.float64).sum() / L.size + 0.5))
    else:
        deg = smooth_np(img)
    return blend_np(deg, img, factor)


def affine_np(img, m, bicubic, fill):
    """Image.transform(size, AFFINE, m, BILINEAR | BICUBIC, fillcolor=fill), same-size output."""
    H, W, _ = img.shape
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    xin = m[0] * (xs + 0.5) + m[1] * (ys + 0.5) + m[2]
    yin = m[3] * (xs + 0.5) + m[4] * (ys + 0.5) + m[5]
    inside = (xin >= 0) & (xin < W) & (yin >= 0) & (yin < H)
    xin, yin = xin - 0.5, yin - 0.5
    fl = lambda v: np.where(v >= 0, np.trunc(v), np.floor(v)).astype(np.int64)  # noqa: E731
    x, y = fl(xin), fl(yin)
    dx, dy = (xin - x)[..., None], (yin - y)[..., None]
    I = img.astype(np.float64)
    xc, yc = (lambda v: np.clip(v, 0, W - 1)), (lambda v: np.clip(v, 0, H - 1))
    has = lambda yy: ((yy >= 0) & (yy < H))[..., None]  # noqa: E731
    if not bicubic:
        rowv = lambda yy: I[yy, xc(x)] + (I[yy, xc(x + 1)] - I[yy, xc(x)]) * dx  # noqa: E731
        v1 = rowv(yc(y))
        v2 = np.where(has(y + 1), rowv(yc(y + 1)), v1)
        out = (v1 + (v2 - v1) * dy).astype(np.int64)
    else:
        def cub(v1, v2, v3, v4, d):
            p1, p2, p3, p4 = v2, -v1 + v3, 2 * (v1 - v2) + v3 - v4, -v1 + v2 - v3 + v4
            return p1 + d * (p2 + d * (p3 + d * p4))
        x0, y0 = x - 1, y - 1
        rowv = lambda yy: cub(I[yy, xc(x0)], I[yy, xc(x0 + 1)], I[yy, xc(x0 + 2)], I[yy, xc(x0 + 3)], dx)  # noqa: E731
        v1 = rowv(yc(y0))
        v2 = np.where(has(y0 + 1), rowv(yc(y0 + 1)), v1)
        v3 = np.where(has(y0 + 2), rowv(yc(y0 + 2)), v2)
        v4 = np.where(has(y0 + 3), rowv(yc(y0 + 3)), v3)
        v = cub(v1, v2, v3, v4, dy)
        out = np.where(v <= 0, 0, np.where(v >= 255, 255, v.astype(np.int64)))
    return np.where(inside[..., None], np.clip(out, 0, 255).astype(np.uint8), np.array(fill, np.uint8)[None, None, :])


def rotation_matrix_np(w, h, degrees):
    a = -math.radians(degrees % 360.0)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2, h / 2
    m[2], m[5] = m[0] * -cx + m[1] * -cy + m[2], m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def aug_images():
    """Seeded test images: uniform noise (odd size), a narrow-range image, a clipped gaussian, a constant."""
    rng = np.random.RandomState(3)
    return [rng.randint(0, 256, (97, 131, 3), dtype=np.uint8), (rng.rand(64, 80, 3) * 120 + 60).astype(np.uint8),
            np.clip(rng.randn(50, 70, 3) * 30 + 128, 0, 255).astype(np.uint8), np.full((9, 11, 3), 77, np.uint8)]


# every op with the argument values the level functions produce at magnitude 0 / 4.3 / 9 / 10, both signs
OP_CASES = ([("AutoContrast", ()), ("Equalize", ()), ("Invert", ())]
            + [("Rotate", (d,)) for d in (27.0, -12.9, 0.0, 180.0, 30.0)]
            + [("PosterizeIncreasing", (b,)) for b in (4, 3, 1, 0, 8)]
            + [("SolarizeIncreasing", (t,)) for t in (256, 146, 26, 0)]
            + [("SolarizeAdd", (a,)) for a in (0, 47, 99, 110)]
            + [(n, (f,)) for n in ("ColorIncreasing", "ContrastIncreasing", "BrightnessIncreasing", "SharpnessIncreasing")
               for f in (0.1, 0.19, 0.613, 1.0, 1.387, 1.81, 1.9)]
            + [(n, (f,)) for n in ("ShearX", "ShearY") for f in (0.27, -0.129, 0.3)]
            + [(n, (f,)) for n in ("TranslateXRel", "TranslateYRel") for f in (0.405, -0.19, 0.45)])
