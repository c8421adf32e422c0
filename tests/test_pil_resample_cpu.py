"""CPU: the host half of the PIL-exact eval transform (teethrt.preproc.pil_coeffs / resized_size) is pinned to Pillow itself:
a numpy walk over the coefficient tables, doing what the CUDA kernel does (22-bit fixed point, round + clip per pass,
horizontal then vertical), must reproduce torchvision Resize + CenterCrop on PIL images bit for bit."""
import numpy as np
import pytest
from PIL import Image
from torchvision import transforms

from teethrt.preproc import pil_coeffs, resized_size

IMODE = {"bicubic": transforms.InterpolationMode.BICUBIC, "bilinear": transforms.InterpolationMode.BILINEAR}


def pil_path(img, short, crop, interp):
    tf = transforms.Compose([transforms.Resize(short, interpolation=IMODE[interp]), transforms.CenterCrop(crop)])
    return np.asarray(tf(Image.fromarray(img)))


def table_pass(img, bounds, coeffs, axis):
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((len(bounds),) + src.shape[1:], np.uint8)
    for o, ((first, taps), k) in enumerate(zip(bounds, coeffs)):
        acc = (1 << 21) + np.tensordot(k[:taps].astype(np.int64), src[first:first + taps], axes=1)
        out[o] = np.clip(acc >> 22, 0, 255)
    return np.moveaxis(out, 0, axis)


def table_path(img, short, crop, interp):
    h, w = img.shape[:2]
    nh, nw = resized_size(h, w, short)
    hb, hc, _ = pil_coeffs(w, nw, interp)
    vb, vc, _ = pil_coeffs(h, nh, interp)
    full = table_pass(table_pass(img, hb, hc, 1), vb, vc, 0)
    top, left = int(round((nh - crop) / 2.0)), int(round((nw - crop) / 2.0))
    return full[top:top + crop, left:left + crop]


CASES = [(1024, 1024, 256, 224, "bicubic"), (480, 640, 256, 224, "bicubic"), (777, 1003, 434, 380, "bicubic"),
         (200, 150, 256, 224, "bicubic"), (256, 300, 256, 224, "bicubic"), (1000, 1003, 512, 480, "bilinear"),
         (512, 512, 512, 480, "bilinear"), (300, 451, 512, 480, "bilinear"), (97, 1024, 64, 56, "bicubic")]


@pytest.mark.parametrize("h,w,short,crop,interp", CASES)
def test_tables_reproduce_pillow(h, w, short, crop, interp):
    rng = np.random.RandomState(h * 7 + w)
    img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
    img[: h // 3] = (img[: h // 3] > 127) * 255                       # hard edges: overshoot must clip like Pillow
    assert np.array_equal(table_path(img, short, crop, interp), pil_path(img, short, crop, interp))


def test_table_shapes_and_identity():
    b, c, k = pil_coeffs(256, 256, "bicubic")
    assert k == 5 and all(int(c[i].sum()) == 1 << 22 for i in range(256))
    assert all(int(c[i, np.flatnonzero(c[i])[0]]) == 1 << 22 for i in range(256))          # same size: pure copy
    b, c, k = pil_coeffs(1024, 256, "bicubic")
    assert k == 17 and b[:, 1].max() <= k and abs(int(c[100].sum()) - (1 << 22)) <= k
    assert resized_size(480, 640, 256) == (256, 341) and resized_size(640, 480, 256) == (341, 256)
