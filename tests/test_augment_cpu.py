"""CPU: the arithmetic of the RandAugment image operations is pinned to Pillow three ways without a
GPU: (1) the numpy restatements in oracle/ref_augment.py equal Pillow; (2) the per-pixel C++ functions the CUDA kernels wrap
(csrc/augment_core.h), compiled for the host by g++ into a throw-away library, equal Pillow; (3) the host-side pieces of the
product (rotation matrix, level functions, crop boxes) equal Pillow / the documented policy."""
import ctypes as C
import math
import os
import random
import subprocess

import numpy as np
import pytest
from PIL import Image

import ref_augment as RA

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMGS = RA.aug_images()


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = tmp_path_factory.mktemp("aug") / "libaug_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "multimodal-teeth-restoration-selection_b200", "csrc"),
                    "-o", str(so), os.path.join(ROOT, "tests", "cpu_harness", "augment_host.cpp")], check=True)
    lib = C.CDLL(str(so))
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    lib.h_enhance_rgb.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_float, u8p]
    lib.h_affine.argtypes = [u8p, C.c_int, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.float64), C.c_int, u8p, u8p]
    lib.h_hist_lut.argtypes = [u8p, C.c_long, C.c_int, C.c_int, u8p]
    return lib


def affine_matrix(name, arg, w, h):
    return {"ShearX": (1, arg, 0, 0, 1, 0), "ShearY": (1, 0, 0, arg, 1, 0), "TranslateXRel": (1, 0, arg * w, 0, 1, 0),
            "TranslateYRel": (1, 0, 0, 0, 1, arg * h)}.get(name) or RA.rotation_matrix_np(w, h, arg)


@pytest.mark.parametrize("k", range(len(IMGS)))
def test_core_functions_and_restatements_equal_pillow(host, k):
    a = IMGS[k]
    im = Image.fromarray(a)
    H, W = a.shape[:2]
    fill = np.array(RA.FILL, np.uint8)
    for name, args in RA.OP_CASES:
        want = np.asarray(RA.OPS[name](im, *args))
        if name in ("ColorIncreasing", "ContrastIncreasing", "BrightnessIncreasing", "SharpnessIncreasing"):
            mode = {"B": 0, "C": None}.get(name[0])
            mode = {"BrightnessIncreasing": 0, "ColorIncreasing": 1, "ContrastIncreasing": 2, "SharpnessIncreasing": 3}[name]
            out = np.empty_like(a)
            host.h_enhance_rgb(a, H, W, mode, args[0], out)
            assert np.array_equal(out, want), (name, args)
            assert np.array_equal(RA.enhance_np(a, mode, args[0]), want), (name, args)
        elif name in ("AutoContrast", "Equalize"):
            lut = np.empty(768, np.uint8)
            host.h_hist_lut(a, H * W, 3, 0 if name == "AutoContrast" else 1, lut)
            assert np.array_equal(np.stack([lut[c * 256:(c + 1) * 256][a[..., c]] for c in range(3)], -1), want), name
        elif name in ("ShearX", "ShearY", "TranslateXRel", "TranslateYRel") or (name == "Rotate" and args[0] % 180 != 0):
            m = np.array(affine_matrix(name, args[0], W, H), np.float64)
            for interp, bic in (("bicubic", 1), ("bilinear", 0)):
                want = np.asarray(RA.OPS[name](im, *args, resample=interp))
                out = np.empty_like(a)
                host.h_affine(a, H, W, 3, m, bic, fill, out)
                assert np.array_equal(out, want), (name, args, interp)
                assert np.array_equal(RA.affine_np(a, m, bic, RA.FILL), want), (name, args, interp)


def test_host_side_pieces():
    from teethrt import augment as A                     # host logic only; no kernel is launched here
    assert A.IMG_MEAN_FILL == RA.FILL and list(A._OP_FN) == list(RA.OPS) == A.RAND_INCREASING_TRANSFORMS
    assert A.rotation_matrix(131, 97, 23.7) == RA.rotation_matrix_np(131, 97, 23.7)
    assert A.rotation_matrix(64, 64, 90.0) is None and A.rotation_matrix(64, 48, 90.0) is not None and A.rotation_matrix(64, 48, 180.0) is None
    # level functions of the 'increasing' family at magnitude 9: rotate 27 deg, enhance 1 +- 0.81, shear 0.27, translate 0.405,
    # posterize 4 - 3 bits, solarize threshold 256 - 230, solarize-add 99
    r = random.Random(0)
    assert abs(abs(A._level_args("Rotate", 9.0, r)[0]) - 27.0) < 1e-12
    assert round(abs(A._level_args("ColorIncreasing", 9.0, r)[0] - 1.0), 6) == 0.81
    assert abs(abs(A._level_args("ShearX", 9.0, r)[0]) - 0.27) < 1e-12 and abs(abs(A._level_args("TranslateYRel", 9.0, r)[0]) - 0.405) < 1e-12
    assert A._level_args("PosterizeIncreasing", 9.0, r) == (1,) and A._level_args("SolarizeIncreasing", 9.0, r) == (26,)
    assert A._level_args("SolarizeAdd", 9.0, r) == (99,) and A._level_args("Equalize", 9.0, r) == ()
    assert A._level_args("BrightnessIncreasing", 10.0, random.Random(3))[0] >= 0.1
    # crop boxes stay inside the image and respect the scale / ratio ranges
    rr = random.Random(1)
    for _ in range(200):
        t, l, ch, cw = A.random_resized_crop_box(300, 451, rr)
        assert 0 <= t and 0 <= l and t + ch <= 300 and l + cw <= 451 and ch > 0 and cw > 0
        assert 0.07 * 300 * 451 <= ch * cw <= 300 * 451 and 0.7 <= cw / ch <= 1.4
    assert A.random_resized_crop_box(10, 1000, rr, scale=(5.0, 6.0)) == (0, 493, 10, 13)        # fallback: central, ratio clamped


def test_point_tables_equal_pillow():
    from teethrt import augment as A
    for a in IMGS:
        im = Image.fromarray(a)
        look = lambda t: t[a]  # noqa: E731
        assert np.array_equal(look(A.point_table("invert")), np.asarray(RA.invert(im)))
        for bits in (0, 1, 3, 4, 7):
            assert np.array_equal(look(A.point_table("posterize", bits)), np.asarray(RA.posterize(im, bits))), bits
        for t in (0, 26, 146, 256):
            assert np.array_equal(look(A.point_table("solarize", t)), np.asarray(RA.solarize(im, t))), t
        for add in (0, 47, 99, 110):
            assert np.array_equal(look(A.point_table("solarize_add", add)), np.asarray(RA.solarize_add(im, add))), add


def test_pillow_transpose_fast_paths_are_rot90():
    import torch
    a = IMGS[0][:97, :97].copy()
    im = Image.fromarray(a)
    for deg in (90.0, 180.0, 270.0, -90.0):
        k = int(round((deg % 360.0) / 90.0))
        assert np.array_equal(torch.rot90(torch.from_numpy(a), k, dims=(0, 1)).numpy(), np.asarray(im.rotate(deg, Image.BICUBIC, fillcolor=RA.FILL)))


def test_device_coefficient_routine_equals_pillow_tables(host):
    """The batched crop+resize builds its per-image coefficient tables ON THE DEVICE with pil_resample_taps (augment_core.h).
    Compiled for the host it must reproduce teethrt.preproc.pil_coeffs, the table builder that tests/test_pil_resample_cpu.py
    pins to Pillow's own resize, for up- and down-scaling and both filters."""
    import teethrt  # noqa: F401  (import alias)
    from teethrt.preproc import pil_coeffs
    host.h_resample_taps.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for in_size, out_size in [(512, 224), (433, 224), (97, 160), (224, 224), (1024, 224), (60, 224), (354, 96), (7, 3)]:
        for interp in ("bicubic", "bilinear"):
            b, c, ks = pil_coeffs(in_size, out_size, interp)
            first, k = C.c_int(), (C.c_int * ks)()
            for xx in range(out_size):
                n = host.h_resample_taps(in_size, out_size, int(interp == "bicubic"), xx, ks, C.byref(first), k)
                assert (first.value, n) == tuple(b[xx]) and list(k) == list(c[xx]), (in_size, out_size, interp, xx)
