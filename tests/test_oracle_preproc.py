"""CPU: pins the preprocessing oracle (oracle/ref_preproc.py) to OpenCV and to the committed golden fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

import ref_preproc as P

cv2 = pytest.importorskip("cv2")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preproc_golden.json")))

# SURVEY.md App. A checksums (sha1 of int32 LE bytes, first 16 hex)
TABLE_SHA = {"gamma": "8a88c35441764553", "cbrt": "1eb5ec72d4085b5a", "yf": "2bcfd7adb3a3178f",
             "abxz": "5f8e53ab76820250", "invgamma": "5d288977e5795f21"}


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_table_checksums():
    for k, v in P.tables().items():
        assert P.table_sha(v) == TABLE_SHA[k] == GOLD["tables"][k], k


def test_lab_conversions_exhaustive_vs_cv2():
    c = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([c & 255, (c >> 8) & 255, (c >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert not (cv2.cvtColor(img, cv2.COLOR_BGR2LAB) != P.bgr2lab_np(img)).any()
    assert not (cv2.cvtColor(img, cv2.COLOR_LAB2BGR) != P.lab2bgr_np(img)).any()


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"{c['name']}-{c['h']}x{c['w']}")
def test_clahe_and_resize_match_cv2_and_golden(case):
    img = P.image_set(case["name"], case["h"], case["w"])
    assert sha(img) == case["input"], "synthetic image set is not reproducible"
    ref = P.apply_clahe_cv2(img)
    assert sha(ref) == case["clahe"], "OpenCV output differs from the fixture minted from the reference"
    if case["h"] * case["w"] <= 1024 * 1024:
        assert not (P.apply_clahe_np(img) != ref).any()
    for s in (224, 512):
        r = P.centre_crop_resize_cv2(ref, s)
        assert sha(r) == case[f"clahe_resize{s}"]
        assert not (P.centre_crop_resize_np(ref, s) != r).any()
        assert sha(P.centre_crop_resize_np(img, s)) == case[f"resize{s}"]


def test_resize_upscale_and_odd_shapes():
    rng = np.random.default_rng(5)
    for (h, w, d) in [(300, 200, 512), (37, 91, 64), (224, 224, 224), (513, 513, 100)]:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert not (P.centre_crop_resize_np(img, d) != P.centre_crop_resize_cv2(img, d)).any()


def test_normalize_flip_matches_torch():
    import torch
    img = P.image_set("noise", 32, 48)
    x = torch.from_numpy(np.ascontiguousarray(img[..., ::-1])).permute(2, 0, 1).float().div(255)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    ref = (x - mean) / std
    assert np.allclose(P.normalize_flip_np(img, 0), ref.numpy(), atol=1e-6)
    assert np.allclose(P.normalize_flip_np(img, 1), torch.flip(ref, [2]).numpy(), atol=1e-6)
    assert np.allclose(P.normalize_flip_np(img, 2), torch.flip(ref, [1]).numpy(), atol=1e-6)


@pytest.mark.reference
def test_oracle_equals_imported_reference():
    import sys
    sys.path.insert(0, "/root/reference")
    from src.preprocessing.normalise import apply_clahe
    from src.preprocessing.pipeline import centre_crop_resize
    for name, h, w in [("radiograph", 512, 512), ("noise", 250, 333)]:
        img = P.image_set(name, h, w)
        assert not (apply_clahe(img) != P.apply_clahe_cv2(img)).any()
        assert not (apply_clahe(img) != P.apply_clahe_np(img)).any()
        assert not (centre_crop_resize(img, 224) != P.centre_crop_resize_np(img, 224)).any()


# ---- deskew: numpy restatements vs OpenCV, the golden fixture of the reference's own deskew, and the
# host half of the product (angle from integer moments, rotation matrix, affine inversion)
DESKEW_GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "deskew_golden.json")))


def _sha(a):
    import hashlib
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("idx", range(len(P.DESKEW_CASES)))
def test_deskew_restatement_matches_cv2_and_golden(idx):
    import cv2
    from teethrt import preproc as pre                   # host-side helpers only; no kernel is launched here
    h, w, seed, tilt = P.DESKEW_CASES[idx]
    g = DESKEW_GOLD["cases"][idx]
    img = P.tooth_image(h, w, seed, tilt)
    assert _sha(img) == g["input"]
    rot, angle, edges = P.deskew_cv2(img)
    assert angle == g["angle"] and _sha(rot) == g["output"]                    # oracle == the reference's function
    gray = P.gray_np(img)
    assert np.array_equal(gray, cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(P.canny_np(gray), edges)
    ys, xs = np.nonzero(edges)
    ys, xs = ys.astype(object), xs.astype(object)                              # exact Python integers
    mom = [len(ys), ys.sum(), xs.sum(), (ys * ys).sum(), (xs * ys).sum(), (xs * xs).sum()]
    mine = pre.edge_angle(mom)
    ref_angle = np.rad2deg(np.arctan2(*_principal(edges)))
    assert abs(mine - ref_angle) < 1e-9
    if g["angle"] != 0.0:
        M = pre.rotation_matrix_2d((w / 2, h / 2), g["angle"])
        assert np.array_equal(M, cv2.getRotationMatrix2D((w / 2, h / 2), g["angle"], 1.0))
        assert np.array_equal(P.rotation_matrix_np((w / 2, h / 2), g["angle"]), M)
        assert _sha(P.warp_affine_np(img, M, w, h)) == g["output"]
        assert np.allclose(pre.invert_affine(M), cv2.invertAffineTransform(M).ravel(), atol=1e-12)
        assert _sha(P.warp_affine_np(img, pre.rotation_matrix_2d((w / 2, h / 2), mine), w, h)) == g["output"]


def _principal(edges):
    coords = np.column_stack(np.where(edges > 0))
    centered = coords - coords.mean(axis=0)
    vals, vecs = np.linalg.eigh(np.cov(centered, rowvar=False))
    p = vecs[:, np.argmax(vals)]
    return p[0], p[1]


def test_warp_restatement_on_extreme_maps():
    import cv2
    img = P.image_set("noise", 97, 131)
    for M in (np.array([[1, 0, 0.5], [0, 1, -0.25]]), np.array([[0.5, 0.1, -30], [-0.2, 1.7, 40.0]]),
              P.rotation_matrix_np((65.5, 48.5), 180.0), np.array([[1, 0, 0], [0, 1, 0.0]]), np.array([[1, 0, 500], [0, 1, 0.0]])):
        want = cv2.warpAffine(img, M, (131, 97), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        assert np.array_equal(P.warp_affine_np(img, M, 131, 97), want)
        want2 = cv2.warpAffine(img, M, (200, 50), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        assert np.array_equal(P.warp_affine_np(img, M, 200, 50), want2)


def test_ab_to_xz_table_is_pure_integer_arithmetic():
    """The CLAHE fast path computes OpenCV's abToXZ_b entries instead of gathering them from a 147 KB table
    (csrc/preproc.cu abxz_at): the closed form must reproduce the digest-checked table entry by entry, in int32 range."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("lab_tables", os.path.join(os.path.dirname(__file__), "..",
                                                  "multimodal-teeth-restoration-selection_b200", "lab_tables.py"))
    lt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lt)
    table = lt._abxz()
    i = np.arange(-8145, -8145 + 36864, dtype=np.int64)
    lo = np.sign(i * 108) * (np.abs(i * 108) // 841) - 290          # C truncating division, as in `(i * 108) / 841 - 290`
    hi = (((i * i) >> 14) * i) >> 14
    assert int((i * i).max()) < 2 ** 31 and int((((i * i) >> 14) * i).max()) < 2 ** 31
    assert np.array_equal(np.where(i <= 3390, lo, hi), table)
