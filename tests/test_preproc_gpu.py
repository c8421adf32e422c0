"""GPU parity, input stage: CUDA CLAHE / resize / normalise against the oracle (the reference's own OpenCV call sequence,
oracle/ref_preproc.py) and the committed golden digests minted from /root/reference.  Bar: 0 differing bytes."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import ref_preproc as P   # oracle (checker only)

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preproc_golden.json")))


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def pre():
    import teethrt
    teethrt.init()
    from teethrt import preproc
    return preproc


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"{c['name']}-{c['h']}x{c['w']}")
def test_clahe_resize_bit_exact_vs_golden_and_opencv(pre, case):
    img = P.image_set(case["name"], case["h"], case["w"])
    out = pre.apply_clahe(img)
    ref = P.apply_clahe_cv2(img)
    assert int((out != ref).sum()) == 0
    assert sha(out) == case["clahe"]
    for s in (224, 512):
        r = pre.centre_crop_resize(out, s)
        assert int((r != P.centre_crop_resize_cv2(ref, s)).sum()) == 0
        assert sha(r) == case[f"clahe_resize{s}"]
        assert sha(pre.centre_crop_resize(img, s)) == case[f"resize{s}"]


def test_lab_roundtrip_exhaustive_all_colours(pre):
    """Every one of the 2^24 BGR colours through the CUDA Lab->CLAHE->BGR path on an image whose CLAHE LUTs are what
    OpenCV computes for it: compared byte for byte with OpenCV (covers both colour conversions exhaustively)."""
    c = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([c & 255, (c >> 8) & 255, (c >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    out = pre.apply_clahe(img)
    assert int((out != P.apply_clahe_cv2(img)).sum()) == 0


@pytest.mark.parametrize("h,w", [(64, 64), (8, 8), (9, 23), (250, 333), (1024, 8)])
def test_clahe_small_and_ragged_shapes(pre, h, w):
    img = P.image_set("noise", h, w, seed=h * 1000 + w)
    assert int((pre.apply_clahe(img) != P.apply_clahe_cv2(img)).sum()) == 0


def test_batched_device_path_and_idempotent_buffers(pre):
    imgs = np.stack([P.image_set(n, 256, 256, seed=i) for i, n in enumerate(["noise", "smooth", "radiograph", "ramp"])])
    dev = torch.from_numpy(imgs).cuda()
    out = pre.apply_clahe(dev)
    assert out.is_cuda and out.shape == dev.shape
    for i in range(4):
        assert int((out[i].cpu().numpy() != P.apply_clahe_cv2(imgs[i])).sum()) == 0
    small = pre.centre_crop_resize(out, 224)
    for i in range(4):
        assert int((small[i].cpu().numpy() != P.centre_crop_resize_cv2(P.apply_clahe_cv2(imgs[i]), 224)).sum()) == 0


@pytest.mark.parametrize("flip", [0, 1, 2])
def test_normalize_flip(pre, flip):
    img = P.image_set("noise", 40, 56)
    ref = P.normalize_flip_np(img, flip)
    out = pre.normalize_flip(img, flip, dtype=torch.float32).cpu().numpy()
    assert np.allclose(out, ref, atol=1e-6)      # fp32: same IEEE operations; tolerance covers division rounding only
    ob = pre.normalize_flip(img, flip, dtype=torch.bfloat16).float().cpu().numpy()
    assert np.allclose(ob, ref, atol=2e-2)       # bf16 storage: 8-bit mantissa


def test_input_stage_full_size_properties(pre):
    """BASELINE config 3 at full size (N x 1024^2): size-independent properties — flips commute with the stage's last step,
    and the batched result equals the per-image result."""
    n = 8
    imgs = torch.from_numpy(np.stack([P.image_set("radiograph", 1024, 1024, seed=i) for i in range(n)])).cuda()
    st = pre.InputStage(n, 1024, 1024, size=224, dtype=torch.float32)
    a = st(imgs, 0).clone()
    b = st(imgs, 1).clone()
    c = st(imgs, 2).clone()
    assert torch.equal(torch.flip(a, [3]), b) and torch.equal(torch.flip(a, [2]), c)
    one = pre.normalize_flip(pre.centre_crop_resize(pre.apply_clahe(imgs[3]), 224), 0)
    assert torch.equal(one, a[3])
    ref = P.normalize_flip_np(P.centre_crop_resize_cv2(P.apply_clahe_cv2(imgs[3].cpu().numpy()), 224), 0)
    assert np.allclose(a[3].cpu().numpy(), ref, atol=1e-6)


# ---- PIL-exact eval transform: the checker is Pillow/torchvision itself, i.e. what the reference calls
RESAMPLE_CASES = [(1024, 1024, 256, 224, "bicubic"), (480, 640, 256, 224, "bicubic"), (777, 1003, 434, 380, "bicubic"),
                  (200, 150, 256, 224, "bicubic"), (256, 300, 256, 224, "bicubic"), (1000, 1003, 512, 480, "bilinear"),
                  (512, 512, 512, 480, "bilinear"), (300, 451, 512, 480, "bilinear"), (97, 1024, 64, 56, "bicubic"),
                  (2048, 3072, 434, 380, "bicubic")]


@pytest.mark.parametrize("h,w,short,crop,interp", RESAMPLE_CASES)
def test_resize_center_crop_is_bit_identical_to_pillow(pre, h, w, short, crop, interp):
    from PIL import Image
    from torchvision import transforms
    mode = {"bicubic": transforms.InterpolationMode.BICUBIC, "bilinear": transforms.InterpolationMode.BILINEAR}[interp]
    rng = np.random.RandomState(h * 7 + w)
    img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
    img[: h // 3] = (img[: h // 3] > 127) * 255                    # hard edges: bicubic overshoot must clip like Pillow
    tf = transforms.Compose([transforms.Resize(short, interpolation=mode), transforms.CenterCrop(crop)])
    want = np.asarray(tf(Image.fromarray(img)))
    got = pre.resize_center_crop(img, short, crop, interp)
    assert got.is_cuda and got.shape == (crop, crop, 3) and int((got.cpu().numpy() != want).sum()) == 0
    swapped = pre.resize_center_crop(torch.from_numpy(img).cuda(), short, crop, interp, swap_channels=True)
    assert int((swapped.cpu().numpy() != want[..., ::-1]).sum()) == 0
    grey = pre.resize_center_crop(img[..., :1].copy(), short, crop, interp)
    want_g = np.asarray(tf(Image.fromarray(img[..., 0])))
    assert int((grey.cpu().numpy()[..., 0] != want_g).sum()) == 0


def test_resize_center_crop_rejects_bad_input(pre):
    with pytest.raises(ValueError):
        pre.resize_center_crop(np.zeros((8, 8, 3), np.float32), 8, 4)
    with pytest.raises(ValueError):
        pre.resize_center_crop(np.zeros((8, 8, 2), np.uint8), 8, 4)
    with pytest.raises(ValueError):
        pre.resize_center_crop(np.zeros((64, 64, 3), np.uint8), 32, 48)
