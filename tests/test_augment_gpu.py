"""GPU parity for the RandAugment image operations: every op of 'rand-m9-mstd0.5-inc1' through the C ABI
against Pillow (the library timm's op functions call), bit for bit, at the argument values the policy's level functions
produce; then the sampling wrapper and the whole train transform (shape / range / determinism — draw-for-draw parity with
timm's random streams is unpinned, see teethrt/augment.py)."""
import numpy as np
import pytest
import torch
from PIL import Image

import ref_augment as RA   # oracle (checker only)

pytestmark = pytest.mark.gpu
IMGS = RA.aug_images()


@pytest.fixture(scope="module")
def A():
    import teethrt
    teethrt.init()
    from teethrt import augment
    return augment


@pytest.mark.parametrize("k", range(len(IMGS)))
def test_every_op_is_bit_identical_to_pillow(A, k):
    a = IMGS[k]
    im = Image.fromarray(a)
    dev = torch.from_numpy(a).cuda()
    for name, args in RA.OP_CASES:
        for interp in (("bicubic", "bilinear") if name in ("Rotate", "ShearX", "ShearY", "TranslateXRel", "TranslateYRel") else ("bicubic",)):
            if k == 0 and name == "Rotate" and args[0] % 180 == 0:
                pass                                                    # 0 / 180 degrees: Pillow's transpose fast paths
            want = np.asarray(RA.OPS[name](im, *args, resample=interp))
            got = A._OP_FN[name](dev, *args, resample=interp, fillcolor=RA.FILL).cpu().numpy()
            assert got.shape == want.shape and int((got != want).sum()) == 0, (k, name, args, interp, int((got != want).sum()))


def test_square_rotations_and_numpy_input(A):
    a = IMGS[0][:97, :97].copy()
    im = Image.fromarray(a)
    for deg in (90.0, 270.0, 45.0, -90.0, 360.0):
        want = np.asarray(im.rotate(deg, Image.BICUBIC, fillcolor=RA.FILL))
        assert int((A.rotate(a, deg).cpu().numpy() != want).sum()) == 0, deg
    with pytest.raises(ValueError):
        A.color(np.zeros((8, 8, 1), np.uint8), 1.2)


def test_randaugment_policy_and_train_transform(A):
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (300, 451, 3), dtype=np.uint8)
    im = Image.fromarray(img)
    seen = set()
    for seed in range(40):
        ra = A.RandAugment(seed=seed)
        out = ra(img)
        # replaying the ops the policy chose through Pillow gives the same image: the wrapper adds nothing to the pixels
        ref = im
        for name, args in ra.last:
            ref = RA.OPS[name](ref, *args)
            seen.add(name)
        assert int((out.cpu().numpy() != np.asarray(ref)).sum()) == 0, (seed, ra.last)
        assert len(ra.last) <= 2
        for name, args in ra.last:                     # magnitudes stay inside the documented ranges
            if name == "Rotate":
                assert abs(args[0]) <= 30.0
            if name.endswith("Increasing") and name[0] in "CBS" and name != "SolarizeIncreasing":
                assert 0.1 <= args[0] <= 1.9
    assert len(seen) >= 10                               # 40 seeds x 2 layers x p=0.5 reach most of the 15 ops
    again = A.RandAugment(seed=7)(img)
    assert torch.equal(again, A.RandAugment(seed=7)(img))
    tf = A.TrainTransform(224, seed=5)
    x = tf(img)
    assert x.shape == (3, 224, 224) and x.dtype == torch.float32 and x.is_cuda and torch.isfinite(x).all()
    assert torch.equal(x, A.TrainTransform(224, seed=5)(img))
    erased = 0
    for seed in range(30):
        y = A.TrainTransform(96, re_prob=1.0, seed=seed)(img)
        erased += int((y.abs() > 2.7).any())             # normal noise exceeds the normalised pixel range somewhere
    assert erased >= 25
    bf = A.TrainTransform(96, dtype=torch.bfloat16, seed=1)(img)
    assert bf.dtype == torch.bfloat16 and bf.shape == (3, 96, 96)


@pytest.mark.parametrize("B,H,W,S", [(12, 300, 451, 224), (7, 512, 512, 96), (5, 130, 97, 160)])
def test_batched_train_transform_equals_the_per_image_path(A, B, H, W, S):
    """One launch per stage over the whole batch must give, bit for bit, what the single-image operations (pinned to Pillow
    above) give when they replay the same draws: crop + resize, flip, two RandAugment layers, ToTensor + Normalize; inside an
    erase box the pixels are the noise, outside they are untouched."""
    from teethrt.preproc import resized_crop, normalize_flip
    rng = np.random.RandomState(B)
    imgs = rng.randint(0, 256, (B, H, W, 3), dtype=np.uint8)
    imgs[0] = (np.add.outer(np.arange(H), np.arange(W))[:, :, None] % 256).astype(np.uint8)      # a smooth image: LUT / equalize paths
    seen = set()
    for seed in range(6):
        tf = A.BatchTrainTransform(S, re_prob=0.5, dtype=torch.float32, seed=seed)
        plan = tf.sample(B, H, W)
        out = tf.apply(imgs, plan)
        assert out.shape == (B, 3, S, S) and out.dtype == torch.float32 and torch.isfinite(out).all()
        for b in range(B):
            top, left, ch, cw = (int(v) for v in plan.boxes[b])
            x = resized_crop(imgs[b], top, left, ch, cw, S, "bicubic")
            if plan.flips[b]:
                x = torch.flip(x, dims=[1]).contiguous()
            for layer in plan.layers:
                if layer[b] is not None:
                    name, args = layer[b]
                    seen.add(name)
                    x = A._OP_FN[name](x, *args, resample="bicubic", fillcolor=A.IMG_MEAN_FILL)
            want = normalize_flip(torch.flip(x, dims=[2]).contiguous(), 0, dtype=torch.float32)
            et, el, eh, ew = (int(v) for v in plan.erase[b])
            mask = torch.ones(S, S, dtype=torch.bool, device="cuda")
            mask[et:et + eh, el:el + ew] = False
            assert torch.equal(out[b][:, mask], want[:, mask]), (seed, b, plan.layers[0][b], plan.layers[1][b])
            if eh:
                box = out[b][:, et:et + eh, el:el + ew]
                assert not torch.equal(box, want[:, et:et + eh, el:el + ew]) and float(box.std()) > 0.5
    assert len(seen) >= 12
    # determinism, dtype, device-tensor input
    a = A.BatchTrainTransform(S, seed=3)(torch.from_numpy(imgs).cuda())
    b_ = A.BatchTrainTransform(S, seed=3)(imgs)
    assert a.dtype == torch.bfloat16 and torch.equal(a, b_)
