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


# ================================================================================================= batched transform
# The reference runs the transform above image by image inside DataLoader workers (train_mm_joint_dualtask.py:72-85,211).
# At 5,000+ images/s per GPU that is the bottleneck again, so the batch form below takes one uint8 batch [B,H,W,3] on the
# device and runs every stage as ONE launch over the images that need it: the host only samples (vectorised numpy) and
# fills three small job tables that are uploaded in one copy.  The kernels call the same per-pixel code as the
# single-image path, so `apply(batch, plan)` equals running `TrainTransform`'s operations image by image with the same
# draws (tests/test_augment_gpu.py) - and that path is pinned to Pillow.  Sampling parity with timm stays UNPINNED.
_AUG_JOB = np.dtype([("src", "<i8"), ("dst", "<i8"), ("op", "<i4"), ("mode", "<i4"), ("factor", "<f4"), ("bicubic", "<i4"),
                     ("m", "<f8", (6,)), ("slot", "<i4"), ("fill", "u1", (4,))], align=True)
_FIN_JOB = np.dtype([("src", "<i8"), ("top", "<i4"), ("left", "<i4"), ("eh", "<i4"), ("ew", "<i4"), ("noise", "<i8")], align=True)
_ENHANCE_MODE = {"BrightnessIncreasing": 0, "ColorIncreasing": 1, "ContrastIncreasing": 2, "SharpnessIncreasing": 3}
_POINT = {"Invert": "invert", "PosterizeIncreasing": "posterize", "SolarizeIncreasing": "solarize", "SolarizeAdd": "solarize_add"}


class BatchPlan:
    """Everything random about one batch, drawn on the host: crop boxes + flips, per layer (op, args) or None, erase boxes."""

    def __init__(self, boxes, flips, layers, erase):
        self.boxes, self.flips, self.layers, self.erase = boxes, flips, layers, erase


class BatchTrainTransform:
    """timm's training transform for a whole batch: uint8 RGB [B,H,W,3] (CUDA tensor, or numpy / pinned tensor that is
    uploaded) -> normalised [B,3,S,S] fp32 / bf16 CUDA tensor."""

    def __init__(self, img_size, re_prob=0.2, dtype=torch.bfloat16, seed=None, magnitude=9.0, magnitude_std=0.5, num_layers=2,
                 prob=0.5, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
        from ._lib import lib
        assert _AUG_JOB.itemsize == lib.trt_aug_job_bytes(), "AugJob layout out of sync with augment.cu"
        self.size, self.re_prob, self.dtype = int(img_size), float(re_prob), dtype
        self.magnitude, self.magnitude_std, self.num_layers, self.prob = magnitude, magnitude_std, num_layers, prob
        self.scale, self.ratio = scale, ratio
        self.rng = np.random.RandomState(seed)
        self.gen, self.seed = None, seed
        self._bufs = {}

    # ---- sampling: one vectorised pass over the batch -------------------------------------------------------------
    def sample(self, B, H, W):
        r = self.rng
        S = self.size
        # RandomResizedCropAndInterpolation.get_params: first of 10 attempts that fits, else the central fallback
        area = H * W
        ta = r.uniform(self.scale[0], self.scale[1], (B, 10)) * area
        asp = np.exp(r.uniform(math.log(self.ratio[0]), math.log(self.ratio[1]), (B, 10)))
        cw = np.rint(np.sqrt(ta * asp)).astype(np.int64)
        ch = np.rint(np.sqrt(ta / asp)).astype(np.int64)
        ok = (cw <= W) & (ch <= H) & (cw > 0) & (ch > 0)
        first = np.where(ok.any(1), ok.argmax(1), -1)
        in_ratio = W / H
        if in_ratio < min(self.ratio):
            fw, fh = W, int(round(W / min(self.ratio)))
        elif in_ratio > max(self.ratio):
            fh, fw = H, int(round(H * max(self.ratio)))
        else:
            fw, fh = W, H
        idx = np.arange(B)
        bw = np.where(first >= 0, cw[idx, np.maximum(first, 0)], fw)
        bh = np.where(first >= 0, ch[idx, np.maximum(first, 0)], fh)
        top = np.where(first >= 0, np.floor(r.uniform(0, 1, B) * (H - bh + 1)).astype(np.int64), (H - bh) // 2)
        left = np.where(first >= 0, np.floor(r.uniform(0, 1, B) * (W - bw + 1)).astype(np.int64), (W - bw) // 2)
        boxes = np.stack([top, left, bh, bw], 1).astype(np.int32)
        flips = r.uniform(0, 1, B) < 0.5
        # RandAugment 'rand-m9-mstd0.5-inc1'
        names = r.choice(len(RAND_INCREASING_TRANSFORMS), (B, self.num_layers), replace=True)
        applied = r.uniform(0, 1, (B, self.num_layers)) <= self.prob if self.prob < 1.0 else np.ones((B, self.num_layers), bool)
        mags = np.clip(r.normal(self.magnitude, self.magnitude_std, (B, self.num_layers)) if self.magnitude_std > 0
                       else np.full((B, self.num_layers), self.magnitude), 0.0, _LEVEL_DENOM)
        signs = r.uniform(0, 1, (B, self.num_layers))

        class _Sign:                                   # _level_args draws the sign through rng.random()
            def __init__(self, v):
                self.v = v

            def random(self):
                return self.v
        layers = []
        for l in range(self.num_layers):
            layer = []
            for b in range(B):
                if not applied[b, l]:
                    layer.append(None)
                    continue
                name = RAND_INCREASING_TRANSFORMS[names[b, l]]
                layer.append((name, _level_args(name, float(mags[b, l]), _Sign(float(signs[b, l])))))
            layers.append(layer)
        # RandomErasing(p, mode='pixel', one box, 10 attempts)
        erase = np.zeros((B, 4), np.int32)
        do = r.uniform(0, 1, B) <= self.re_prob
        et = r.uniform(0.02, 1 / 3, (B, 10)) * (S * S)
        ea = np.exp(r.uniform(math.log(0.3), math.log(1 / 0.3), (B, 10)))
        eh = np.rint(np.sqrt(et * ea)).astype(np.int64)
        ew = np.rint(np.sqrt(et / ea)).astype(np.int64)
        eok = (ew < S) & (eh < S)
        ef = np.where(eok.any(1), eok.argmax(1), -1)
        u1, u2 = r.uniform(0, 1, B), r.uniform(0, 1, B)
        for b in np.nonzero(do & (ef >= 0))[0]:
            h_, w_ = int(eh[b, ef[b]]), int(ew[b, ef[b]])
            erase[b] = (int(u1[b] * (S - h_ + 1)), int(u2[b] * (S - w_ + 1)), h_, w_)
        return BatchPlan(boxes, flips, layers, erase)

    # ---- execution ------------------------------------------------------------------------------------------------
    def _scratch(self, dev, B, H, S, kmax):
        key = (str(dev), B, H, S, kmax)
        if key not in self._bufs:
            u8 = lambda *s: torch.empty(s, device=dev, dtype=torch.uint8)
            self._bufs = {key: dict(bounds=torch.empty((B, 2, S, 2), device=dev, dtype=torch.int32),
                                    coeffs=torch.empty((B, 2, S, kmax), device=dev, dtype=torch.int32),
                                    tmp=u8(B, H, S, 3), a=u8(B, S, S, 3), b=u8(B, S, S, 3),
                                    luts=u8(2 * B, 768), stats=torch.zeros(2 * B * 769, device=dev, dtype=torch.int64))}
        return self._bufs[key]

    def _pinned(self, dev, B, blob_bytes):
        """Page-locked staging for the job tables and the point-operation LUTs: a ring of four sets, each guarded by the event
        recorded behind its upload.  (A fresh `pin_memory()` per call page-locks every step, and an upload from PAGEABLE memory
        blocks the host until the stream has drained - the previous train step - so the GPU idled while the rest of the
        transform and the step were still being enqueued: 13.7 instead of 12.0 ms per step end to end on a slow host.)"""
        ring = self.__dict__.setdefault("_pin_ring", {})
        key = (str(dev), B)
        if key not in ring or ring[key]["cap"] < blob_bytes:
            ring[key] = dict(cap=max(blob_bytes, B * (64 + 8 * _AUG_JOB.itemsize)), i=0, sets=[])
            for _ in range(4):
                ring[key]["sets"].append(dict(blob=torch.empty(ring[key]["cap"], dtype=torch.uint8).pin_memory(),
                                              luts=torch.empty((2 * B, 768), dtype=torch.uint8).pin_memory(), ev=None))
        r = ring[key]
        st = r["sets"][r["i"] % 4]
        r["i"] += 1
        if st["ev"] is not None:
            st["ev"].synchronize()
        return st

    def apply(self, imgs, plan):
        from ._lib import lib, check, stream as cur_stream
        if isinstance(imgs, np.ndarray):
            imgs = torch.from_numpy(imgs)
        imgs = imgs.cuda(non_blocking=True) if not imgs.is_cuda else imgs
        if imgs.dtype != torch.uint8 or imgs.dim() != 4 or imgs.shape[3] != 3:
            raise ValueError("expected a uint8 RGB batch [B,H,W,3]")
        imgs = imgs.contiguous()
        B, H, W, _ = imgs.shape
        S, dev = self.size, imgs.device
        hmax, wmax = int(plan.boxes[:, 2].max()), int(plan.boxes[:, 3].max())
        kmax = int(math.ceil(2.0 * max(1.0, max(hmax, wmax) / S))) * 2 + 1
        sc = self._scratch(dev, B, H, S, kmax)
        bufs = (sc["a"], sc["b"])
        cur = np.zeros(B, np.int64)                                        # which buffer holds image b right now
        base = [t.data_ptr() for t in bufs]
        img_bytes = S * S * 3
        # ---- job tables (host), one upload
        crop = np.zeros((B, 8), np.int32)
        crop[:, :4] = plan.boxes
        crop[:, 4] = plan.flips
        pin = self._pinned(dev, B, B * (crop.itemsize * 8 + len(plan.layers) * _AUG_JOB.itemsize + _FIN_JOB.itemsize))
        layer_jobs, luts_host, fallbacks = [], pin["luts"].numpy(), []
        slot = 0
        for l, layer in enumerate(plan.layers):
            jobs, need_stats = [], False
            for b, item in enumerate(layer):
                if item is None:
                    continue
                name, args = item
                j = np.zeros((), _AUG_JOB)
                j["slot"] = slot
                if name in _POINT:
                    if name == "PosterizeIncreasing" and args[0] >= 8:
                        continue                                           # Pillow returns the image unchanged
                    j["op"] = 0
                    luts_host[slot] = np.tile(point_table(_POINT[name], *args), 3)
                elif name in ("AutoContrast", "Equalize"):
                    j["op"], j["mode"], need_stats = 1, int(name == "Equalize"), True
                elif name in _ENHANCE_MODE:
                    j["op"], j["mode"], j["factor"] = 2, _ENHANCE_MODE[name], args[0]
                    need_stats |= name == "ContrastIncreasing"
                else:
                    if name == "Rotate":
                        m = rotation_matrix(S, S, args[0])
                        if m is None:                                      # exact multiples of 90 degrees: Pillow's transpose paths
                            fallbacks.append((l, b, args[0]))
                            continue
                    elif name == "ShearX":
                        m = (1, args[0], 0, 0, 1, 0)
                    elif name == "ShearY":
                        m = (1, 0, 0, args[0], 1, 0)
                    elif name == "TranslateXRel":
                        m = (1, 0, args[0] * S, 0, 1, 0)
                    else:
                        m = (1, 0, 0, 0, 1, args[0] * S)
                    j["op"], j["bicubic"], j["m"] = 3, 1, m
                    j["fill"][:3] = IMG_MEAN_FILL
                j["src"] = base[cur[b]] + b * img_bytes
                cur[b] ^= 1
                j["dst"] = base[cur[b]] + b * img_bytes
                jobs.append((b, j))
                slot += 1
            layer_jobs.append((jobs, need_stats))
        # erase noise only for the images that erase
        er = np.nonzero(plan.erase[:, 2] > 0)[0]
        noise = None
        if len(er):
            if self.gen is None:
                self.gen = torch.Generator(device=dev)
                if self.seed is not None:
                    self.gen.manual_seed(self.seed + 2)
            noise = torch.empty((len(er), 3, S, S), device=dev, dtype=torch.float32).normal_(generator=self.gen)
        # rotate fast paths are rare (the magnitude is continuous): run them between the layers on the single-image op
        table = [crop.tobytes()] + [b"".join(j.tobytes() for _, j in jobs) for jobs, _ in layer_jobs]
        offs = np.cumsum([0] + [len(t) for t in table])
        fin_off = int(offs[-1])
        blob = pin["blob"][:fin_off + B * _FIN_JOB.itemsize]
        host = blob.numpy()
        for t, o in zip(table, offs[:-1]):
            host[o:o + len(t)] = np.frombuffer(t, np.uint8)
        # the fin jobs depend on the final buffer of every image (known now that all layers are planned)
        for l, b, deg in fallbacks:
            cur[b] ^= 1
        fin = np.zeros(B, _FIN_JOB)
        fin["src"] = [base[cur[b]] + b * img_bytes for b in range(B)]
        fin["top"], fin["left"], fin["eh"], fin["ew"] = plan.erase.T
        for i, b in enumerate(er):
            fin["noise"][b] = noise[i].data_ptr()
        host[fin_off:] = np.frombuffer(fin.tobytes(), np.uint8)
        dblob = blob.to(dev, non_blocking=True)
        dluts = sc["luts"]
        dluts.copy_(pin["luts"], non_blocking=True)
        pin["ev"] = torch.cuda.Event()
        pin["ev"].record(torch.cuda.current_stream(dev))
        p0 = dblob.data_ptr()
        st = cur_stream()
        check(lib.trt_crop_resize_batch_u8(imgs.data_ptr(), B, H, W, p0, S, 1, kmax, hmax, sc["bounds"].data_ptr(),
                                           sc["coeffs"].data_ptr(), sc["tmp"].data_ptr(), sc["a"].data_ptr(), st))
        state = np.zeros(B, np.int64)
        for l, (jobs, need_stats) in enumerate(layer_jobs):
            if jobs:
                if need_stats:
                    sc["stats"].zero_()
                hist, luma = sc["stats"].data_ptr(), sc["stats"].data_ptr() + 2 * B * 768 * 8
                check(lib.trt_aug_layer_batch_u8(p0 + int(offs[1 + l]), len(jobs), S, int(need_stats), hist, luma, dluts.data_ptr(), st))
                for b, _ in jobs:
                    state[b] ^= 1
            for fl, b, deg in fallbacks:
                if fl == l:
                    src = bufs[state[b]][b]
                    state[b] ^= 1
                    bufs[state[b]][b].copy_(rotate(src, deg))
        out = torch.empty((B, 3, S, S), device=dev, dtype=self.dtype)
        check(lib.trt_normalize_erase_batch(p0 + fin_off, B, S, out.data_ptr(), int(self.dtype == torch.bfloat16), st))
        self._keep = (dblob, noise, imgs)             # alive until the next call: the launches above read them asynchronously
        return out

    def __call__(self, imgs):
        B, H, W = imgs.shape[:3]
        return self.apply(imgs, self.sample(B, H, W))
