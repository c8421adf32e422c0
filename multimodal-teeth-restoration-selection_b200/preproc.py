"""On-device input stage with the reference's function names.

    apply_clahe(img_bgr)            <- src/preprocessing/normalise.py:10-16
    centre_crop_resize(img, size)   <- src/preprocessing/pipeline.py:23-29
    normalize_flip(img_u8, flip)    <- ToTensor + Normalize (+ torch.flip) of train_mm_joint_dualtask.py:83-84,328-333
    InputStage                      <- the three chained on the device, batch at a time

Inputs may be numpy uint8 HWC arrays (the reference's type: copied to the GPU, result copied back) or CUDA uint8 tensors
[H,W,3] / [N,H,W,3] (result stays on the device).  All arithmetic runs in libteethrt kernels; there is no OpenCV/CPU
fallback."""
import numpy as np
import torch

from . import lab_tables
from ._lib import lib, check, ptr, stream, init

CLAHE_CLIP = 3.0          # src/config.py:15
CLAHE_TILEGR = (8, 8)     # src/config.py:16 (compiled into the kernels)
OUTPUT_SIZE = 512         # src/config.py:14

_tables = {}


def device_tables(device):
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _tables:
        init(key)
        _tables[key] = torch.from_numpy(lab_tables.build_packed()).to(f"cuda:{key}")
    return _tables[key]


def _to_dev(img):
    """-> (uint8 CUDA tensor [N,H,W,3], was_numpy, had_batch_dim)"""
    was_np = isinstance(img, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(img)).cuda() if was_np else img
    if t.dtype != torch.uint8 or t.shape[-1] != 3 or t.dim() not in (3, 4):
        raise ValueError("expected a uint8 image [H,W,3] or batch [N,H,W,3]")
    batched = t.dim() == 4
    if not batched:
        t = t.unsqueeze(0)
    return t.contiguous(), was_np, batched


def _back(t, was_np, batched):
    if not batched:
        t = t[0]
    return t.cpu().numpy() if was_np else t


def apply_clahe(img_bgr, clip=CLAHE_CLIP):
    x, was_np, batched = _to_dev(img_bgr)
    n, h, w, _ = x.shape
    out = torch.empty_like(x)
    ws = torch.empty(lib.trt_clahe_workspace_bytes(n), device=x.device, dtype=torch.uint8)
    check(lib.trt_clahe_bgr_u8(ptr(x), ptr(out), n, h, w, clip, ptr(device_tables(x.device)), ptr(ws), ws.numel(), stream()))
    return _back(out, was_np, batched)


def centre_crop_resize(img, size=OUTPUT_SIZE):
    x, was_np, batched = _to_dev(img)
    n, h, w, _ = x.shape
    out = torch.empty((n, size, size, 3), device=x.device, dtype=torch.uint8)
    check(lib.trt_resize_linear_u8(ptr(x), ptr(out), n, h, w, 1, size, size, stream()))
    return _back(out, was_np, batched)


def normalize_flip(img_bgr_u8, flip=0, dtype=torch.float32):
    """uint8 BGR HWC -> RGB CHW (u8/255 - mean)/std on the device; flip 0/1/2 = none / W-reverse / H-reverse."""
    x, _, batched = _to_dev(img_bgr_u8)
    n, h, w, _ = x.shape
    out = torch.empty((n, 3, h, w), device=x.device, dtype=dtype)
    check(lib.trt_normalize_flip_u8(ptr(x), ptr(out), n, h, w, flip, int(dtype == torch.bfloat16), stream()))
    return out if batched else out[0]


class InputStage:
    """CLAHE -> centre crop + resize -> normalise (+flip) for a fixed batch geometry, buffers allocated once."""

    def __init__(self, n, h, w, size=224, device="cuda", dtype=torch.bfloat16, clahe=True):
        self.n, self.h, self.w, self.size, self.dtype, self.clahe = n, h, w, size, dtype, clahe
        dev = torch.device(device)
        self.tables = device_tables(dev)
        self.buf_clahe = torch.empty((n, h, w, 3), device=dev, dtype=torch.uint8)
        self.buf_small = torch.empty((n, size, size, 3), device=dev, dtype=torch.uint8)
        self.out = torch.empty((n, 3, size, size), device=dev, dtype=dtype)
        self.ws = torch.empty(lib.trt_clahe_workspace_bytes(n), device=dev, dtype=torch.uint8)

    def __call__(self, x_u8, flip=0):
        n, h, w = self.n, self.h, self.w
        assert x_u8.shape == (n, h, w, 3) and x_u8.dtype == torch.uint8 and x_u8.is_cuda and x_u8.is_contiguous()
        src = x_u8
        if self.clahe:
            check(lib.trt_clahe_bgr_u8(ptr(src), ptr(self.buf_clahe), n, h, w, CLAHE_CLIP, ptr(self.tables), ptr(self.ws),
                                       self.ws.numel(), stream()))
            src = self.buf_clahe
        check(lib.trt_resize_linear_u8(ptr(src), ptr(self.buf_small), n, h, w, 1, self.size, self.size, stream()))
        check(lib.trt_normalize_flip_u8(ptr(self.buf_small), ptr(self.out), n, self.size, self.size, flip,
                                        int(self.dtype == torch.bfloat16), stream()))
        return self.out
