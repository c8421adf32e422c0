"""On-device input stage with the reference's function names.

    apply_clahe(img_bgr)            <- src/preprocessing/normalise.py:10-16
    centre_crop_resize(img, size)   <- src/preprocessing/pipeline.py:23-29
    normalize_flip(img_u8, flip)    <- ToTensor + Normalize (+ torch.flip) of train_mm_joint_dualtask.py:83-84,328-333
    InputStage                      <- the three chained on the device, batch at a time
    resize_center_crop(img, R, S)   <- PIL Resize + CenterCrop of the eval transforms (infer_mm.py:12-17, infer_mil.py:116-119)
    deskew(img_bgr)                 <- src/preprocessing/normalise.py:19-57 (BGR2GRAY, Canny, PCA angle, warpAffine)

Inputs may be numpy uint8 HWC arrays (the reference's type: copied to the GPU, result copied back) or CUDA uint8 tensors
[H,W,3] / [N,H,W,3] (result stays on the device).  All arithmetic runs in libteethrt kernels; there is no OpenCV/CPU
fallback."""
import ctypes as C
import functools
import math

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


# ------------------------------------------------------------------------------------------------ PIL-exact eval transform
# SURVEY.md §8 row f2 (deterministic half): Resize(shorter edge -> R, antialiased PIL filter) + CenterCrop(S) on the device,
# bit-identical to torchvision-on-PIL, which is what timm.create_transform(is_training=False) builds for the reference
# (ui/gradio_app/infer_mm.py:12-17: bicubic, R = floor(S / 0.875); ui/gradio_app/infer_mil.py:116-119: bilinear 512 -> 480).
_PIL_PRECISION_BITS = 32 - 8 - 2


def _bicubic_filter(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear_filter(x):
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


_PIL_FILTERS = {"bicubic": (_bicubic_filter, 2.0), "bilinear": (_bilinear_filter, 1.0)}


@functools.lru_cache(maxsize=64)
def pil_coeffs(in_size, out_size, interpolation="bicubic"):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for a full-axis resize in_size -> out_size, in the same double
    arithmetic and evaluation order -> (bounds int32 [out,2] = first tap / tap count, coeffs int32 [out,ksize], ksize)."""
    filt, fsupport = _PIL_FILTERS[interpolation]
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    one = float(1 << _PIL_PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ws, ww = [], 0.0
        for x in range(xmax):
            w = filt((x + xmin - center + 0.5) * ss)
            ws.append(w)
            ww += w
        for x in range(xmax):
            w = ws[x] / ww if ww != 0.0 else ws[x]
            coeffs[xx, x] = int(-0.5 + w * one) if w < 0 else int(0.5 + w * one)
        bounds[xx] = (xmin, xmax)
    return bounds, coeffs, ksize


def resized_size(h, w, short):
    """torchvision _compute_resized_output_size for Resize(int): shorter edge -> `short`, longer = int(short * long / short0)."""
    if w <= h:
        return int(short * h / w), short
    return short, int(short * w / h)


_coeff_dev = {}


def _coeffs_on(device, in_size, out_size, interpolation, lo, hi):
    """Device copies of the table rows [lo, hi) (the crop window), cached per geometry."""
    key = (str(device), in_size, out_size, interpolation, lo, hi)
    if key not in _coeff_dev:
        b, c, ksize = pil_coeffs(in_size, out_size, interpolation)
        b, c = b[lo:hi].copy(), c[lo:hi].copy()
        first = int(b[:, 0].min())
        last = int((b[:, 0] + b[:, 1]).max())
        b[:, 0] -= first
        if len(_coeff_dev) > 256:
            _coeff_dev.clear()
        _coeff_dev[key] = (torch.from_numpy(b).to(device), torch.from_numpy(c).to(device), ksize, first, last)
    return _coeff_dev[key]


def _as_u8_image(img_u8):
    t = torch.from_numpy(np.array(img_u8, copy=True)).cuda() if isinstance(img_u8, np.ndarray) else img_u8
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] not in (1, 3) or not t.is_cuda:
        raise ValueError("expected a uint8 image [H,W,1|3] (numpy array or CUDA tensor)")
    return t.contiguous()


def _resample_window(t, box, out_hw, window, interpolation, swap_channels):
    """Pillow resize of the source box (top, left, h, w) of `t` to out_hw, of which only `window` (top, left, h, w in
    output coordinates) is computed: horizontal pass over the source rows those output rows need, then vertical pass."""
    top, left, h, w = box
    wt, wl, wh, ww = window
    ch, pitch = t.shape[2], t.shape[1] * t.shape[2]
    vb, vc, vk, r0, r1 = _coeffs_on(t.device, h, out_hw[0], interpolation, wt, wt + wh)
    hb, hc, hk, c0, _ = _coeffs_on(t.device, w, out_hw[1], interpolation, wl, wl + ww)
    tmp = torch.empty((r1 - r0, ww, ch), device=t.device, dtype=torch.uint8)
    out = torch.empty((wh, ww, ch), device=t.device, dtype=torch.uint8)
    src = t[top + r0:, left + c0:]                                      # tables are relative to (r0, c0) inside the box
    check(lib.trt_resample_u8(src.data_ptr(), pitch, ch, ptr(tmp), r1 - r0, ww, ptr(hb), ptr(hc), hk, 0, 0, stream()))
    check(lib.trt_resample_u8(ptr(tmp), ww * ch, ch, ptr(out), wh, ww, ptr(vb), ptr(vc), vk, 1, int(swap_channels), stream()))
    return out


def resize_center_crop(img_u8, short, crop, interpolation="bicubic", swap_channels=False):
    """transforms.Resize(short, interpolation) -> transforms.CenterCrop(crop) on a uint8 [H,W,C] image (numpy or CUDA
    tensor, C in {1,3}) -> CUDA uint8 [crop, crop, C], bit-identical to the PIL path.  Only the crop window is computed:
    the horizontal pass runs over the source rows the cropped output rows need and the cropped columns, the vertical pass
    over the result.  swap_channels writes the channels reversed (RGB <-> BGR)."""
    t = _as_u8_image(img_u8)
    h, w, _ = t.shape
    nh, nw = resized_size(h, w, short)
    if crop > nh or crop > nw:
        raise ValueError(f"crop {crop} larger than the resized image {nh}x{nw} (the reference would zero-pad; not built)")
    top, left = int(round((nh - crop) / 2.0)), int(round((nw - crop) / 2.0))
    return _resample_window(t, (0, 0, h, w), (nh, nw), (top, left, crop, crop), interpolation, swap_channels)


def resized_crop(img_u8, top, left, height, width, size, interpolation="bicubic", swap_channels=False):
    """torchvision F.resized_crop on a PIL image (= img.crop(box).resize(size)): the deterministic part of timm's
    RandomResizedCropAndInterpolation in the train transform (train_mm_joint_dualtask.py:75-84), given the sampled box.
    size: int or (h, w).  Bit-identical to the PIL path; the box must lie inside the image."""
    t = _as_u8_image(img_u8)
    H, W, _ = t.shape
    oh, ow = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
    if top < 0 or left < 0 or height <= 0 or width <= 0 or top + height > H or left + width > W:
        raise ValueError(f"crop box ({top}, {left}, {height}, {width}) outside the {H}x{W} image (PIL would zero-pad; not built)")
    return _resample_window(t, (top, left, height, width), (oh, ow), (0, 0, oh, ow), interpolation, swap_channels)


# ------------------------------------------------------------------------------------------------ deskew
ROT_TOLERANCE = 15        # src/config.py:17


def canny(img_bgr, low=50, high=150, want_edges=True):
    """cv2.Canny(cv2.cvtColor(img, COLOR_BGR2GRAY), low, high) on the device, bit-identical
    -> (edges uint8 [H,W] CUDA tensor or None, moments int64 [6] = N, Σy, Σx, Σyy, Σxy, Σxx of the edge pixels)."""
    x, _, batched = _to_dev(img_bgr)
    if batched:
        raise ValueError("canny works on one image [H,W,3]")
    x = x[0]
    h, w, _ = x.shape
    label = torch.empty((h, w), device=x.device, dtype=torch.uint8)
    changed = torch.zeros(4, device=x.device, dtype=torch.int32)      # one flag per pass of a round
    check(lib.trt_canny_nms_bgr_u8(ptr(x), h, w, int(low), int(high), ptr(label), stream()))
    for _ in range(4 * (h + w)):                       # bound: a path can cross at most every tile once per pass
        changed.zero_()
        for i in range(4):
            check(lib.trt_canny_hysteresis_pass(ptr(label), h, w, changed.data_ptr() + 4 * i, stream()))
        if int(changed[3].item()) == 0:                # the round's last pass promoted nothing: fixed point reached
            break
    edges = torch.empty((h, w), device=x.device, dtype=torch.uint8) if want_edges else None
    mom = torch.empty(6, device=x.device, dtype=torch.int64)
    check(lib.trt_canny_finish(ptr(label), h, w, ptr(edges), ptr(mom), stream()))
    return edges, mom


def rotation_matrix_2d(center, angle_deg, scale=1.0):
    """cv2.getRotationMatrix2D in the same double arithmetic."""
    a = angle_deg * (math.pi / 180)            # OpenCV: angle *= CV_PI/180 (the constant is folded first)
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * center[0] - beta * center[1]],
                     [-beta, alpha, beta * center[0] + (1 - alpha) * center[1]]], dtype=np.float64)


def invert_affine(M):
    """The inversion cv2.warpAffine applies to a forward map (no WARP_INVERSE_MAP), statement for statement."""
    m = [float(v) for v in np.asarray(M, dtype=np.float64).ravel()]
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_affine(img_u8, M, dsize):
    """cv2.warpAffine(img, M, (w, h), flags=INTER_LINEAR, borderMode=BORDER_REPLICATE), bit-identical; [H,W,1|3] uint8."""
    t = torch.from_numpy(np.array(img_u8, copy=True)).cuda() if isinstance(img_u8, np.ndarray) else img_u8.contiguous()
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] not in (1, 3):
        raise ValueError("expected a uint8 image [H,W,1|3]")
    dw, dh = int(dsize[0]), int(dsize[1])
    out = torch.empty((dh, dw, t.shape[2]), device=t.device, dtype=torch.uint8)
    inv = (C.c_double * 6)(*invert_affine(M))
    check(lib.trt_warp_affine_linear_u8(ptr(t), t.shape[0], t.shape[1], t.shape[2], ptr(out), dh, dw, inv, stream()))
    return out


def edge_angle(moments):
    """normalise.py:31-43 from the exact integer moments: covariance of the (y, x) edge coordinates (N-1 normalisation, as
    np.cov), principal axis by np.linalg.eigh — the same LAPACK call, so the same eigenvector sign — angle in degrees."""
    n, sy, sx, syy, sxy, sxx = (int(v) for v in moments)
    cyy = (syy * n - sy * sy) / (n * (n - 1))           # exact integers until the final division
    cxy = (sxy * n - sx * sy) / (n * (n - 1))
    cxx = (sxx * n - sx * sx) / (n * (n - 1))
    eigvals, eigvecs = np.linalg.eigh(np.array([[cyy, cxy], [cxy, cxx]], dtype=np.float64))
    principal = eigvecs[:, np.argmax(eigvals)]
    return float(np.rad2deg(np.arctan2(principal[0], principal[1])))


def deskew(img_bgr):
    """src/preprocessing/normalise.py:19-57 -> (rotated image, applied angle in degrees); numpy in -> numpy out, CUDA
    tensor in -> CUDA tensor out.  Fewer than 10 edge points or |angle| < ROT_TOLERANCE returns the input and 0.0."""
    was_np = isinstance(img_bgr, np.ndarray)
    x, _, batched = _to_dev(img_bgr)
    if batched:
        raise ValueError("deskew works on one image [H,W,3]")
    _, mom = canny(x[0], 50, 150, want_edges=False)
    mom = mom.cpu().numpy()
    if int(mom[0]) < 10:
        return img_bgr, 0.0
    angle = edge_angle(mom)
    if abs(angle) < ROT_TOLERANCE:
        return img_bgr, 0.0
    h, w = x.shape[1], x.shape[2]
    out = warp_affine(x[0], rotation_matrix_2d((w / 2, h / 2), angle, 1.0), (w, h))
    return (out.cpu().numpy() if was_np else out), angle


MIN_EDGE_PX = 400         # src/config.py:13


def preprocess_image(img_bgr, rotate=True, size=OUTPUT_SIZE):
    """The image path of Preprocessor.process_file without segmentation (src/preprocessing/pipeline.py:80-117, crop_mode
    'none'): size check -> CLAHE -> deskew (optional) -> centre_crop_resize, chained on the device with one small read-back
    (the edge moments).  -> (uint8 [size,size,3], info dict with the reference's keys); numpy in -> numpy out."""
    was_np = isinstance(img_bgr, np.ndarray)
    x, _, batched = _to_dev(img_bgr)
    if batched:
        raise ValueError("preprocess_image works on one image [H,W,3]")
    if min(x.shape[1], x.shape[2]) < MIN_EDGE_PX:
        raise ValueError("Image too small (<400 px)")
    img = apply_clahe(x[0])
    info = {"rotation_deg": 0.0, "crop_mode": "none"}
    if rotate:
        img, info["rotation_deg"] = deskew(img)
    out = centre_crop_resize(img, size)
    return (out.cpu().numpy() if was_np else out), info
