"""ORACLE (test infrastructure): the reference's on-CPU input stage.

Two forms of each function:
  * `*_cv2`  — the reference's own call sequence on OpenCV (src/preprocessing/normalise.py:10-16,
    src/preprocessing/pipeline.py:23-29).  OpenCV 4.13.0 is part of this image (and of the GPU box's),
    so this is the reference's arithmetic itself, not a port.
  * `*_np`   — a numpy restatement of OpenCV's published integer/fixed-point algorithm (SURVEY.md
    App. A), the spec the CUDA kernels and the device LUTs are written from.  Pinned to `*_cv2`
    exhaustively over all 2^24 colours and on the image set of SURVEY.md §8d (tests/test_oracle_preproc.py).

The table builders here are also what the product uploads to the GPU?  NO — the product has its own copy in
`<pkg>/lab_tables.py`; this file is checker-only.
"""
import hashlib

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

CLAHE_CLIP = 3.0          # src/config.py:15
CLAHE_TILEGR = (8, 8)     # src/config.py:16
OUTPUT_SIZE = 512         # src/config.py:14

f32 = np.float32


def _rint(x):
    return np.rint(x).astype(np.int64)


def descale(x, n):
    return (x + (1 << (n - 1))) >> n


# ------------------------------------------------------------------ tables (App. A.1 / A.3)
def srgb_gamma_tab_b():
    i = np.arange(256, dtype=f32)
    x = (i / f32(255)).astype(f32)
    xd = x.astype(np.float64)
    g = np.where(x <= f32(0.04045), xd / 12.92, ((xd + 0.055) / 1.055) ** 2.4)
    return _rint(f32(2040) * g.astype(f32)).astype(np.int32)


def lab_cbrt_tab_b():
    j = np.arange(3072, dtype=f32)
    x = (j / f32(2040)).astype(f32)
    lin = (x * f32(841.0 / 108.0) + f32(16.0 / 116.0)).astype(f32)
    cb = np.cbrt(x).astype(f32)  # numpy cbrt on float32 input = cbrtf
    f = np.where(x < f32(216.0 / 24389.0), lin, cb).astype(f32)
    return _rint(f32(32768) * f).astype(np.int32)


def lab_to_yf_b():
    BASE = 16384
    out = np.zeros(512, dtype=np.int32)
    for L in range(256):
        if L <= 20:
            y = int(np.rint(f32(L * BASE * 20 * 9) / f32(17 * 29 ** 3)))
            ify = int(np.rint(f32(BASE) * (f32(16) / f32(116) + f32(L * 5) / f32(3 * 17 * 29))))
        else:
            fy = f32(f32(L * 100 * BASE) / f32(255 * 116)) + f32(f32(16 * BASE) / f32(116))
            ify = int(np.rint(fy))
            y = int(np.rint(f32(f32(f32(fy * fy) * fy) / f32(BASE * BASE))))
        out[2 * L], out[2 * L + 1] = y, ify
    return out


def ab_to_xz_b():
    BASE, minAB = 16384, -8145
    i = np.arange(minAB, minAB + 36864, dtype=np.int64)

    def cdiv(a, b):  # C truncating division
        return np.where(a >= 0, a // b, -((-a) // b))

    lo = cdiv(i * 108, 841) - (BASE * 16 // 116) * 108 // 841
    hi = cdiv(cdiv(i * i, BASE) * i, BASE)
    return np.where(i <= 3390, lo, hi).astype(np.int32)


def srgb_inv_gamma_tab_b():
    t = np.arange(4096, dtype=f32)
    x = (t / f32(4096)).astype(f32)
    xd = x.astype(np.float64)
    g = np.where(x <= f32(0.0031308), 12.92 * xd, 1.055 * np.power(xd, 1.0 / 2.4) - 0.055)
    return _rint(f32(255) * g.astype(f32)).astype(np.int32)


def table_sha(t):
    return hashlib.sha1(np.ascontiguousarray(t, dtype='<i4').tobytes()).hexdigest()[:16]


_T = {}


def tables():
    if not _T:
        _T.update(gamma=srgb_gamma_tab_b(), cbrt=lab_cbrt_tab_b(), yf=lab_to_yf_b(), abxz=ab_to_xz_b(),
                  invgamma=srgb_inv_gamma_tab_b())
    return _T


# ------------------------------------------------------------------ colour conversions
def bgr2lab_np(img):
    """cv2.cvtColor(BGR2LAB) on uint8 (normalise.py:11), integer path of App. A.1."""
    t = tables()
    b = t["gamma"][img[..., 0]].astype(np.int64)
    g = t["gamma"][img[..., 1]].astype(np.int64)
    r = t["gamma"][img[..., 2]].astype(np.int64)
    cb = t["cbrt"].astype(np.int64)
    fX = cb[descale(r * 1777 + g * 1541 + b * 778, 12)]
    fY = cb[descale(r * 871 + g * 2929 + b * 296, 12)]
    fZ = cb[descale(r * 73 + g * 448 + b * 3575, 12)]
    L = descale(296 * fY - 1336934, 15)
    a = descale(500 * (fX - fY) + 128 * 32768, 15)
    bb = descale(200 * (fY - fZ) + 128 * 32768, 15)
    return np.clip(np.stack([L, a, bb], -1), 0, 255).astype(np.uint8)


def lab2bgr_np(lab):
    """cv2.cvtColor(LAB2BGR) on uint8 (normalise.py:16), integer path of App. A.3."""
    t = tables()
    BASE, minAB = 16384, -8145
    L = lab[..., 0].astype(np.int64)
    a = lab[..., 1].astype(np.int64)
    b = lab[..., 2].astype(np.int64)
    y = t["yf"][2 * L].astype(np.int64)
    ify = t["yf"][2 * L + 1].astype(np.int64)
    adiv = ((5 * a * 53687 + 128) >> 13) - 128 * BASE // 500
    bdiv = ((b * 41943 + 16) >> 9) - 128 * BASE // 200 + 1
    X = t["abxz"][ify + adiv - minAB].astype(np.int64)
    Z = t["abxz"][ify - bdiv - minAB].astype(np.int64)
    ro = np.clip(descale(12615 * X - 6296 * y - 2223 * Z, 14), 0, 4095)
    go = np.clip(descale(-3773 * X + 7684 * y + 185 * Z, 14), 0, 4095)
    bo = np.clip(descale(217 * X - 836 * y + 4715 * Z, 14), 0, 4095)
    ig = t["invgamma"]
    return np.stack([ig[bo], ig[go], ig[ro]], -1).astype(np.uint8)


# ------------------------------------------------------------------ CLAHE (App. A.2)
def reflect101_pad(src, pad_b, pad_r):
    return np.pad(src, ((0, pad_b), (0, pad_r)), mode="reflect")


def clahe_luts_np(L, clip=CLAHE_CLIP, grid=CLAHE_TILEGR):
    H, W = L.shape
    gy, gx = grid[1], grid[0]
    if H % gy or W % gx:
        src = reflect101_pad(L, gy - (H % gy), gx - (W % gx))
    else:
        src = L
    Hp, Wp = src.shape
    th, tw = Hp // gy, Wp // gx
    area = th * tw
    clip_limit = max(int(clip * area / 256), 1)
    lut_scale = f32(255) / f32(area)
    luts = np.zeros((gy, gx, 256), dtype=np.uint8)
    for ty in range(gy):
        for tx in range(gx):
            tile = src[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            h = np.bincount(tile.ravel(), minlength=256).astype(np.int64)
            clipped = int(np.maximum(h - clip_limit, 0).sum())
            h = np.minimum(h, clip_limit)
            h += clipped // 256
            residual = clipped % 256
            if residual:
                step = max(256 // residual, 1)
                i = 0
                while i < 256 and residual > 0:
                    h[i] += 1
                    i += step
                    residual -= 1
            cs = np.cumsum(h).astype(f32)
            luts[ty, tx] = np.clip(_rint((cs * lut_scale).astype(f32)), 0, 255).astype(np.uint8)
    return luts, th, tw


def clahe_np(L, clip=CLAHE_CLIP, grid=CLAHE_TILEGR):
    """cv2.createCLAHE(clip, grid).apply(L) (normalise.py:13-14)."""
    H, W = L.shape
    luts, th, tw = clahe_luts_np(L, clip, grid)
    gy, gx = grid[1], grid[0]
    inv_th, inv_tw = f32(1.0) / f32(th), f32(1.0) / f32(tw)
    tyf = (np.arange(H, dtype=f32) * inv_th - f32(0.5)).astype(f32)
    txf = (np.arange(W, dtype=f32) * inv_tw - f32(0.5)).astype(f32)
    ty1 = np.floor(tyf).astype(np.int64)
    tx1 = np.floor(txf).astype(np.int64)
    ya = (tyf - ty1.astype(f32)).astype(f32)
    xa = (txf - tx1.astype(f32)).astype(f32)
    ya1, xa1 = (f32(1) - ya).astype(f32), (f32(1) - xa).astype(f32)
    ty2 = np.minimum(ty1 + 1, gy - 1)
    tx2 = np.minimum(tx1 + 1, gx - 1)
    ty1 = np.maximum(ty1, 0)
    tx1 = np.maximum(tx1, 0)
    v = L.astype(np.int64)
    Y1, Y2 = ty1[:, None], ty2[:, None]
    X1, X2 = tx1[None, :], tx2[None, :]
    l11 = luts[Y1, X1, v].astype(f32)
    l12 = luts[Y1, X2, v].astype(f32)
    l21 = luts[Y2, X1, v].astype(f32)
    l22 = luts[Y2, X2, v].astype(f32)
    XA, XA1 = xa[None, :], xa1[None, :]
    YA, YA1 = ya[:, None], ya1[:, None]
    top = ((l11 * XA1).astype(f32) + (l12 * XA).astype(f32)).astype(f32)
    bot = ((l21 * XA1).astype(f32) + (l22 * XA).astype(f32)).astype(f32)
    res = ((top * YA1).astype(f32) + (bot * YA).astype(f32)).astype(f32)
    return np.clip(_rint(res), 0, 255).astype(np.uint8)


def apply_clahe_np(img_bgr):
    """src/preprocessing/normalise.py:10-16."""
    lab = bgr2lab_np(img_bgr)
    lab2 = lab.copy()
    lab2[..., 0] = clahe_np(lab[..., 0])
    return lab2bgr_np(lab2)


def apply_clahe_cv2(img_bgr):
    """src/preprocessing/normalise.py:10-16, the reference's own OpenCV call sequence."""
    lab = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2LAB)
    l, a, b = cv2.split(lab)
    l2 = cv2.createCLAHE(clipLimit=CLAHE_CLIP, tileGridSize=CLAHE_TILEGR).apply(l)
    return cv2.cvtColor(cv2.merge((l2, a, b)), cv2.COLOR_LAB2BGR)


# ------------------------------------------------------------------ resize (App. A.4)
def _resize_axis(ssize, dsize, zero_edges):
    scale = np.float64(ssize) / np.float64(dsize)
    d = np.arange(dsize, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)
    s = np.floor(fx).astype(np.int64)
    fx = (fx - s.astype(f32)).astype(f32)
    if zero_edges:
        lo = s < 0
        fx[lo] = 0
        s[lo] = 0
        hi = s >= ssize - 1
        fx[hi] = 0
        s[hi] = ssize - 1
    w1 = _rint((fx * f32(2048)).astype(f32))
    w0 = _rint(((f32(1) - fx).astype(f32) * f32(2048)).astype(f32))
    return s, w0, w1


def resize_linear_np(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=INTER_LINEAR) on uint8 HWC (pipeline.py:29), App. A.4."""
    sh, sw = src.shape[:2]
    sx, wx0, wx1 = _resize_axis(sw, dw, True)
    sy, wy0, wy1 = _resize_axis(sh, dh, False)
    sx1 = np.minimum(sx + 1, sw - 1)
    y0 = np.clip(sy, 0, sh - 1)
    y1 = np.clip(sy + 1, 0, sh - 1)
    S = src.astype(np.int64)
    if S.ndim == 2:
        S = S[..., None]
    w0 = wx0[None, :, None]
    w1 = wx1[None, :, None]
    r0 = S[y0][:, sx] * w0 + S[y0][:, sx1] * w1
    r1 = S[y1][:, sx] * w0 + S[y1][:, sx1] * w1
    b0 = wy0[:, None, None]
    b1 = wy1[:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if src.ndim == 3 else out[..., 0]


def centre_crop(img):
    """src/preprocessing/pipeline.py:25-28."""
    h, w = img.shape[:2]
    d = min(h, w)
    y0, x0 = (h - d) // 2, (w - d) // 2
    return img[y0:y0 + d, x0:x0 + d]


def centre_crop_resize_np(img, size=OUTPUT_SIZE):
    return resize_linear_np(centre_crop(img), size, size)


def centre_crop_resize_cv2(img, size=OUTPUT_SIZE):
    """src/preprocessing/pipeline.py:23-29."""
    return cv2.resize(np.ascontiguousarray(centre_crop(img)), (size, size), interpolation=cv2.INTER_LINEAR)


# ------------------------------------------------------------------ ToTensor + Normalize + flips
MEAN = np.array([0.485, 0.456, 0.406], dtype=f32)
STD = np.array([0.229, 0.224, 0.225], dtype=f32)


def normalize_flip_np(img_bgr_u8, flip=0):
    """ToTensor (u8/255 in fp32) + Normalize(mean,std) on RGB CHW (train_mm_joint_dualtask.py:83-84,91-92),
    optional torch.flip dims=[3] (flip=1, W-reverse) / dims=[2] (flip=2, H-reverse) (:328-333).
    Input is BGR HWC as produced by apply_clahe; the RGB swap is SURVEY.md q12."""
    rgb = img_bgr_u8[..., ::-1].astype(f32)
    x = ((rgb / f32(255)).astype(f32) - MEAN) / STD
    x = np.ascontiguousarray(x.transpose(2, 0, 1).astype(f32))
    if flip == 1:
        x = x[:, :, ::-1]
    elif flip == 2:
        x = x[:, ::-1, :]
    return np.ascontiguousarray(x)


# ------------------------------------------------------------------ the §8d synthetic image set
def image_set(name, h=1024, w=1024, seed=0):
    rng = np.random.default_rng(seed)
    if name == "noise":
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    if name == "smooth":
        n = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        return cv2.GaussianBlur(n, (0, 0), 8)
    if name == "radiograph":
        n = rng.random((h, w)).astype(np.float32)
        n = cv2.GaussianBlur(n, (0, 0), 15)
        n = (n - n.min()) / max(float(n.max() - n.min()), 1e-12)
        g = np.repeat((n * 255.0)[..., None], 3, axis=2) + rng.normal(0, 3, size=(h, w, 3))
        return np.clip(np.rint(g), 0, 255).astype(np.uint8)
    if name.startswith("const"):
        return np.full((h, w, 3), int(name[5:]), dtype=np.uint8)
    if name == "ramp":
        r = (np.arange(w) * 255 // max(w - 1, 1)).astype(np.uint8)
        return np.ascontiguousarray(np.broadcast_to(r[None, :, None], (h, w, 3)))
    raise KeyError(name)


# ------------------------------------------------------------------ deskew (src/preprocessing/normalise.py:19-57)
ROT_TOLERANCE = 15        # src/config.py:17


def deskew_cv2(img_bgr):
    """The reference's own call sequence (normalise.py:19-57) -> (image, angle, edges)."""
    gray = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2GRAY)
    edges = cv2.Canny(gray, 50, 150)
    coords = np.column_stack(np.where(edges > 0))
    if coords.shape[0] < 10:
        return img_bgr, 0.0, edges
    centered = coords - coords.mean(axis=0)
    cov = np.cov(centered, rowvar=False)
    eigvals, eigvecs = np.linalg.eigh(cov)
    principal = eigvecs[:, np.argmax(eigvals)]
    angle_deg = np.rad2deg(np.arctan2(principal[0], principal[1]))
    if abs(angle_deg) < ROT_TOLERANCE:
        return img_bgr, 0.0, edges
    (h, w) = img_bgr.shape[:2]
    M = cv2.getRotationMatrix2D((w / 2, h / 2), angle_deg, 1.0)
    return cv2.warpAffine(img_bgr, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE), float(angle_deg), edges


def gray_np(img):
    """OpenCV 8-bit BGR2GRAY: 15-bit fixed point, coefficients 3735 / 19235 / 9798."""
    b, g, r = [img[..., i].astype(np.int64) for i in range(3)]
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def canny_np(gray, low=50, high=150):
    """cv2.Canny(gray, low, high) (aperture 3, L1 magnitude): Sobel with replicated border, magnitudes outside the image 0,
    OpenCV's TG22 fixed-point sector test with its asymmetric comparisons, hysteresis = 8-connected components of the
    candidates that hold a pixel above `high`."""
    import scipy.ndimage as ndi
    p = np.pad(gray.astype(np.int32), 1, mode='edge')
    dx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    dy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    m = np.abs(dx) + np.abs(dy)
    H, W = gray.shape
    mp = np.pad(m, 1, mode='constant')
    c = lambda oy, ox: mp[1 + oy:1 + oy + H, 1 + ox:1 + ox + W]  # noqa: E731
    x = np.abs(dx).astype(np.int64)
    y = np.abs(dy).astype(np.int64) << 15
    tg22x = x * 13573
    tg67x = tg22x + (x << 16)
    horiz = y < tg22x
    vert = ~horiz & (y > tg67x)
    diag = ~horiz & ~vert
    nm_d = np.where((dx ^ dy) < 0, (m > c(-1, 1)) & (m > c(1, -1)), (m > c(-1, -1)) & (m > c(1, 1)))
    cand = (m > low) & ((horiz & (m > c(0, -1)) & (m >= c(0, 1))) | (vert & (m > c(-1, 0)) & (m >= c(1, 0))) | (diag & nm_d))
    lab, n = ndi.label(cand, structure=np.ones((3, 3)))
    keep = np.zeros(n + 1, bool)
    keep[np.unique(lab[cand & (m > high)])] = True
    keep[0] = False
    return (keep[lab] * 255).astype(np.uint8)


def rotation_matrix_np(center, angle_deg):
    import math
    a = angle_deg * (math.pi / 180)            # OpenCV: angle *= CV_PI/180 (the constant is folded first)
    alpha, beta = math.cos(a), math.sin(a)
    return np.array([[alpha, beta, (1 - alpha) * center[0] - beta * center[1]], [-beta, alpha, beta * center[0] + (1 - alpha) * center[1]]])


def warp_affine_np(img, M, w, h):
    """cv2.warpAffine(img, M, (w, h), INTER_LINEAR, BORDER_REPLICATE): inverse map, 10-bit fixed-point coordinates rounded
    half-to-even, 5-bit sub-pixel position, 15-bit weights from OpenCV's saturated short table."""
    M = np.asarray(M, np.float64).copy().ravel()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2] = b1; M[5] = b2
    rnd = lambda v: np.rint(v).astype(np.int64)  # noqa: E731
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    adelta, bdelta = rnd(M[0] * xs * 1024), rnd(M[3] * xs * 1024)
    X0, Y0 = rnd((M[1] * ys + M[2]) * 1024) + 16, rnd((M[4] * ys + M[5]) * 1024) + 16
    X, Y = (X0[:, None] + adelta[None, :]) >> 5, (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
    w00, w01, w10, w11 = (32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32
    sat = (fx == 0) & (fy == 0)
    w00, w11 = np.where(sat, 32767, w00), np.where(sat, 1, w11)
    H, W = img.shape[:2]
    x0, x1, y0, y1 = np.clip(sx, 0, W - 1), np.clip(sx + 1, 0, W - 1), np.clip(sy, 0, H - 1), np.clip(sy + 1, 0, H - 1)
    I = img.astype(np.int64)
    out = (I[y0, x0] * w00[..., None] + I[y0, x1] * w01[..., None] + I[y1, x0] * w10[..., None] + I[y1, x1] * w11[..., None]
           + (1 << 14)) >> 15
    return np.clip(out, 0, 255).astype(np.uint8)


def tooth_image(h, w, seed, angle):
    """A rotated bright body with a row of bars on a dark noisy background: something Canny + PCA can orient."""
    import math
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    cy, cx = h / 2 + rng.uniform(-20, 20), w / 2 + rng.uniform(-20, 20)
    t = math.radians(angle)
    u = (xx - cx) * math.cos(t) + (yy - cy) * math.sin(t)
    v = -(xx - cx) * math.sin(t) + (yy - cy) * math.cos(t)
    body = ((u / (0.38 * w)) ** 2 + (v / (0.16 * h)) ** 2 < 1).astype(np.float64)
    bars = ((np.abs(v) < 0.05 * h) & (np.abs(u) < 0.3 * w) & ((u // 25) % 2 == 0)).astype(np.float64)
    img = cv2.GaussianBlur(40 + 120 * body + 60 * bars + rng.randn(h, w) * 6, (0, 0), 1.2)
    out = np.stack([img * 0.9, img, img * 1.05], -1) + rng.randn(h, w, 3) * 2
    return np.clip(out, 0, 255).astype(np.uint8)


DESKEW_CASES = [(512, 512, 0, 17.0), (480, 640, 1, -33.0), (600, 400, 2, 71.0), (1024, 1024, 3, 5.5), (700, 900, 4, -58.0),
                (400, 400, 5, 0.0)]
