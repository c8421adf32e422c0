"""Host-side construction of the integer lookup tables of OpenCV's 8-bit BGR<->Lab conversion,
packed in the layout of include/teethrt.h (TRT_TAB_*).  Built once with numpy in the exact float32/float64 mix OpenCV
uses, uploaded to the device once per process, and checksummed against the published table digests."""
import hashlib

import numpy as np

F = np.float32
GAMMA_OFF, CBRT_OFF, YF_OFF, ABXZ_OFF, INVGAMMA_OFF, TAB_BYTES = 0, 512, 6656, 8704, 156160, 160256

# sha1(int32 little-endian)[:16] of each table
DIGESTS = {"gamma": "8a88c35441764553", "cbrt": "1eb5ec72d4085b5a", "yf": "2bcfd7adb3a3178f",
           "abxz": "5f8e53ab76820250", "invgamma": "5d288977e5795f21"}


def _gamma():
    x = (np.arange(256, dtype=F) / F(255)).astype(F)
    xd = x.astype(np.float64)
    g = np.where(x <= F(0.04045), xd / 12.92, ((xd + 0.055) / 1.055) ** 2.4).astype(F)
    return np.rint(F(2040) * g).astype(np.int32)


def _cbrt():
    x = (np.arange(3072, dtype=F) / F(2040)).astype(F)
    lin = (x * F(841.0 / 108.0) + F(16.0 / 116.0)).astype(F)
    f = np.where(x < F(216.0 / 24389.0), lin, np.cbrt(x).astype(F)).astype(F)
    return np.rint(F(32768) * f).astype(np.int32)


def _yf():
    base = 16384
    out = np.empty(512, dtype=np.int32)
    for L in range(256):
        if L <= 20:
            y = np.rint(F(L * base * 20 * 9) / F(17 * 29 ** 3))
            ify = np.rint(F(base) * (F(16) / F(116) + F(L * 5) / F(3 * 17 * 29)))
        else:
            fy = F(F(L * 100 * base) / F(255 * 116)) + F(F(16 * base) / F(116))
            ify = np.rint(fy)
            y = np.rint(F(F(F(fy * fy) * fy) / F(base * base)))
        out[2 * L], out[2 * L + 1] = int(y), int(ify)
    return out


def _abxz():
    base, min_ab = 16384, -8145
    i = np.arange(min_ab, min_ab + 36864, dtype=np.int64)
    tdiv = lambda a, b: np.sign(a) * (np.abs(a) // b)       # C truncating division
    lo = tdiv(i * 108, 841) - (base * 16 // 116) * 108 // 841
    hi = tdiv(tdiv(i * i, base) * i, base)
    return np.where(i <= 3390, lo, hi).astype(np.int32)


def _invgamma():
    x = (np.arange(4096, dtype=F) / F(4096)).astype(F)
    xd = x.astype(np.float64)
    g = np.where(x <= F(0.0031308), 12.92 * xd, 1.055 * np.power(xd, 1.0 / 2.4) - 0.055).astype(F)
    return np.rint(F(255) * g).astype(np.int32)


def _digest(t):
    return hashlib.sha1(np.ascontiguousarray(t, dtype="<i4").tobytes()).hexdigest()[:16]


def build_packed():
    """-> uint8 numpy array of TAB_BYTES in device layout; raises if any table deviates from OpenCV's."""
    tabs = {"gamma": _gamma(), "cbrt": _cbrt(), "yf": _yf(), "abxz": _abxz(), "invgamma": _invgamma()}
    for k, t in tabs.items():
        if _digest(t) != DIGESTS[k]:
            raise RuntimeError(f"Lab table {k} does not match OpenCV's (digest {_digest(t)})")
    buf = np.zeros(TAB_BYTES, dtype=np.uint8)
    buf[GAMMA_OFF:GAMMA_OFF + 512] = tabs["gamma"].astype("<u2").view(np.uint8)
    buf[CBRT_OFF:CBRT_OFF + 6144] = tabs["cbrt"].astype("<u2").view(np.uint8)
    buf[YF_OFF:YF_OFF + 2048] = tabs["yf"].astype("<i4").view(np.uint8)
    buf[ABXZ_OFF:ABXZ_OFF + 147456] = tabs["abxz"].astype("<i4").view(np.uint8)
    buf[INVGAMMA_OFF:INVGAMMA_OFF + 4096] = tabs["invgamma"].astype(np.uint8)
    return buf
