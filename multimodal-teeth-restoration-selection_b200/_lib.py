"""ctypes binding of libteethrt.so — the only way the Python host side reaches the GPU kernels.

There is NO fallback: if the shared library is missing or fails to load, importing the product path raises.
Pointers handed to the library are device pointers of torch tensors (torch is plumbing: memory, streams, NCCL).
"""
import ctypes as C
import os
import re
import threading

import torch

from ._build import LIB_PATH

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "teethrt.h")


class TeethRTError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise TeethRTError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). teethrt has no CPU or PyTorch fallback.")
    return C.CDLL(LIB_PATH)


_cdll = _load()
_tl = threading.local()          # .dev = ordinal of the device the most recent operand lives on (set by ptr())


class _Lib:
    """The loaded library.  Every entry point launches on the device its operands live on: the reference lets the caller
    pick any device (`MMEnsemble(device='cuda:1')`), while a CUDA launch needs the CURRENT device to own the stream, so a
    call whose operands sit on another device runs under a device guard (one process per GPU never takes that branch)."""

    def __init__(self, cdll):
        self._cdll = cdll

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)       # AttributeError here = library/header mismatch: fail loudly

        def call(*args):
            dev = getattr(_tl, "dev", None)
            if dev is None or dev == torch.cuda.current_device():
                return fn(*args)
            with torch.cuda.device(dev):
                return fn(*args)
        call.__name__ = name
        call.raw = fn
        setattr(self, name, call)
        return call


lib = _Lib(_cdll)

vp, i32, i64, f32, f64, u64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_ulonglong, C.c_size_t

_SIGS = {
    "trt_version": (i32, []),
    "trt_last_error_string": (C.c_char_p, []),
    "trt_init": (i32, [i32]),
    "trt_launch_count": (u64, []),
    "trt_set_pdl": (i32, [i32]),
    "trt_stat_replicas": (i32, []),
    "trt_gemm_bf16": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]),
    "trt_gemm_bf16_bnbwd": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "trt_gemm_wgrad_bf16": (i32, [vp, vp, vp, i32, i32, i32, i64, i64, i32, i32, i32, i32, vp]),
    "trt_clahe_workspace_bytes": (sz, [i32]),
    "trt_clahe_bgr_u8": (i32, [vp, vp, i32, i32, i32, f32, vp, vp, sz, vp]),
    "trt_resize_linear_u8": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "trt_normalize_flip_u8": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "trt_bn_finalize": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, f64, f32, f32, vp]),
    "trt_bn_fold_eval": (i32, [vp, vp, vp, vp, vp, i32, f32, vp]),
    "trt_bn_bwd_finalize": (i32, [vp, vp, vp, vp, vp, vp, i32, f64, vp]),
    "trt_bn_apply": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "trt_pool_act": (i32, [vp, vp, vp, i32, vp, i32, i32, i32, i32, vp]),
    "trt_se_fwd": (i32, [vp, f32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "trt_gate_apply": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    "trt_bn_bwd_reduce": (i32, [vp, vp, vp, vp, i32, i32, vp]),
    "trt_affine2": (i32, [vp, vp, vp, vp, vp, i32, i32, vp]),
    "trt_se_bwd_reduce": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "trt_act_bwd_apply": (i32, [vp, vp, vp, f32, vp, vp, vp, vp, i32, i32, i32, vp]),
    "trt_se_bwd": (i32, [vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp]),
    "trt_se_workspace_bytes": (sz, [i32, i32, i32]),
    "trt_se_fwd_fused": (i32, [vp, f32, vp, vp, vp, vp, vp, vp, vp, sz, i32, i32, i32, vp]),
    "trt_se_bwd_fused": (i32, [vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, i32, i32, i32, vp]),
    "trt_act_bwd": (i32, [vp, vp, vp, f32, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "trt_scale_f32": (i32, [vp, sz, f32, vp]),
    "trt_pack_w1x1": (i32, [vp, vp, vp, i32, i32, vp]),
    "trt_pack_w1x1_batch": (i32, [vp, i32, i32, vp]),
    "trt_dwconv_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "trt_dwconv_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "trt_stem_fwd": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "trt_stem_wgrad": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, vp]),
    "trt_stem_im2col": (i32, [vp, i32, vp, i32, i32, i32, vp]),
    "trt_stem_pack_w": (i32, [vp, vp, i32, vp]),
    "trt_mil_attn_smem_bytes": (sz, [i32, i32, i32, i32]),
    "trt_mil_attn_fwd": (i32, [vp] * 11 + [i32, i32, i32, i32, vp]),
    "trt_mil_attn_bwd": (i32, [vp] * 15 + [i32, i32, i32, i32, vp]),
    "trt_mil_attn_tc_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "trt_mil_attn_fwd_tc": (i32, [vp] * 11 + [i32, i32, i32, i32, vp, sz, vp]),
    "trt_linear1_fwd": (i32, [vp, vp, vp, vp, i32, i32, f32, u64, vp, vp]),
    "trt_linear1_bwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, f32, u64, vp, vp]),
    "trt_bce_logits": (i32, [vp, vp, vp, vp, vp, i32, vp]),
    "trt_tab_heads_scratch_floats": (sz, [i32, i32]),
    "trt_tab_heads_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32,
                                f32, f32, u64, vp, vp]),
    "trt_tab_heads_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, u64, vp, vp]),
    "trt_optim_state_bytes": (sz, []),
    "trt_optim_advance": (i32, [vp, vp]),
    "trt_grad_sumsq": (i32, [vp, sz, vp, vp]),
    "trt_adamw_step": (i32, [vp, vp, vp, vp, sz, vp, vp, vp, f32, f32, f32, f32, vp]),
    "trt_resample_u8": (i32, [vp, sz, i32, vp, i32, i32, vp, vp, i32, i32, i32, vp]),
    "trt_canny_nms_bgr_u8": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "trt_canny_hysteresis_pass": (i32, [vp, i32, i32, vp, vp]),
    "trt_canny_finish": (i32, [vp, i32, i32, vp, vp, vp]),
    "trt_warp_affine_linear_u8": (i32, [vp, i32, i32, i32, vp, i32, i32, vp, vp]),
    "trt_hist_u8": (i32, [vp, sz, i32, vp, vp]),
    "trt_lut_build_u8": (i32, [vp, i32, i32, vp, vp]),
    "trt_lut_apply_u8": (i32, [vp, vp, sz, i32, vp, vp]),
    "trt_enhance_rgb_u8": (i32, [vp, i32, i32, i32, f32, vp, vp, vp]),
    "trt_affine_pil_u8": (i32, [vp, i32, i32, i32, vp, i32, vp, vp, vp]),
    "trt_crop_resize_batch_u8": (i32, [vp, i32, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "trt_aug_layer_batch_u8": (i32, [vp, i32, i32, i32, vp, vp, vp, vp]),
    "trt_normalize_erase_batch": (i32, [vp, i32, i32, vp, i32, vp]),
    "trt_aug_job_bytes": (i32, []),
    "trt_temperature_nll": (i32, [vp, vp, vp, vp, i32, vp]),
    "trt_scaled_sigmoid": (i32, [vp, f32, vp, i32, vp]),
    "trt_binary_metrics": (i32, [vp, i32, vp, i32, vp, i32, vp, vp, vp]),
    "trt_logreg_fit": (i32, [vp, vp, i32, i32, f64, i32, f64, vp, vp, vp]),
    "trt_logreg_predict": (i32, [vp, i32, i32, vp, vp, vp]),
}

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(_cdll, _name)     # AttributeError here = library/header mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

EPI_SCALE_SHIFT, EPI_SILU, EPI_RESIDUAL, EPI_STATS = 1, 2, 4, 8


class BnFin(C.Structure):
    """trt_bn_fin_t (include/teethrt.h): a lazy train-mode BatchNorm - the first consumer derives scale/shift from the
    producer's statistics and publishes the record."""
    _fields_ = [("stats", vp), ("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp), ("num_batches_tracked", vp),
                ("rec", vp), ("count", f64), ("eps", f32), ("momentum", f32)]


class BnBwdFin(C.Structure):
    """trt_bn_bwd_fin_t: lazy BatchNorm backward - trt_affine2 derives dx = a*dy + b*x + c from the sums and writes dgamma/dbeta."""
    _fields_ = [("bstats", vp), ("rec", vp), ("gamma", vp), ("dgamma", vp), ("dbeta", vp), ("coef", vp), ("count", f64), ("raw_x", i32)]


class SeBn(C.Structure):
    """trt_se_bn_t: BatchNorm backward of the gated depthwise activation folded into the SE MLP backward."""
    _fields_ = [("sums", vp), ("rec", vp), ("gamma", vp), ("coef", vp), ("dgamma", vp), ("dbeta", vp), ("count", f64)]


def header_symbols():
    """Every function name include/teethrt.h declares (used by the CPU test that checks the .so exports them all)."""
    txt = open(_HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", txt)))


def ptr(t):
    """Device pointer of a tensor (None -> NULL); remembers the operand's device for stream() and the launch guard."""
    if t is None:
        return None
    if t.is_cuda:
        _tl.dev = t.device.index
    return t.data_ptr()


def stream():
    """The current stream OF THE OPERANDS' DEVICE (ptr() of every operand is evaluated before this in a call's argument
    list), not of whatever device happens to be current."""
    return torch.cuda.current_stream(getattr(_tl, "dev", None)).cuda_stream


def check(rc):
    if rc != 0:
        raise TeethRTError(f"libteethrt error {rc}: {lib.trt_last_error_string().decode()}")


_inited = set()


def init(device=None):
    """trt_init for the current (or given) CUDA device; raises without a B200-class GPU."""
    if not torch.cuda.is_available():
        raise TeethRTError("teethrt needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = None if device is None else (device if isinstance(device, int) else torch.device(device).index)
    dev = torch.cuda.current_device() if idx is None else idx
    if dev not in _inited:
        check(_cdll.trt_init(dev))      # leaves the caller's current device as it was
        _inited.add(dev)
    return dev
