"""Thin tensor-level wrappers over the C ABI (include/teethrt.h).  Each function only checks dtypes/contiguity, hands
device pointers + the current CUDA stream to libteethrt and returns the output tensors; all arithmetic is in the kernels.
Activations: NHWC bf16 viewed as [rows, C]; parameters/gradients fp32."""
import ctypes as C
import math
import os

import torch

from ._lib import lib, check, ptr, stream, EPI_SCALE_SHIFT, EPI_SILU, EPI_RESIDUAL, EPI_STATS, BnFin, BnBwdFin, SeBn  # noqa: F401

bf16 = torch.bfloat16


STAT_REPLICAS = int(lib.trt_stat_replicas())


def new_stats(C_, device):
    """Zeroed BatchNorm statistics buffer [STAT_REPLICAS, 2, C] (fp64): blocks spread their atomics over the replicas."""
    return torch.zeros((STAT_REPLICAS, 2, C_), device=device, dtype=torch.float64)


def stats_total(st, C_=None):
    """[2, C] totals of a replicated statistics buffer."""
    return st.view(STAT_REPLICAS, 2, -1).sum(0)


def _c(t, dtype=None):
    assert t.is_cuda and t.is_contiguous(), "teethrt ops need contiguous CUDA tensors"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return t


def same_out(i, s):
    return (i + s - 1) // s


# ------------------------------------------------------------------------------------------------ GEMMs (tcgen05)
def gemm(A, B, flags=0, scale=None, shift=None, residual=None, stats=None, out=None, block_n=0):
    """C[M,N] = epi(A[M,K] @ B[N,K]^T), bf16 in/out, fp32 accumulate in TMEM."""
    _c(A, bf16), _c(B, bf16)
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=bf16)
    check(lib.trt_gemm_bf16(ptr(A), ptr(B), ptr(out), M, N, K, flags, ptr(scale), ptr(shift), ptr(residual), ptr(stats),
                            block_n, stream()))
    return out


def bn_fin(stats, gamma, beta, rm, rv, nbt, rec, count, eps, momentum=0.1):
    """Host record of a LAZY train-mode BatchNorm: `stats` are complete, `rec` is not written yet.  Pass it to the first
    consumer (dwconv_fwd / pool_act / bn_apply): that kernel derives scale/shift itself and publishes rec + running stats."""
    return BnFin(ptr(stats), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(nbt), ptr(rec), float(count), eps, momentum)


def gemm_bnbwd(A, B, bn_x, bstats, residual=None, out=None):
    """Data-gradient GEMM C = A @ B^T (+ residual) that also accumulates {sum C, sum C * bn_x} per output channel: the
    backward sums of the BatchNorm (input bn_x) that C flows into next (pass raw_x=True to bn_bwd_fin)."""
    _c(A, bf16), _c(B, bf16), _c(bn_x, bf16)
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and tuple(bn_x.shape) == (M, N)
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=bf16)
    check(lib.trt_gemm_bf16_bnbwd(ptr(A), ptr(B), ptr(out), M, N, K, EPI_RESIDUAL if residual is not None else 0, ptr(residual),
                                  ptr(bn_x), ptr(bstats), stream()))
    return out


def bn_bwd_fin(bstats, rec, gamma, dgamma, dbeta, count, coef=None, raw_x=False):
    """Lazy BatchNorm backward: pass to affine2 instead of a coefficient tensor.  coef: optional [3,C] scratch that lets the
    library finalise with its own launch where that is cheaper (small tensors)."""
    return BnBwdFin(ptr(bstats), ptr(rec), ptr(gamma), ptr(dgamma), ptr(dbeta), ptr(coef), float(count), int(raw_x))


def _ref(st):
    return None if st is None else C.byref(st)


def gemm_wgrad(P, Q, out, so_p=None, so_q=None, lbo=0, sbo=0, kstep=0, q_store=0):
    """out[p*so_p + q*so_q] += sum_m P[m,p] * Q[m,q]  (fp32 accumulate/atomics); only columns q < q_store are stored."""
    _c(P, bf16), _c(Q, bf16), _c(out, torch.float32)
    M, Cp = P.shape
    Cq = Q.shape[1]
    assert Q.shape[0] == M
    if so_p is None:
        so_p, so_q = Cq, 1
    check(lib.trt_gemm_wgrad_bf16(ptr(P), ptr(Q), ptr(out), M, Cp, Cq, so_p, so_q, q_store, lbo, sbo, kstep, stream()))
    return out


def pack_w1x1(w, w_bf16, wt_bf16=None):
    N, K = w.shape[0], w.shape[1]
    check(lib.trt_pack_w1x1(ptr(w), ptr(w_bf16), ptr(wt_bf16), N, K, stream()))


def pack_w1x1_batch(table, total_tiles):
    check(lib.trt_pack_w1x1_batch(ptr(table), table.shape[0], total_tiles, stream()))


# ------------------------------------------------------------------------------------------------ BN / SE / pool
def bn_finalize(stats, gamma, beta, rm, rv, nbt, rec, count, eps, momentum=0.1):
    check(lib.trt_bn_finalize(ptr(stats), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(nbt), ptr(rec), gamma.numel(),
                              float(count), eps, momentum, stream()))


def bn_fold_eval(gamma, beta, rm, rv, rec, eps):
    check(lib.trt_bn_fold_eval(ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(rec), gamma.numel(), eps, stream()))


def bn_bwd_finalize(bstats, rec, gamma, coef, dgamma, dbeta, count):
    check(lib.trt_bn_bwd_finalize(ptr(bstats), ptr(rec), ptr(gamma), ptr(coef), ptr(dgamma), ptr(dbeta), gamma.numel(),
                                  float(count), stream()))


def bn_apply(x, rec, out, residual=None, act=0, fin=None):
    rows, Cc = x.shape
    check(lib.trt_bn_apply(ptr(x), ptr(rec), ptr(residual), ptr(out), _ref(fin), rows, Cc, act, stream()))
    return out


def pool_act(x, rec, pooled, N, HW, act=1, zeroed=False, fin=None):
    check(lib.trt_pool_act(ptr(x), ptr(rec), ptr(pooled), int(zeroed), _ref(fin), N, HW, x.shape[-1], act, stream()))
    return pooled


def se_workspace(N, Cc, rd, device):
    """Zeroed workspace of the one-launch SE MLP kernels (barrier state + split-K partials); owned by ONE stream at a time."""
    return torch.zeros(int(lib.trt_se_workspace_bytes(N, Cc, rd)), device=device, dtype=torch.uint8)


def se_fwd(pooled, inv_hw, Wr, br, We, be, s1, gate, apply_x=None, HW=0, ws=None):
    """apply_x: activated bf16 [N*HW, C] tensor gated IN PLACE by the same launch (small inference feature maps).
    ws: workspace from se_workspace() -> the one-launch kernel (training batches)."""
    N, Cc = pooled.shape
    if ws is not None and apply_x is None:
        check(lib.trt_se_fwd_fused(ptr(pooled), inv_hw, ptr(Wr), ptr(br), ptr(We), ptr(be), ptr(s1), ptr(gate), ptr(ws), ws.numel(),
                                   N, Cc, Wr.shape[0], stream()))
        return
    check(lib.trt_se_fwd(ptr(pooled), inv_hw, ptr(Wr), ptr(br), ptr(We), ptr(be), ptr(s1), ptr(gate), ptr(apply_x), HW, N, Cc,
                         Wr.shape[0], stream()))


def gate_apply(x, rec, gate, out, N, HW):
    check(lib.trt_gate_apply(ptr(x), ptr(rec), ptr(gate), ptr(out), N, HW, x.shape[-1], stream()))
    return out


def scale_f32(x, alpha):
    check(lib.trt_scale_f32(ptr(x), x.numel(), alpha, stream()))
    return x


def bn_bwd_reduce(dy, x, rec, bstats):
    rows, Cc = x.shape
    check(lib.trt_bn_bwd_reduce(ptr(dy), ptr(x), ptr(rec), ptr(bstats), rows, Cc, stream()))


def affine2(dy, x, coef, out, fin=None):
    """out = a*dy + b*x + c with coef [3,C], or (coef None) the lazy BatchNorm-backward record `fin`."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    check(lib.trt_affine2(ptr(dy), ptr(x), ptr(coef), ptr(out), _ref(fin), rows, Cc, stream()))
    return out


def se_bwd_reduce(dA, x, rec, sums, N, HW, zeroed=False, full=False):
    """sums: [N,C] (dgate_pre) or, with full=True, [5,N,C] (dgate_pre + the four sums the merged backward needs)."""
    check(lib.trt_se_bwd_reduce(ptr(dA), ptr(x), ptr(rec), ptr(sums), int(zeroed), int(full), N, HW, x.shape[-1], stream()))


def se_bn(sums, rec, gamma, coef, dgamma, dbeta, count):
    return SeBn(ptr(sums), ptr(rec), ptr(gamma), ptr(coef), ptr(dgamma), ptr(dbeta), float(count))


def se_bwd(dgate_pre, gate, s1, pooled, inv_hw, Wr, We, ds2, ds1, dmean, dWr, dbr, dWe, dbe, ds1_zeroed=False, bn=None, ws=None):
    N, Cc = gate.shape
    if ws is not None:
        check(lib.trt_se_bwd_fused(ptr(dgate_pre), ptr(gate), ptr(s1), ptr(pooled), inv_hw, ptr(Wr), ptr(We), ptr(ds2), ptr(ds1),
                                   ptr(dmean), ptr(dWr), ptr(dbr), ptr(dWe), ptr(dbe), _ref(bn), ptr(ws), ws.numel(), N, Cc,
                                   Wr.shape[0], stream()))
        return
    check(lib.trt_se_bwd(ptr(dgate_pre), ptr(gate), ptr(s1), ptr(pooled), inv_hw, ptr(Wr), ptr(We), ptr(ds2), ptr(ds1),
                         ptr(dmean), ptr(dWr), ptr(dbr), ptr(dWe), ptr(dbe), int(ds1_zeroed), _ref(bn), N, Cc, Wr.shape[0], stream()))


def act_bwd_apply(dA, gate, dmean, inv_hw, x, rec, coef, out, N, HW):
    check(lib.trt_act_bwd_apply(ptr(dA), ptr(gate), ptr(dmean), inv_hw, ptr(x), ptr(rec), ptr(coef), ptr(out), N, HW, x.shape[-1],
                                stream()))
    return out


def act_bwd(dA, gate, dmean, inv_hw, x, rec, g_out, bstats, N, HW, act=1):
    check(lib.trt_act_bwd(ptr(dA), ptr(gate), ptr(dmean), inv_hw, ptr(x), ptr(rec), ptr(g_out), ptr(bstats), N, HW,
                          x.shape[-1], act, stream()))
    return g_out


# ------------------------------------------------------------------------------------------------ spatial convs
def dwconv_fwd(x, in_rec, w, out, N, H, W, k, s, out_rec=None, pooled=None, stats=None, in_fin=None, pooled_zeroed=False, act_out=None):
    """in_fin: lazy BatchNorm record of the INPUT (in_rec is derived from its statistics and published by this launch).
    act_out (stride 1, with in_rec): receives silu(bn(x)) for the weight-gradient kernel of the backward pass."""
    check(lib.trt_dwconv_fwd(ptr(x), ptr(in_rec), ptr(w), ptr(out), ptr(out_rec), ptr(pooled), int(pooled_zeroed), ptr(stats),
                             _ref(in_fin), ptr(act_out), N, H, W, x.shape[-1], k, s, stream()))
    return out


def dwconv_bwd(dD, w, x_raw, x_rec, g_out, bstats, dw, N, H, W, k, s):
    """dD: gradient w.r.t. the raw depthwise output (BN-backward affine already applied, see affine2)."""
    check(lib.trt_dwconv_bwd(ptr(dD), ptr(w), ptr(x_raw), ptr(x_rec), ptr(g_out), ptr(bstats), ptr(dw), N, H, W,
                             x_raw.shape[-1], k, s, stream()))


def stem_fwd(x, w, out, out_rec=None, stats=None):
    N, _, H, W = x.shape
    assert x.dtype in (torch.float32, bf16)
    check(lib.trt_stem_fwd(ptr(x), int(x.dtype == bf16), ptr(w), ptr(out), ptr(out_rec), ptr(stats), N, H, W, w.shape[0],
                           stream()))
    return out


def stem_im2col(x, patches):
    """patches [N*OH*OW, 32] bf16 = im2col of the 3x3 stride-2 'same' stem (27 taps + 5 zero columns)."""
    N, _, H, W = x.shape
    check(lib.trt_stem_im2col(ptr(x), int(x.dtype == bf16), ptr(patches), N, H, W, stream()))
    return patches


def stem_pack_w(w, w_bf16):
    check(lib.trt_stem_pack_w(ptr(w), ptr(w_bf16), w.shape[0], stream()))
    return w_bf16


def stem_wgrad(x, ds, dw):
    N, _, H, W = x.shape
    check(lib.trt_stem_wgrad(ptr(x), int(x.dtype == bf16), ptr(ds), ptr(dw), N, H, W, dw.shape[0], stream()))


# ------------------------------------------------------------------------------------------------ MIL pooling
_mil_ws = {}


def mil_attn_fwd(H, Vw, Vb, Uw, Ub, ww, wb, save=False, tensor_core=None):
    B, K, D = H.shape
    hid = Vw.shape[0]
    M = torch.empty((B, D), device=H.device, dtype=torch.float32)
    A = torch.empty((B, K), device=H.device, dtype=torch.float32)
    gV = gU = None
    if save:
        gV = torch.empty((B, K, hid), device=H.device, dtype=torch.float32)
        gU = torch.empty_like(gV)
    if tensor_core is None:
        # measured (profiles/r02_microbench.jsonl): 6 bags x 16 = 96 instance rows are one 128-row tile - the CUDA-core kernel
        # (45 us) beats split + GEMM + pool (72 us, launch-bound); from 64 bags on the tensor-core path wins (70 vs 115 us,
        # 128 vs 942 us at 1024 bags)
        tensor_core = B * K >= 512 and D % 64 == 0 and hid % 8 == 0 and os.environ.get("TEETHRT_MIL_TC", "1") != "0"
    if tensor_core:
        nbytes = int(lib.trt_mil_attn_tc_workspace_bytes(B, K, D, hid))
        key = (str(H.device), nbytes)
        ws = _mil_ws.get(key)
        if ws is None:           # one per shape, never dropped: a captured train step keeps replaying into its workspace
            ws = _mil_ws[key] = torch.empty(nbytes + 256, device=H.device, dtype=torch.uint8)
        off = (-ws.data_ptr()) % 256
        check(lib.trt_mil_attn_fwd_tc(ptr(H), ptr(Vw), ptr(Vb), ptr(Uw), ptr(Ub), ptr(ww), ptr(wb), ptr(M), ptr(A), ptr(gV),
                                      ptr(gU), B, K, D, hid, ws.data_ptr() + off, nbytes, stream()))
    else:
        check(lib.trt_mil_attn_fwd(ptr(H), ptr(Vw), ptr(Vb), ptr(Uw), ptr(Ub), ptr(ww), ptr(wb), ptr(M), ptr(A), ptr(gV),
                                   ptr(gU), B, K, D, hid, stream()))
    return M, A, gV, gU


def mil_attn_bwd(dM, H, A, gV, gU, Vw, Uw, ww, dVw, dVb, dUw, dUb, dww, dwb):
    B, K, D = H.shape
    dH = torch.empty_like(H)
    check(lib.trt_mil_attn_bwd(ptr(dM), ptr(H), ptr(A), ptr(gV), ptr(gU), ptr(Vw), ptr(Uw), ptr(ww), ptr(dH), ptr(dVw),
                               ptr(dVb), ptr(dUw), ptr(dUb), ptr(dww), ptr(dwb), B, K, D, Vw.shape[0], stream()))
    return dH


def linear1_fwd(M, w, b, drop_p=0.0, seed=0, step=None):
    B, D = M.shape
    logit = torch.empty(B, device=M.device, dtype=torch.float32)
    check(lib.trt_linear1_fwd(ptr(M), ptr(w), ptr(b), ptr(logit), B, D, drop_p, seed, ptr(step), stream()))
    return logit


def linear1_bwd(dlogit, M, w, dw, db, drop_p=0.0, seed=0, step=None):
    B, D = M.shape
    dM = torch.empty_like(M)
    check(lib.trt_linear1_bwd(ptr(dlogit), ptr(M), ptr(w), ptr(dM), ptr(dw), ptr(db), B, D, drop_p, seed, ptr(step), stream()))
    return dM


def bce_logits(logit, y, sample_w=None, want_grad=True):
    B = logit.numel()
    loss = torch.empty(1, device=logit.device, dtype=torch.float32)
    dlogit = torch.empty_like(logit) if want_grad else None
    check(lib.trt_bce_logits(ptr(logit), ptr(y), ptr(sample_w), ptr(loss), ptr(dlogit), B, stream()))
    return loss, dlogit


# ------------------------------------------------------------------------------------------------ tab + heads + loss
def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def tab_heads_scratch(B, Hd, device):
    return torch.empty(lib.trt_tab_heads_scratch_floats(B, Hd), device=device, dtype=torch.float32)


def tab_heads_fwd(feat, xtab, params, bn_rm, bn_rv, bn_nbt, scratch, train, drop_p=0.0, targets=None, alpha=1.0, beta=0.3,
                  seed=0, step=None, out=None):
    """params: the 10 tensors in the order of teethrt.h.  targets = (y_hard, y_soft, sample_w|None) fuses the loss."""
    B, F = feat.shape
    T, Hd = xtab.shape[1], params[0].shape[0]
    dev = feat.device
    if out is None:
        out = {k: torch.empty(B, device=dev, dtype=torch.float32) for k in ("logit", "reg", "dlogit", "dreg")}
        out["loss"] = torch.zeros(1, device=dev, dtype=torch.float32)
    yh = ys = sw = None
    if targets is not None:
        yh, ys, sw = targets
    arr = _ptr_array(params)
    check(lib.trt_tab_heads_fwd(ptr(feat), ptr(xtab), arr, ptr(bn_rm), ptr(bn_rv), ptr(bn_nbt), ptr(yh), ptr(ys), ptr(sw),
                                ptr(out["logit"]), ptr(out["reg"]), ptr(out["loss"]), ptr(out["dlogit"]), ptr(out["dreg"]),
                                ptr(scratch), B, T, Hd, F, int(train), drop_p, alpha, beta, seed, ptr(step), stream()))
    return out


def tab_heads_bwd(feat, xtab, params, bn_rm, bn_rv, dlogit, dreg, dfeat, grads, scratch, train, drop_p=0.0, seed=0,
                  step=None):
    B, F = feat.shape
    T, Hd = xtab.shape[1], params[0].shape[0]
    check(lib.trt_tab_heads_bwd(ptr(feat), ptr(xtab), _ptr_array(params), ptr(bn_rm), ptr(bn_rv), ptr(dlogit), ptr(dreg),
                                ptr(dfeat), _ptr_array(grads), ptr(scratch), B, T, Hd, F, int(train), drop_p, seed,
                                ptr(step), stream()))
    return dfeat


# ------------------------------------------------------------------------------------------------ optimiser
class OptimState:
    """Device-resident {step, lr schedule, bias corrections} (layout: optim.cu OptState)."""

    def __init__(self, device, lr, t_max=0, betas=(0.9, 0.999)):
        import struct
        raw = struct.pack("<Qddddffffd", 0, float(lr), float(t_max), betas[0], betas[1], float(lr), 0.0, 0.0, 0.0, 0.0)
        assert len(raw) == lib.trt_optim_state_bytes()
        self.buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)

    def advance(self):
        check(lib.trt_optim_advance(ptr(self.buf), stream()))

    def read(self):
        import struct
        vals = struct.unpack("<Qddddffffd", bytes(self.buf.cpu().numpy().tobytes()))
        return dict(step=vals[0], lr0=vals[1], t_max=vals[2], lr=vals[5], bc1=vals[6], bc2=vals[7], skipped=int(vals[9]))


def grad_sumsq(g, out):
    check(lib.trt_grad_sumsq(ptr(g), g.numel(), ptr(out), stream()))


def adamw_step(p, g, m, v, state, normsq=None, norm_out=None, grad_scale=1.0, max_norm=1.0, eps=1e-8, weight_decay=1e-4):
    check(lib.trt_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(state.buf), ptr(normsq), ptr(norm_out),
                             grad_scale, max_norm, eps, weight_decay, stream()))


# ------------------------------------------------------------------------------------------------ fold calibration
def temperature_nll(logits, targets, log_T, out=None):
    """out[0] = mean BCE(logits / exp(log_T), targets), out[1] = d/dlog_T (train_mm_joint_dualtask.py:168-170,276-281)."""
    out = torch.empty(2, device=logits.device, dtype=torch.float32) if out is None else out
    check(lib.trt_temperature_nll(ptr(logits), ptr(targets), ptr(log_T), ptr(out), logits.numel(), stream()))
    return out


def scaled_sigmoid(logits, T, out=None):
    out = torch.empty_like(logits) if out is None else out
    check(lib.trt_scaled_sigmoid(ptr(logits), float(T), ptr(out), logits.numel(), stream()))
    return out


def binary_metrics(prob, y, thr):
    """-> (counts [nthr, 4] int64 = tp, fp, fn, tn per threshold; auc [4] int64 = 2*wins+ties, #pos, #neg, #bad labels)."""
    nthr = 0 if thr is None else thr.numel()
    counts = torch.empty(max(nthr, 1), 4, device=prob.device, dtype=torch.int64)
    auc = torch.empty(4, device=prob.device, dtype=torch.int64)
    if prob.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"scores must be fp32 or fp64, got {prob.dtype}")
    check(lib.trt_binary_metrics(ptr(prob), int(prob.dtype == torch.float64), ptr(y), prob.numel(), ptr(thr), nthr, ptr(counts),
                                 ptr(auc), stream()))
    return counts[:nthr], auc


def logreg_fit(X, y, C=1.0, max_iter=100, tol=None):
    """L2 logistic regression by Newton on the device -> (coef fp64 [d+1] = w..., b ; info fp64 [3])."""
    n, d = X.shape
    coef = torch.empty(d + 1, device=X.device, dtype=torch.float64)
    info = torch.empty(3, device=X.device, dtype=torch.float64)
    tol = 1e-10 * max(1.0, C * n) if tol is None else tol
    check(lib.trt_logreg_fit(ptr(X), ptr(y), n, d, float(C), int(max_iter), float(tol), ptr(coef), ptr(info), stream()))
    return coef, info


def logreg_predict(X, coef):
    n, d = X.shape
    out = torch.empty(n, device=X.device, dtype=torch.float64)
    check(lib.trt_logreg_predict(ptr(X), n, d, ptr(coef), ptr(out), stream()))
    return out


# ------------------------------------------------------------------------------------------------ input stage
def same_pad(i, k, s):
    total = max((math.ceil(i / s) - 1) * s + k - i, 0)
    return total // 2, total - total // 2


# ------------------------------------------------------------------------------------------------ Pillow-exact augment ops
def hist_lut_u8(img, mode):
    """ImageOps.autocontrast (mode 0) / equalize (mode 1) lookup table of a uint8 HWC image, built on the device."""
    ch = img.shape[2]
    hist = torch.empty(ch * 256, device=img.device, dtype=torch.int64)
    lut = torch.empty(ch * 256, device=img.device, dtype=torch.uint8)
    check(lib.trt_hist_u8(ptr(img), img.shape[0] * img.shape[1], ch, ptr(hist), stream()))
    check(lib.trt_lut_build_u8(ptr(hist), ch, int(mode), ptr(lut), stream()))
    return lut


def lut_apply_u8(img, lut):
    out = torch.empty_like(img)
    check(lib.trt_lut_apply_u8(ptr(img), ptr(lut), img.shape[0] * img.shape[1], img.shape[2], ptr(out), stream()))
    return out


def enhance_rgb_u8(img, mode, factor):
    out = torch.empty_like(img)
    scratch = torch.empty(1, device=img.device, dtype=torch.int64)
    check(lib.trt_enhance_rgb_u8(ptr(img), img.shape[0], img.shape[1], int(mode), float(factor), ptr(scratch), ptr(out), stream()))
    return out


def affine_pil_u8(img, matrix, bicubic, fill):
    out = torch.empty_like(img)
    m = (C.c_double * 6)(*[float(v) for v in matrix])
    f = (C.c_ubyte * 4)(*([int(v) for v in fill] + [0] * (4 - len(fill))))
    check(lib.trt_affine_pil_u8(ptr(img), img.shape[0], img.shape[1], img.shape[2], m, int(bool(bicubic)), f, ptr(out), stream()))
    return out
