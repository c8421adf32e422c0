"""EfficientNet encoder (`tf_efficientnet_b{0,4}_ns`) on libteethrt kernels, with timm's state-dict keys.

This is what `timm.create_model(name, pretrained, num_classes=0, global_pool='avg')` returns in the reference
(experiments/multimodal_v1/train_mm_joint_dualtask.py:138; ui/gradio_app/infer_mm.py:22;
experiments/vision_v2/train_mil_attention_v1.py:135; ui/gradio_app/infer_mil.py:75).  Parameters live in ordinary
torch-layout fp32 `nn.Parameter`s under timm's names (conv_stem, bn1, blocks.S.I.{conv_pw,bn1,conv_dw,bn2,se.conv_reduce,
se.conv_expand,conv_pwl,bn3}, conv_head, bn2), so reference checkpoints load with strict=True.  All arithmetic runs in the
CUDA library: activations are NHWC bf16, 1x1 convs are tcgen05 GEMMs with BN statistics / folded BN + SiLU + residual in
the epilogue, depthwise convs apply the producer's BN+SiLU on load.  No PyTorch compute fallback exists.
"""
import math
import os

import torch
import torch.nn as nn

from . import ops
from ._lib import init, lib

bf16 = torch.bfloat16
BN_EPS, BN_MOM = 1e-3, 0.1

_B0 = [("ds", 1, 3, 1, 1, 16), ("ir", 2, 3, 2, 6, 24), ("ir", 2, 5, 2, 6, 40), ("ir", 3, 3, 2, 6, 80),
       ("ir", 3, 5, 1, 6, 112), ("ir", 4, 5, 2, 6, 192), ("ir", 1, 3, 1, 6, 320)]
ARCHS = {"tf_efficientnet_b0_ns": (1.0, 1.0), "tf_efficientnet_b0": (1.0, 1.0), "tf_efficientnet_b0.ns_jft_in1k": (1.0, 1.0),
         "tf_efficientnet_b4_ns": (1.4, 1.8), "tf_efficientnet_b4": (1.4, 1.8), "tf_efficientnet_b4.ns_jft_in1k": (1.4, 1.8)}


def _round_ch(c, mult, div=8):
    c = c * mult
    n = max(div, int(c + div / 2) // div * div)
    return n + div if n < 0.9 * c else n


def arch_spec(name):
    if name not in ARCHS:
        raise ValueError(f"teethrt: backbone {name!r} is not built (have {sorted(ARCHS)})")
    wm, dm = ARCHS[name]
    stem = _round_ch(32, wm)
    stages, cin = [], stem
    for typ, r, k, s, e, c in _B0:
        cout = _round_ch(c, wm)
        blocks = []
        for i in range(int(math.ceil(r * dm))):
            blocks.append(dict(type=typ, k=k, s=s if i == 0 else 1, cin=cin, cout=cout, mid=cin * e,
                               rd=int(round(cin * 0.25))))
            cin = cout
        stages.append(blocks)
    return stem, stages, _round_ch(1280, wm)


# ------------------------------------------------------------------------------------------------ parameter holders
class _Conv(nn.Module):
    def __init__(self, cout, cin_g, k, bias=False, fan_out=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin_g, k, k))
        nn.init.normal_(self.weight, 0.0, math.sqrt(2.0 / (fan_out or k * k * cout)))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None


class _BN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _SE(nn.Module):
    def __init__(self, c, rd):
        super().__init__()
        self.conv_reduce = _Conv(rd, c, 1, bias=True)
        self.conv_expand = _Conv(c, rd, 1, bias=True)


class _DSBlock(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.cfg = b
        self.conv_dw = _Conv(b["cin"], 1, b["k"], fan_out=b["k"] * b["k"])
        self.bn1 = _BN(b["cin"])
        self.se = _SE(b["cin"], b["rd"])
        self.conv_pw = _Conv(b["cout"], b["cin"], 1)
        self.bn2 = _BN(b["cout"])


class _IRBlock(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.cfg = b
        self.conv_pw = _Conv(b["mid"], b["cin"], 1)
        self.bn1 = _BN(b["mid"])
        self.conv_dw = _Conv(b["mid"], 1, b["k"], fan_out=b["k"] * b["k"])
        self.bn2 = _BN(b["mid"])
        self.se = _SE(b["mid"], b["rd"])
        self.conv_pwl = _Conv(b["cout"], b["mid"], 1)
        self.bn3 = _BN(b["cout"])


class _Arena:
    """Bump allocator over one pre-zeroed tensor (per-BN statistics: one memset per pass instead of one per layer).
    Every slot starts on a 32-byte boundary (the kernels read them with 16- and 32-byte vector loads); `sizes` lists the
    slots that will be taken, so the padding is part of the allocation."""
    ALIGN = 8

    @classmethod
    def _pad(cls, n):
        return (n + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN

    def __init__(self, sizes, dtype, device):
        self.buf = torch.zeros(sum(self._pad(int(n)) for n in sizes), dtype=dtype, device=device)
        self.off = 0

    def take(self, n, shape=None):
        v = self.buf[self.off:self.off + n]
        assert v.numel() == n, "teethrt: arena overflow (slot list and takes disagree)"
        self.off += self._pad(n)
        return v.view(shape) if shape else v


class EfficientNet(nn.Module):
    """timm-compatible EfficientNet feature extractor (num_classes=0)."""

    def __init__(self, name="tf_efficientnet_b4_ns", global_pool="avg"):
        super().__init__()
        stem, stages, feat = arch_spec(name)
        self.arch, self.num_features, self.global_pool_type = name, feat, global_pool
        self.conv_stem = _Conv(stem, 3, 3)
        self.bn1 = _BN(stem)
        self.blocks = nn.Sequential(*[nn.Sequential(*[(_DSBlock if b["type"] == "ds" else _IRBlock)(b) for b in st])
                                      for st in stages])
        self.conv_head = _Conv(feat, stages[-1][-1]["cout"], 1)
        self.bn2 = _BN(feat)
        self.classifier = nn.Identity()
        self._eval_cache = None
        self.grad_sink = None      # optional dict name -> fp32 tensor that receives (accumulates) parameter gradients

    # ---- cache invalidation: anything that can change weights or BN buffers outside a train step
    def train(self, mode=True):
        self._eval_cache = None
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._eval_cache = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._eval_cache = None
        return super().load_state_dict(*a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._eval_cache = None
        return super()._load_from_state_dict(*a, **k)

    def block_list(self):
        return [(f"blocks.{si}.{bi}", blk) for si, st in enumerate(self.blocks) for bi, blk in enumerate(st)]

    def bn_list(self):
        out = [("bn1", self.bn1)]
        for name, blk in self.block_list():
            for bn in ("bn1", "bn2", "bn3"):
                if hasattr(blk, bn):
                    out.append((f"{name}.{bn}", getattr(blk, bn)))
        return out + [("bn2", self.bn2)]

    def conv1x1_list(self):
        out = []
        for name, blk in self.block_list():
            for cv in ("conv_pw", "conv_pwl"):
                if hasattr(blk, cv):
                    out.append((f"{name}.{cv}", getattr(blk, cv)))
        return out + [("conv_head", self.conv_head)]

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, x):
        if self.global_pool_type == "avg":
            return encoder_forward(self, x)
        if self.global_pool_type == "":
            # timm's un-pooled feature map [N, F, h, w], as the MIL twin's encoder returns it (ui/gradio_app/infer_mil.py:75,
            # 85-92).  Inference only: the twin never trains, and its own AdaptiveAvgPool2d is fused away by forward_pooled()
            if self.training and torch.is_grad_enabled():
                raise NotImplementedError("teethrt EfficientNet(global_pool=''): the un-pooled feature map is an inference "
                                          "output; train with global_pool='avg'")
            return forward_eval(self, x, feature_map=True)
        raise ValueError(f"teethrt EfficientNet: global_pool={self.global_pool_type!r} is not built ('avg' or '')")

    def forward_pooled(self, x):
        return encoder_forward(self, x)


def create_model(model_name, pretrained=False, num_classes=0, global_pool="avg", **kw):
    """Drop-in for the reference's `timm.create_model(...)` call; `pretrained` needs a network and is ignored (q11)."""
    if num_classes != 0:
        raise ValueError("teethrt.create_model: only num_classes=0 (feature extractor) is on the hot path")
    return EfficientNet(model_name, global_pool=global_pool)


# ================================================================================================= runtime
def _packed_weights(enc, dev, transposed):
    """bf16 copies of every 1x1 conv weight ([Cout,Cin] and, for training, [Cin,Cout]), refreshed by ONE batched launch.
    The bf16 buffers and the device-side pointer table persist on the encoder (keyed by the parameters' storage, so
    re-homing them into a flat buffer or moving the module rebuilds the table)."""
    convs = enc.conv1x1_list()
    key = (transposed, str(dev)) + tuple(cv.weight.data_ptr() for _, cv in convs)
    # one table per (transposed, device, storage) key and none is ever dropped while the encoder lives: a captured train
    # step keeps replaying into its table after an eval pass built the other one
    caches = enc.__dict__.setdefault("_pack_cache", {})
    cache = caches.get(key)
    if cache is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("teethrt: run one eager step before capturing a CUDA graph (weight-pack table not built yet)")
        out, rows, tiles = {}, [], 0
        for name, cv in convs:
            n, k = cv.weight.shape[0], cv.weight.shape[1]
            wb = torch.empty((n, k), device=dev, dtype=bf16)
            wt = torch.empty((k, n), device=dev, dtype=bf16) if transposed else None
            out[name] = (wb, wt)
            rows.append([cv.weight.data_ptr(), wb.data_ptr(), wt.data_ptr() if transposed else 0, n, k, tiles])
            tiles += ((n + 31) // 32) * ((k + 31) // 32)
        cache = dict(key=key, w=out, table=torch.tensor(rows, dtype=torch.int64).to(dev), tiles=tiles)
        caches[key] = cache
    ops.pack_w1x1_batch(cache["table"], cache["tiles"])
    return cache["w"]


def _version_key(enc):
    """In-place updates made through torch (an optimiser stepping while the module is in eval mode, `copy_` into a buffer)
    bump these counters; updates made by libteethrt kernels do not, and those paths clear the cache themselves."""
    return (enc.conv_stem.weight._version, enc.conv_head.weight._version, enc.bn1.running_mean._version,
            enc.bn2.running_var._version, enc.bn2.weight._version)


def _eval_cache(enc, dev):
    if enc._eval_cache is not None and enc._eval_cache["version"] != _version_key(enc):
        enc._eval_cache = None
    if enc._eval_cache is None:
        recs = {}
        for name, bn in enc.bn_list():
            rec = torch.empty((4, bn.weight.numel()), device=dev, dtype=torch.float32)
            ops.bn_fold_eval(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, rec, BN_EPS)
            recs[name] = rec
        enc._eval_cache = dict(recs=recs, w=_packed_weights(enc, dev, False), version=_version_key(enc))
    return enc._eval_cache


def _check_input(enc, x):
    if not x.is_cuda:
        raise RuntimeError("teethrt EfficientNet runs on CUDA (sm_100a) only; there is no CPU fallback")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("expected an image batch [N,3,H,W]")
    if x.dtype not in (torch.float32, bf16):
        x = x.float()
    init(x.device)
    return x.contiguous()


def _se_workspace(enc, N, dev):
    """Workspace of the one-launch SE MLP kernels (ops.se_workspace), sized for the encoder's largest block at batch N.  One per
    (device, size): captured graphs keep pointing at theirs, so a workspace is never freed or shrunk while the encoder lives.
    Measured per layer (tools/se_probe.py, batch 64): a grid barrier costs 1.5-2 us on B200 - more than a launch boundary inside
    a CUDA graph - so the one-launch BACKWARD only wins where the two-launch kernels are compute-bound (>= 1632 channels:
    22.6 vs 27 us, 30.7 vs 47 us at 2688) and the one-launch FORWARD nowhere (opt-in: TEETHRT_SE_FUSED_FWD=1).
    TEETHRT_SE_FUSED=0 keeps the two-launch kernels everywhere (A/B switch)."""
    if os.environ.get("TEETHRT_SE_FUSED", "1") == "0" or N < 32:
        return None
    need = max(int(lib.trt_se_workspace_bytes(N, blk.conv_dw.weight.shape[0], blk.cfg["rd"])) for _, blk in enc.block_list())
    cache = enc.__dict__.setdefault("_se_ws", {})
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), need)
    if key not in cache:
        cache[key] = torch.zeros(need, device=dev, dtype=torch.uint8)
    return cache[key]


def _se_fused_min_c():
    return int(os.environ.get("TEETHRT_SE_FUSED_MIN_C", "1632"))


def _se_gate(blk, pooled, inv_hw, N, C, dev, save, apply_x=None, HW=0, ws=None):
    rd = blk.cfg["rd"]
    s1 = torch.empty((N, rd), device=dev, dtype=torch.float32)
    gate = torch.empty((N, C), device=dev, dtype=torch.float32)
    se = blk.se
    ops.se_fwd(pooled, inv_hw, se.conv_reduce.weight.detach(), se.conv_reduce.bias.detach(), se.conv_expand.weight.detach(),
               se.conv_expand.bias.detach(), s1, gate, apply_x=apply_x, HW=HW, ws=ws)
    return s1, gate


def forward_eval(enc, x, feature_map=False):
    """Inference: BN folded from running statistics; 4 launches per MBConv block (+1 for the SE MLP).
    feature_map=True returns the head's feature map as [N, F, h, w] fp32 (timm's global_pool='')."""
    x = _check_input(enc, x)
    dev = x.device
    cache = _eval_cache(enc, dev)
    prev_pdl = lib.trt_set_pdl(1)         # the eval chain is launch-latency-bound: overlap each launch with its predecessor's tail
    try:
        return _forward_eval(enc, x, cache, feature_map)
    finally:
        lib.trt_set_pdl(prev_pdl)


def _forward_eval(enc, x, cache, feature_map):
    dev = x.device
    R, Wp = cache["recs"], cache["w"]
    N, _, H, W = x.shape
    h, w = ops.same_out(H, 2), ops.same_out(W, 2)
    # every SE pooling target of the pass in one zeroed arena: one memset node instead of one per block (and no memset
    # between two kernels, so the whole chain can use programmatic dependent launch)
    pool_arena = _Arena([N * blk.conv_dw.weight.shape[0] for _, blk in enc.block_list()], torch.float32, dev)
    cur = torch.empty((N * h * w, enc.conv_stem.weight.shape[0]), device=dev, dtype=bf16)
    ops.stem_fwd(x, enc.conv_stem.weight.detach(), cur, out_rec=R["bn1"])
    for name, blk in enc.block_list():
        c = blk.cfg
        k, s = c["k"], c["s"]
        skip = cur if (s == 1 and c["cin"] == c["cout"]) else None
        if c["type"] == "ir":
            rec = R[name + ".bn1"]
            e = ops.gemm(cur, Wp[name + ".conv_pw"][0], ops.EPI_SCALE_SHIFT | ops.EPI_SILU, rec[0], rec[1])
            cm, dwrec, pw, outrec = c["mid"], R[name + ".bn2"], Wp[name + ".conv_pwl"][0], R[name + ".bn3"]
        else:
            e, cm, dwrec, pw, outrec = cur, c["cin"], R[name + ".bn1"], Wp[name + ".conv_pw"][0], R[name + ".bn2"]
        oh, ow = ops.same_out(h, s), ops.same_out(w, s)
        d = torch.empty((N * oh * ow, cm), device=dev, dtype=bf16)
        pooled = pool_arena.take(N * cm, (N, cm))
        ops.dwconv_fwd(e, None, blk.conv_dw.weight.detach(), d, N, h, w, k, s, out_rec=dwrec, pooled=pooled, pooled_zeroed=True)
        # Opt-in (TEETHRT_SE_APPLY=1): on small maps the SE expand launch can gate the activation itself.  Measured SLOWER
        # (batch-1 forward 1.119 vs 1.067 ms, A/B on one box): the expand kernel has C/64 blocks, the separate gate_apply
        # launch spreads the same bytes over the whole GPU and its launch latency is hidden by PDL.
        if N * oh * ow <= 1024 and os.environ.get("TEETHRT_SE_APPLY", "0") == "1":
            _se_gate(blk, pooled, 1.0 / (oh * ow), N, cm, dev, False, apply_x=d, HW=oh * ow)
        else:
            _, gate = _se_gate(blk, pooled, 1.0 / (oh * ow), N, cm, dev, False)
            ops.gate_apply(d, None, gate, d, N, oh * ow)
        flags = ops.EPI_SCALE_SHIFT | (ops.EPI_RESIDUAL if skip is not None else 0)
        cur = ops.gemm(d, pw, flags, outrec[0], outrec[1], residual=skip)
        h, w = oh, ow
    rec = R["bn2"]
    hd = ops.gemm(cur, Wp["conv_head"][0], ops.EPI_SCALE_SHIFT | ops.EPI_SILU, rec[0], rec[1])
    if feature_map:
        return hd.view(N, h, w, enc.num_features).permute(0, 3, 1, 2).float()      # layout + dtype only: NHWC bf16 -> NCHW fp32
    feat = torch.empty((N, enc.num_features), device=dev, dtype=torch.float32)
    ops.pool_act(hd, None, feat, N, h * w, act=0)
    ops.scale_f32(feat, 1.0 / (h * w))
    return feat


class _LazyRec:
    """A train-mode BatchNorm whose statistics are complete but whose record {scale, shift, mean, rstd} is not written yet:
    the FIRST consumer kernel takes `fin()` (a host struct) and finalises it in its prologue - no finalise launch between
    producer and consumer.  Later consumers read `rec` as usual."""

    def __init__(self, rec, fin):
        self.rec, self._fin = rec, fin

    def fin(self):
        f, self._fin = self._fin, None
        return f


def _lazy_bn():
    """TEETHRT_LAZY_BN bit mask: 1 = depthwise conv consumes its input's BatchNorm lazily, 2 = the streaming forward
    consumers (pool_act / bn_apply), 4 = affine2 in the backward pass.  0 = one finalise launch per BatchNorm (round-1
    path; the A/B switch).  Default 7: the library decides per call whether the in-prologue form pays."""
    return int(os.environ.get("TEETHRT_LAZY_BN", "7"))


def forward_train(enc, x, save=True):
    """Train-mode forward (batch statistics, running-stat update).  Returns (feat [N,F] fp32, ctx for backward)."""
    x = _check_input(enc, x)
    dev = x.device
    enc._eval_cache = None          # this pass updates the running statistics (and usually precedes a parameter update)
    N, _, H, W = x.shape
    # the bf16 weight packing (one launch, ~70 us) runs beside the stem / first depthwise layer: its first consumer is the
    # first 1x1 conv, which waits on the event below
    main = torch.cuda.current_stream(dev)
    side = _SideQueue.stream_for(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        Wp = _packed_weights(enc, dev, True)
        packed = torch.cuda.Event()
        packed.record(side)
    pack_pending = [True]

    def need_packed():
        if pack_pending[0]:
            torch.cuda.current_stream(dev).wait_event(packed)
            pack_pending[0] = False
    bns = dict(enc.bn_list())
    total_c = sum(b.weight.numel() for b in bns.values())
    chans = [b.weight.numel() for b in bns.values()]
    stats = _Arena([ops.STAT_REPLICAS * 2 * c for c in chans], torch.float64, dev)
    recs = _Arena([4 * c for c in chans], torch.float32, dev)
    ctx = dict(x=x, N=N, Wp=Wp, rec={}, blocks=[], dims=[])
    se_ws = _se_workspace(enc, N, dev)
    se_fwd_fused, se_min_c = os.environ.get("TEETHRT_SE_FUSED_FWD") == "1", _se_fused_min_c()
    ctx["se_ws"] = se_ws

    pool_arena = _Arena([N * blk.conv_dw.weight.shape[0] for _, blk in enc.block_list()] + [N * enc.num_features],
                        torch.float32, dev)                # every SE / global pooling target of the pass: one memset

    lazy = _lazy_bn()

    def bn_stage(bn_name, count, st, consumer=2):
        """The statistics `st` of BatchNorm `bn_name` have just been produced -> its (lazy) record.  consumer: which kind
        of kernel reads it first (1 = depthwise conv, 2 = pool_act / bn_apply)."""
        bn = bns[bn_name]
        c = bn.weight.numel()
        rec = recs.take(4 * c, (4, c))
        ctx["rec"][bn_name] = rec
        if lazy & consumer:
            return _LazyRec(rec, ops.bn_fin(st, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                            bn.num_batches_tracked, rec, count, BN_EPS, BN_MOM))
        ops.bn_finalize(st, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.num_batches_tracked,
                        rec, count, BN_EPS, BN_MOM)
        return _LazyRec(rec, None)

    h, w = ops.same_out(H, 2), ops.same_out(W, 2)
    cs = enc.conv_stem.weight.shape[0]
    st = stats.take(ops.STAT_REPLICAS * 2 * cs)
    # the stem as tcgen05 GEMMs over an explicit im2col (K = 27 padded to 32); the patches are kept for the weight gradient
    patches = ops.stem_im2col(x, torch.empty((N * h * w, 32), device=dev, dtype=bf16))
    w_stem = ops.stem_pack_w(enc.conv_stem.weight.detach(), torch.empty((cs, 32), device=dev, dtype=bf16))
    s_raw = torch.empty((N * h * w, cs), device=dev, dtype=bf16)
    ops.gemm(patches, w_stem, ops.EPI_STATS, stats=st, out=s_raw)
    first = enc.block_list()[0][1].cfg
    stem_to_dw = first["type"] == "ds" and not (first["s"] == 1 and first["cin"] == first["cout"])     # stem BN consumed by a depthwise conv
    rec_s = bn_stage("bn1", N * h * w, st, 1 if stem_to_dw else 2)
    cur, cur_rec = s_raw, rec_s                                   # lazy: (raw, pending BN+SiLU)
    ctx["stem"] = dict(raw=s_raw, h=h, w=w, patches=patches)
    for name, blk in enc.block_list():
        c = blk.cfg
        k, s = c["k"], c["s"]
        has_skip = s == 1 and c["cin"] == c["cout"]
        sv = dict(name=name, h=h, w=w, in_rec=None)
        if c["type"] == "ir" or has_skip:
            if cur_rec is not None:        # materialise the pending activation (GEMM operand / residual need a real tensor)
                cur = ops.bn_apply(cur, cur_rec.rec, torch.empty_like(cur), act=1, fin=cur_rec.fin())
                sv["materialised_from"] = True
                cur_rec = None
        sv["x"] = cur
        if c["type"] == "ir":
            cm = c["mid"]
            st = stats.take(ops.STAT_REPLICAS * 2 * cm)
            need_packed()
            e_raw = torch.empty((N * h * w, cm), device=dev, dtype=bf16)
            ops.gemm(cur, Wp[name + ".conv_pw"][0], ops.EPI_STATS, stats=st, out=e_raw)
            dw_in, dw_rec, bn_dw, pw_name, bn_out = e_raw, bn_stage(name + ".bn1", N * h * w, st, 1), name + ".bn2", name + ".conv_pwl", name + ".bn3"
            sv["e_raw"] = e_raw
        else:
            cm = c["cin"]
            dw_in, dw_rec, bn_dw, pw_name, bn_out = cur, cur_rec, name + ".bn1", name + ".conv_pw", name + ".bn2"
            sv["in_rec"] = None if cur_rec is None else cur_rec.rec
        oh, ow = ops.same_out(h, s), ops.same_out(w, s)
        d_raw = torch.empty((N * oh * ow, cm), device=dev, dtype=bf16)
        st = stats.take(ops.STAT_REPLICAS * 2 * cm)
        # stride 1: the forward also writes its activated input once, and the weight-gradient kernel of the backward pass reads
        # that instead of recomputing silu(bn(x)) over its halo'd tiles (TEETHRT_DW_SAVE_ACT=0: recompute, the round-1 path)
        dw_act = None
        if save and s == 1 and dw_rec is not None and os.environ.get("TEETHRT_DW_SAVE_ACT", "1") != "0":
            dw_act = torch.empty_like(dw_in)
        ops.dwconv_fwd(dw_in, None if dw_rec is None else dw_rec.rec, blk.conv_dw.weight.detach(), d_raw, N, h, w, k, s, stats=st,
                       in_fin=None if dw_rec is None else dw_rec.fin(), act_out=dw_act)
        rec_d = bn_stage(bn_dw, N * oh * ow, st)
        pooled = pool_arena.take(N * cm, (N, cm))
        ops.pool_act(d_raw, rec_d.rec, pooled, N, oh * ow, act=1, zeroed=True, fin=rec_d.fin())
        s1, gate = _se_gate(blk, pooled, 1.0 / (oh * ow), N, cm, dev, True, ws=se_ws if se_fwd_fused and cm >= se_min_c else None)
        a = ops.gate_apply(d_raw, rec_d.rec, gate, torch.empty_like(d_raw), N, oh * ow)
        st = stats.take(ops.STAT_REPLICAS * 2 * c["cout"])
        need_packed()
        p_raw = torch.empty((N * oh * ow, c["cout"]), device=dev, dtype=bf16)
        ops.gemm(a, Wp[pw_name][0], ops.EPI_STATS, stats=st, out=p_raw)
        rec_o = bn_stage(bn_out, N * oh * ow, st)
        y = ops.bn_apply(p_raw, rec_o.rec, torch.empty_like(p_raw), residual=cur if has_skip else None, act=0, fin=rec_o.fin())
        sv.update(d_raw=d_raw, pooled=pooled, s1=s1, gate=gate, a=a, p_raw=p_raw, oh=oh, ow=ow, has_skip=has_skip,
                  bn_dw=bn_dw, pw_name=pw_name, bn_out=bn_out, dw_act=dw_act)
        ctx["blocks"].append((blk, sv))
        cur, cur_rec, h, w = y, None, oh, ow
    st = stats.take(ops.STAT_REPLICAS * 2 * enc.num_features)
    need_packed()
    hd_raw = torch.empty((N * h * w, enc.num_features), device=dev, dtype=bf16)
    ops.gemm(cur, Wp["conv_head"][0], ops.EPI_STATS, stats=st, out=hd_raw)
    rec_h = bn_stage("bn2", N * h * w, st)
    feat = pool_arena.take(N * enc.num_features, (N, enc.num_features))
    ops.pool_act(hd_raw, rec_h.rec, feat, N, h * w, act=1, zeroed=True, fin=rec_h.fin())
    ops.scale_f32(feat, 1.0 / (h * w))
    ctx.update(head=dict(x=cur, raw=hd_raw, h=h, w=w))
    return feat, ctx


_ones_coef = {}


def _add_coef(C, dev):
    key = (C, dev)
    if key not in _ones_coef:
        t = torch.zeros((3, C), device=dev, dtype=torch.float32)
        t[:2] = 1.0
        _ones_coef[key] = t
    return _ones_coef[key]


class _SideQueue:
    """Weight-gradient GEMMs only feed the optimiser, so they run on a second stream beside the data-gradient chain (in a
    captured graph: a parallel branch).  Every tensor they read is kept alive until join() so the caching allocator cannot
    hand its memory to later work on the main stream."""
    _streams = {}

    @staticmethod
    def stream_for(dev):
        key = (dev.type, dev.index)
        if key not in _SideQueue._streams:
            _SideQueue._streams[key] = torch.cuda.Stream(device=dev)
        return _SideQueue._streams[key]

    def __init__(self, dev):
        self.enabled = os.environ.get("TEETHRT_WGRAD_STREAM", "1") != "0"
        self.keep = []
        if self.enabled:
            self.side = _SideQueue.stream_for(dev)
            self.dev = dev

    @property
    def main(self):
        # looked up at every use: a backward pass split over several captured graphs sees a different capture stream in each
        return torch.cuda.current_stream(self.dev)

    def wgrad(self, P, Q, out, **kw):
        if not self.enabled:
            return ops.gemm_wgrad(P, Q, out, **kw)
        self.keep += [P, Q]
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            ops.gemm_wgrad(P, Q, out, **kw)

    def dw_wgrad(self, dD, w, x_raw, x_rec, dw, N, h, w_, k, s):
        """Depthwise weight gradient: independent of the data gradient (both read dD), so it joins the side stream."""
        if not self.enabled:
            return ops.dwconv_bwd(dD, w, x_raw, x_rec, None, None, dw, N, h, w_, k, s)
        self.keep += [dD, x_raw]
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            ops.dwconv_bwd(dD, w, x_raw, x_rec, None, None, dw, N, h, w_, k, s)

    def join(self):
        if self.enabled:
            self.main.wait_stream(self.side)
        self.keep.clear()


def _stem_wgrad(ctx, ds, grads, sq):
    g = grads["conv_stem.weight"]
    sq.wgrad(ds, ctx["stem"]["patches"], g.view(g.shape[0], 27), so_p=27, so_q=1, q_store=27)


def backward_train(enc, ctx, dfeat, grads):
    """Backward of forward_train.  `grads`: dict name -> fp32 tensor (torch layout) for every parameter of the encoder;
    1x1-conv / depthwise / stem weight gradients are ACCUMULATED (+=) — pass zeroed tensors; the others are written."""
    for _ in backward_train_iter(enc, ctx, dfeat, grads):
        pass
    return grads


def backward_train_iter(enc, ctx, dfeat, grads, split_after=()):
    """Generator form of backward_train: yields the block name after finishing every block listed in `split_after` (reverse
    execution order), with all side-stream work joined - the gradients of that block and of everything executed before it
    are final, so a data-parallel trainer can start their all-reduce while the rest of the backward pass runs."""
    dev = dfeat.device
    N, Wp, REC = ctx["N"], ctx["Wp"], ctx["rec"]
    bns = dict(enc.bn_list())
    total_c = sum(b.weight.numel() for b in bns.values())
    bstats = _Arena([ops.STAT_REPLICAS * 2 * b.weight.numel() for b in bns.values()], torch.float64, dev)
    dfeat = dfeat.contiguous().float()
    sq = _SideQueue(dev)

    merged = os.environ.get("TEETHRT_SE_BWD_MERGED", "1") != "0"     # 0: round-1 path (reduce, act_bwd, affine2: three passes)
    se_min_c = _se_fused_min_c()
    zeros = _Arena([n for blk, sv in ctx["blocks"] for n in (N * (5 if merged else 1) * sv["d_raw"].shape[1], N * blk.cfg["rd"])],
                   torch.float32, dev)

    lazy = _lazy_bn()

    # Optional (TEETHRT_GEMM_BNBWD=1): the data-gradient GEMM that produces a block's input gradient also accumulates the
    # backward sums of the BatchNorm that gradient flows into next (the previous block's project BN), so the separate pass
    # over (dy, p_raw) disappears.  Measured neutral: the 31 reduce launches (0.30 ms serialised) go away, the GEMM epilogue
    # (an extra 16-byte global load per store, spills at 128 registers) gets 0.3-0.4 ms slower - 11.69 ms/step with it,
    # 11.52-11.70 without (A/B on one box).  Off by default; needs the lazy affine2.
    fuse_bnbwd = bool(lazy & 4) and os.environ.get("TEETHRT_GEMM_BNBWD", "0") != "0"

    def bn_dx(bn_name, count, bst, g, x_raw, out, raw_x=False):
        """BatchNorm backward apply: out = a*g + b*x_raw + c with the coefficients of BatchNorm `bn_name`, whose sums `bst`
        the producer of `g` has just accumulated; also lands dgamma / dbeta in `grads`.  Lazy (default): trt_affine2 derives
        the coefficients itself - no finalise launch in between."""
        bn = bns[bn_name]
        dgm, dbt = grads[bn_name + ".weight"], grads[bn_name + ".bias"]
        coef = torch.empty((3, bn.weight.numel()), device=dev, dtype=torch.float32)
        if lazy & 4:
            return ops.affine2(g, x_raw, None, out, fin=ops.bn_bwd_fin(bst, REC[bn_name], bn.weight.detach(), dgm, dbt, count, coef, raw_x))
        assert not raw_x
        ops.bn_bwd_finalize(bst, REC[bn_name], bn.weight.detach(), coef, dgm, dbt, count)
        return ops.affine2(g, x_raw, coef, out)

    # ---- head: feat = mean_hw silu(bn2(conv_head(y)))
    hd = ctx["head"]
    hw = hd["h"] * hd["w"]
    F_ = enc.num_features
    bst = bstats.take(ops.STAT_REPLICAS * 2 * F_)
    g = ops.act_bwd(None, None, dfeat, 1.0 / hw, hd["raw"], REC["bn2"], torch.empty_like(hd["raw"]), bst, N, hw, act=1)
    d_raw = bn_dx("bn2", N * hw, bst, g, hd["raw"], g)
    blocks = ctx["blocks"]

    def dgrad(A, Wt, idx, residual=None):
        """dx of the GEMM whose output is the input gradient of block `idx` - i.e. the output gradient of block idx - 1, whose
        project BatchNorm consumes it next.  -> (dx, bst of that BatchNorm or None)."""
        if fuse_bnbwd and idx >= 1:
            psv = blocks[idx - 1][1]
            bst_prev = bstats.take(ops.STAT_REPLICAS * 2 * psv["p_raw"].shape[1])
            return ops.gemm_bnbwd(A, Wt, psv["p_raw"], bst_prev, residual=residual), bst_prev
        return ops.gemm(A, Wt, ops.EPI_RESIDUAL if residual is not None else 0, residual=residual), None

    dy, dy_bst = dgrad(d_raw, Wp["conv_head"][1], len(blocks))
    sq.wgrad(d_raw, hd["x"], grads["conv_head.weight"])
    del g, d_raw

    for bidx in range(len(blocks) - 1, -1, -1):
        blk, sv = blocks[bidx]
        c, name = blk.cfg, sv["name"]
        k, s = c["k"], c["s"]
        h, w, oh, ow = sv["h"], sv["w"], sv["oh"], sv["ow"]
        ohw = oh * ow
        cm = sv["d_raw"].shape[1]
        # project conv + its BN (no activation)
        if dy_bst is not None:        # the GEMM that produced dy already summed {dy, dy * p_raw}
            dp = bn_dx(sv["bn_out"], N * ohw, dy_bst, dy, sv["p_raw"], torch.empty_like(dy), raw_x=True)
        else:
            bst = bstats.take(ops.STAT_REPLICAS * 2 * c["cout"])
            ops.bn_bwd_reduce(dy, sv["p_raw"], REC[sv["bn_out"]], bst)
            dp = bn_dx(sv["bn_out"], N * ohw, bst, dy, sv["p_raw"], torch.empty_like(dy))
        dy_bst = None
        dA = ops.gemm(dp, Wp[sv["pw_name"]][1])
        sq.wgrad(dp, sv["a"], grads[sv["pw_name"] + ".weight"])
        # squeeze-excite + activation + BN of the depthwise output
        rec_d = REC[sv["bn_dw"]]
        ds2, dmean = torch.empty((N, cm), device=dev, dtype=torch.float32), torch.empty((N, cm), device=dev, dtype=torch.float32)
        ds1 = zeros.take(N * c["rd"], (N, c["rd"]))
        se = blk.se
        se_args = (sv["gate"], sv["s1"], sv["pooled"], 1.0 / ohw, se.conv_reduce.weight.detach(), se.conv_expand.weight.detach(),
                   ds2, ds1, dmean, grads[name + ".se.conv_reduce.weight"], grads[name + ".se.conv_reduce.bias"],
                   grads[name + ".se.conv_expand.weight"], grads[name + ".se.conv_expand.bias"])
        if merged:
            # two passes over (dA, d_raw) instead of three: pass 1 leaves five per-(image, channel) sums, the SE MLP backward
            # turns them into the BatchNorm-backward coefficients, pass 2 writes dD directly (no g tensor, no affine pass)
            sums = zeros.take(5 * N * cm, (5, N, cm))
            ops.se_bwd_reduce(dA, sv["d_raw"], rec_d, sums, N, ohw, zeroed=True, full=True)
            bn = bns[sv["bn_dw"]]
            coef_d = torch.empty((3, cm), device=dev, dtype=torch.float32)
            ops.se_bwd(sums[0], *se_args, ds1_zeroed=True, ws=ctx.get("se_ws") if cm >= se_min_c else None,
                       bn=ops.se_bn(sums, rec_d, bn.weight.detach(), coef_d, grads[sv["bn_dw"] + ".weight"], grads[sv["bn_dw"] + ".bias"], N * ohw))
            dD = ops.act_bwd_apply(dA, sv["gate"], dmean, 1.0 / ohw, sv["d_raw"], rec_d, coef_d, dA, N, ohw)
        else:
            dgate_pre = zeros.take(N * cm, (N, cm))
            ops.se_bwd_reduce(dA, sv["d_raw"], rec_d, dgate_pre, N, ohw, zeroed=True)
            ops.se_bwd(dgate_pre, *se_args, ds1_zeroed=True, ws=ctx.get("se_ws") if cm >= se_min_c else None)
            bst = bstats.take(ops.STAT_REPLICAS * 2 * cm)
            g2 = ops.act_bwd(dA, sv["gate"], dmean, 1.0 / ohw, sv["d_raw"], rec_d, dA, bst, N, ohw, act=1)
            dD = bn_dx(sv["bn_dw"], N * ohw, bst, g2, sv["d_raw"], g2)           # gradient w.r.t. the raw depthwise output
        dw_grad = grads[name + ".conv_dw.weight"]
        if c["type"] == "ir":
            e_raw, rec1 = sv["e_raw"], REC[name + ".bn1"]
            bst = bstats.take(ops.STAT_REPLICAS * 2 * cm)
            g1 = torch.empty_like(e_raw)
            if sv.get("dw_act") is not None:       # the forward saved silu(bn1(e_raw)): no activation in the weight-gradient kernel
                sq.dw_wgrad(dD, blk.conv_dw.weight.detach(), sv["dw_act"], None, dw_grad, N, h, w, k, s)
            else:
                sq.dw_wgrad(dD, blk.conv_dw.weight.detach(), e_raw, rec1, dw_grad, N, h, w, k, s)
            ops.dwconv_bwd(dD, blk.conv_dw.weight.detach(), e_raw, rec1, g1, bst, None, N, h, w, k, s)
            de = bn_dx(name + ".bn1", N * h * w, bst, g1, e_raw, g1)
            dx, dy_bst = dgrad(de, Wp[name + ".conv_pw"][1], bidx, residual=dy if sv["has_skip"] else None)
            sq.wgrad(de, sv["x"], grads[name + ".conv_pw.weight"])
            dy = dx
        else:
            x_in, in_rec = sv["x"], sv["in_rec"]
            is_first = blk is ctx["blocks"][0][0]
            if in_rec is not None:                       # input was the (lazy) stem output: BN+SiLU applied on load
                bst = bstats.take(ops.STAT_REPLICAS * 2 * c["cin"])
                g_in = torch.empty_like(x_in)
                if sv.get("dw_act") is not None:
                    sq.dw_wgrad(dD, blk.conv_dw.weight.detach(), sv["dw_act"], None, dw_grad, N, h, w, k, s)
                else:
                    sq.dw_wgrad(dD, blk.conv_dw.weight.detach(), x_in, in_rec, dw_grad, N, h, w, k, s)
                ops.dwconv_bwd(dD, blk.conv_dw.weight.detach(), x_in, in_rec, g_in, bst, None, N, h, w, k, s)
                ds = bn_dx("bn1", N * h * w, bst, g_in, x_in, g_in)
                _stem_wgrad(ctx, ds, grads, sq)
                dy = None
            else:
                g_in = torch.empty_like(x_in)
                sq.dw_wgrad(dD, blk.conv_dw.weight.detach(), x_in, None, dw_grad, N, h, w, k, s)
                ops.dwconv_bwd(dD, blk.conv_dw.weight.detach(), x_in, None, g_in, None, None, N, h, w, k, s)
                if sv["has_skip"]:
                    g_in = ops.affine2(g_in, dy, _add_coef(c["cin"], dev), g_in)
                dy = g_in
                if sv.get("materialised_from") and is_first:
                    # the block input was silu(bn1(stem)) materialised by bn_apply: continue into the stem
                    st_raw = ctx["stem"]["raw"]
                    hw0 = ctx["stem"]["h"] * ctx["stem"]["w"]
                    bst = bstats.take(ops.STAT_REPLICAS * 2 * c["cin"])
                    g_s = ops.act_bwd(dy, None, None, 0.0, st_raw, REC["bn1"], torch.empty_like(st_raw), bst, N, hw0, act=1)
                    ds = bn_dx("bn1", N * hw0, bst, g_s, st_raw, g_s)
                    _stem_wgrad(ctx, ds, grads, sq)
                    dy = None
        if c["type"] == "ir" and sv.get("materialised_from"):
            st_raw = ctx["stem"]["raw"]
            hw0 = ctx["stem"]["h"] * ctx["stem"]["w"]
            bst = bstats.take(ops.STAT_REPLICAS * 2 * st_raw.shape[1])
            g_s = ops.act_bwd(dy, None, None, 0.0, st_raw, REC["bn1"], torch.empty_like(st_raw), bst, N, hw0, act=1)
            ds = bn_dx("bn1", N * hw0, bst, g_s, st_raw, g_s)
            _stem_wgrad(ctx, ds, grads, sq)
            dy = None
        if name in split_after:
            sq.join()
            yield name
    sq.join()


class _EncoderFn(torch.autograd.Function):
    """autograd bridge for the reference's own training loop (loss.backward() at train_mm_joint_dualtask.py:248)."""

    @staticmethod
    def forward(ctx, enc, x, *params):
        feat, saved = forward_train(enc, x)
        ctx.enc, ctx.saved = enc, saved
        ctx.names = [n for n, _ in enc.named_parameters()]
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        enc = ctx.enc
        named = dict(enc.named_parameters())
        sink = enc.grad_sink
        grads = {n: (sink[n] if sink is not None else torch.zeros_like(p)) for n, p in named.items()}
        backward_train(enc, ctx.saved, dfeat, grads)
        ctx.saved = None
        out = tuple(grads[n] if named[n].requires_grad else None for n in ctx.names)
        return (None, None) + out


def encoder_forward(enc, x):
    if enc.training:
        params = [p for _, p in enc.named_parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _EncoderFn.apply(enc, x, *params)
        return forward_train(enc, x)[0]
    if torch.is_grad_enabled() and any(p.requires_grad for p in enc.parameters()) and x.requires_grad:
        raise NotImplementedError("teethrt: gradients w.r.t. the input image in eval mode are not on the hot path")
    return forward_eval(enc, x)
