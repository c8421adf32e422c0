"""GPU parity tests, kernel level: every C-ABI entry point of libteethrt against a plain PyTorch fp32 restatement of the
same op (the floating-point kernels) — byte/integer kernels are in test_preproc_gpu.py against the oracle.
Tolerances: bf16 storage => relative 1e-2 of the tensor's max magnitude; fp32 kernels 1e-4/1e-5."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    import teethrt
    teethrt.init()
    from teethrt import ops as o
    return o


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def close(a, b, rtol, atol=1e-5):
    """max-norm closeness with an absolute floor (some gradients are mathematically zero, e.g. a bias in front of a
    BatchNorm or the scalar bias in front of a softmax)."""
    a, b = a.float(), b.float()
    return float((a - b).abs().max()) <= rtol * float(b.abs().max()) + atol


def rnd(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


# ------------------------------------------------------------------------------------------------ tcgen05 GEMMs
GEMM_SHAPES = [(256, 144, 24), (1000, 32, 144), (647, 336, 56), (3136, 1792, 448), (64, 24, 144), (12544, 48, 24),
               (130, 272, 960), (49, 448, 2688), (8192, 256, 64),
               # several n-blocks with the weights of ONE n-block resident per CTA (grid = a multiple of the n-block count)
               (12544, 960, 160), (3136, 1632, 272), (3136, 2688, 448), (50176, 336, 56), (19000, 672, 112), (129, 1000, 72)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(ops, M, N, K):
    A, B = rnd(M, K, seed=1, dtype=bf16), rnd(N, K, scale=K ** -0.5, seed=2, dtype=bf16)
    C = ops.gemm(A, B)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert rel_err(C, ref) < 1e-2


# Many tiles per persistent CTA (several rounds of every epilogue group / TMEM accumulator): the path on which mbarrier
# parity aliasing or a staging-buffer race would show up as a hang or wrong rows; small shapes never reach it.
@pytest.mark.parametrize("N,K", [(24, 24), (144, 24), (192, 32), (56, 192), (272, 160), (672, 112), (960, 160)])
def test_gemm_many_tiles_per_cta(ops, N, K):
    M = 148 * 128 * 5 + 77
    A, B = rnd(M, K, seed=11, dtype=bf16), rnd(N, K, scale=K ** -0.5, seed=12, dtype=bf16)
    stats = ops.new_stats(N, "cuda")
    for _ in range(3):                       # repeated launches: the hang this guards against was intermittent
        stats.zero_()
        C = ops.gemm(A, B, ops.EPI_STATS, stats=stats)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert rel_err(C, ref) < 1e-2
    cf = C.double()
    tot = ops.stats_total(stats)
    assert torch.allclose(tot[0], cf.sum(0), rtol=1e-4, atol=5e-2)
    assert torch.allclose(tot[1], (cf * cf).sum(0), rtol=1e-4, atol=5e-2)
    dA = ops.gemm(C, B.t().contiguous())     # dgrad orientation (K and N swap roles)
    assert rel_err(dA, C.float() @ B.float()) < 1e-2


@pytest.mark.parametrize("M,N,K", [(1000, 144, 24), (647, 336, 56), (3136, 672, 112), (12544, 960, 160), (3136, 1632, 272),
                                   # M <= 64 (batch-1 7x7 maps): half-height A box unless the epilogue takes statistics
                                   (49, 272, 1632), (64, 160, 960), (17, 112, 672)])
def test_gemm_epilogues(ops, M, N, K):
    A, B = rnd(M, K, seed=3, dtype=bf16), rnd(N, K, scale=K ** -0.5, seed=4, dtype=bf16)
    sc, sh = rnd(N, seed=5) * 0.2 + 1.0, rnd(N, seed=6) * 0.3
    res = rnd(M, N, seed=7, dtype=bf16)
    acc = A.float() @ B.float().t()
    C = ops.gemm(A, B, ops.EPI_SCALE_SHIFT | ops.EPI_SILU, sc, sh)
    assert rel_err(C, F.silu(acc * sc + sh)) < 1e-2
    C = ops.gemm(A, B, ops.EPI_SCALE_SHIFT | ops.EPI_RESIDUAL, sc, sh, residual=res)
    assert rel_err(C, acc * sc + sh + res.float()) < 1e-2
    C = ops.gemm(A, B, ops.EPI_RESIDUAL, residual=res)
    assert rel_err(C, acc + res.float()) < 1e-2
    stats = ops.new_stats(N, "cuda")
    C = ops.gemm(A, B, ops.EPI_STATS, stats=stats)
    torch.cuda.synchronize()
    assert rel_err(C, acc) < 1e-2
    cf = C.double()
    tot = ops.stats_total(stats)
    assert torch.allclose(tot[0], cf.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[1], (cf * cf).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("M,N,K,res", [(3136, 448, 2688, True), (50176, 56, 336, False), (200, 24, 144, True), (12544, 160, 960, True)])
def test_gemm_with_batchnorm_backward_sums(ops, M, N, K, res):
    """Data-gradient GEMM whose epilogue also accumulates {sum C, sum C * x} per output channel over the stored (bf16) C: the
    backward sums of the BatchNorm (input x) that C flows into next - equal to the separate bn_bwd_reduce pass, and through
    the lazy affine2 (raw_x) equal to bn_bwd_reduce + bn_bwd_finalize + affine2."""
    A, B = rnd(M, K, seed=61, dtype=bf16), rnd(N, K, seed=62, dtype=bf16, scale=K ** -0.5)
    x = rnd(M, N, seed=63, dtype=bf16) * 1.5 + 0.4
    r = rnd(M, N, seed=64, dtype=bf16) if res else None
    bst = ops.new_stats(N, "cuda")
    C = ops.gemm_bnbwd(A, B, x, bst, residual=r)
    want = ops.gemm(A, B, ops.EPI_RESIDUAL if res else 0, residual=r)
    assert torch.equal(C, want)
    Cd, xd = C.double(), x.double()
    tot = ops.stats_total(bst)
    assert torch.allclose(tot[0], Cd.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(tot[1], (Cd * xd).sum(0), rtol=1e-4, atol=5e-2)
    # through the lazy BatchNorm-backward apply
    gamma, beta = rnd(N, seed=65) * 0.1 + 1, rnd(N, seed=66) * 0.1
    rec, _, _, _ = make_rec(ops, x, gamma, beta)
    bs2 = ops.new_stats(N, "cuda")
    ops.bn_bwd_reduce(C, x, rec, bs2)
    coef, dg0, db0 = torch.empty(3, N, device="cuda"), torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    ops.bn_bwd_finalize(bs2, rec, gamma, coef, dg0, db0, M)
    ref = ops.affine2(C, x, coef, torch.empty_like(C))
    dg, db = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    got = ops.affine2(C, x, None, torch.empty_like(C), fin=ops.bn_bwd_fin(bst, rec, gamma, dg, db, M, torch.empty(3, N, device="cuda"), raw_x=True))
    assert rel_err(got, ref) < 2e-3 and rel_err(dg, dg0) < 2e-3 and rel_err(db, db0) < 1e-4


@pytest.mark.parametrize("M,Cp,Cq", [(4096, 144, 24), (1000, 24, 144), (50000, 32, 192), (3136, 448, 2688), (777, 1632, 272),
                                     (64, 48, 24)])
def test_gemm_wgrad(ops, M, Cp, Cq):
    P, Q = rnd(M, Cp, seed=8, dtype=bf16), rnd(M, Cq, seed=9, dtype=bf16)
    out = torch.zeros(Cp, Cq, device="cuda")
    ops.gemm_wgrad(P, Q, out)
    ref = P.float().t() @ Q.float()
    assert rel_err(out, ref) < 1e-2
    out_t = torch.zeros(Cq, Cp, device="cuda")           # transposed output through strides, accumulating twice
    ops.gemm_wgrad(P, Q, out_t, so_p=1, so_q=Cp)
    ops.gemm_wgrad(P, Q, out_t, so_p=1, so_q=Cp)
    assert rel_err(out_t, 2 * ref.t()) < 1e-2


def test_pack_w1x1(ops):
    w = rnd(144, 24, 1, 1, seed=10)
    o, ot = torch.empty(144, 24, device="cuda", dtype=bf16), torch.empty(24, 144, device="cuda", dtype=bf16)
    ops.pack_w1x1(w, o, ot)
    assert torch.equal(o, w.view(144, 24).to(bf16)) and torch.equal(ot, w.view(144, 24).t().contiguous().to(bf16))


@pytest.mark.parametrize("count", [1, 5, 40, 70])
def test_pack_w1x1_batch(ops, count):
    """All 1x1 weights of a model in one launch: every block finds its table entry (first tiles ascending, more than one
    warp's worth of entries at count = 40 / 70), ragged N / K, entries with and without the transposed copy."""
    g = torch.Generator().manual_seed(count)
    shapes = [(int(torch.randint(1, 150, (1,), generator=g)), int(torch.randint(1, 150, (1,), generator=g))) for _ in range(count)]
    ws = [rnd(n, k, seed=100 + i) for i, (n, k) in enumerate(shapes)]
    outs = [torch.zeros(n, k, device="cuda", dtype=bf16) for n, k in shapes]
    outs_t = [torch.zeros(k, n, device="cuda", dtype=bf16) if i % 3 else None for i, (n, k) in enumerate(shapes)]
    rows, tiles = [], 0
    for w, o, ot, (n, k) in zip(ws, outs, outs_t, shapes):
        rows.append([w.data_ptr(), o.data_ptr(), ot.data_ptr() if ot is not None else 0, n, k, tiles])
        tiles += ((n + 31) // 32) * ((k + 31) // 32)
    ops.pack_w1x1_batch(torch.tensor(rows, dtype=torch.int64).cuda(), tiles)
    for w, o, ot in zip(ws, outs, outs_t):
        assert torch.equal(o, w.to(bf16))
        if ot is not None:
            assert torch.equal(ot, w.t().contiguous().to(bf16))


# ------------------------------------------------------------------------------------------------ BN + SE composite
def make_rec(ops, xr, gamma, beta, eps=1e-3):
    """rec via trt_bn_finalize from fp64 sums of the stored (bf16) tensor; also returns the running stats it updated."""
    C = xr.shape[-1]
    x2 = xr.reshape(-1, C).double()
    stats = ops.new_stats(x2.shape[1], "cuda")
    stats[0] = torch.stack([x2.sum(0), (x2 * x2).sum(0)])
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.int64)
    rec = torch.empty(4, C, device="cuda")
    ops.bn_finalize(stats, gamma, beta, rm, rv, nbt, rec, x2.shape[0], eps)
    return rec, rm, rv, nbt


@pytest.mark.parametrize("N,HW,C,rd", [(4, 49, 144, 6), (3, 196, 24, 6), (2, 64, 2688, 112), (5, 33, 336, 14), (70, 9, 48, 12), (96, 4, 1632, 68)])
@pytest.mark.parametrize("narrow", ["0", "1"], ids=["full-width", "narrow-slabs"])
def test_bn_se_block_forward_backward(ops, N, HW, C, rd, narrow, monkeypatch):
    monkeypatch.setenv("TEETHRT_NARROW_SLABS", narrow)      # opt-in geometry of the per-image reductions (one pass, plain stores)
    x_raw = rnd(N, HW, C, seed=11, dtype=bf16) * 1.5 + 0.3
    gamma, beta = rnd(C, seed=12) * 0.1 + 1, rnd(C, seed=13) * 0.1
    Wr, br = rnd(rd, C, seed=14, scale=C ** -0.5), rnd(rd, seed=15, scale=0.1)
    We, be = rnd(C, rd, seed=16, scale=rd ** -0.5), rnd(C, seed=17, scale=0.1)
    dA = rnd(N, HW, C, seed=18, dtype=bf16)
    # ---- torch reference (fp32 autograd on the same bf16-rounded inputs)
    xt = x_raw.float().requires_grad_(True)
    ps = [t.clone().requires_grad_(True) for t in (gamma, beta, Wr, br, We, be)]
    g_, b_, Wr_, br_, We_, be_ = ps
    rm_ref, rv_ref = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    y = F.batch_norm(xt.permute(0, 2, 1), rm_ref, rv_ref, g_, b_, True, 0.1, 1e-3).permute(0, 2, 1)
    y = F.silu(y)
    s = y.mean(1)
    s1_ref = s @ Wr_.t() + br_
    gate_ref = torch.sigmoid(F.silu(s1_ref) @ We_.t() + be_)
    A_ref = y * gate_ref[:, None, :]
    A_ref.backward(dA.float())
    # ---- kernels
    rec, rm, rv, nbt = make_rec(ops, x_raw, gamma, beta)
    assert torch.allclose(rm, rm_ref, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5) and int(nbt) == 1
    x2 = x_raw.view(N * HW, C)
    pooled = torch.empty(N, C, device="cuda")
    ops.pool_act(x2, rec, pooled, N, HW, act=1)
    assert rel_err(pooled / HW, s) < 2e-3
    s1, gate = torch.empty(N, rd, device="cuda"), torch.empty(N, C, device="cuda")
    ops.se_fwd(pooled, 1.0 / HW, Wr, br, We, be, s1, gate)
    assert rel_err(s1, s1_ref) < 2e-3 and rel_err(gate, gate_ref) < 2e-3
    A = ops.gate_apply(x2, rec, gate, torch.empty_like(x2), N, HW)
    assert rel_err(A, A_ref.reshape(N * HW, C)) < 1e-2
    y_only = ops.bn_apply(x2, rec, torch.empty_like(x2), act=1)
    assert rel_err(y_only, y.reshape(N * HW, C)) < 1e-2
    # inference flavour for small maps: the SE expand launch gates the (already activated) tensor in place
    gated = y_only.clone()
    s1b, gate_b = torch.empty_like(s1), torch.empty_like(gate)
    ops.se_fwd(pooled, 1.0 / HW, Wr, br, We, be, s1b, gate_b, apply_x=gated, HW=HW)
    assert torch.allclose(gate_b, gate, rtol=1e-5, atol=1e-6)       # N <= 4 takes the small-batch kernels without apply_x
    assert rel_err(gated.view(N, HW, C), y_only.float().view(N, HW, C) * gate[:, None, :]) < 1e-2
    y_res = ops.bn_apply(x2, rec, torch.empty_like(x2), residual=dA.view(N * HW, C), act=0)
    ref_res = (x_raw.float() * rec[0] + rec[1] + dA.float()).view(N * HW, C)
    assert rel_err(y_res, ref_res) < 1e-2
    # backward
    dgate_pre = torch.empty(N, C, device="cuda")
    ops.se_bwd_reduce(dA.view(N * HW, C), x2, rec, dgate_pre, N, HW)
    ds2, ds1, dmean = torch.empty(N, C, device="cuda"), torch.empty(N, rd, device="cuda"), torch.empty(N, C, device="cuda")
    dWr, dbr, dWe, dbe = torch.empty_like(Wr), torch.empty_like(br), torch.empty_like(We), torch.empty_like(be)
    ops.se_bwd(dgate_pre, gate, s1, pooled, 1.0 / HW, Wr, We, ds2, ds1, dmean, dWr, dbr, dWe, dbe)
    for got, want in ((dWr, Wr_.grad), (dbr, br_.grad), (dWe, We_.grad), (dbe, be_.grad)):
        assert rel_err(got, want) < 2e-2
    bstats = ops.new_stats(C, "cuda")
    g = ops.act_bwd(dA.view(N * HW, C), gate, dmean, 1.0 / HW, x2, rec, torch.empty_like(x2), bstats, N, HW, act=1)
    coef, dgamma, dbeta = torch.empty(3, C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_bwd_finalize(bstats, rec, gamma, coef, dgamma, dbeta, N * HW)
    dx = ops.affine2(g, x2, coef, torch.empty_like(x2))
    assert rel_err(dgamma, g_.grad) < 2e-2 and rel_err(dbeta, b_.grad) < 2e-2
    assert rel_err(dx, xt.grad.reshape(N * HW, C)) < 2e-2
    # merged path: pass 1 leaves five per-(image, channel) sums, the SE MLP backward derives the BatchNorm-backward
    # coefficients from them, pass 2 writes dx directly (two passes over the tensor instead of three)
    sums = torch.empty(5, N, C, device="cuda")
    ops.se_bwd_reduce(dA.view(N * HW, C), x2, rec, sums, N, HW, full=True)
    assert torch.allclose(sums[0], dgate_pre, rtol=1e-4, atol=1e-3)
    ds2b, ds1b, dmeanb = torch.empty_like(ds2), torch.empty_like(ds1), torch.empty_like(dmean)
    gW = [torch.empty_like(t) for t in (dWr, dbr, dWe, dbe)]
    coef2, dgamma2, dbeta2 = torch.empty(3, C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.se_bwd(sums[0], gate, s1, pooled, 1.0 / HW, Wr, We, ds2b, ds1b, dmeanb, *gW,
               bn=ops.se_bn(sums, rec, gamma, coef2, dgamma2, dbeta2, N * HW))
    assert rel_err(dmeanb, dmean) < 1e-3 and rel_err(gW[0], dWr) < 1e-3 and rel_err(gW[2], dWe) < 1e-3
    dx2 = ops.act_bwd_apply(dA.view(N * HW, C), gate, dmeanb, 1.0 / HW, x2, rec, coef2, torch.empty_like(x2), N, HW)
    assert rel_err(dgamma2, g_.grad) < 2e-2 and rel_err(dbeta2, b_.grad) < 2e-2
    assert rel_err(dx2, xt.grad.reshape(N * HW, C)) < 2e-2
    assert rel_err(dx2, dx) < 1e-2 and rel_err(coef2, coef) < 1e-2
    # standalone reduce == the fused sums
    bs2 = torch.zeros_like(bstats)
    ops.bn_bwd_reduce(g, x2, rec, bs2)
    assert torch.allclose(ops.stats_total(bs2), ops.stats_total(bstats), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("N,C,rd", [(64, 2688, 112), (64, 144, 6), (64, 1632, 68), (8, 48, 12), (70, 336, 14), (130, 960, 40), (256, 24, 6)])
def test_se_mlp_one_launch_matches_two_launch(ops, N, C, rd):
    """trt_se_fwd_fused / trt_se_bwd_fused (one launch, split-K over channel chunks behind a grid barrier) against the two-launch
    kernels on the same inputs, and the forward against fp64 torch.  The workspace is reused for several calls in a row: the
    barrier state must come back to a usable state by itself."""
    HW = 49
    pooled = (rnd(N, C, seed=51).abs() + 0.2) * HW
    Wr, br = rnd(rd, C, seed=52, scale=C ** -0.5), rnd(rd, seed=53, scale=0.1)
    We, be = rnd(C, rd, seed=54, scale=rd ** -0.5), rnd(C, seed=55, scale=0.1)
    ws = ops.se_workspace(N, C, rd, "cuda")
    s1a, ga = torch.empty(N, rd, device="cuda"), torch.empty(N, C, device="cuda")
    ops.se_fwd(pooled, 1.0 / HW, Wr, br, We, be, s1a, ga)
    for _ in range(3):
        s1b, gb = torch.full((N, rd), float("nan"), device="cuda"), torch.full((N, C), float("nan"), device="cuda")
        ops.se_fwd(pooled, 1.0 / HW, Wr, br, We, be, s1b, gb, ws=ws)
        assert torch.allclose(s1b, s1a, rtol=1e-4, atol=1e-5) and torch.allclose(gb, ga, rtol=1e-4, atol=1e-5)
    s1_ref = (pooled.double() / HW) @ Wr.double().t() + br.double()
    gate_ref = torch.sigmoid(F.silu(s1_ref) @ We.double().t() + be.double())
    assert rel_err(s1b, s1_ref.float()) < 1e-4 and rel_err(gb, gate_ref.float()) < 1e-4
    # backward, without and with the BatchNorm tail
    dgate_pre = rnd(N, C, seed=56)
    sums = rnd(5, N, C, seed=57)
    sums[0] = dgate_pre
    rec = torch.stack([rnd(C, seed=58) * 0.1 + 1, rnd(C, seed=59) * 0.1, rnd(C, seed=60) * 0.2, rnd(C, seed=61).abs() + 0.5])
    gamma = rnd(C, seed=62) * 0.1 + 1

    def run(ws_):
        outs = dict(ds2=torch.empty(N, C, device="cuda"), ds1=torch.zeros(N, rd, device="cuda"), dmean=torch.empty(N, C, device="cuda"),
                    dWr=torch.empty_like(Wr), dbr=torch.empty_like(br), dWe=torch.empty_like(We), dbe=torch.empty_like(be),
                    coef=torch.empty(3, C, device="cuda"), dgamma=torch.empty(C, device="cuda"), dbeta=torch.empty(C, device="cuda"))
        ops.se_bwd(sums[0], ga, s1a, pooled, 1.0 / HW, Wr, We, outs["ds2"], outs["ds1"], outs["dmean"], outs["dWr"], outs["dbr"],
                   outs["dWe"], outs["dbe"], ds1_zeroed=True,
                   bn=ops.se_bn(sums, rec, gamma, outs["coef"], outs["dgamma"], outs["dbeta"], N * HW), ws=ws_)
        return outs
    want = run(None)
    for _ in range(2):
        got = run(ws)
        for k in ("ds2", "dmean", "dWr", "dbr", "dWe", "dbe", "coef", "dgamma", "dbeta"):
            assert rel_err(got[k], want[k]) < 2e-4, k
    # ds1: the two-launch path leaves the pre-silu' accumulator in its buffer, the one-launch path the finished gradient
    sg = torch.sigmoid(s1a.double())
    ds1_ref = ((dgate_pre * ga * (1 - ga)).double() @ We.double()) * (sg * (1 + s1a.double() * (1 - sg)))
    assert rel_err(got["ds1"], ds1_ref.float()) < 1e-4
    # plain (no BatchNorm tail), ds2 not requested
    dmean2, dWr2 = torch.empty(N, C, device="cuda"), torch.empty_like(Wr)
    ops.se_bwd(dgate_pre, ga, s1a, pooled, 1.0 / HW, Wr, We, None, torch.empty(N, rd, device="cuda"), dmean2, dWr2,
               torch.empty_like(br), torch.empty_like(We), torch.empty_like(be), ws=ws)
    assert rel_err(dmean2, want["dmean"]) < 2e-4 and rel_err(dWr2, want["dWr"]) < 2e-4


@pytest.mark.parametrize("force", ["1", "0"])
@pytest.mark.parametrize("N,H,W,C", [(3, 14, 14, 144), (2, 9, 7, 2688), (5, 6, 6, 24), (2, 5, 5, 4352)])
def test_lazy_batchnorm_records(ops, N, H, W, C, force, monkeypatch):
    """Consumer-side finalisation: dwconv_fwd / pool_act / bn_apply given the producer's statistics (trt_bn_fin_t) must
    (a) compute what they compute from a finalised record, bit for bit, and (b) publish the same record and running
    statistics trt_bn_finalize writes; affine2 given the backward sums (trt_bn_bwd_fin_t) must equal bn_bwd_finalize + affine2.
    force = 1: the in-prologue form; force = 0: the entry point's own finalise launch (what it picks for small tensors)."""
    monkeypatch.setenv("TEETHRT_LAZY_FORCE", force)
    HW = H * W
    x = rnd(N, H, W, C, seed=41, dtype=bf16) * 1.3 + 0.4
    gamma, beta = rnd(C, seed=42) * 0.1 + 1, rnd(C, seed=43) * 0.1
    x2 = x.view(N * HW, C)
    xd = x2.double()
    stats = ops.new_stats(C, "cuda")
    # spread the sums over the replicas the way producers do
    for r in range(ops.STAT_REPLICAS):
        part = xd[r::ops.STAT_REPLICAS]
        stats[r] = torch.stack([part.sum(0), (part * part).sum(0)])

    def fresh():
        return (torch.full((C,), 0.25, device="cuda"), torch.full((C,), 2.0, device="cuda"), torch.zeros((), device="cuda", dtype=torch.int64),
                torch.zeros(4, C, device="cuda"))
    rm0, rv0, nbt0, rec0 = fresh()
    ops.bn_finalize(stats, gamma, beta, rm0, rv0, nbt0, rec0, N * HW, 1e-3)

    def check_published(rm, rv, nbt, rec):
        assert torch.equal(rec, rec0) and torch.equal(rm, rm0) and torch.equal(rv, rv0) and int(nbt) == 1

    # bn_apply (with residual)
    res = rnd(N * HW, C, seed=44, dtype=bf16)
    want = ops.bn_apply(x2, rec0, torch.empty_like(x2), residual=res, act=1)
    rm, rv, nbt, rec = fresh()
    got = ops.bn_apply(x2, rec, torch.empty_like(x2), residual=res, act=1,
                       fin=ops.bn_fin(stats, gamma, beta, rm, rv, nbt, rec, N * HW, 1e-3))
    assert torch.equal(got, want)
    check_published(rm, rv, nbt, rec)
    # pool_act
    want = ops.pool_act(x2, rec0, torch.empty(N, C, device="cuda"), N, HW, act=1)
    rm, rv, nbt, rec = fresh()
    got = ops.pool_act(x2, rec, torch.empty(N, C, device="cuda"), N, HW, act=1, fin=ops.bn_fin(stats, gamma, beta, rm, rv, nbt, rec, N * HW, 1e-3))
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-4)       # atomics: summation order differs between launches
    check_published(rm, rv, nbt, rec)
    # dwconv_fwd with a lazy input BatchNorm
    if C <= 2688:
        w = rnd(C, 1, 3, 3, seed=45, scale=1 / 3)
        st_a, st_b = ops.new_stats(C, "cuda"), ops.new_stats(C, "cuda")
        want = ops.dwconv_fwd(x, rec0, w, torch.empty_like(x), N, H, W, 3, 1, stats=st_a)
        rm, rv, nbt, rec = fresh()
        got = ops.dwconv_fwd(x, rec, w, torch.empty_like(x), N, H, W, 3, 1, stats=st_b,
                             in_fin=ops.bn_fin(stats, gamma, beta, rm, rv, nbt, rec, N * HW, 1e-3))
        assert torch.equal(got, want)
        check_published(rm, rv, nbt, rec)
    # backward: affine2 with lazy coefficients
    g = rnd(N * HW, C, seed=46, dtype=bf16)
    bst = ops.new_stats(C, "cuda")
    ops.bn_bwd_reduce(g, x2, rec0, bst)
    coef, dg0, db0 = torch.empty(3, C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_bwd_finalize(bst, rec0, gamma, coef, dg0, db0, N * HW)
    want = ops.affine2(g, x2, coef, torch.empty_like(x2))
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    got = ops.affine2(g, x2, None, torch.empty_like(x2), fin=ops.bn_bwd_fin(bst, rec0, gamma, dg, db, N * HW, torch.empty(3, C, device="cuda")))
    assert torch.equal(got, want) and torch.equal(dg, dg0) and torch.equal(db, db0)


def test_bn_fold_eval_and_act0(ops):
    C = 56
    gamma, beta, rm, rv = rnd(C, seed=1) + 2, rnd(C, seed=2), rnd(C, seed=3), rnd(C, seed=4).abs() + 0.5
    rec = torch.empty(4, C, device="cuda")
    ops.bn_fold_eval(gamma, beta, rm, rv, rec, 1e-3)
    sc = gamma / torch.sqrt(rv + 1e-3)
    assert torch.allclose(rec[0], sc, rtol=1e-5) and torch.allclose(rec[1], beta - rm * sc, rtol=1e-5, atol=1e-6)
    # act=0 path of act_bwd (BN without activation, e.g. head pooling backward uses dmean only)
    N, HW = 3, 10
    x = rnd(N * HW, C, seed=5, dtype=bf16)
    dmean = rnd(N, C, seed=6)
    bst = ops.new_stats(C, "cuda")
    g = ops.act_bwd(None, None, dmean, 1.0 / HW, x, rec, torch.empty_like(x), bst, N, HW, act=0)
    ref = (dmean / HW)[:, None, :].expand(N, HW, C).reshape(N * HW, C)
    assert rel_err(g, ref) < 1e-2


# ------------------------------------------------------------------------------------------------ depthwise conv
def same_pad_t(x, k, s):
    H, W = x.shape[-2:]
    ph = max((math.ceil(H / s) - 1) * s + k - H, 0)
    pw = max((math.ceil(W / s) - 1) * s + k - W, 0)
    return F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))


@pytest.mark.parametrize("k,s", [(3, 1), (3, 2), (5, 1), (5, 2)])
@pytest.mark.parametrize("N,H,W,C", [(2, 14, 14, 144), (3, 17, 13, 24), (1, 9, 20, 72), (2, 33, 33, 64)])
def test_dwconv_forward_backward(ops, k, s, N, H, W, C):
    x_raw = rnd(N, H, W, C, seed=21, dtype=bf16) + 0.2
    g1, b1 = rnd(C, seed=22) * 0.1 + 1, rnd(C, seed=23) * 0.1
    w = rnd(C, 1, k, k, seed=24, scale=1.0 / k)
    rec1, _, _, _ = make_rec(ops, x_raw, g1, b1)
    OH, OW = ops.same_out(H, s), ops.same_out(W, s)
    # reference
    xt = x_raw.float().requires_grad_(True)
    wt = w.clone().requires_grad_(True)
    xa = F.silu(xt * rec1[0] + rec1[1])                       # same folded scale/shift as the kernel
    y_ref = F.conv2d(same_pad_t(xa.permute(0, 3, 1, 2), k, s), wt, stride=s, groups=C).permute(0, 2, 3, 1)
    # forward (train flavour: raw output + statistics)
    stats = ops.new_stats(C, "cuda")
    y = torch.empty(N, OH, OW, C, device="cuda", dtype=bf16)
    ops.dwconv_fwd(x_raw, rec1, w, y, N, H, W, k, s, stats=stats)
    assert y_ref.shape == y.shape
    assert rel_err(y, y_ref) < 1e-2
    yd = y.double().view(-1, C)
    tot = ops.stats_total(stats)
    assert torch.allclose(tot[0], yd.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(tot[1], (yd * yd).sum(0), rtol=1e-4, atol=1e-2)
    # forward (eval flavour: folded BN + SiLU + SE pooling in the epilogue), no input transform
    g2, b2 = rnd(C, seed=25) * 0.1 + 1, rnd(C, seed=26) * 0.1
    rec2, _, _, _ = make_rec(ops, y, g2, b2)
    pooled = torch.empty(N, C, device="cuda")
    y_eval = torch.empty_like(y)
    ops.dwconv_fwd(x_raw, None, w, y_eval, N, H, W, k, s, out_rec=rec2, pooled=pooled)
    ref_eval = F.silu(F.conv2d(same_pad_t(x_raw.float().permute(0, 3, 1, 2), k, s), w, stride=s, groups=C).permute(0, 2, 3, 1)
                      * rec2[0] + rec2[1])
    assert rel_err(y_eval, ref_eval) < 1e-2
    assert rel_err(pooled, y_eval.float().sum((1, 2))) < 2e-3
    # backward: upstream dD = a*gy + b*y + c
    gy = rnd(N, OH, OW, C, seed=27, dtype=bf16)
    coef = torch.stack([rnd(C, seed=28) * 0.2 + 1, rnd(C, seed=29) * 0.05, rnd(C, seed=30) * 0.05])
    dD = coef[0] * gy.float() + coef[1] * y.float() + coef[2]
    dD = dD.to(bf16).float()                                  # the kernel rounds the staged tile to bf16
    y_ref.backward(dD)
    g_out = torch.empty_like(x_raw)
    bst = ops.new_stats(C, "cuda")
    dw = torch.zeros_like(w)
    dD_dev = ops.affine2(gy, y, coef, torch.empty_like(gy))   # the BN-backward affine runs as its own streaming pass
    ops.dwconv_bwd(dD_dev, w, x_raw, rec1, g_out, bst, dw, N, H, W, k, s)
    # autograd gives d/dx_raw = dIn * silu'(.) * scale ; the kernel stops before the BN scale (that is affine2's job)
    want_g = xt.grad / rec1[0]
    assert rel_err(g_out, want_g) < 2e-2
    assert rel_err(dw, wt.grad) < 2e-2
    gd = g_out.double().view(-1, C)
    xh = (x_raw.double().view(-1, C) - rec1[2].double()) * rec1[3].double()
    bt = ops.stats_total(bst)
    assert torch.allclose(bt[0], gd.sum(0), rtol=1e-3, atol=1e-2) and torch.allclose(bt[1], (gd * xh).sum(0), rtol=1e-3, atol=1e-2)
    if s == 1:
        # the forward can also emit its activated input; the weight gradient taken from that tensor (no activation in the
        # kernel) equals the one that recomputes silu(bn(x)) from the raw input
        act = torch.full_like(x_raw, float("nan"))
        y_b = torch.empty_like(y)
        ops.dwconv_fwd(x_raw, rec1, w, y_b, N, H, W, k, s, stats=ops.new_stats(C, "cuda"), act_out=act)
        assert torch.equal(y_b, y)
        want_act = ops.bn_apply(x_raw.view(-1, C), rec1, torch.empty_like(x_raw.view(-1, C)), act=1).view_as(x_raw)
        assert torch.equal(act, want_act)
        dw_b = torch.zeros_like(w)
        ops.dwconv_bwd(dD_dev, w, act, None, None, None, dw_b, N, H, W, k, s)
        assert rel_err(dw_b, dw) < 1e-3
    # no-transform / no-coef flavour (DS block whose input is already an activation)
    xt2 = x_raw.float().requires_grad_(True)
    y2 = F.conv2d(same_pad_t(xt2.permute(0, 3, 1, 2), k, s), w, stride=s, groups=C).permute(0, 2, 3, 1)
    y2.backward(gy.float())
    g2_out = torch.empty_like(x_raw)
    dw2 = torch.zeros_like(w)
    ops.dwconv_bwd(gy, w, x_raw, None, g2_out, None, dw2, N, H, W, k, s)
    assert rel_err(g2_out, xt2.grad) < 2e-2


# ------------------------------------------------------------------------------------------------ stem
@pytest.mark.parametrize("CS,N,H,W,dt", [(48, 2, 32, 32, torch.float32), (32, 3, 33, 47, torch.float32), (48, 1, 64, 64, bf16)])
def test_stem_forward_wgrad(ops, CS, N, H, W, dt):
    x = rnd(N, 3, H, W, seed=31).to(dt)
    w = rnd(CS, 3, 3, 3, seed=32, scale=0.2)
    xt = x.to(bf16).float()
    wt = w.clone().requires_grad_(True)
    ref = F.conv2d(same_pad_t(xt, 3, 2), wt, stride=2).permute(0, 2, 3, 1)
    OH, OW = ref.shape[1:3]
    out = torch.empty(N, OH, OW, CS, device="cuda", dtype=bf16)
    stats = ops.new_stats(CS, "cuda")
    ops.stem_fwd(x, w, out, stats=stats)
    assert rel_err(out, ref) < 1e-2
    od = out.double().view(-1, CS)
    tot = ops.stats_total(stats)
    assert torch.allclose(tot[0], od.sum(0), rtol=1e-4, atol=1e-2) and torch.allclose(tot[1], (od * od).sum(0), rtol=1e-4, atol=1e-2)
    rec = torch.stack([rnd(CS, seed=33) * 0.1 + 1, rnd(CS, seed=34) * 0.1, torch.zeros(CS, device="cuda"), torch.ones(CS, device="cuda")])
    out2 = torch.empty_like(out)
    ops.stem_fwd(x, w, out2, out_rec=rec)
    assert rel_err(out2, F.silu(ref * rec[0] + rec[1])) < 1e-2
    ds = rnd(N, OH, OW, CS, seed=35, dtype=bf16)
    ref.backward(ds.float())
    dw = torch.zeros_like(w)
    ops.stem_wgrad(x, ds, dw)
    assert rel_err(dw, wt.grad) < 1e-2
    # train path: explicit im2col (27 taps padded to 32) + tcgen05 GEMMs for the forward and the weight gradient
    patches = ops.stem_im2col(x, torch.empty(N * OH * OW, 32, device="cuda", dtype=bf16))
    cols = F.unfold(same_pad_t(xt, 3, 2), 3, stride=2).transpose(1, 2).reshape(-1, 27)      # column = ci*9 + kh*3 + kw
    assert torch.equal(patches[:, :27].float(), cols) and float(patches[:, 27:].abs().max()) == 0.0
    wp = ops.stem_pack_w(w, torch.empty(CS, 32, device="cuda", dtype=bf16))
    stats2 = ops.new_stats(CS, "cuda")
    out3 = ops.gemm(patches, wp, ops.EPI_STATS, stats=stats2).view(N, OH, OW, CS)
    assert rel_err(out3, ref) < 1e-2
    dw2 = torch.zeros_like(w)
    ops.gemm_wgrad(ds.view(-1, CS), patches, dw2.view(CS, 27), so_p=27, so_q=1, q_store=27)
    assert rel_err(dw2, wt.grad) < 1e-2


# ------------------------------------------------------------------------------------------------ MIL pooling
@pytest.mark.parametrize("tc", [True, False], ids=["tcgen05", "cuda-cores"])
@pytest.mark.parametrize("B,K,D,hid", [(6, 16, 1280, 128), (1, 5, 1280, 256), (3, 12, 64, 32), (2, 31, 1280, 128), (160, 7, 64, 40),
                                       (70, 16, 1280, 128), (9, 16, 1280, 256)])
def test_mil_attention_forward_backward(ops, B, K, D, hid, tc):
    """tc: the score projection as a tcgen05 GEMM over split-bf16 operands (fp32-level accuracy) with the gate math in the
    epilogue; otherwise the round-1 CUDA-core kernel.  Same tolerances for both."""
    H = rnd(B, K, D, seed=41)
    Vw, Vb = rnd(hid, D, seed=42, scale=D ** -0.5), rnd(hid, seed=43, scale=0.1)
    Uw, Ub = rnd(hid, D, seed=44, scale=D ** -0.5), rnd(hid, seed=45, scale=0.1)
    ww, wb = rnd(hid, seed=46, scale=hid ** -0.5), rnd(1, seed=47, scale=0.1)
    ps = [t.clone().requires_grad_(True) for t in (H, Vw, Vb, Uw, Ub, ww, wb)]
    Ht, Vw_, Vb_, Uw_, Ub_, ww_, wb_ = ps
    g = torch.tanh(Ht @ Vw_.t() + Vb_) * torch.sigmoid(Ht @ Uw_.t() + Ub_)
    a = torch.softmax(g @ ww_ + wb_, dim=1)
    M_ref = torch.einsum('bkd,bk->bd', Ht, a)
    M, A, gV, gU = ops.mil_attn_fwd(H, Vw, Vb, Uw, Ub, ww, wb, save=True, tensor_core=tc)
    assert torch.allclose(M, M_ref, atol=1e-4, rtol=1e-4) and torch.allclose(A, a, atol=1e-5, rtol=1e-4)
    assert torch.allclose(gV, torch.tanh(Ht @ Vw_.t() + Vb_).detach(), atol=1e-4) and torch.allclose(gU, torch.sigmoid(Ht @ Uw_.t() + Ub_).detach(), atol=1e-4)
    M2, A2, _, _ = ops.mil_attn_fwd(H, Vw, Vb, Uw, Ub, ww, wb, save=False, tensor_core=tc)          # inference flavour: no gate tensors
    assert torch.allclose(M2, M, atol=1e-6) and torch.allclose(A2, A, atol=1e-6)
    dM = rnd(B, D, seed=48)
    M_ref.backward(dM)
    grads = [torch.zeros_like(t) for t in (Vw, Vb, Uw, Ub, ww, wb)]
    dH = ops.mil_attn_bwd(dM, H, A, gV, gU, Vw, Uw, ww, *grads)
    assert rel_err(dH, Ht.grad) < 1e-4
    for got, want in zip(grads, (Vw_.grad, Vb_.grad, Uw_.grad, Ub_.grad, ww_.grad, wb_.grad)):
        # the last one (scalar bias in front of the softmax) is mathematically zero: what is left is the rounding noise of
        # B * K terms, so its floor grows with the batch
        assert close(got, want, 1e-3, atol=1e-5 * max(1.0, B * K / 96))


# ------------------------------------------------------------------------------------------------ tab MLP + heads + loss
class TabHeadsRef(torch.nn.Module):
    """MMJointDualHead minus the backbone (train_mm_joint_dualtask.py:140-159), dropout 0."""

    def __init__(self, F_, T=9, Hd=64):
        super().__init__()
        nn = torch.nn
        self.tab = nn.Sequential(nn.Linear(T, Hd), nn.BatchNorm1d(Hd), nn.ReLU(), nn.Dropout(0.0), nn.Linear(Hd, Hd), nn.ReLU())
        self.cls_head, self.reg_head = nn.Linear(F_ + Hd, 1), nn.Linear(F_ + Hd, 1)

    def forward(self, feat, xt):
        f = torch.cat([feat, self.tab(xt)], 1)
        return self.cls_head(f).squeeze(1), self.reg_head(f).squeeze(1)


@pytest.mark.parametrize("B,Fdim,train", [(8, 1792, True), (64, 1792, True), (5, 1280, False), (2, 96, True)])
def test_tab_heads_forward_backward(ops, B, Fdim, train):
    torch.manual_seed(0)
    ref = TabHeadsRef(Fdim).cuda()
    ref.tab[1].running_mean.normal_(0, 0.1)
    ref.tab[1].running_var.uniform_(0.5, 1.5)
    ref.train(train)
    names = ["tab.0.weight", "tab.0.bias", "tab.1.weight", "tab.1.bias", "tab.4.weight", "tab.4.bias", "cls_head.weight",
             "cls_head.bias", "reg_head.weight", "reg_head.bias"]
    sd = dict(ref.named_parameters())
    params = [sd[n].detach().clone().contiguous() for n in names]
    rm, rv = ref.tab[1].running_mean.clone(), ref.tab[1].running_var.clone()
    nbt = torch.zeros((), device="cuda", dtype=torch.int64)
    feat, xt = rnd(B, Fdim, seed=51), rnd(B, 9, seed=52)
    yh = (rnd(B, seed=53) > 0).float()
    ys = torch.rand(B, device="cuda")
    sw = torch.rand(B, device="cuda") + 0.5
    ft = feat.clone().requires_grad_(True)
    lg, rg = ref(ft, xt)
    loss_ref = 1.0 * F.binary_cross_entropy_with_logits(lg, yh, weight=sw) + 0.3 * F.binary_cross_entropy_with_logits(rg, ys, weight=sw)
    loss_ref.backward()
    scratch = ops.tab_heads_scratch(B, 64, "cuda")
    out = ops.tab_heads_fwd(feat, xt, params, rm, rv, nbt, scratch, train, 0.0, targets=(yh, ys, sw))
    assert torch.allclose(out["logit"], lg, atol=1e-4, rtol=1e-4) and torch.allclose(out["reg"], rg, atol=1e-4, rtol=1e-4)
    assert abs(float(out["loss"]) - float(loss_ref)) < 1e-5
    if train:
        assert torch.allclose(rm, ref.tab[1].running_mean, atol=1e-5) and torch.allclose(rv, ref.tab[1].running_var, atol=1e-5)
        assert int(nbt) == 1
    grads = [torch.empty_like(p) for p in params]
    dfeat = torch.empty_like(feat)
    ops.tab_heads_bwd(feat, xt, params, rm, rv, out["dlogit"], out["dreg"], dfeat, grads, scratch, train, 0.0)
    assert rel_err(dfeat, ft.grad) < 1e-4
    for n, g in zip(names, grads):
        assert close(g, sd[n].grad, 1e-3), n


def test_tab_heads_batch1_train_raises(ops):
    from teethrt import TeethRTError
    ref = TabHeadsRef(96).cuda()
    params = [p.detach().contiguous() for p in ref.parameters()]
    with pytest.raises(TeethRTError):
        ops.tab_heads_fwd(rnd(1, 96), rnd(1, 9), params, torch.zeros(64, device="cuda"), torch.ones(64, device="cuda"), None,
                          ops.tab_heads_scratch(1, 64, "cuda"), True)


def test_dropout_is_deterministic_and_scaled(ops):
    ref = TabHeadsRef(96).cuda().train()
    params = [p.detach().contiguous() for p in ref.parameters()]
    feat, xt = rnd(16, 96, seed=1), rnd(16, 9, seed=2)
    mk = lambda: ops.tab_heads_fwd(feat, xt, params, torch.zeros(64, device="cuda"), torch.ones(64, device="cuda"), None,
                                   ops.tab_heads_scratch(16, 64, "cuda"), True, 0.2, seed=7)["logit"].clone()
    a, b = mk(), mk()
    assert torch.equal(a, b)
    c = ops.tab_heads_fwd(feat, xt, params, torch.zeros(64, device="cuda"), torch.ones(64, device="cuda"), None,
                          ops.tab_heads_scratch(16, 64, "cuda"), True, 0.2, seed=8)["logit"]
    assert not torch.equal(a, c)


# ------------------------------------------------------------------------------------------------ optimiser
def test_adamw_clip_cosine_matches_torch(ops):
    n = 100_003
    p0 = rnd(n, seed=61)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=3e-4, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    st = ops.OptimState("cuda", 3e-4, t_max=10)
    nsq = torch.zeros(1, device="cuda", dtype=torch.float64)
    nout = torch.zeros(1, device="cuda")
    for step in range(5):
        g = rnd(n, seed=70 + step, scale=0.01 * (step + 1))
        p_ref.grad = g.clone()
        gn = torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        sched.step()
        st.advance()
        ops.grad_sumsq(g, nsq)
        ops.adamw_step(p, g, m, v, st, nsq, nout, 1.0, 1.0, 1e-8, 1e-4)
        assert abs(float(nout) - float(gn)) < 1e-3 * float(gn)
        assert torch.allclose(p, p_ref.detach(), atol=2e-6, rtol=1e-5), step
    s = st.read()
    assert s["step"] == 5 and abs(s["lr"] - 3e-4 * (1 + math.cos(math.pi * 4 / 10)) / 2) < 1e-9
