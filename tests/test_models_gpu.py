"""GPU parity, model level: the teethrt modules (C-ABI kernels underneath) against the oracle (oracle/ref_models.py, the
restatement pinned to the reference's own classes) and the golden fixtures minted from the reference.

Tolerances (SURVEY.md §8c, bf16 activations / fp32 accumulate vs the fp32 oracle):
  eval logits |d| <= 5e-2, probabilities |d| <= 1e-2; train step: loss |d| <= 3e-2, per-tensor gradient cosine >= 0.99
  for tensors carrying real signal, grad-norm within 5 %.
"""
import os

import pytest
import torch

import ref_models as R   # oracle (checker only)

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "models_golden.pt"), weights_only=False)


@pytest.fixture(scope="module")
def T():
    import teethrt
    teethrt.init()
    from teethrt import modules, train, backbone  # noqa: F401
    return teethrt


def mm_inputs(B, img, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, img, img, generator=g)
    xt = torch.randn(B, 9, generator=g)
    yh = (torch.rand(B, generator=g) < 0.6).float()
    ys = (yh * 0.8 + 0.2 * torch.rand(B, generator=g)).clamp(0, 1)
    return x, xt, yh, ys


def test_state_dict_contract(T):
    from teethrt.modules import MMJointDualHead, MMNet, MILNet, MILNetTwin
    ora = R.MMJointDualHead()
    mine = MMJointDualHead()
    a, b = ora.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    assert mine.load_state_dict(a, strict=True)
    assert set(MMNet().state_dict().keys()) == set(a.keys())
    assert list(R.MILNet().state_dict().keys()) == list(MILNet().state_dict().keys())
    assert set(R.MILNetTwin().state_dict().keys()) == set(MILNetTwin().state_dict().keys())
    assert sum(p.numel() for p in mine.parameters()) == 17_557_258


def test_mm_eval_forward_config0_vs_golden_and_oracle(T):
    """BASELINE.json configs[0]: MM dual-task fwd, B4, batch 8 @224."""
    from teethrt.modules import MMJointDualHead
    ora = R.seeded_model("mm", seed=0, warm=2, img=64)
    m = MMJointDualHead().cuda()
    m.load_state_dict(ora.state_dict(), strict=True)
    m.eval()
    x, xt, _, _ = mm_inputs(8, 224, 100)
    with torch.no_grad():
        logit, reg = m(x.cuda(), xt.cuda())
        lo, ro = ora(x, xt)
    g = GOLD["mm_b4_fwd224"]
    assert torch.allclose(lo, g["logit"], atol=1e-5)                       # oracle still equals the reference's output
    assert (logit.cpu() - g["logit"]).abs().max() < 5e-2 and (reg.cpu() - g["reg"]).abs().max() < 5e-2
    assert (torch.sigmoid(logit.cpu()) - torch.sigmoid(g["logit"])).abs().max() < 1e-2
    # state dict round trip is lossless (fp32 masters untouched by the bf16 compute path)
    sd = m.state_dict()
    assert all(torch.equal(sd[k].cpu(), v) for k, v in ora.state_dict().items())


def test_mm_batch1_and_tta(T):
    from teethrt.modules import MMNet
    from teethrt.infer import tta_logit
    ora = R.seeded_model("mm", seed=0, warm=2, img=64)
    m = MMNet().cuda()
    m.load_state_dict(ora.state_dict(), strict=True)
    m.eval()
    x, xt, _, _ = mm_inputs(8, 224, 100)
    with torch.no_grad():
        l1, _ = m(x[:1].cuda(), xt[:1].cuda())
        tta = tta_logit(m, x[:2].cuda(), xt[:2].cuda())
    assert abs(float(l1[0]) - float(GOLD["mm_b4_fwd224"]["logit"][0])) < 5e-2
    g = GOLD["mm_b4_tta224"]
    assert (tta.cpu() - g["logit"]).abs().max() < 5e-2
    assert (torch.sigmoid(tta.cpu() / 2.5) - g["prob_T2p5"]).abs().max() < 1e-2


def grad_report(mine_named, ora_named):
    """cosine / relative-norm per tensor between two gradient dicts"""
    rep = {}
    for n, go in ora_named.items():
        gm = mine_named[n].detach().float().cpu().flatten()
        go = go.flatten()
        no, nm = float(go.norm()), float(gm.norm())
        cos = float(torch.dot(gm, go) / (nm * no + 1e-30))
        rep[n] = (cos, nm, no)
    return rep


@pytest.mark.parametrize("backbone,B,img", [("tf_efficientnet_b0_ns", 16, 128), ("tf_efficientnet_b4_ns", 8, 96)])
def test_mm_train_forward_backward_autograd_path(T, backbone, B, img):
    """The reference's own loop shape: logits via module.forward in train mode, loss by torch, loss.backward().
    Gradient bar: the reference trains under AMP (train_mm_joint_dualtask.py:242), so the yardstick is how close torch's own
    bf16-autocast run of the ORACLE gets to the fp32 oracle on the same batch — teethrt must be at least as close
    (batch-statistic BatchNorm over few samples amplifies 8-bit-mantissa rounding; the absolute cosine depends on the
    config, the comparison does not)."""
    import copy
    from teethrt.modules import MMJointDualHead
    ora = R.seeded_model("mm", seed=1, warm=1, img=64, backbone=backbone, drop=0.0).train()
    sd = copy.deepcopy(ora.state_dict())
    x, xt, yh, ys = [t.cuda() for t in mm_inputs(B, img, 200)]
    ref = copy.deepcopy(ora).cuda()
    lo, ro = ref(x, xt)
    R.dual_bce_loss(lo, ro, yh, ys).backward()
    amp = copy.deepcopy(ora).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la, ra = amp(x, xt)
    R.dual_bce_loss(la.float(), ra.float(), yh, ys).backward()
    m = MMJointDualHead(backbone=backbone, drop=0.0).cuda()
    m.load_state_dict(sd, strict=True)
    m.train()
    lm, rm = m(x, xt)
    R.dual_bce_loss(lm, rm, yh, ys).backward()
    err_mine = max(float((lm - lo).abs().max()), float((rm - ro).abs().max()))
    err_amp = max(float((la.float() - lo).abs().max()), float((ra.float() - ro).abs().max()))
    assert err_mine < max(5e-2, 2.0 * err_amp), (err_mine, err_amp)
    # BN running statistics updated like torch's
    assert torch.allclose(m.backbone.bn1.running_mean, ref.backbone.bn1.running_mean, atol=2e-3)
    assert torch.allclose(m.tab[1].running_var, ref.tab[1].running_var, atol=1e-4)
    assert int(m.backbone.bn1.num_batches_tracked) == int(ref.backbone.bn1.num_batches_tracked)
    names = [n for n, _ in ref.named_parameters()]
    gr, ga, gm = (dict((n, p.grad.detach().float()) for n, p in mod.named_parameters()) for mod in (ref, amp, m))
    flat = lambda g: torch.cat([g[n].flatten() for n in names])
    cos = lambda a, b: float(torch.dot(a.flatten(), b.flatten()) / (a.norm() * b.norm() + 1e-30))
    fr, fa, fm = flat(gr), flat(ga), flat(gm)
    assert abs(float(fm.norm()) - float(fr.norm())) / float(fr.norm()) < 0.05
    assert cos(fm, fr) > cos(fa, fr) - 0.02, (cos(fm, fr), cos(fa, fr))
    big = [n for n in names if float(gr[n].norm()) > 1e-3 * float(fr.norm())]
    cm = sorted(cos(gm[n], gr[n]) for n in big)
    ca = sorted(cos(ga[n], gr[n]) for n in big)
    assert len(big) > 50
    assert cm[len(cm) // 2] > ca[len(ca) // 2] - 0.02, (cm[len(cm) // 2], ca[len(ca) // 2])
    assert cm[len(cm) // 10] > ca[len(ca) // 10] - 0.05, (cm[len(cm) // 10], ca[len(ca) // 10])   # 10th percentile
    if backbone.endswith("b0_ns"):
        assert cos(fm, fr) > 0.99


def test_fused_trainer_matches_golden_train_steps(T):
    """3 steps of the reference loop (train_mm_joint_dualtask.py:241-256) minted into tests/golden from the reference's
    classes, replayed by DualTaskTrainer (eager steps, then a captured CUDA-graph step)."""
    from teethrt.modules import MMJointDualHead
    from teethrt.train import DualTaskTrainer
    g = GOLD["mm_b0_train64"]
    ora = R.seeded_model("mm", seed=1, warm=1, img=64, backbone="tf_efficientnet_b0_ns", drop=0.0)
    m = MMJointDualHead(backbone="tf_efficientnet_b0_ns", drop=0.0).cuda()
    m.load_state_dict(ora.state_dict(), strict=True)
    tr = DualTaskTrainer(m, lr=3e-4, weight_decay=1e-4, t_max=10, alpha=1.0, beta=0.3, grad_clip=1.0, graph=True)
    tr.graph_warmup = 2
    for s in range(3):
        x, xt, yh, ys = mm_inputs(8, 64, 200 + s)
        loss = tr.step(x.cuda(), xt.cuda(), yh.cuda(), ys.cuda())
        torch.cuda.synchronize()
        assert abs(float(loss) - float(g["losses"][s])) < 3e-2, (s, float(loss), float(g["losses"][s]))
        assert abs(float(tr.grad_norm) - float(g["grad_norms"][s])) < 0.05 * float(g["grad_norms"][s]) + 1e-3
    assert tr._graphs is not None                                          # third step ran as a graph replay
    assert abs(tr.lr() - 3e-4 * (1 + __import__("math").cos(__import__("math").pi * 2 / 10)) / 2) < 1e-9
    m.eval()
    x, xt, _, _ = mm_inputs(8, 64, 299)
    with torch.no_grad():
        lg, rg = m(x.cuda(), xt.cuda())
    assert (lg.cpu() - g["logit_after"]).abs().max() < 5e-2 and (rg.cpu() - g["reg_after"]).abs().max() < 5e-2
    assert torch.allclose(m.backbone.bn1.running_mean.cpu(), g["bn1_running_mean"], atol=5e-3)
    # checkpoint layout of the reference (train_mm_joint_dualtask.py:302-313) round-trips through torch.save/load
    import io
    buf = io.BytesIO()
    torch.save({"model": m.state_dict(), "scaler_mean": None, "scaler_scale": None, "thr": 0.5, "T": 1.0,
                "args": {"backbone": "tf_efficientnet_b0_ns", "img_size": 64, "tab_hidden": 64, "dropout": 0.0}, "epoch": 1}, buf)
    buf.seek(0)
    ck = torch.load(buf, map_location="cpu", weights_only=False)
    assert R.MMJointDualHead(backbone="tf_efficientnet_b0_ns").load_state_dict(ck["model"], strict=True)


def test_mil_forward_and_twin_vs_golden(T):
    from teethrt.modules import MILNet, MILNetTwin
    ora = R.seeded_model("mil", seed=2, warm=1, img=64)
    m = MILNet(drop=0.0).cuda()
    m.load_state_dict(ora.state_dict(), strict=True)
    m.eval()
    gen = torch.Generator().manual_seed(300)
    bags = torch.randn(2, 16, 3, 96, 96, generator=gen)
    H = torch.randn(6, 16, 1280, generator=gen)
    with torch.no_grad():
        lg, A = m(bags.cuda())
        M, A2 = m.mil(H.cuda())
    assert (lg.cpu() - GOLD["mil_b0_fwd96"]["logit"]).abs().max() < 5e-2
    assert (A.cpu() - GOLD["mil_b0_fwd96"]["A"]).abs().max() < 1e-2
    assert torch.allclose(M.cpu(), GOLD["mil_pool"]["M"], atol=1e-4) and torch.allclose(A2.cpu(), GOLD["mil_pool"]["A"], atol=1e-5)
    tw_o = R.seeded_model("mil_twin", seed=3, warm=1, img=64)
    tw = MILNetTwin().cuda()
    tw.load_state_dict(tw_o.state_dict(), strict=True)
    tw.eval()
    with torch.no_grad():
        out = tw(bags[0].cuda())
    assert abs(float(out) - float(GOLD["mil_twin_fwd96"]["logit"])) < 5e-2


def test_mil_trainer_step_vs_oracle(T):
    from teethrt.modules import MILNet
    from teethrt.train import MILTrainer
    ora = R.seeded_model("mil", seed=2, warm=1, img=64, drop=0.0).train()
    m = MILNet(drop=0.0).cuda()
    m.load_state_dict(ora.state_dict(), strict=True)
    opt, sched = torch.optim.AdamW(ora.parameters(), lr=2e-4, weight_decay=1e-4), None
    tr = MILTrainer(m, lr=2e-4, weight_decay=1e-4, t_max=0, graph=True)
    gen = torch.Generator().manual_seed(5)
    for s in range(3):
        bags = torch.randn(2, 4, 3, 64, 64, generator=gen)
        y = (torch.rand(2, generator=gen) < 0.5).float()
        lo, gn = R.mil_train_step(ora, opt, sched, bags, y)
        loss = tr.step(bags.cuda(), y.cuda())
        torch.cuda.synchronize()
        assert abs(float(loss) - lo) < 3e-2
        assert abs(float(tr.grad_norm) - gn) < 0.08 * gn + 1e-3


def test_ensembles_mirror_reference_behaviour(T, tmp_path):
    import numpy as np
    from PIL import Image
    from teethrt.infer import MMEnsemble, MILEnsemble
    # --- MM: two folds with different temperatures
    ora = [R.seeded_model("mm", seed=s, warm=1, img=64, backbone="tf_efficientnet_b0_ns") for s in (0, 1)]
    Ts, mean, scale = [2.5, 3.1], np.arange(9, dtype=np.float64), np.array([1, 2, 0, 1, 1, 2, 1, 1, 3], dtype=np.float64)
    for f, (o, Tf) in enumerate(zip(ora, Ts)):
        torch.save({"model": o.state_dict(), "scaler_mean": mean, "scaler_scale": scale, "thr": 0.5, "T": Tf,
                    "args": {"backbone": "tf_efficientnet_b0_ns", "img_size": 64, "tab_hidden": 64, "dropout": 0.2}, "epoch": 3},
                   tmp_path / f"mm_dualtask_fold{f}.pt")
    rng = np.random.default_rng(0)
    img_path = tmp_path / "tooth.png"
    Image.fromarray(rng.integers(0, 256, size=(90, 120, 3), dtype=np.uint8)).save(img_path)
    ens = MMEnsemble(tmp_path, device="cuda")
    assert ens.num_folds == 2
    tab = {k: float(i) for i, k in enumerate(R.TAB_FEATURES)}
    import timm  # the oracle's shim provides the reference's eval transform
    tf = timm.data.create_transform(input_size=64, is_training=False, interpolation="bicubic")
    xi = tf(Image.open(img_path).convert("RGB")).unsqueeze(0)
    for td in (tab, None):
        want, _ = R.mm_ensemble_prob(list(zip(ora, Ts)), [(mean, scale)] * 2, xi, td)
        got, dbg = ens.predict(img_path, td)
        assert abs(got - want) < 1e-2 and "fold_probs" in dbg
    assert MMEnsemble(tmp_path / "nothing", device="cuda").predict(img_path, None) == (0.5, "MM not loaded")
    # --- MIL: trainer-format checkpoints ('model' key, hid 128) load and run (the reference twin silently cannot, q9)
    mo = R.seeded_model("mil", seed=2, warm=1, img=64)
    mil_dir = tmp_path / "mil"
    mil_dir.mkdir()
    torch.save({"model": mo.state_dict(), "args": {}, "thr": 0.5, "epoch": 1}, mil_dir / "mil_v1_fold0.pt")
    inst = tmp_path / "inst"
    inst.mkdir()
    for i in range(3):
        Image.fromarray(rng.integers(0, 256, size=(530, 540, 3), dtype=np.uint8)).save(inst / f"{i}.png")
    me = MILEnsemble(mil_dir, device="cuda")
    assert me.num_folds == 1
    prob, dbg = me.predict(inst)
    tw = R.MILNetTwin(hid_dim=128)
    tw.load_state_dict(R.remap_mil_keys(mo.state_dict()), strict=True)
    tw.eval()
    from torchvision import transforms
    tfm = transforms.Compose([transforms.Resize(512), transforms.CenterCrop(480), transforms.ToTensor()])
    xs = torch.stack([tfm(Image.open(p).convert("RGB")) for p in sorted(inst.iterdir())])
    with torch.no_grad():
        want = float(torch.sigmoid(tw(xs)))
    assert abs(prob - want) < 1e-2 and "Instances=3" in dbg
    assert me.predict(tmp_path / "empty_dir_that_does_not_exist")[0] is None


def test_separate_bn_finalise_and_prefetch_match_default_path(T, monkeypatch):
    """The default path finalises every BatchNorm in the prologue of its first consumer (lazy records).  The round-1 path
    (one finalise launch per BatchNorm, TEETHRT_LAZY_BN=0; three-pass SE/BatchNorm backward, TEETHRT_SE_BWD_MERGED=0) and
    Trainer.prefetch() must give the same training trajectory."""
    from teethrt.modules import MMJointDualHead
    from teethrt.train import DualTaskTrainer

    def run(fuse, prefetch):
        monkeypatch.setenv("TEETHRT_LAZY_BN", fuse)
        monkeypatch.setenv("TEETHRT_SE_BWD_MERGED", fuse)
        monkeypatch.setenv("TEETHRT_GEMM_BNBWD", "1" if (fuse == "1" and prefetch) else "0")      # opt-in variants ride along with one arm
        monkeypatch.setenv("TEETHRT_NARROW_SLABS", "1" if (fuse == "1" and prefetch) else "0")
        torch.manual_seed(0)
        m = MMJointDualHead("tf_efficientnet_b0_ns", 9, 64, 0.0).cuda()
        tr = DualTaskTrainer(m, t_max=10, graph=False)
        batches = [[t.pin_memory() for t in mm_inputs(8, 96, 300 + i)] for i in range(3)]
        losses = []
        for i, b in enumerate(batches):
            loss = tr.step(*b)
            if prefetch and i + 1 < len(batches):
                tr.prefetch(*batches[i + 1])
            losses.append(float(loss))
        return losses, torch.cat([p.detach().flatten() for p in m.parameters()]).double()

    l0, p0 = run("1", False)
    l0b, p0b = run("1", False)               # run-to-run noise of the default path (atomic summation order, bf16 rounding)
    l1, p1 = run("0", False)
    l2, p2 = run("1", True)
    dl = lambda a, b: max(abs(x - y) for x, y in zip(a, b))
    dp = lambda a, b: float((a - b).norm() / a.norm())
    print("loss diffs", dl(l0, l0b), dl(l0, l1), dl(l0, l2), "param diffs", dp(p0, p0b), dp(p0, p1), dp(p0, p2))
    # floors = the noise two runs of the SAME path show on this tiny batch (loss up to 3e-3, parameters ~9e-4 relative: bf16
    # rounding on top of atomic summation order); a single noise sample can come out 10x smaller than the next one
    tol_l, tol_p = max(6e-3, 4 * dl(l0, l0b)), max(3e-3, 4 * dp(p0, p0b))
    assert dl(l0, l1) < tol_l and dl(l0, l2) < tol_l
    assert dp(p0, p1) < tol_p and dp(p0, p2) < tol_p


def test_config1_full_size_step_vs_oracle_on_the_gpu(T):
    """BASELINE.json configs[1] at its full size (B4, batch 64, 224x224): one fused train step against the fp32 oracle run
    on the same GPU (TF32 off), plus two size-independent properties — an eval batch equals the concatenation of its halves,
    and the TTA logit equals the mean of three separately flipped forwards."""
    from teethrt.modules import MMJointDualHead
    from teethrt.train import DualTaskTrainer
    from teethrt.infer import tta_logit
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ora = R.seeded_model("mm", seed=3, warm=1, img=64, drop=0.0).cuda().train()
    m = MMJointDualHead(drop=0.0).cuda()
    m.load_state_dict(ora.state_dict(), strict=True)
    x, xt, yh, ys = (t.cuda() for t in mm_inputs(64, 224, 777))
    # --- oracle step (train_mm_joint_dualtask.py:241-256 without AMP), fp32 on the device
    opt, _ = R.make_optimizer(ora, lr=3e-4, weight_decay=1e-4)
    loss_o, gn_o = R.mm_train_step(ora, opt, None, x, xt, yh, ys, alpha=1.0, beta=0.3, grad_clip=1.0)
    # --- ours
    tr = DualTaskTrainer(m, lr=3e-4, weight_decay=1e-4, t_max=0, alpha=1.0, beta=0.3, grad_clip=1.0, graph=False)
    loss = float(tr.step(x, xt, yh, ys))
    assert abs(loss - float(loss_o)) < 3e-2, (loss, float(loss_o))
    assert abs(float(tr.grad_norm) - gn_o) < 0.05 * gn_o, (float(tr.grad_norm), gn_o)
    # whole-model gradient direction (AdamW's first update is lr*sign(g) per weight, so the weights themselves only say
    # how many noise-level gradients kept their sign; the flat gradient is the meaningful comparison)
    go = torch.cat([p.grad.flatten() for _, p in ora.named_parameters()])
    gm = torch.cat([tr.flat.grads[n].flatten() for n, _ in ora.named_parameters()])
    cos_mine = float(torch.nn.functional.cosine_similarity(go, gm, dim=0))
    # the bar for a bf16-activation path: at least as close to the fp32 gradient as torch's own bf16 autocast of the oracle
    # (the reference trains under AMP, train_mm_joint_dualtask.py:242)
    amp = R.seeded_model("mm", seed=3, warm=1, img=64, drop=0.0).cuda().train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la, ra = amp(x, xt)
        loss_a = R.dual_bce_loss(la.float(), ra.float(), yh, ys, 1.0, 0.3)
    loss_a.backward()
    ga = torch.cat([p.grad.flatten() for _, p in amp.named_parameters()])
    cos_amp = float(torch.nn.functional.cosine_similarity(go, ga, dim=0))
    assert cos_mine > 0.9 and cos_mine >= cos_amp - 0.02, (cos_mine, cos_amp)
    del amp, ga
    # --- eval-mode properties at full size
    m.eval(); ora.eval()
    with torch.no_grad():
        full, _ = m(x, xt)
        halves = torch.cat([m(x[:32], xt[:32])[0], m(x[32:], xt[32:])[0]])
        assert (full - halves).abs().max() < 2e-3
        lo, _ = ora(x, xt)
        # the two models took one optimiser step each from the same weights: their logits still agree at bf16 level
        assert (full - lo).abs().max() < 1e-1
        t3 = tta_logit(m, x[:8], xt[:8])
        sep = torch.stack([m(x[:8], xt[:8])[0], m(torch.flip(x[:8], dims=[3]), xt[:8])[0], m(torch.flip(x[:8], dims=[2]), xt[:8])[0]]).mean(0)
        assert (t3 - sep).abs().max() < 2e-3
