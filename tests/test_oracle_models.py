"""CPU: pins the model oracle (oracle/timm shim + oracle/ref_models.py) to
  (1) an independent EfficientNet implementation (transformers.EfficientNetModel, TF semantics),
  (2) the committed golden fixtures minted from the reference's own classes (tests/golden/make_golden.py),
  (3) the reference classes themselves when /root/reference is present."""
import os

import pytest
import torch

import ref_models as R
import timm  # the oracle shim
from conftest import load_reference_module

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "models_golden.pt"), weights_only=False)


def mm_inputs(B, img, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, img, img, generator=g)
    xt = torch.randn(B, 9, generator=g)
    yh = (torch.rand(B, generator=g) < 0.6).float()
    ys = (yh * 0.8 + 0.2 * torch.rand(B, generator=g)).clamp(0, 1)
    return x, xt, yh, ys


def test_param_counts_and_key_inventory():
    b4 = timm.create_model("tf_efficientnet_b4_ns", num_classes=0, global_pool="avg")
    b0 = timm.create_model("tf_efficientnet_b0_ns", num_classes=0, global_pool="avg")
    assert sum(p.numel() for p in b4.parameters()) == 17_548_616 and len(b4.state_dict()) == 704
    assert sum(p.numel() for p in b0.parameters()) == 4_007_548 and len(b0.state_dict()) == 358
    assert b4.num_features == 1792 and b0.num_features == 1280
    sd = b4.state_dict()
    assert sd["conv_stem.weight"].shape == (48, 3, 3, 3)          # reference keys on this: predict_vision.py:9-14
    assert sd["blocks.0.0.se.conv_reduce.weight"].shape == (12, 48, 1, 1)
    assert sd["blocks.1.0.conv_pw.weight"].shape == (144, 24, 1, 1)
    assert sd["blocks.1.0.se.conv_reduce.weight"].shape == (6, 144, 1, 1)
    assert sd["blocks.6.1.conv_pwl.weight"].shape == (448, 2688, 1, 1)
    mm = R.MMJointDualHead()
    assert sum(p.numel() for p in mm.parameters()) == 17_557_258


@pytest.mark.parametrize("name,wc,dc,res", [("tf_efficientnet_b0_ns", 1.0, 1.0, 96), ("tf_efficientnet_b4_ns", 1.4, 1.8, 128)])
def test_shim_matches_hf_efficientnet(name, wc, dc, res):
    tr = pytest.importorskip("transformers")
    torch.manual_seed(0)
    mine = timm.create_model(name, num_classes=0, global_pool="avg")
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for m in mine.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_var.copy_(1 + 0.2 * torch.rand(m.bias.shape, generator=g))
            if isinstance(m, torch.nn.Conv2d) and m.bias is not None:
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    cfg = tr.EfficientNetConfig(width_coefficient=wc, depth_coefficient=dc, image_size=res, dropout_rate=0.0,
                                drop_connect_rate=0.0, batch_norm_eps=1e-3, hidden_dim=mine.num_features)
    hf = tr.EfficientNetModel(cfg).eval()
    a = [(k, v) for k, v in mine.state_dict().items()]
    b = [(k, v) for k, v in hf.state_dict().items()]
    assert len(a) == len(b)
    new = {}
    for (ka, va), (kb, vb) in zip(a, b):
        assert va.shape == vb.shape, (ka, kb)
        new[kb] = va
    hf.load_state_dict(new, strict=True)
    x = torch.randn(2, 3, res, res, generator=g)
    with torch.no_grad():
        y = mine.eval()(x)
        yh = hf(pixel_values=x).pooler_output
    assert (y - yh).abs().max().item() < 1e-5


def test_golden_mm_forward_config0():
    m = R.seeded_model("mm", seed=0, warm=2, img=64)
    x, xt, _, _ = mm_inputs(8, 224, 100)
    with torch.no_grad():
        logit, reg = m(x, xt)
    assert torch.allclose(logit, GOLD["mm_b4_fwd224"]["logit"], atol=1e-5)
    assert torch.allclose(reg, GOLD["mm_b4_fwd224"]["reg"], atol=1e-5)
    tta = R.mm_tta_logit(m, x[:2], xt[:2])
    assert torch.allclose(tta, GOLD["mm_b4_tta224"]["logit"], atol=1e-5)
    assert torch.allclose(torch.sigmoid(tta / 2.5), GOLD["mm_b4_tta224"]["prob_T2p5"], atol=1e-6)


def test_golden_mm_train_steps_b0():
    g = GOLD["mm_b0_train64"]
    m = R.seeded_model("mm", seed=1, warm=1, img=64, backbone="tf_efficientnet_b0_ns", drop=0.0).train()
    opt, sched = R.make_optimizer(m, t_max=10)
    for s in range(3):
        x, xt, yh, ys = mm_inputs(8, 64, 200 + s)
        loss, gn = R.mm_train_step(m, opt, sched, x, xt, yh, ys)
        assert abs(loss - g["losses"][s].item()) < 1e-5
        assert abs(gn - g["grad_norms"][s].item()) < 1e-3 * max(1.0, gn)
    m.eval()
    x, xt, _, _ = mm_inputs(8, 64, 299)
    with torch.no_grad():
        lg, _ = m(x, xt)
    assert torch.allclose(lg, g["logit_after"], atol=1e-4)
    assert torch.allclose(m.backbone.bn1.running_mean, g["bn1_running_mean"], atol=1e-6)


def test_golden_mil():
    m = R.seeded_model("mil", seed=2, warm=1, img=64)
    g = torch.Generator().manual_seed(300)
    bags = torch.randn(2, 16, 3, 96, 96, generator=g)
    H = torch.randn(6, 16, 1280, generator=g)
    with torch.no_grad():
        lg, A = m(bags)
        M, A2 = m.mil(H)
    assert torch.allclose(lg, GOLD["mil_b0_fwd96"]["logit"], atol=1e-5)
    assert torch.allclose(A, GOLD["mil_b0_fwd96"]["A"], atol=1e-6)
    assert torch.allclose(M, GOLD["mil_pool"]["M"], atol=1e-5)
    assert torch.allclose(A2, GOLD["mil_pool"]["A"], atol=1e-6)
    tw = R.seeded_model("mil_twin", seed=3, warm=1, img=64)
    with torch.no_grad():
        assert torch.allclose(tw(bags[0]), GOLD["mil_twin_fwd96"]["logit"], atol=1e-5)


def test_mil_key_remap_and_prep_tab():
    m = R.MILNet()
    sd = R.remap_mil_keys(m.state_dict())
    assert "enc.conv_stem.weight" in sd and "mil.V.weight" in sd and "mil.w.bias" in sd
    z = R.prep_tab(None, [1.0] * 9, [0.0] * 9)
    assert z.shape == (1, 9) and float(z.abs().max()) == 0.0
    d = {k: 2.0 for k in R.TAB_FEATURES}
    z = R.prep_tab(d, [1.0] * 9, [0.5] * 9)
    assert torch.allclose(z, torch.full((1, 9), 2.0))


@pytest.mark.reference
def test_restatement_equals_imported_reference_classes():
    ref_mm = load_reference_module("experiments/multimodal_v1/train_mm_joint_dualtask.py", "ref_mm_t")
    ref_mil = load_reference_module("experiments/vision_v2/train_mil_attention_v1.py", "ref_mil_t")
    ref_imil = load_reference_module("ui/gradio_app/infer_mil.py", "ref_imil_t")
    mine = R.seeded_model("mm", seed=0, warm=1, img=64, backbone="tf_efficientnet_b0_ns")
    theirs = ref_mm.MMJointDualHead(backbone="tf_efficientnet_b0_ns").eval()
    theirs.load_state_dict(mine.state_dict(), strict=True)
    x, xt, yh, ys = mm_inputs(3, 64, 11)
    with torch.no_grad():
        a, b = mine(x, xt), theirs(x, xt)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    l1 = R.dual_bce_loss(a[0], a[1], yh, ys)
    l2 = 1.0 * ref_mm.bce_logits_with_soft_targets(b[0], yh) + 0.3 * ref_mm.bce_logits_with_soft_targets(b[1], ys)
    assert torch.equal(l1, l2)
    mil = R.seeded_model("mil", seed=0, warm=1, img=64)
    tm = ref_mil.MILNet().eval()
    tm.load_state_dict(mil.state_dict(), strict=True)
    bags = torch.randn(2, 4, 3, 64, 64)
    with torch.no_grad():
        a, b = mil(bags), tm(bags)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    tw = R.seeded_model("mil_twin", seed=0, warm=1, img=64)
    tt = ref_imil.MILNet().eval()
    tt.load_state_dict(tw.state_dict(), strict=True)
    with torch.no_grad():
        assert torch.equal(tw(bags[0]), tt(bags[0]))
    assert R.remap_mil_keys(mil.state_dict()).keys() == ref_imil._remap_state_dict_keys(mil.state_dict()).keys()
