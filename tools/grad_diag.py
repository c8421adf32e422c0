"""Diagnostic: per-tensor gradient cosine of (a) teethrt, (b) torch bf16-autocast on GPU, both against the fp32 oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import copy, torch
import ref_models as R
import teethrt
from teethrt.modules import MMJointDualHead
teethrt.init()

def inputs(B, img, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, img, img, generator=g); xt = torch.randn(B, 9, generator=g)
    yh = (torch.rand(B, generator=g) < 0.6).float(); ys = (yh * 0.8 + 0.2 * torch.rand(B, generator=g)).clamp(0, 1)
    return x, xt, yh, ys

def cos(a, b):
    a, b = a.flatten().float().cpu(), b.flatten().float().cpu()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))

for backbone, B, img in [("tf_efficientnet_b4_ns", 4, 96), ("tf_efficientnet_b4_ns", 16, 128), ("tf_efficientnet_b0_ns", 16, 128)]:
    ora = R.seeded_model("mm", seed=1, warm=1, img=64, backbone=backbone, drop=0.0).train()
    sd = copy.deepcopy(ora.state_dict())
    x, xt, yh, ys = inputs(B, img, 200)
    ref = copy.deepcopy(ora).cuda()
    lo, ro = ref(x.cuda(), xt.cuda()); R.dual_bce_loss(lo, ro, yh.cuda(), ys.cuda()).backward()
    ac = copy.deepcopy(ora).cuda(); ac.load_state_dict(sd)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        la, ra = ac(x.cuda(), xt.cuda()); loss = R.dual_bce_loss(la.float(), ra.float(), yh.cuda(), ys.cuda())
    loss.backward()
    m = MMJointDualHead(backbone=backbone, drop=0.0).cuda(); m.load_state_dict(sd); m.train()
    lm, rm = m(x.cuda(), xt.cuda()); R.dual_bce_loss(lm, rm, yh.cuda(), ys.cuda()).backward()
    print(f"=== {backbone} B={B} img={img}: logit diff mine {float((lm-lo).abs().max()):.4f} autocast {float((la.float()-lo).abs().max()):.4f}")
    rows = []
    gr = dict(ref.named_parameters()); ga = dict(ac.named_parameters()); gm = dict(m.named_parameters())
    tot = sum(float(p.grad.norm())**2 for p in gr.values())**0.5
    for n in gr:
        rows.append((cos(gm[n].grad, gr[n].grad), cos(ga[n].grad, gr[n].grad), float(gm[n].grad.norm()), float(gr[n].grad.norm()), n))
    allm = torch.cat([gm[n].grad.flatten() for n in gr]); allr = torch.cat([gr[n].grad.flatten() for n in gr]); alla = torch.cat([ga[n].grad.flatten().float() for n in gr])
    print(f"global cos mine {cos(allm, allr):.4f}  autocast {cos(alla, allr):.4f}; norms mine {float(allm.norm()):.3f} ref {float(allr.norm()):.3f} autocast {float(alla.norm()):.3f}")
    rows = [r for r in rows if r[3] > 1e-3 * tot]
    rows.sort()
    for r in rows[:25]:
        print(f"  mine {r[0]:.4f}  autocast {r[1]:.4f}  |g| mine {r[2]:.4f} ref {r[3]:.4f}  {r[4]}")
    import statistics
    print("  median cos mine", statistics.median(r[0] for r in rows), "autocast", statistics.median(r[1] for r in rows))
