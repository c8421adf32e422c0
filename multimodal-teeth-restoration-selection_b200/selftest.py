"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle (the only product-side file that
may touch oracle/, as the checker — never as the thing shipped or measured)."""
import os
import sys

import numpy as np
import torch


def smoke():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import ref_models as R
    import ref_preproc as P
    import teethrt
    from teethrt import preproc
    from teethrt.modules import MMJointDualHead
    from teethrt.train import DualTaskTrainer
    torch.cuda.set_device(0)
    teethrt.init(0)
    # 1. input stage: byte-exact CLAHE + resize
    img = P.image_set("radiograph", 256, 256, seed=1)
    out = preproc.centre_crop_resize(preproc.apply_clahe(img), 64)
    ref = P.centre_crop_resize_cv2(P.apply_clahe_cv2(img), 64)
    assert int((out != ref).sum()) == 0, "CLAHE/resize not byte-exact"
    # 2. one eval forward + one fused train step of the dual-task model (B0 backbone, 64 px) vs the fp32 oracle
    ora = R.seeded_model("mm", seed=1, warm=1, img=64, backbone="tf_efficientnet_b0_ns", drop=0.0)
    m = MMJointDualHead(backbone="tf_efficientnet_b0_ns", drop=0.0).cuda()
    m.load_state_dict(ora.state_dict(), strict=True)
    g = torch.Generator().manual_seed(7)
    x, xt = torch.randn(4, 3, 64, 64, generator=g), torch.randn(4, 9, generator=g)
    yh = (torch.rand(4, generator=g) < 0.6).float()
    ys = (yh * 0.8 + 0.2 * torch.rand(4, generator=g)).clamp(0, 1)
    m.eval()
    with torch.no_grad():
        lg, rg = m(x.cuda(), xt.cuda())
        lo, ro = ora(x, xt)
    assert (lg.cpu() - lo).abs().max() < 5e-2 and (rg.cpu() - ro).abs().max() < 5e-2, "eval logits off"
    ora.train()
    opt, sched = R.make_optimizer(ora, t_max=10)
    loss_o, gn_o = R.mm_train_step(ora, opt, sched, x, xt, yh, ys)
    tr = DualTaskTrainer(m, t_max=10, graph=False)
    loss = tr.step(x.cuda(), xt.cuda(), yh.cuda(), ys.cuda())
    torch.cuda.synchronize()
    assert abs(float(loss) - loss_o) < 3e-2, (float(loss), loss_o)
    assert abs(float(tr.grad_norm) - gn_o) < 0.08 * gn_o + 1e-3, (float(tr.grad_norm), gn_o)
    print(f"smoke ok: clahe byte-exact, logit diff {(lg.cpu() - lo).abs().max():.2e}, loss {float(loss):.4f} vs oracle {loss_o:.4f}, "
          f"grad-norm {float(tr.grad_norm):.4f} vs {gn_o:.4f}")
