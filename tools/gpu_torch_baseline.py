#!/usr/bin/env python
"""The GPU bar the survey names: the reference's own module (oracle restatement of
MMJointDualHead on the timm shim = what the reference runs, with cuDNN / cuBLAS underneath) on the SAME B200, same recipe
(dual BCE, clip 1.0, AdamW, cosine), B = 64 @224, timed like bench.py (warm-up, CUDA events over K steps):

    eager_amp      torch.autocast(bf16) + GradScaler-free loop (train_mm_joint_dualtask.py:241-256 under --amp)
    eager_amp_cl   the same with channels_last weights and inputs
    compile_cl     channels_last + torch.compile(mode="max-autotune-no-cudagraphs")
    compile_cl_cg  channels_last + torch.compile(mode="max-autotune")  (CUDA graphs)
    infer_b1_*     batch-1 eval forward (eager / channels_last / compiled)

One JSON line per arm on stdout (also appended to --out).  This is a measurement of LIBRARY code, kept beside our number;
it is never on the product path.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def synth(B, img, dev):
    import torch
    g = torch.Generator().manual_seed(1000)
    x = torch.randn(B, 3, img, img, generator=g)
    xt = torch.randn(B, 9, generator=g)
    yh = (torch.rand(B, generator=g) < 0.6).float()
    ys = (yh * 0.8 + 0.2 * torch.rand(B, generator=g)).clamp(0, 1)
    return [t.to(dev) for t in (x, xt, yh, ys)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--arms", default="eager_amp,eager_amp_cl,compile_cl,compile_cl_cg,infer_b1")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import ref_models as R
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    lines = []

    def emit(d):
        d.update(gpu=torch.cuda.get_device_name(0), torch=torch.__version__, batch=args.batch, img=args.img)
        print(json.dumps(d), flush=True)
        lines.append(d)

    def train_arm(name, channels_last, compile_mode):
        torch.manual_seed(0)
        model = R.MMJointDualHead().to(dev).train()
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
        opt, sched = R.make_optimizer(model, t_max=args.steps + args.warmup + 8)
        x, xt, yh, ys = synth(args.batch, args.img, dev)
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        fwd = model
        t_compile = 0.0
        if compile_mode:
            fwd = torch.compile(model, mode=compile_mode)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logit, reg = fwd(x, xt)
            loss = R.dual_bce_loss(logit.float(), reg.float(), yh, ys, 1.0, 0.3)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            sched.step()
            return loss
        try:
            t0 = time.time()
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            t_compile = time.time() - t0
            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            emit({"arm": name, "kind": "train", "ms_per_step": ms, "images_per_s": args.batch / ms * 1e3, "steps": args.steps,
                  "warmup": args.warmup + 3, "first_3_steps_s": t_compile, "last_loss": float(loss),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
        except Exception as e:  # an arm that cannot run here is reported, not hidden
            emit({"arm": name, "kind": "train", "error": f"{type(e).__name__}: {str(e)[:300]}"})
        del model, opt, fwd
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        if compile_mode:
            torch._dynamo.reset()

    def infer_arm():
        torch.manual_seed(0)
        model = R.MMJointDualHead().to(dev).eval()
        x, xt, _, _ = synth(1, args.img, dev)

        def p50(fn, n=200, warm=30):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            return ts[len(ts) // 2], ts[int(len(ts) * 0.95)]
        with torch.no_grad():
            a = p50(lambda: model(x, xt))
            emit({"arm": "infer_b1_eager_fp32", "kind": "infer", "p50_ms": a[0], "p95_ms": a[1]})
            with torch.autocast("cuda", dtype=torch.bfloat16):
                a = p50(lambda: model(x, xt))
            emit({"arm": "infer_b1_eager_amp", "kind": "infer", "p50_ms": a[0], "p95_ms": a[1]})
            mcl = model.to(memory_format=torch.channels_last)
            xcl = x.contiguous(memory_format=torch.channels_last)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                a = p50(lambda: mcl(xcl, xt))
            emit({"arm": "infer_b1_eager_amp_cl", "kind": "infer", "p50_ms": a[0], "p95_ms": a[1]})
            try:
                half = R.MMJointDualHead().to(dev).eval().to(torch.bfloat16).to(memory_format=torch.channels_last)
                cm = torch.compile(half, mode="max-autotune")
                xb, xtb = xcl.to(torch.bfloat16), xt.to(torch.bfloat16)
                a = p50(lambda: cm(xb, xtb))
                emit({"arm": "infer_b1_compile_cg_bf16", "kind": "infer", "p50_ms": a[0], "p95_ms": a[1]})
            except Exception as e:
                emit({"arm": "infer_b1_compile_cg_bf16", "kind": "infer", "error": f"{type(e).__name__}: {str(e)[:300]}"})

    arms = args.arms.split(",")
    if "eager_amp" in arms:
        train_arm("eager_amp", False, None)
    if "eager_amp_cl" in arms:
        train_arm("eager_amp_cl", True, None)
    if "compile_cl" in arms:
        train_arm("compile_cl", True, "max-autotune-no-cudagraphs")
    if "compile_cl_cg" in arms:
        train_arm("compile_cl_cg", True, "max-autotune")
    if "infer_b1" in arms:
        infer_arm()
    if args.out:
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
