#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config: mm dual-task TRAIN images/s (tf_efficientnet_b4_ns + tabular MLP,
synthetic 224x224 images + 9 clinical features, batch 64 per GPU, bf16 compute / fp32 masters, dropout 0.2, dual BCE,
clip 1.0, AdamW 3e-4 / wd 1e-4, cosine per iteration) and, with --infer, batch-1 inference p50 latency.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N>1, one rank per GPU)
    python bench.py --impl reference --gpus N ...            # the reference's own CPU path (oracle = reference classes
                                                             # on the timm shim), rank 0 only

One JSON line on stdout (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same through
DualTaskTrainer.step() with pinned HOST inputs (H2D inside the timed region) and a D2H read of the loss every step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG, BATCH, TAB = 224, 64, 9
GLOBAL_BATCH_DP = 512                            # BASELINE.json configs[4]: data-parallel training, batch 512 GLOBAL at 2/4/8 GPUs
FLOP_PER_IMG_FWD = 3.004e9                       # tf_efficientnet_b4_ns @224
FLOP_PER_IMG_TRAIN = 3 * FLOP_PER_IMG_FWD        # fwd + dgrad + wgrad
BYTES_PER_IMG_TRAIN = 245e6                      # layer-fused bf16 lower bound
BYTES_PER_STEP_OPT = 492e6                       # AdamW: 28 B x 17.56 M params
WORKLOAD = ("configs[1]: mm dual-task train step, tf_efficientnet_b4_ns + tab MLP(9->64->64) + dual heads, "
            "224x224, batch 64 on one GPU, dropout 0.2, dual BCE, clip 1.0, AdamW, cosine/iter")
WORKLOAD_DP = ("configs[4]: data-parallel mm dual-task training, global batch 512 (per-GPU 512/N), NCCL gradient all-reduce; "
               "same model and recipe as configs[1]")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(B, seed, device="cpu", pin=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, IMG, IMG, generator=g)
    xt = torch.randn(B, TAB, generator=g)
    yh = (torch.rand(B, generator=g) < 0.6).float()
    ys = (yh * 0.8 + 0.2 * torch.rand(B, generator=g)).clamp(0, 1)
    out = [x, xt, yh, ys]
    if pin:
        out = [t.pin_memory() for t in out]
    if device != "cpu":
        out = [t.to(device) for t in out]
    return out


# ====================================================================================================== reference arm
def oracle_train_throughput(batch, steps, warmup, threads):
    """The reference's hot loop (train_mm_joint_dualtask.py:241-256) on its own classes over the timm shim, fp32, CPU."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_models as R
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = R.MMJointDualHead().train()
    opt, sched = R.make_optimizer(model, t_max=1000)
    x, xt, yh, ys = synth_batch(batch, 1000)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        R.mm_train_step(model, opt, sched, x, xt, yh, ys)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return batch / med, med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    sample_b = 16
    steps, warmup = min(args.steps, 6), min(args.warmup, 2)
    ips, med = oracle_train_throughput(sample_b, steps, max(warmup, 1), cores)
    line = {"impl": "reference", "metric": "mm_dualtask_train_images_per_s", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": max(warmup, 1), "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.gpus == 1 else WORKLOAD_DP,
                       "global_batch": BATCH if args.gpus == 1 else GLOBAL_BATCH_DP, "parallelism": f"dp{args.gpus}",
                       "per_step_sample": f"CPU reference path, fp32, all host threads: each step is a batch of {sample_b} "
                                          "(bounded sample of the batch-64 step; images/s is batch-size normalised)"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} fp32 train steps of batch {sample_b} (not 64: bounded sample) through oracle/ref_models.py's "
                                       "restatement of MMJointDualHead on the oracle timm shim (torch CPU kernels, all host threads)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ====================================================================================================== our arm
def time_dominant_kernel(torch, ops, pk):
    """Roofline of the dominant kernel class (tcgen05 1x1-conv GEMM) on its largest launch of the step: blocks.1.0.conv_pw,
    A[64*112*112, 24] x W[144, 24]^T with the BN-statistics epilogue.  Algorithmic bytes = read A + W, write C (bf16)."""
    M, K, N = BATCH * 112 * 112, 24, 144
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Wt = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    stats = ops.new_stats(N, "cuda")
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)      # > 126 MB L2
    for _ in range(3):
        ops.gemm(A, Wt, ops.EPI_STATS, stats=stats, out=C)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, Wt, ops.EPI_STATS, stats=stats, out=C)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    t = statistics.mean(times)
    alg = (M * K + N * K + M * N) * 2
    ach = alg / t / 1e9
    return {"kernel": "gemm_kmajor_kernel (blocks.1.0.conv_pw fwd, M=802816 K=24 N=144, BN-stats epilogue)", "bound": "hbm",
            "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": ncu_traffic(),
            "peak_source": pk["src"], "avg_launch_us": t * 1e6, "algorithmic_bytes_per_launch": alg,
            "write_cap_gbs": 3880.0, "frac_of_write_floor": (M * N * 2 / 3880e9) / t,
            "note": "write-heavy launch: its floor is bytes_written / the measured 3.88 TB/s pure-write cap (tools/ubench/bw_probe.py); "
                    "traffic = dram read+write bytes of this launch from profiles/r02_ncu_full_top_kernels.csv (ncu --set full)"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the same launch from the committed ncu --set full capture (bytes)."""
    import csv
    p = os.path.join(ROOT, "profiles", "r02_ncu_full_top_kernels.csv")
    try:
        rows = list(csv.reader(open(p)))
        h, units = rows[0], rows[1]
        r = next(r for r in rows[2:] if "gemm_kmajor" in r[0])
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(k)
            tot += float(r[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        return tot
    except Exception:
        return None


def gemm_class_shapes(B):
    """(M, K, N, epilogue flags) of every 1x1-conv / stem GEMM launch of one B4 @224 train step: forward (BN-statistics
    epilogue) and data gradient (residual add where the block has a skip) — the 127 `gemm` rows of the per-op step table."""
    from teethrt import ops
    from teethrt.backbone import arch_spec
    stem, stages, feat = arch_spec("tf_efficientnet_b4_ns")
    h = ops.same_out(IMG, 2)
    out = [(B * h * h, 32, stem, ops.EPI_STATS)]
    for st in stages:
        for b in st:
            oh = ops.same_out(h, b["s"])
            skip = ops.EPI_RESIDUAL if (b["s"] == 1 and b["cin"] == b["cout"]) else 0
            if b["type"] == "ir":
                out += [(B * h * h, b["cin"], b["mid"], ops.EPI_STATS), (B * oh * oh, b["mid"], b["cout"], ops.EPI_STATS),
                        (B * oh * oh, b["cout"], b["mid"], 0), (B * h * h, b["mid"], b["cin"], skip)]
            else:
                out += [(B * oh * oh, b["cin"], b["cout"], ops.EPI_STATS), (B * oh * oh, b["cout"], b["cin"], 0)]
            h = oh
    cl = stages[-1][-1]["cout"]
    return out + [(B * h * h, cl, feat, ops.EPI_STATS), (B * h * h, feat, cl, 0)]


def time_gemm_class(torch, ops, pk):
    """Class average of the dominant kernel: every GEMM launch of the step on its real shape, in step order, timed with CUDA
    events (3 passes, last 2 averaged; no flush - consecutive launches evict each other as in the step, 4.8 GB per pass)."""
    shapes = gemm_class_shapes(BATCH)
    bufs = {}
    for (M, K, N, fl) in set(shapes):
        bufs[(M, K, N, fl)] = (torch.randn(M, K, device="cuda").to(torch.bfloat16), torch.randn(N, K, device="cuda").to(torch.bfloat16),
                               torch.empty(M, N, device="cuda", dtype=torch.bfloat16),
                               ops.new_stats(N, "cuda") if fl & ops.EPI_STATS else None,
                               torch.randn(M, N, device="cuda").to(torch.bfloat16) if fl & ops.EPI_RESIDUAL else None)
    tot_t, tot_b, n = 0.0, 0.0, 0
    for ps in range(3):
        evs = []
        for key in shapes:
            A, W, C, st, res = bufs[key]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(A, W, key[3], stats=st, residual=res, out=C)
            e1.record()
            evs.append((key, e0, e1))
        torch.cuda.synchronize()
        if ps == 0:
            continue
        for (M, K, N, fl), e0, e1 in evs:
            tot_t += e0.elapsed_time(e1) * 1e-3
            tot_b += (M * K + N * K + M * N * (2 if fl & ops.EPI_RESIDUAL else 1)) * 2
            n += 1
    ach = tot_b / tot_t / 1e9
    return {"launches_per_step": len(shapes), "algorithmic_gb_per_step": tot_b / 2 / 1e9, "us_per_step": tot_t / 2 * 1e6,
            "achieved": ach, "unit": "GB/s", "frac": ach / pk["hbm"],
            "what": "all 127 forward + data-gradient GEMM launches of one step (stem, conv_pw, conv_pwl, conv_head), real shapes, "
                    "step order, CUDA events per launch"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # optional: NCCL's internal stream at high priority (the train step's main chain is captured on a high-priority
        # stream).  Measured identical at N=2 (12.33 ms/step either way), so the default stays NCCL's own default
        opts = None
        if os.environ.get("TEETHRT_NCCL_HIPRIO", "0") != "0":
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    import teethrt
    from teethrt import ops
    from teethrt._lib import lib
    from teethrt.modules import MMJointDualHead
    from teethrt.train import DualTaskTrainer
    teethrt.init(local)
    pk = peaks()
    dev = torch.device("cuda", local)
    # N = 1 is configs[1] (batch 64); N > 1 is configs[4] AS WRITTEN: global batch 512, so 256 / 128 / 64 per GPU at 2 / 4 / 8.
    # The round-1 weak-scaling arm (64 per GPU at every N) is still measured and reported under "weak_scaling".
    if args.batch > 0:
        per_gpu = args.batch
    elif world == 1:
        per_gpu = BATCH
    else:
        if GLOBAL_BATCH_DP % world:
            raise SystemExit(f"global batch {GLOBAL_BATCH_DP} does not divide over {world} GPUs")
        per_gpu = GLOBAL_BATCH_DP // world
    torch.manual_seed(0)
    model = MMJointDualHead('tf_efficientnet_b4_ns', tab_in=TAB, tab_hidden=64, drop=0.2).to(dev)
    total_steps = 100000
    tr = DualTaskTrainer(model, lr=3e-4, weight_decay=1e-4, t_max=total_steps, alpha=1.0, beta=0.3, grad_clip=1.0,
                         graph=os.environ.get('TEETHRT_NO_GRAPH') != '1', seed=1234)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(B, steps):
        """-> (seconds device-resident, seconds end to end, last loss, h2d bytes per step) for `steps` steps at per-GPU batch B."""
        dev_batches = [synth_batch(B, 1000 + rank * 17 + i, device=dev) for i in range(2)]
        host_batches = [synth_batch(B, 2000 + rank * 17 + i, pin=True) for i in range(2)]
        # warm-up: eager steps + graph capture + replays (a new batch size builds its own plan and graphs)
        for i in range(max(args.warmup, 3) + tr.graph_warmup):
            tr.step(*dev_batches[i % 2])
        barrier()
        # ---- (1) device-resident inputs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            tr.step(*dev_batches[i % 2])
        e1.record()
        barrier()
        t_dev = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        # ---- (2) end to end: pinned host inputs -> H2D -> step -> D2H loss, every step
        # untimed warm-up of the host-input path itself: the first prefetch() allocates the staging set, copy stream and events,
        # the first loss_async() pins the read-back ring (page-locking takes milliseconds and used to land in the timed region)
        for i in range(2):
            tr.step(*host_batches[i % 2])
            tr.loss_value(tr.loss_async())
            tr.prefetch(*host_batches[(i + 1) % 2])
        barrier()
        losses = []
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        pending = None
        for i in range(steps):
            tr.step(*host_batches[i % 2])
            ticket = tr.loss_async()                     # D2H copy of THIS step's loss into pinned memory, queued behind the step
            tr.prefetch(*host_batches[(i + 1) % 2])      # next step's H2D (from pinned memory) overlaps this step's kernels
            if pending is not None:
                losses.append(tr.loss_value(pending))    # read step i-1's loss on the host while step i runs
            pending = ticket
        losses.append(tr.loss_value(pending))            # every step's result has been read inside the timed region
        e3.record()
        barrier()
        t_e2e = torch.tensor([e2.elapsed_time(e3) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        h2d = sum(t.numel() * t.element_size() for t in host_batches[0])
        return float(t_dev), float(t_e2e), losses[-1], h2d, dev_batches

    launches_before = lib.trt_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_dev, t_e2e, last_loss, h2d, dev_batches = measure(per_gpu, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    per_step = tr.launches_per_step or 0
    timed_steps = 2 * args.steps

    # ---- sustained arm (N = 1): the 0.25 s burst above never reaches the clocks a long job settles at
    sustained = None
    if world == 1 and args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds / (t_dev / args.steps)) + 1)
        s2 = ClockSampler(local)
        s2.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_sus):
            tr.step(*dev_batches[i % 2])
        e1.record()
        barrier()
        ts = e0.elapsed_time(e1) * 1e-3
        sustained = {"seconds": ts, "steps": n_sus, "value": per_gpu * n_sus / ts, "unit": "images/s", "ms_per_step": ts / n_sus * 1e3,
                     "clocks": s2.stop()}
        timed_steps += n_sus
    del dev_batches

    # ---- end to end from uint8 images (N = 1): pinned uint8 batch -> H2D -> batched train transform on the device
    # (RandomResizedCrop + flip + RandAugment + normalise + erase, teethrt.augment.BatchTrainTransform) -> step -> D2H loss.
    # What the reference's DataLoader workers do on the host per image (train_mm_joint_dualtask.py:72-85) runs here as ~10
    # launches per batch, and the copy is 12.6 MB of uint8 instead of 38.5 MB of fp32.
    e2e_u8 = None
    if world == 1 and not args.no_u8:
        from teethrt.augment import BatchTrainTransform
        import numpy as np
        SRC = 256
        rs = np.random.RandomState(7)
        u8_batches = [torch.from_numpy(rs.randint(0, 256, (per_gpu, SRC, SRC, 3), dtype=np.uint8)).pin_memory() for _ in range(2)]
        small = [[t.pin_memory() for t in synth_batch(per_gpu, 3000 + i)[1:]] for i in range(2)]
        btf = BatchTrainTransform(IMG, dtype=torch.bfloat16, seed=11)
        copy_stream = torch.cuda.Stream(device=dev)

        def u8_step(i):
            with torch.cuda.stream(copy_stream):
                raw = u8_batches[i % 2].to(dev, non_blocking=True)
                rest = [t.to(dev, non_blocking=True) for t in small[i % 2]]
            main = torch.cuda.current_stream(dev)
            main.wait_stream(copy_stream)
            for t in [raw] + rest:
                t.record_stream(main)            # allocated on the copy stream, read on the main stream
            x = btf(raw)
            tr.step(x, *rest)
            return tr.loss_async()
        for i in range(4):                      # new input dtype (bf16) -> its own plan: eager warm-up + capture
            tr.loss_value(u8_step(i))
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        pending = None
        for i in range(args.steps):
            ticket = u8_step(i)
            if pending is not None:
                last_u8 = tr.loss_value(pending)
            pending = ticket
        last_u8 = tr.loss_value(pending)
        e5.record()
        barrier()
        t_u8 = e4.elapsed_time(e5) * 1e-3
        e2e_u8 = {"value": per_gpu * args.steps / t_u8, "unit": "images/s", "ms_per_step": t_u8 / args.steps * 1e3,
                  "h2d_bytes_per_step": int(u8_batches[0].numel()) + sum(t.numel() * 4 for t in small[0]), "d2h_bytes_per_step": 4,
                  "last_loss": last_u8,
                  "what": f"pinned uint8 [{per_gpu},{SRC},{SRC},3] -> H2D -> BatchTrainTransform({IMG}) on the device -> DualTaskTrainer.step -> loss "
                          "read back; transform sampling on the host (vectorised), pixels Pillow-exact, sampling parity with timm unpinned"}
        timed_steps += args.steps

    # ---- weak-scaling arm of round 1 (64 per GPU at every N), when the main line ran another per-GPU batch
    weak = None
    if world > 1 and per_gpu != BATCH and args.batch <= 0:
        tw_dev, tw_e2e, _, _, _ = measure(BATCH, args.steps)
        weak = {"per_gpu_batch": BATCH, "global_batch": BATCH * world, "value": BATCH * world * args.steps / tw_dev,
                "e2e_value": BATCH * world * args.steps / tw_e2e, "unit": "images/s", "ms_per_step": tw_dev / args.steps * 1e3,
                "scaling": "weak"}
        timed_steps += 2 * args.steps

    # ---- data-parallel correctness on the hardware: after all those NCCL steps every rank must hold the same parameters
    ddp_identical = None
    if world > 1:
        p = tr.flat.p
        mine = torch.stack([p.double().sum(), p.double().abs().sum(), p.view(torch.int32).long().sum().double()])
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        ddp_identical = all(torch.equal(allv[0], v) for v in allv)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if ddp_identical is False:
            sys.exit(3)
        return
    imgs = per_gpu * world * args.steps
    value, e2e_value = imgs / t_dev, imgs / t_e2e
    scale_b = per_gpu / BATCH
    step_flops = FLOP_PER_IMG_TRAIN * per_gpu
    step_bytes = BYTES_PER_IMG_TRAIN * per_gpu + BYTES_PER_STEP_OPT
    ms = t_dev / args.steps * 1e3
    roof = time_dominant_kernel(torch, ops, pk)
    roof["class_average"] = time_gemm_class(torch, ops, pk)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        ips, med = oracle_train_throughput(8, 4, 1, cores)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "4 fp32 train steps of batch 8 (bounded sample of the batch-64 step) through oracle/ref_models.py's "
                         "restatement of the reference's MMJointDualHead on the oracle timm shim, all host threads"}
    infer = None
    if world == 1 and not args.no_infer:
        del tr
        torch.cuda.empty_cache()
        infer = infer_leg(max(10, args.steps))
    line = {"metric": "mm_dualtask_train_images_per_s", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3) + 2, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD if world == 1 else WORKLOAD_DP,
                       "global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "parallelism": f"dp{world}",
                       "l2": "per-step activations (~7 GB at batch 64) and parameters+moments (281 MB) exceed the 126 MB L2; no flush needed",
                       "cuda_graph": os.environ.get('TEETHRT_NO_GRAPH') != '1'},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": t_e2e / args.steps * 1e3, "last_loss": last_loss},
            "gpu_launches": int(per_step * timed_steps) if per_step else int(lib.trt_launch_count() - launches_before),
            "launches_per_step": per_step,
            "clocks": clocks,
            "roofline": roof,
            "step_roofline": {"tensor_frac_of_measured_sustained": step_flops / (ms * 1e-3) / 1e12 / pk["tf_sust"],
                              "achieved_tflops": step_flops / (ms * 1e-3) / 1e12,
                              "hbm_frac_layer_fused_bytes": step_bytes / (ms * 1e-3) / 1e9 / pk["hbm"],
                              "algorithmic_gb_per_step": step_bytes / 1e9, "gflop_per_step": step_flops / 1e9},
            "cpu_baseline": cpu}
    if sustained is not None:
        line["sustained"] = sustained
    if e2e_u8 is not None:
        line["e2e_u8_transform"] = e2e_u8
    if weak is not None:
        line["weak_scaling"] = weak
    if ddp_identical is not None:
        line["ddp_params_identical"] = ddp_identical
    if infer is not None:
        line["infer"] = infer
    lib_base = gpu_library_baseline()
    if lib_base is not None:
        line["gpu_library_baseline"] = lib_base
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if ddp_identical is False:
        sys.exit(3)


def gpu_library_baseline():
    """The reference's module under PyTorch/cuDNN on the same GPU class (tools/gpu_torch_baseline.py, recorded run): the bar
    SURVEY.md 2.2 / BASELINE.md 4 name.  An extra key; the `--impl reference` arm stays the CPU path."""
    p = os.path.join(ROOT, "profiles", "r02_gpu_torch_baseline.jsonl")
    if not os.path.exists(p):
        return None
    rows = [json.loads(l) for l in open(p) if l.strip()]
    out = {"source": "profiles/r02_gpu_torch_baseline.jsonl (tools/gpu_torch_baseline.py on a B200 of this pool; recorded, not re-run here)"}
    for r in rows:
        if "error" in r:
            out[r["arm"]] = {"error": r["error"]}
        elif r.get("kind") == "train":
            out[r["arm"]] = {"images_per_s": r["images_per_s"], "ms_per_step": r["ms_per_step"]}
        else:
            out[r["arm"]] = {"p50_ms": r["p50_ms"]}
    return out


def infer_leg(steps):
    """batch-1 inference p50 latency (the metric's second half): (A) one MMNet forward, (B) MMEnsemble semantics = 5 folds x
    3 TTA flips, (C) end to end through MMEnsemble.predict_image from a decoded host image.  -> dict"""
    import torch
    import teethrt
    from teethrt.modules import MMNet
    from teethrt.infer import _GraphedForward
    torch.cuda.set_device(0)
    teethrt.init(0)
    torch.manual_seed(0)
    folds = [MMNet().cuda().eval() for _ in range(5)]
    x1, t1 = torch.randn(1, 3, IMG, IMG, device="cuda"), torch.randn(1, TAB, device="cuda")
    x3, t3 = torch.randn(3, 3, IMG, IMG, device="cuda"), torch.randn(3, TAB, device="cuda")
    g1 = _GraphedForward(lambda a, b: folds[0](a, b)[0], [x1, t1])
    g3 = [_GraphedForward(lambda a, b, m=m: m(a, b)[0], [x3, t3]) for m in folds]

    def timeit(fn, n, warm):
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

    from teethrt.infer import replay_concurrently
    fold_streams = [torch.cuda.Stream() for _ in g3]

    def ens():          # MMEnsemble.predict_tensor's schedule: the five folds replay side by side, one stream each
        outs = replay_concurrently(g3, [(x3, t3)] * len(g3), fold_streams)
        return torch.stack([torch.sigmoid(o.mean() / 2.5) for o in outs]).mean()

    def ens_serial():
        out = [torch.sigmoid(g(x3, t3).mean() / 2.5) for g in g3]
        return torch.stack(out).mean()

    p50a, p95a = timeit(lambda: g1(x1, t1), steps * 10, 50)
    p50b, p95b = timeit(ens, steps * 5, 20)
    p50s, _ = timeit(ens_serial, steps * 2, 10)
    # (C) end to end through the public API: decoded 1024x1024 RGB image on the HOST -> MMEnsemble.predict_image (upload,
    # PIL-exact eval transform, 3 TTA flips, 5 folds, calibrated probabilities) -> host numpy, wall clock
    import tempfile
    import time
    import numpy as np
    from teethrt.infer import MMEnsemble, TAB_FEATURES
    with tempfile.TemporaryDirectory() as d:
        for f, m in enumerate(folds):
            torch.save({"model": m.state_dict(), "scaler_mean": np.zeros(TAB), "scaler_scale": np.ones(TAB), "thr": 0.5, "T": 2.5,
                        "args": {"backbone": "tf_efficientnet_b4_ns", "img_size": IMG, "tab_hidden": 64, "dropout": 0.2}, "epoch": 1},
                       os.path.join(d, f"mm_dualtask_fold{f}.pt"))
        ens_api = MMEnsemble(d, device="cuda")
    rgb = np.random.default_rng(0).integers(0, 256, size=(1024, 1024, 3), dtype=np.uint8)
    tab = {k: 0.5 for k in TAB_FEATURES}
    for _ in range(5):
        ens_api.predict_image(rgb, tab)
    ws = []
    for _ in range(steps * 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ens_api.predict_image(rgb, tab)
        ws.append((time.perf_counter() - t0) * 1e3)
    ws.sort()
    return {"single_p50_ms": p50a, "single_p95_ms": p95a, "ensemble_p50_ms": p50b, "ensemble_p95_ms": p95b,
            "ensemble_serial_p50_ms": p50s,
            "e2e_p50_ms": ws[len(ws) // 2], "e2e_p95_ms": ws[int(len(ws) * 0.95)],
            "what": "single = batch-1 MMNet forward (B4 @224 + tab), CUDA-graph replay, CUDA events; ensemble = 5 folds x 3 TTA flips, the folds replayed concurrently on five streams (ensemble_serial: one after the other); "
                    "e2e = MMEnsemble.predict_image wall clock: 1024x1024 RGB host array -> upload -> PIL-exact eval transform -> "
                    "TTA -> 5 folds -> probabilities on the host",
            "h2d_bytes": int(rgb.nbytes) + 5 * 3 * TAB * 4, "d2h_bytes": 5 * 4}


def run_infer(args):
    r = infer_leg(args.steps)
    print(json.dumps({"metric": "mm_batch1_infer_p50_latency", "value": r["single_p50_ms"], "unit": "ms", "higher_is_better": False,
                      "n_gpus": 1, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": "batch-1 MMNet forward (B4 @224 + tab), CUDA-graph replay"},
                      "single_forward": {"p50_ms": r["single_p50_ms"], "p95_ms": r["single_p95_ms"]},
                      "ensemble_5fold_3tta": {"p50_ms": r["ensemble_p50_ms"], "p95_ms": r["ensemble_p95_ms"]},
                      "e2e": {"value": r["e2e_p50_ms"], "unit": "ms", "p95_ms": r["e2e_p95_ms"], "what": r["what"],
                              "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": r["d2h_bytes"]}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--infer", action="store_true", help="batch-1 inference latency instead of the train step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-infer", action="store_true", help="skip the batch-1 inference leg of the default line (N=1)")
    ap.add_argument("--no-u8", action="store_true", help="skip the uint8 -> device train transform -> step end-to-end leg (N=1)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: 64 at N=1, 512/N at N>1)")
    ap.add_argument("--sustain-seconds", type=float, default=5.0, help="length of the sustained arm at N=1 (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.infer:
        run_infer(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
