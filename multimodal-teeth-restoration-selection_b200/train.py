"""Fused train steps for the two hot loops of the reference, with the same recipe and none of its per-step host work:

    DualTaskTrainer.step(x_img, x_tab, y_h, y_s, w)   <- experiments/multimodal_v1/train_mm_joint_dualtask.py:236-256
    MILTrainer.step(bags, y)                          <- experiments/vision_v2/train_mil_attention_v1.py:177-189

zero_grad -> forward -> loss -> backward -> clip_grad_norm_(1.0) -> AdamW(lr, wd) -> cosine LR per iteration, all on
libteethrt kernels over FLAT fp32 parameter/gradient/moment buffers (the nn.Parameters of the model become views of the
flat buffer, so state_dict()/checkpoints are unchanged).  The whole step is captured once into a CUDA graph and replayed;
the step counter, LR schedule and dropout seeds live on the device.  Data parallel: one process per GPU, the flat gradient
is all-reduced over NCCL in reverse-execution buckets launched on a side stream while the rest of backward runs.
"""
import torch
import torch.distributed as dist

from . import ops
from .backbone import forward_train, backward_train, backward_train_iter
from .ddp import GradSync
from .modules import TAB_PARAM_KEYS

ALIGN = 8  # floats: every parameter starts on a 32-byte boundary inside the flat buffer


class FlatParams:
    """Re-homes every parameter of `module` into one flat fp32 buffer (forward/registration order)."""

    def __init__(self, module):
        self.module = module
        named = list(module.named_parameters())
        dev = named[0][1].device
        self.offsets, off = {}, 0
        for n, p in named:
            if p.dtype != torch.float32:
                raise TypeError(f"{n}: fp32 master parameters expected")
            self.offsets[n] = (off, p.numel(), tuple(p.shape))
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.p = torch.zeros(off, device=dev)
        self.g = torch.zeros(off, device=dev)
        self.m = torch.zeros(off, device=dev)
        self.v = torch.zeros(off, device=dev)
        with torch.no_grad():
            for n, p in named:
                o, k, shp = self.offsets[n]
                view = self.p[o:o + k].view(shp)
                view.copy_(p.data)
                p.data = view
        self.grads = {n: self.g[o:o + k].view(shp) for n, (o, k, shp) in self.offsets.items()}
        self.names = [n for n, _ in named]

    def attach_grads(self):
        """Expose the flat gradient through .grad (for inspection / torch optimisers)."""
        for n, p in self.module.named_parameters():
            p.grad = self.grads[n]


class _FusedTrainer:
    def __init__(self, model, lr, weight_decay, t_max, grad_clip, graph, process_group, num_buckets, graph_warmup=2):
        self.model = model.train()
        self.flat = FlatParams(model)
        dev = self.flat.p.device
        self.dev = dev
        self.state = ops.OptimState(dev, lr, t_max=t_max)
        self.wd, self.clip = float(weight_decay), float(grad_clip)
        self.normsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.grad_norm = torch.zeros(1, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.sync = GradSync(self.flat, process_group)
        self.world = self.sync.world
        self.num_buckets = max(1, num_buckets)
        self.use_graph = graph
        self.graph_warmup = graph_warmup
        self._graphs = None
        self._nsteps = 0
        self._static = None
        self._stage = None
        self._staged = None
        self.launches_per_step = None
        self._loss_ring = None
        # one plan (static inputs, scratch, captured graphs) per input shape: the last batch of an epoch is usually ragged
        # (DataLoader(shuffle=True) without drop_last, train_mm_joint_dualtask.py:211)
        self._plans, self._plan_key, self._plan_steps, self._plan_attrs = {}, None, 0, None
        self.sync.sync_initial_state(self.model)

    def _bucket_ranges(self, boundaries):
        return self.sync.bucket_ranges(boundaries)

    # ---- to be provided by subclasses --------------------------------------------------------------------------------
    def _make_static(self, *inputs):
        raise NotImplementedError

    def _segments(self):
        """List of callables; segment i finishes the gradients of bucket i (reverse order)."""
        raise NotImplementedError

    # ---- one optimiser step ------------------------------------------------------------------------------------------
    def _optimizer(self):
        self.state.advance()
        ops.grad_sumsq(self.flat.g, self.normsq)
        ops.adamw_step(self.flat.p, self.flat.g, self.flat.m, self.flat.v, self.state, self.normsq, self.grad_norm,
                       self.sync.grad_scale, self.clip, 1e-8, self.wd)

    def _run_eager(self):
        segs, ranges = self._segments()
        self.flat.g.zero_()
        for i, seg in enumerate(segs):
            seg()
            self._reduce_bucket(ranges[i] if self.world > 1 else None)
        self._finish_comm()
        self._optimizer()

    def _reduce_bucket(self, rng):
        self.sync.reduce(rng)

    def _finish_comm(self):
        self.sync.finish()

    def _capture(self):
        from ._lib import lib
        segs, ranges = self._segments()
        c0 = lib.trt_launch_count()
        graphs = []
        pool = None
        first = True
        # the data-gradient chain is captured on a HIGH-priority stream, so its kernel nodes win the block scheduler over
        # the weight-gradient branch (side stream, default = lowest priority) whenever both have blocks pending
        import os
        prio = int(os.environ.get("TEETHRT_MAIN_PRIORITY", "-1"))
        cap = torch.cuda.Stream(device=self.dev, priority=prio) if prio != 0 else None
        for seg in segs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=cap):
                if first:
                    self.flat.g.zero_()
                seg()
            pool = g.pool()
            graphs.append(g)
            first = False
        gopt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gopt, pool=pool, stream=cap):
            self._optimizer()
        self._graphs = (graphs, gopt, ranges)
        self.launches_per_step = int(lib.trt_launch_count() - c0)     # kernels recorded into the replayed graphs

    def _replay(self):
        graphs, gopt, ranges = self._graphs
        for i, g in enumerate(graphs):
            g.replay()
            self._reduce_bucket(ranges[i] if self.world > 1 else None)
        self._finish_comm()
        gopt.replay()

    def prefetch(self, *inputs):
        """Start the host->device copy of the NEXT step's (pinned) inputs on a copy stream, so it overlaps the step that is
        running; the matching step() call then only does a device-to-device copy.  (What the reference gets from DataLoader
        workers + pin_memory + non_blocking copies, train_mm_joint_dualtask.py:211,238-240.)"""
        if self._static is None or self._plan_key != self._shape_key(inputs):
            return
        if self._stage is None:
            self._stage = [torch.empty_like(t) for t in self._static]
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._stage_free = torch.cuda.Event()
            self._stage_ready = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(self.dev))
        self._copy_stream.wait_event(self._stage_free)            # the previous step has finished reading the staging set
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._stage, inputs):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
            self._stage_ready.record(self._copy_stream)
        self._staged = tuple(inputs)               # the tensors themselves: CPython recycles the id() of a freed tensor

    @staticmethod
    def _shape_key(inputs):
        return tuple(None if t is None else tuple(t.shape) for t in inputs)

    _BASE_PLAN_ATTRS = ("_static", "_graphs", "_stage", "_staged", "_copy_stream", "_stage_free", "_stage_ready",
                        "_plan_steps", "launches_per_step")

    def _switch_plan(self, key):
        if self._plan_key is not None:
            self._plans[self._plan_key] = {k: self.__dict__.get(k) for k in self._plan_attrs}
        plan = self._plans.get(key)
        if plan is None:
            self._static = self._graphs = self._stage = self._staged = None
            self._plan_steps = 0
        else:
            self.__dict__.update(plan)
        self._plan_key = key

    def _step(self, *inputs):
        with torch.cuda.device(self.dev):       # streams, events and graphs below belong to the trainer's device
            return self._step_on_device(*inputs)

    def _step_on_device(self, *inputs):
        key = self._shape_key(inputs)
        if key != self._plan_key:
            self._switch_plan(key)
        if self._static is None:
            before = set(self.__dict__)
            self._make_static(*inputs)
            self._plan_attrs = tuple(sorted(set(self._BASE_PLAN_ATTRS) | (set(self.__dict__) - before) | set(self._plan_attrs or ())))
        staged, self._staged = self._staged, None   # consumed or dropped: a step with other tensors never sees it again
        if staged is not None and len(staged) == len(inputs) and all(a is b for a, b in zip(staged, inputs)):
            main = torch.cuda.current_stream(self.dev)
            main.wait_event(self._stage_ready)
            for dst, src, given in zip(self._static, self._stage, inputs):
                if given is not None:
                    dst.copy_(src, non_blocking=True)
            self._stage_free.record(main)
        else:
            for dst, src in zip(self._static, inputs):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
        if self.use_graph and self._plan_steps >= self.graph_warmup:
            if self._graphs is None:
                torch.cuda.synchronize()
                self._capture()
            self._replay()
        else:
            self._run_eager()
        self._nsteps += 1
        self._plan_steps += 1
        self._invalidate_eval_caches()
        return self.loss

    def _invalidate_eval_caches(self):
        """The step rewrote the flat parameters and the BatchNorm running statistics in place: folded-BN records and
        packed bf16 weights cached for eval are stale, whatever mode the module is in."""
        for m in self.model.modules():
            if hasattr(m, "_eval_cache"):
                m._eval_cache = None

    def lr(self):
        return self.state.read()["lr"]

    # ---- loss read-back without stalling the launch queue -----------------------------------------------------------
    def loss_async(self):
        """Queue a device->host copy of the step that was just launched into a pinned ring; -> ticket for loss_value().
        Reading ticket i after launching step i+1 keeps one step in flight (the reference's `loss.item()` per step,
        train_mm_joint_dualtask.py:256, drains the GPU every iteration)."""
        if self._loss_ring is None:
            depth = 16
            self._loss_ring = (torch.empty(depth, dtype=torch.float32).pin_memory(), [torch.cuda.Event() for _ in range(depth)])
            self._loss_tickets = 0
        buf, evs = self._loss_ring
        slot = self._loss_tickets % len(evs)
        buf[slot:slot + 1].copy_(self.loss, non_blocking=True)
        evs[slot].record(torch.cuda.current_stream(self.dev))
        self._loss_tickets += 1
        return self._loss_tickets - 1

    def loss_value(self, ticket):
        buf, evs = self._loss_ring
        if not 0 <= self._loss_tickets - 1 - ticket < len(evs):
            raise ValueError(f"loss ticket {ticket} is no longer in the ring (latest {self._loss_tickets - 1}, depth {len(evs)})")
        evs[ticket % len(evs)].synchronize()
        return float(buf[ticket % len(evs)])


class DualTaskTrainer(_FusedTrainer):
    """MMJointDualHead + dual BCE + clip + AdamW + cosine (train_mm_joint_dualtask.py:217-256)."""

    def __init__(self, model, lr=3e-4, weight_decay=1e-4, t_max=0, alpha=1.0, beta=0.3, grad_clip=1.0,
                 use_sample_weights=False, graph=True, process_group=None, num_buckets=4, seed=0):
        super().__init__(model, lr, weight_decay, t_max, grad_clip, graph, process_group, num_buckets)
        self.alpha, self.beta, self.use_w, self.seed = float(alpha), float(beta), bool(use_sample_weights), int(seed)

    def _make_static(self, x_img, x_tab, y_h, y_s, w):
        dev = self.dev
        B = x_img.shape[0]
        self._static = [torch.empty(tuple(x_img.shape), device=dev, dtype=x_img.dtype if x_img.dtype == torch.bfloat16 else torch.float32),
                        torch.empty((B, x_tab.shape[1]), device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev),
                        torch.ones(B, device=dev)]
        self.scratch = ops.tab_heads_scratch(B, self.model.tab_hidden, dev)
        self.head_out = {k: torch.empty(B, device=dev) for k in ("logit", "reg", "dlogit", "dreg")}
        self.head_out["loss"] = self.loss
        self.dfeat = torch.empty((B, self.model.backbone.num_features), device=dev)

    def _segments(self):
        m, fl = self.model, self.flat
        x_img, x_tab, y_h, y_s, w = self._static
        enc = m.backbone
        step_ptr = self.state.buf            # first 8 bytes = u64 step counter: mixes into the dropout seed
        bn = m.tab[1]
        params = [fl.p[fl.offsets[k][0]:fl.offsets[k][0] + fl.offsets[k][1]].view(fl.offsets[k][2]) for k in TAB_PARAM_KEYS]
        grads = [fl.grads[k] for k in TAB_PARAM_KEYS]
        enc_grads = {n[len("backbone."):]: g for n, g in fl.grads.items() if n.startswith("backbone.")}
        holder = {}

        def fwd_and_heads():
            feat, ctx = forward_train(enc, x_img)
            holder["feat"], holder["ctx"] = feat, ctx
            ops.tab_heads_fwd(feat, x_tab, params, bn.running_mean, bn.running_var, bn.num_batches_tracked, self.scratch, True,
                              m.drop_p, targets=(y_h, y_s, w if self.use_w else None), alpha=self.alpha, beta=self.beta,
                              seed=self.seed, step=step_ptr, out=self.head_out)
            ops.tab_heads_bwd(feat, x_tab, params, bn.running_mean, bn.running_var, self.head_out["dlogit"],
                              self.head_out["dreg"], self.dfeat, grads, self.scratch, True, m.drop_p, seed=self.seed,
                              step=step_ptr)

        # bucket 0 = tab + heads (everything after the backbone in the flat buffer); the backbone is cut where the parameter
        # mass is: the last two stages + head hold ~80 % of the weights but are the FIRST quarter of the backward pass, so
        # their all-reduce (bucket 1) hides behind the backward of the early, activation-heavy stages.  With 4 buckets those
        # are cut once more in front of stage 2: the LAST bucket (stem + stages 0-1, ~1 % of the weights) is the only
        # all-reduce nothing can hide, so it should be a latency-sized message (round 1: 14 MB exposed, 0.24 ms/step at N=8)
        first_head = min(fl.offsets[k][0] for k in TAB_PARAM_KEYS)
        splits = []
        if self.world > 1 and self.num_buckets >= 3 and len(enc.blocks) >= 3:
            splits.append(f"blocks.{len(enc.blocks) - 2}.0")    # first block of the second-to-last stage
            if self.num_buckets >= 4 and len(enc.blocks) >= 5:
                splits.append("blocks.2.0")                     # first block of stage 2
        if not splits:
            def enc_backward():
                backward_train(enc, holder["ctx"], self.dfeat, enc_grads)
                holder.clear()
            return [fwd_and_heads, enc_backward], self._bucket_ranges([first_head])

        def first_segment():
            holder["it"] = backward_train_iter(enc, holder["ctx"], self.dfeat, enc_grads, split_after=tuple(splits))
            assert next(holder["it"]) == splits[0]

        def middle_segment(expected):
            def run():
                assert next(holder["it"]) == expected
            return run

        def last_segment():
            for _ in holder["it"]:
                pass
            holder.clear()

        segs = [fwd_and_heads, first_segment] + [middle_segment(sp) for sp in splits[1:]] + [last_segment]
        cuts = [min(o for n, (o, _, _) in fl.offsets.items() if n.startswith("backbone." + sp + ".")) for sp in splits]
        return segs, self._bucket_ranges([first_head] + cuts)

    def prefetch(self, x_img, x_tab, y_h, y_s, w=None):
        super().prefetch(x_img, x_tab, y_h, y_s, w)

    def step(self, x_img, x_tab, y_h, y_s, w=None):
        """One train step; returns the (device-resident) loss tensor — read it whenever convenient, no per-step sync."""
        if x_img.shape[0] < 2:
            raise ValueError("Expected more than 1 value per channel when training (BatchNorm1d needs batch > 1)")
        return self._step(x_img, x_tab, y_h, y_s, w)


class MILTrainer(_FusedTrainer):
    """MILNet + BCE + clip + AdamW + cosine (train_mil_attention_v1.py:168-189)."""

    def __init__(self, model, lr=2e-4, weight_decay=1e-4, t_max=0, grad_clip=1.0, graph=True, process_group=None, seed=0):
        super().__init__(model, lr, weight_decay, t_max, grad_clip, graph, process_group, 2)
        self.seed = int(seed)

    def _make_static(self, bags, y):
        dev = self.dev
        self._static = [torch.empty(tuple(bags.shape), device=dev, dtype=bags.dtype if bags.dtype == torch.bfloat16 else torch.float32),
                        torch.empty(bags.shape[0], device=dev)]

    def _segments(self):
        m, fl = self.model, self.flat
        bags, y = self._static
        B, K = bags.shape[:2]
        enc = m.encoder
        step_ptr = self.state.buf
        P = lambda k: fl.p[fl.offsets[k][0]:fl.offsets[k][0] + fl.offsets[k][1]].view(fl.offsets[k][2])
        G = fl.grads
        enc_grads = {n[len("encoder."):]: g for n, g in G.items() if n.startswith("encoder.")}
        holder = {}

        def fwd_and_pool():
            feat, ctx = forward_train(enc, bags.view(B * K, *bags.shape[2:]))
            H = feat.view(B, K, -1)
            M, A, gV, gU = ops.mil_attn_fwd(H, P("mil.attention_V.weight"), P("mil.attention_V.bias"),
                                            P("mil.attention_U.weight"), P("mil.attention_U.bias"),
                                            P("mil.attention_w.weight").view(-1), P("mil.attention_w.bias"), save=True)
            logit = ops.linear1_fwd(M, P("head.weight").view(-1), P("head.bias"), m.drop_p, self.seed, step_ptr)
            loss, dlogit = ops.bce_logits(logit, y)
            self.loss.copy_(loss)
            dM = ops.linear1_bwd(dlogit, M, P("head.weight").view(-1), G["head.weight"], G["head.bias"], m.drop_p, self.seed,
                                 step_ptr)
            dH = ops.mil_attn_bwd(dM, H, A, gV, gU, P("mil.attention_V.weight"), P("mil.attention_U.weight"),
                                  P("mil.attention_w.weight").view(-1), G["mil.attention_V.weight"], G["mil.attention_V.bias"],
                                  G["mil.attention_U.weight"], G["mil.attention_U.bias"], G["mil.attention_w.weight"],
                                  G["mil.attention_w.bias"])
            holder["ctx"], holder["dfeat"] = ctx, dH.view(B * K, -1)

        def enc_backward():
            backward_train(enc, holder["ctx"], holder["dfeat"], enc_grads)
            holder.clear()

        first_head = min(o for n, (o, _, _) in fl.offsets.items() if not n.startswith("encoder."))
        return [fwd_and_pool, enc_backward], self._bucket_ranges([first_head])

    def step(self, bags, y):
        return self._step(bags, y)
