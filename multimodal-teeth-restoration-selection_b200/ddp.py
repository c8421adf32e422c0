"""Data-parallel gradient exchange for the fused trainers: the ONE collective of the hot path.

The reference is single-process (train_mm_joint_dualtask.py:189).  Here each rank owns a full replica; after (and while)
backward runs, the flat fp32 gradient buffer is all-reduced (sum) over NCCL in contiguous buckets, last-executed parameters
first, on a side stream so the transfer of one bucket overlaps the compute of the next segment.  The 1/world averaging is
folded into the AdamW kernel's grad_scale.  BatchNorm statistics stay local.  Device-agnostic on purpose (gloo on CPU in the
tests, NCCL on GPUs)."""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, flat, process_group=None):
        self.flat = flat
        self.pg = process_group
        active = process_group is not None or (dist.is_available() and dist.is_initialized())
        self.world = dist.get_world_size(process_group) if active else 1
        if self.world > 1 and self.pg is None:
            self.pg = dist.group.WORLD
        self.cuda = flat.p.is_cuda
        self.stream = torch.cuda.Stream(device=flat.p.device) if (self.cuda and self.world > 1) else None

    def sync_initial_state(self, module):
        """Every rank starts from rank 0's parameters and buffers (BN running statistics, counters)."""
        if self.world == 1:
            return
        dist.broadcast(self.flat.p, src=dist.get_global_rank(self.pg, 0), group=self.pg)
        for b in module.buffers():
            dist.broadcast(b, src=dist.get_global_rank(self.pg, 0), group=self.pg)

    def bucket_ranges(self, boundaries):
        """Contiguous ranges of the flat gradient in REVERSE forward order: bucket i is complete after segment i."""
        edges = sorted(set([0, self.flat.numel] + [b for b in boundaries if 0 < b < self.flat.numel]))
        return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)][::-1]

    def reduce(self, rng):
        if self.world == 1 or rng is None:
            return
        view = self.flat.g[rng[0]:rng[1]]
        if self.stream is None:
            dist.all_reduce(view, group=self.pg)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            dist.all_reduce(view, group=self.pg)

    def finish(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

    @property
    def grad_scale(self):
        return 1.0 / self.world
