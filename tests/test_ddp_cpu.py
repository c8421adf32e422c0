"""CPU, world_size 2 over gloo: the host-side data-parallel logic of the fused trainers (flat parameter re-homing, initial
state broadcast, reverse-order gradient buckets, sum all-reduce + 1/world scale).  The CUDA kernels are not involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from teethrt.train import FlatParams
        from teethrt.ddp import GradSync
        torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights on purpose
        model = nn.Sequential(nn.Linear(9, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Linear(64, 3))
        before = [p.detach().clone() for p in model.parameters()]
        flat = FlatParams(model)
        # re-homing keeps values, shapes and state_dict keys; parameters are views of the flat buffer
        assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
        assert all(p.data_ptr() >= flat.p.data_ptr() and p.data_ptr() < flat.p.data_ptr() + flat.p.numel() * 4
                   for p in model.parameters())
        assert all(o % 8 == 0 for o, _, _ in flat.offsets.values())
        sync = GradSync(flat)
        assert sync.world == 2
        model[1].running_mean.fill_(float(rank))
        sync.sync_initial_state(model)
        gathered = [torch.zeros_like(flat.p) for _ in range(world)]
        dist.all_gather(gathered, flat.p)
        assert torch.equal(gathered[0], gathered[1])       # everyone holds rank 0's parameters
        assert float(model[1].running_mean[0]) == 0.0      # ... and buffers
        # bucketed all-reduce in reverse forward order
        head_start = flat.offsets["3.weight"][0]
        ranges = sync.bucket_ranges([head_start])
        assert ranges == [(head_start, flat.numel), (0, head_start)]
        mid = flat.offsets["1.weight"][0]                  # three buckets: heads | late encoder | early encoder
        assert sync.bucket_ranges([head_start, mid]) == [(head_start, flat.numel), (mid, head_start), (0, mid)]
        flat.g.fill_(float(rank + 1))
        for r in ranges:
            sync.reduce(r)
        sync.finish()
        assert torch.all(flat.g == 3.0) and sync.grad_scale == 0.5      # sum over ranks; AdamW applies the 1/world
        assert torch.all(flat.grads["0.weight"] == 3.0)                 # per-parameter views see the reduced values
        q.put((rank, "ok"))
    except Exception as e:  # noqa
        q.put((rank, f"FAIL {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_flat_params_and_bucketed_allreduce_world2_gloo():
    import __graft_entry__ as g
    g.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
