#!/usr/bin/env python
"""Bring-up aid: per-phase clock64 cycles of the one-launch SE MLP kernels (library built with -DTRT_SE_TIMING).
Slots fwd: 0 phase 1, 1 barrier, 2 reduce, 3 barrier, 4 phase 2; bwd: 8 phase 1, 9 barrier, 10 reduce, 11 barrier, 12 phase 2."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, teethrt
from teethrt import ops
from teethrt._lib import lib
teethrt.init()
raw = lib._cdll
raw.trt_se_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
raw.trt_se_debug_read.restype = None
N, REPS = 64, 50
for C, rd in [(144, 6), (672, 28), (1632, 68), (2688, 112)]:
    g = torch.Generator(device="cuda").manual_seed(C)
    R = lambda *s: torch.randn(*s, device="cuda", generator=g)
    pooled, Wr, br, We, be = R(N, C).abs() * 49, R(rd, C) * C ** -0.5, R(rd) * 0.1, R(C, rd) * rd ** -0.5, R(C) * 0.1
    s1, gate = torch.empty(N, rd, device="cuda"), torch.empty(N, C, device="cuda")
    ws = ops.se_workspace(N, C, rd, "cuda")
    sums, rec, gamma = R(5, N, C), torch.stack([R(C) * 0.1 + 1, R(C) * 0.1, R(C) * 0.2, R(C).abs() + 0.5]), R(C) * 0.1 + 1
    o = dict(ds2=torch.empty(N, C, device="cuda"), ds1=torch.zeros(N, rd, device="cuda"), dmean=torch.empty(N, C, device="cuda"),
             dWr=torch.empty_like(Wr), dbr=torch.empty_like(br), dWe=torch.empty_like(We), dbe=torch.empty_like(be),
             coef=torch.empty(3, C, device="cuda"), dgamma=torch.empty(C, device="cuda"), dbeta=torch.empty(C, device="cuda"))
    bn = ops.se_bn(sums, rec, gamma, o["coef"], o["dgamma"], o["dbeta"], N * 49)
    for _ in range(3):
        ops.se_fwd(pooled, 1 / 49, Wr, br, We, be, s1, gate, ws=ws)
    raw.trt_se_debug_read(None, 1)
    for _ in range(REPS):
        ops.se_fwd(pooled, 1 / 49, Wr, br, We, be, s1, gate, ws=ws)
        ops.se_bwd(sums[0], gate, s1, pooled, 1 / 49, Wr, We, o["ds2"], o["ds1"], o["dmean"], o["dWr"], o["dbr"], o["dWe"], o["dbe"],
                   ds1_zeroed=True, bn=bn, ws=ws)
    buf = (ctypes.c_ulonglong * 16)()
    raw.trt_se_debug_read(buf, 1)
    cyc = [buf[i] / REPS for i in range(16)]
    print(json.dumps({"C": C, "rd": rd, "fwd_cycles": {"phase1": cyc[0], "bar1": cyc[1], "reduce": cyc[2], "bar2": cyc[3], "phase2": cyc[4]},
                      "bwd_cycles": {"phase1": cyc[8], "bar1": cyc[9], "reduce": cyc[10], "bar2": cyc[11], "phase2": cyc[12]}}), flush=True)
