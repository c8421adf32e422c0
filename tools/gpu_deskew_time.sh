#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_preproc_gpu.py -m gpu -q --timeout 200 -k "canny or deskew or preprocess" 2>&1 | tail -2
timeout 200 python - <<'PY'
import sys, time, statistics
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch, teethrt
from teethrt import preproc
import ref_preproc as P
teethrt.init()
img = torch.from_numpy(P.tooth_image(1024, 1024, 3, 33.0)).cuda()
def wall(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3
print("deskew 1024^2 device-resident ms", wall(lambda: preproc.deskew(img)), "canny ms", wall(lambda: preproc.canny(img, want_edges=False)))
PY
