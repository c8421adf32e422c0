#!/bin/bash
# GPU suite, CLAHE before/after, batch-1 inference: programmatic dependent launch A/B + launch list of one forward.
mkdir -p gpurun_out
T=${TAG:-r02e}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=8 TMO=900 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
cat > /tmp/clahe_t.py <<'PY'
import sys, os, json, statistics
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, torch, teethrt
from teethrt import preproc
from teethrt._lib import check, ptr, stream
import ref_preproc as P
teethrt.init()
for kind in ("radiograph", "noise"):
    for n in (1, 64):
        imgs = torch.from_numpy(np.stack([P.image_set(kind, 1024, 1024, seed=i % 4) for i in range(n)])).cuda()
        dst = torch.empty_like(imgs)
        ws = torch.empty(teethrt.lib.trt_clahe_workspace_bytes(n), device="cuda", dtype=torch.uint8)
        tab = preproc.device_tables("cuda")
        fn = lambda: check(teethrt.lib.trt_clahe_bgr_u8(ptr(imgs), ptr(dst), n, 1024, 1024, 3.0, ptr(tab), ptr(ws), ws.numel(), stream()))
        for _ in range(3): fn()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        t = statistics.median(ts)
        print(json.dumps({"kind": kind, "n": n, "slow": os.environ.get("TEETHRT_CLAHE_SLOW", "0"), "us": t, "hbm_frac_6Bpx": 6 * n * 1024 * 1024 / (t * 1e-6) / 1e9 / 6527.1}))
PY
TAILN=4 run clahe_fast python /tmp/clahe_t.py
TEETHRT_CLAHE_SLOW=1 TAILN=4 run clahe_slow python /tmp/clahe_t.py
TEETHRT_PDL=1 TAILN=1 CUT=700 run infer_pdl1 python bench.py --infer --steps 20
TEETHRT_PDL=0 TAILN=1 CUT=700 run infer_pdl0 python bench.py --infer --steps 20
TEETHRT_PDL=1 TAILN=1 CUT=700 run infer_pdl1b python bench.py --infer --steps 20
Q="--steps 30 --warmup 3 --no-infer --no-cpu-baseline --sustain-seconds 0 --no-u8"
TEETHRT_PDL=1 TAILN=1 CUT=330 run train_pdl1 python bench.py $Q
TEETHRT_PDL=0 TAILN=1 CUT=330 run train_pdl0 python bench.py $Q
timeout 120 python tools/prof_infer.py > /dev/null 2>&1 && \
TEETHRT_PDL=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_infer_launches.csv python tools/prof_infer.py > gpurun_out/${T}_ncu_infer.log 2>&1
echo "infer launch list exit=$?"
