#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm or mil" 2>&1 | tail -3
timeout 300 python tools/gemm_probe.py 2>&1 | tail -1
GEMM_B=1 timeout 300 python tools/gemm_probe.py 2>&1 | tail -1
timeout 300 python tools/knockout.py --infer 1 --only gemm 2>&1 | tail -3
for i in 1 2; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-infer --no-u8 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02t_tmp.log
  python - gpurun_out/r02t_tmp.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('train', round(d['ms_per_step'],3), round(d['value'],1))
PY
done
