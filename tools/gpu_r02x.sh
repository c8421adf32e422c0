#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dwconv or lazy" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
for cfg in "TEETHRT_DW_SAVE_ACT=1" "TEETHRT_DW_SAVE_ACT=0"; do
  env $cfg timeout 300 python bench.py --steps 40 --warmup 5 --no-infer --no-u8 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02x_tmp.log
  python - "$cfg" gpurun_out/r02x_tmp.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['ms_per_step'],3), round(d['value'],1))
PY
done; done
