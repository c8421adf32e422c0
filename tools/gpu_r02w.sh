#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_augment_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 --no-infer --no-cpu-baseline --sustain-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02w_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1])
print('dev', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'u8', round(d['e2e_u8_transform']['ms_per_step'],3), round(d['e2e_u8_transform']['value'],1))
PY
