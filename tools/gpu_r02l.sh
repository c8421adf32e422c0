#!/bin/bash
# fused SE MLP kernels: unit test, model-level tests, then A/B of the train step in one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "se_mlp_one_launch or bn_se_block" 2>&1 | tail -15 > gpurun_out/r02l_unit.log
cat gpurun_out/r02l_unit.log
grep -q "failed\|error" gpurun_out/r02l_unit.log && exit 1
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02l_models.log
cat gpurun_out/r02l_models.log
for i in 1 2; do
  TEETHRT_SE_FUSED=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-infer --no-u8 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02l_ab_fused_$i.log
  TEETHRT_SE_FUSED=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-infer --no-u8 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02l_ab_two_$i.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02l_ab_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'])
    except Exception as e: print(f, 'ERR', e)
PY
