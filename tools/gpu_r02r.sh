#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "se_ or bn_se" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
for cfg in "TEETHRT_SE_SMALL=1" "TEETHRT_SE_SMALL=0"; do
  echo $cfg; env $cfg timeout 300 python tools/knockout.py --infer 1 --only se_fwd_mlp 2>&1 | tail -3 | head -2
done; done
