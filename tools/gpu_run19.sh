#!/bin/bash
mkdir -p gpurun_out
for f in 11 10 01 00; do
  export TEETHRT_FUSE_BN_FWD=${f:0:1} TEETHRT_FUSE_BN_BWD=${f:1:1}
  echo "== fuse fwd=$TEETHRT_FUSE_BN_FWD bwd=$TEETHRT_FUSE_BN_BWD"
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-140
done
unset TEETHRT_FUSE_BN_FWD TEETHRT_FUSE_BN_BWD
timeout 300 python -m pytest tests/test_models_gpu.py -q --timeout 300 -x 2>&1 | tail -2
