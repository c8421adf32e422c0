#!/bin/bash
# Runs each GPU test group under its own timeout so one hang cannot eat the whole gpurun call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n 25 gpurun_out/$name.log; }
run gemm python -m pytest tests/test_kernels_gpu.py -q -k "gemm or pack" -x --timeout 120
run wgrad_sweep python tools/wgrad_sweep.py
run eltwise python -m pytest tests/test_kernels_gpu.py -q -k "bn_ or se_block" --timeout 120
run dwconv python -m pytest tests/test_kernels_gpu.py -q -k "dwconv or stem" --timeout 300
run small python -m pytest tests/test_kernels_gpu.py -q -k "mil or tab or dropout or adamw" --timeout 120
run preproc python -m pytest tests/test_preproc_gpu.py -q --timeout 300
