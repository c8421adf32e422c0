#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_models_gpu.py -q --timeout 600 -k "autograd_path" > gpurun_out/models2.log 2>&1; echo "models exit=$?"; tail -5 gpurun_out/models2.log
TEETHRT_NO_GRAPH=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_nograph.log 2>&1 && \
TEETHRT_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 1300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/plain_nograph.log | cut -c1-300; tail -3 gpurun_out/ncu.log | cut -c1-300; wc -l gpurun_out/launches_r1.csv
