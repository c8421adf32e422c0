#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-330; }
TAILN=3 run kernels python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -x -k "gemm or stem"
TAILN=1 run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
TEETHRT_WGRAD_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/step_launches.csv python tools/step_profile.py --log gpurun_out/step_ops.json > gpurun_out/stepprof_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/stepprof_ncu.log
