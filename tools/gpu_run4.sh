#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-400; }
run kernels python -m pytest tests/test_kernels_gpu.py -q --timeout 300
run models python -m pytest tests/test_models_gpu.py -q --timeout 600
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TEETHRT_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 1300 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?"
