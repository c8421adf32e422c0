#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-330; }
TAILN=8 run kernels python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -x -k "se_block"
TAILN=4 run models python -m pytest tests/test_models_gpu.py -q --timeout 300 -x
TAILN=1 run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
