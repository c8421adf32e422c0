#!/bin/bash
# GPU run of the "next"-row tests (f1-f4) plus the full suite and a short bench.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-330; }
TAILN=60 TMO=500 run new_tests python -m pytest tests/test_preproc_gpu.py tests/test_stack_gpu.py tests/test_calib_gpu.py -m gpu -q --timeout 200
TAILN=15 TMO=600 run gpu_tests python -m pytest tests/test_models_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x
TAILN=1 run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
