#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-300; }
TAILN=15 run kernels python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -x -k dwconv
TAILN=20 run dw_bench python tools/dw_bench.py
