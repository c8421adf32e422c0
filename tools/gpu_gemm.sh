#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-330; }
TAILN=3 run kernels python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -x -k "gemm or stem"
TAILN=24 run gemm_bench python tools/gemm_bench.py
