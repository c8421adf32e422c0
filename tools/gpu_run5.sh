#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-300; }
run kernels python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -k "se_block or dwconv"
run bench python bench.py --steps 20 --warmup 3
run infer python bench.py --infer --steps 20
TAILN=20 run micro python tools/microbench.py
timeout 300 python tools/prof_gemm.py > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 8 -c 4 -f -o gpurun_out/prof_gemm_r1 python tools/prof_gemm.py > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/prof_ncu.log
