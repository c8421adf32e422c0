#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kmajor -s 4 -c 1 -f -o gpurun_out/prof_gemm_r1b python tools/prof_gemm.py > gpurun_out/prof_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/prof_ncu.log
