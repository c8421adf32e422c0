#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_res1.log 2>&1; echo "exit=$?"
TEETHRT_GEMM_RES_TILED=0 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_res0.log 2>&1; echo "exit=$?"
