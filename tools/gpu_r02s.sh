#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/knockout.py --infer 1 --only gemm 2>&1 | tail -3
timeout 300 python tools/knockout.py --infer 3 --only gemm 2>&1 | tail -3
