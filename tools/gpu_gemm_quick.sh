#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" 2>&1 | tail -2
timeout 300 python tools/gemm_probe.py 2>&1 | tail -1
GEMM_B=1 timeout 300 python tools/gemm_probe.py 2>&1 | tail -1
