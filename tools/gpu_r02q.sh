#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "se_ or bn_se" 2>&1 | tail -3
timeout 300 python tools/se_probe.py 2>&1 | tail -10
