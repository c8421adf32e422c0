#!/bin/bash
# Per-layer depthwise probe on one box: new library (first TMA tile issued before the prologue) against the previous commit (built as in gpu_r02y3.sh).
mkdir -p gpurun_out
LIB=multimodal-teeth-restoration-selection_b200/libteethrt.so
cp $LIB /tmp/new.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dwconv" 2>&1 | tail -2
timeout 300 python tools/dw_probe.py 2>&1 | tail -15 > gpurun_out/r02y5_dw_new.jsonl
cp tools/ab/libteethrt_base.so $LIB
timeout 300 python tools/dw_probe.py 2>&1 | tail -15 > gpurun_out/r02y5_dw_base.jsonl
cp /tmp/new.so $LIB
tail -1 gpurun_out/r02y5_dw_new.jsonl; tail -1 gpurun_out/r02y5_dw_base.jsonl
