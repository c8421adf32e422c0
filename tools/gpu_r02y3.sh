#!/bin/bash
# Per-layer depthwise probe (tools/dw_probe.py, in-graph, HBM-cold) on one box: new library with tall weight-gradient tiles,
# the same with TEETHRT_DW_WGRAD_TALL=0, and the previous commit's library (tools/ab/libteethrt_base.so: stride-2 data
# gradient with 8 pixels per pass = spills, 8-row weight-gradient tiles).
# The base library is the previous commit built the same way: git stash (or checkout) -> python -c "import __graft_entry__ as g; g.build()" -> cp the .so to tools/ab/libteethrt_base.so -> restore and rebuild.
mkdir -p gpurun_out
LIB=multimodal-teeth-restoration-selection_b200/libteethrt.so
cp $LIB /tmp/new.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dwconv or lazy" 2>&1 | tail -3
TEETHRT_DW_WGRAD_TALL=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dwconv" 2>&1 | tail -2
timeout 300 python tools/dw_probe.py 2>&1 | tail -15 > gpurun_out/r02y3_dw_new.jsonl
TEETHRT_DW_WGRAD_TALL=0 timeout 300 python tools/dw_probe.py 2>&1 | tail -15 > gpurun_out/r02y3_dw_new_tall0.jsonl
cp tools/ab/libteethrt_base.so $LIB
timeout 300 python tools/dw_probe.py 2>&1 | tail -15 > gpurun_out/r02y3_dw_base.jsonl
cp /tmp/new.so $LIB
tail -1 gpurun_out/r02y3_dw_new.jsonl; tail -1 gpurun_out/r02y3_dw_new_tall0.jsonl; tail -1 gpurun_out/r02y3_dw_base.jsonl
