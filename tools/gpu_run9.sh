#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 3 -c 4 -f -o gpurun_out/prof_dw_r1c python tools/dw_bench.py --only 3 --reps 1 > gpurun_out/prof_dw_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/prof_dw_ncu.log
