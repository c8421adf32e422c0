#!/bin/bash
mkdir -p gpurun_out
python tools/dw_bench.py > gpurun_out/dw_bench.log 2>&1; echo "dw_bench exit=$?"; cat gpurun_out/dw_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 2 -c 3 -f -o gpurun_out/prof_dw_r1 python tools/dw_bench.py --only 3 --reps 1 > gpurun_out/prof_dw_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/prof_dw_ncu.log
