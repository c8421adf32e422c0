#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/prof_next_rows.py > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/next_rows python tools/prof_next_rows.py > gpurun_out/ncu_next.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/ncu_next.log
