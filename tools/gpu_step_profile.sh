#!/bin/bash
mkdir -p gpurun_out
TEETHRT_WGRAD_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/step_launches.csv python tools/step_profile.py --log gpurun_out/step_ops.json > gpurun_out/stepprof_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/stepprof_ncu.log
