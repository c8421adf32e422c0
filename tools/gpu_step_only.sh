#!/bin/bash
# per-op step profile only (ncu launch durations of ONE eager step, joined with the op log)
mkdir -p gpurun_out
T=${TAG:-r02n}
timeout 300 python tools/step_profile.py --log gpurun_out/${T}_step_ops_plain.json > /dev/null 2>&1 && \
TEETHRT_WGRAD_STREAM=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/${T}_step_launches.csv python tools/step_profile.py --log gpurun_out/${T}_step_ops.json > gpurun_out/${T}_stepprof.log 2>&1
echo "step profile exit=$?"
python tools/join_profile.py gpurun_out/${T}_step_ops.json gpurun_out/${T}_step_launches.csv --top 400 > gpurun_out/${T}_step_per_op.txt 2>&1
head -36 gpurun_out/${T}_step_per_op.txt
