#!/bin/bash
# ncu durations of the tab/heads kernels inside one eager train step (cold-cache, serialised).
mkdir -p gpurun_out
timeout 300 python tools/step_profile.py --log gpurun_out/h_ops.json > /dev/null 2>&1 && \
timeout 300 ncu --kernel-name regex:heads --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/heads_launches.csv python tools/step_profile.py --log gpurun_out/h_ops.json > gpurun_out/heads_prof.log 2>&1
echo "exit=$?"; grep -v "^==" gpurun_out/heads_launches.csv | cut -d, -f5,13- | cut -c1-160
timeout 300 python -m pytest tests -m gpu -q --timeout 300 -k "ragged" 2>&1 | tail -2
