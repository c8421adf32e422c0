#!/bin/bash
# Short evidence refresh: bench (no ncu), then the ncu launch list of the same command and the per-op step profile.
mkdir -p gpurun_out
T=${TAG:-r01}
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { echo "bench failed"; tail -3 gpurun_out/${T}_bench.err; exit 1; }
tail -c 300 gpurun_out/${T}_bench.json
TEETHRT_NO_GRAPH=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
TEETHRT_NO_GRAPH=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 1900 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
echo "launch list exit=$?"
timeout 300 python tools/step_profile.py --log gpurun_out/${T}_step_ops_plain.json > /dev/null 2>&1 && \
TEETHRT_WGRAD_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/${T}_step_launches.csv python tools/step_profile.py --log gpurun_out/${T}_step_ops.json > gpurun_out/${T}_stepprof.log 2>&1
echo "step profile exit=$?"
