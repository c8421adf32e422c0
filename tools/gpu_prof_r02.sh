#!/bin/bash
# Round-2 evidence run for profiles/: (1) bench without ncu, (2) ncu launch list of the same command (eager, 2 steps),
# (3) per-op step profile, (4) ncu --set full of the top kernels.  Every ncu pass only after its command exited 0 without ncu.
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { echo "bench failed"; tail -3 gpurun_out/${T}_bench.err; exit 1; }
tail -c 400 gpurun_out/${T}_bench.json
Q="--steps 2 --warmup 3 --no-cpu-baseline --no-infer --no-u8 --sustain-seconds 0"
TEETHRT_NO_GRAPH=1 timeout 300 python bench.py $Q > /dev/null 2>&1 && \
TEETHRT_NO_GRAPH=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py $Q > gpurun_out/${T}_ncu_bench.log 2>&1
echo "launch list exit=$?"
timeout 300 python tools/step_profile.py --log gpurun_out/${T}_step_ops_plain.json > /dev/null 2>&1 && \
TEETHRT_WGRAD_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/${T}_step_launches.csv python tools/step_profile.py --log gpurun_out/${T}_step_ops.json > gpurun_out/${T}_stepprof.log 2>&1
echo "step profile exit=$?"
timeout 200 python tools/prof_top.py > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_top_kernels python tools/prof_top.py > gpurun_out/${T}_ncu_top.log 2>&1
echo "ncu full exit=$?"; tail -1 gpurun_out/${T}_ncu_top.log
