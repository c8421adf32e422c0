#!/bin/bash
# Round-2 closing evidence on the final code: full GPU test suite, bench WITHOUT ncu, the per-op step profile and the
# --set full capture of the top kernels (each ncu pass only after its command exited 0 without ncu).
mkdir -p gpurun_out
T=${TAG:-r02z}
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${T}_gpu_tests.log; cat gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { echo "bench failed"; tail -3 gpurun_out/${T}_bench.err; exit 1; }
tail -c 300 gpurun_out/${T}_bench.json
timeout 300 python tools/step_profile.py --log gpurun_out/${T}_step_ops_plain.json > /dev/null 2>&1 && \
TEETHRT_WGRAD_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/${T}_step_launches.csv python tools/step_profile.py --log gpurun_out/${T}_step_ops.json > gpurun_out/${T}_stepprof.log 2>&1
echo "step profile exit=$?"
timeout 200 python tools/prof_top.py > gpurun_out/${T}_prof_top_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_top_kernels python tools/prof_top.py > gpurun_out/${T}_ncu_top.log 2>&1
echo "ncu full exit=$?"; tail -1 gpurun_out/${T}_ncu_top.log; tail -2 gpurun_out/${T}_prof_top_plain.log
