#!/bin/bash
# Round-2 final evidence: full GPU test suite, bench WITHOUT ncu, then the per-op step profile and the --set full capture of
# the top kernels (each ncu pass only after its command exited 0 without ncu).  The launch list of bench.py itself is taken
# by tools/gpu_prof_r02.sh (20 minutes under ncu); it is not repeated here.
mkdir -p gpurun_out
T=${TAG:-r02g}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${T}_gpu_tests.log; cat gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { echo "bench failed"; tail -3 gpurun_out/${T}_bench.err; exit 1; }
tail -c 300 gpurun_out/${T}_bench.json
timeout 300 python tools/step_profile.py --log gpurun_out/${T}_step_ops_plain.json > /dev/null 2>&1 && \
TEETHRT_WGRAD_STREAM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/${T}_step_launches.csv python tools/step_profile.py --log gpurun_out/${T}_step_ops.json > gpurun_out/${T}_stepprof.log 2>&1
echo "step profile exit=$?"
timeout 200 python tools/prof_top.py > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_top_kernels python tools/prof_top.py > gpurun_out/${T}_ncu_top.log 2>&1
echo "ncu full exit=$?"; tail -1 gpurun_out/${T}_ncu_top.log
timeout 300 python tools/knockout.py --out gpurun_out/${T}_knockout.jsonl 2>&1 | tail -18
timeout 300 python tools/dw_probe.py 2>&1 | tail -15 > gpurun_out/${T}_dw_probe.jsonl; tail -1 gpurun_out/${T}_dw_probe.jsonl
timeout 300 python tools/gemm_probe.py 2>&1 | tail -23 > gpurun_out/${T}_gemm_probe.jsonl; tail -1 gpurun_out/${T}_gemm_probe.jsonl
timeout 300 python tools/knockout.py --infer 1 --out gpurun_out/${T}_knockout_infer1.jsonl 2>&1 | tail -9
