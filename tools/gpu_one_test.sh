#!/bin/bash
# Run the tests selected by $K (pytest -k expression) on the GPU box, then (BENCH=1) a short bench.
mkdir -p gpurun_out
timeout ${TMO:-600} python -m pytest tests -m gpu -q --timeout 300 -k "$K" > gpurun_out/one_test.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-40} gpurun_out/one_test.log | cut -c1-300
if [ -n "$BENCH" ]; then timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-260; fi
