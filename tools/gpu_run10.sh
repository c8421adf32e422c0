#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; }
TAILN=4 run gpu_tests python -m pytest tests -m gpu -q --timeout 600 -x
TAILN=16 run dw_bench python tools/dw_bench.py
TAILN=2 run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
