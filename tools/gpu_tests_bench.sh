#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-330; }
TAILN=15 run gpu_tests python -m pytest tests -m gpu -q --timeout 300 -x
TAILN=1 run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
