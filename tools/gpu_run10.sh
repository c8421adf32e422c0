#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
TAILN=25 run gpu_tests python -m pytest tests -m gpu -q --timeout 600 -x
TAILN=2 run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
