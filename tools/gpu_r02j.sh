#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02j}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=10 TMO=900 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
TAILN=1 CUT=900 run infer python bench.py --infer --steps 20
TAILN=1 CUT=900 run infer2 python bench.py --infer --steps 20
