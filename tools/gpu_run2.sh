#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-30} gpurun_out/$name.log; }
run small python -m pytest tests/test_kernels_gpu.py -q -k "mil or tab or dropout or adamw" --timeout 120
TAILN=80 run models python -m pytest tests/test_models_gpu.py -q --timeout 600
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --steps 10 --warmup 3
run infer python bench.py --infer --steps 20
