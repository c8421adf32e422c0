#!/bin/bash
# Full GPU suite + smoke + infer bench + short train bench (what the driver runs at round end, in one call).
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-${CUT:-330}; }
TAILN=15 TMO=900 run gpu_tests python -m pytest tests -m gpu -q --timeout 300 -x
TAILN=3 TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=1 CUT=1500 TMO=300 run infer python bench.py --infer --steps 20
TAILN=1 TMO=300 run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
