#!/bin/bash
# GPU suite + A/B of the lazy BatchNorm records and the merged SE/BatchNorm backward on the train step.
mkdir -p gpurun_out
T=${TAG:-r02b}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=25 TMO=1200 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
Q="--steps 30 --warmup 3 --no-infer --no-cpu-baseline --sustain-seconds 0"
TAILN=1 CUT=330 run ab_default python bench.py $Q
TEETHRT_LAZY_BN=0 TAILN=1 CUT=330 run ab_nolazy python bench.py $Q
TEETHRT_SE_BWD_MERGED=0 TAILN=1 CUT=330 run ab_nomerge python bench.py $Q
TEETHRT_LAZY_BN=0 TEETHRT_SE_BWD_MERGED=0 TAILN=1 CUT=330 run ab_r01path python bench.py $Q
TAILN=3 TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=1 CUT=8000 TMO=600 run bench python bench.py --steps 20 --warmup 3
