#!/bin/bash
# GPU suite + A/B of the lazy BatchNorm variants on the train step.
mkdir -p gpurun_out
T=${TAG:-r02c}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=25 TMO=1200 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
Q="--steps 30 --warmup 3 --no-infer --no-cpu-baseline --sustain-seconds 0"
TAILN=1 CUT=330 run ab_default python bench.py $Q
TEETHRT_LAZY_BN=0 TAILN=1 CUT=330 run ab_lazy0 python bench.py $Q
TEETHRT_LAZY_BN=1 TAILN=1 CUT=330 run ab_lazy1 python bench.py $Q
TEETHRT_LAZY_BN=3 TAILN=1 CUT=330 run ab_lazy3 python bench.py $Q
TEETHRT_LAZY_BN=4 TAILN=1 CUT=330 run ab_lazy4 python bench.py $Q
TEETHRT_LAZY_BN=0 TAILN=1 CUT=330 run ab_lazy0_again python bench.py $Q
TAILN=1 CUT=330 run ab_default_again python bench.py $Q
