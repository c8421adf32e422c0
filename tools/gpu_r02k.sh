#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02k}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=10 TMO=900 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
Q="--steps 40 --warmup 3 --no-infer --no-cpu-baseline --sustain-seconds 0 --no-u8"
for i in 1 2; do
TAILN=1 CUT=330 run ab_default_$i python bench.py $Q
TEETHRT_SE_REDUCE_V8=1 TAILN=1 CUT=330 run ab_v8_$i python bench.py $Q
done
for i in 1 2; do
TAILN=1 CUT=420 run infer_default_$i python bench.py --infer --steps 20
TEETHRT_SE_APPLY=0 TAILN=1 CUT=420 run infer_noapply_$i python bench.py --infer --steps 20
done
