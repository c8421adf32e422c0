#!/bin/bash
# Batch-1 inference: programmatic dependent launch A/B, launch list of one forward.
mkdir -p gpurun_out
T=${TAG:-r02e}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=8 TMO=900 run gpu_tests python -m pytest tests -m gpu -q --timeout 600 -x
TEETHRT_PDL=1 TAILN=1 CUT=700 run infer_pdl1 python bench.py --infer --steps 20
TEETHRT_PDL=0 TAILN=1 CUT=700 run infer_pdl0 python bench.py --infer --steps 20
TEETHRT_PDL=1 TAILN=1 CUT=700 run infer_pdl1b python bench.py --infer --steps 20
Q="--steps 30 --warmup 3 --no-infer --no-cpu-baseline --sustain-seconds 0 --no-u8"
TEETHRT_PDL=1 TAILN=1 CUT=330 run train_pdl1 python bench.py $Q
TEETHRT_PDL=0 TAILN=1 CUT=330 run train_pdl0 python bench.py $Q
timeout 120 python tools/prof_infer.py > /dev/null 2>&1 && \
TEETHRT_PDL=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_infer_launches.csv python tools/prof_infer.py > gpurun_out/${T}_ncu_infer.log 2>&1
echo "infer launch list exit=$?"
