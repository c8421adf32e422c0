#!/bin/bash
# GPU suite (MIL tensor-core path), microbench, ncu --set full of the small kernels (new CLAHE + MIL kernels).
mkdir -p gpurun_out
T=${TAG:-r02f}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=12 TMO=900 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
TAILN=8 CUT=600 TMO=600 run micro python tools/microbench.py
TEETHRT_MIL_TC=0 TAILN=4 CUT=600 TMO=600 run micro_notc python tools/microbench.py --only-mil
timeout 200 python tools/prof_small.py > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_small_kernels python tools/prof_small.py > gpurun_out/${T}_ncu_small.log 2>&1
echo "ncu small exit=$?"; tail -2 gpurun_out/${T}_ncu_small.log
TAILN=1 CUT=600 run infer python bench.py --infer --steps 20
