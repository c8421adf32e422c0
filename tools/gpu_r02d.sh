#!/bin/bash
# GPU suite (incl. the batched transform), default bench line with the uint8 leg, microbench, ncu of the small kernels.
mkdir -p gpurun_out
T=${TAG:-r02d}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=12 TMO=1200 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
TAILN=1 CUT=9000 TMO=600 run bench python bench.py --steps 20 --warmup 3
TAILN=40 CUT=700 TMO=600 run micro python tools/microbench.py
timeout 200 python tools/prof_small.py > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_small_kernels python tools/prof_small.py > gpurun_out/${T}_ncu_small.log 2>&1
echo "ncu small exit=$?"; tail -2 gpurun_out/${T}_ncu_small.log
