#!/bin/bash
# Round-2 first evidence run: full GPU suite, smoke, default bench line, and the torch/cuDNN GPU baseline the survey names.
mkdir -p gpurun_out
T=${TAG:-r02a}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/${T}_$name.log | cut -c1-${CUT:-400}; }
TAILN=25 TMO=1200 run gpu_tests python -m pytest tests -m gpu -q --timeout 600
TAILN=3 TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=1 CUT=6000 TMO=600 run bench python bench.py --steps 20 --warmup 3
TAILN=8 CUT=600 TMO=400 run torch_eager python tools/gpu_torch_baseline.py --arms eager_amp,eager_amp_cl,infer_b1 --out gpurun_out/${T}_torch_eager.jsonl
TAILN=4 CUT=600 TMO=600 run torch_compile python tools/gpu_torch_baseline.py --arms compile_cl,compile_cl_cg --out gpurun_out/${T}_torch_compile.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
