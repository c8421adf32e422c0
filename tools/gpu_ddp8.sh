#!/bin/bash
# 8-GPU data-parallel bench (torchrun, NCCL over NVSwitch) + the reference arm launched the same way.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-700; }
N=${N:-8}
TAILN=1 TMO=400 run bench$N python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline
