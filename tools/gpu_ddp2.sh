#!/bin/bash
# 2 GPUs: configs[4] as written (global batch 512 -> 256 per GPU) through torchrun + NCCL, the cross-rank parameter check,
# the weak-scaling arm, the reference arm launched the same way, and the non-current-device test.
mkdir -p gpurun_out
T=${TAG:-r02}
N=${N:-2}
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-400} "$@" > gpurun_out/${T}_$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-3} gpurun_out/${T}_$name.log | cut -c1-${CUT:-2500}; }
TAILN=1 run ddp$N python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3
TAILN=4 CUT=300 run dev1_test python -m pytest tests/test_configs_gpu.py -q -k non_current_device
