#!/bin/bash
# 2-GPU sanity: the preprocessing tests on one GPU, then the data-parallel bench at N=2 (torchrun, NCCL over NVLink).
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
TAILN=8 TMO=300 run pre_tests python -m pytest tests/test_preproc_gpu.py tests/test_stack_gpu.py -m gpu -q --timeout 200
TAILN=1 TMO=400 run bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline
TAILN=1 TMO=300 run ref2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1
