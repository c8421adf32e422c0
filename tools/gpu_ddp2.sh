#!/bin/bash
# 2-GPU A/B of the NCCL stream priority with the data-parallel bench (torchrun, NCCL over NVLink).
mkdir -p gpurun_out
one() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/ddp2_$1.log 2>&1; echo "$1 exit=$? $(tail -1 gpurun_out/ddp2_$1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1))")"; }
TEETHRT_NCCL_HIPRIO=0 one nccl_default 29531
TEETHRT_NCCL_HIPRIO=1 one nccl_high 29532
timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1gpu', round(d['ms_per_step'],3), round(d['value'],1))"
