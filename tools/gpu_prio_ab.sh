#!/bin/bash
# A/B on one box: main chain captured on a default- vs high-priority stream.
mkdir -p gpurun_out
for P in 0 -1 0 -1; do
  TEETHRT_MAIN_PRIORITY=$P timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/prio_$P.log 2>&1
  echo "prio=$P exit=$? $(tail -1 gpurun_out/prio_$P.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1))")"
done
