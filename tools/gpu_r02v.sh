#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for cfg in "TEETHRT_PDL=1" "TEETHRT_PDL=2"; do
  env $cfg timeout 300 python bench.py --steps 40 --warmup 5 --no-infer --no-u8 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02v_tmp.log
  python - "$cfg" gpurun_out/r02v_tmp.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['ms_per_step'],3), round(d['value'],1))
PY
done; done
