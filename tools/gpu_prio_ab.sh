#!/bin/bash
# A/B on one box: wgrad split multiplier (CTA granularity of the side-stream weight-gradient GEMMs) and side stream on/off.
mkdir -p gpurun_out
one() { timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/ab.log 2>&1; echo "$1 exit=$? $(tail -1 gpurun_out/ab.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1))")"; }
TEETHRT_WGRAD_SPLIT_MULT=2 one mult2
TEETHRT_WGRAD_SPLIT_MULT=4 one mult4
TEETHRT_WGRAD_SPLIT_MULT=8 one mult8
TEETHRT_WGRAD_SPLIT_MULT=1 one mult1
TEETHRT_WGRAD_STREAM=0 one no_side_stream
