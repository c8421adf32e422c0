#!/bin/bash
# one box: launch-cost probe under the carve-out / smem-floor / PDL switches
mkdir -p gpurun_out
for cfg in "PROBE_X=0" "PROBE_PDL=1"; do
  env $cfg timeout 300 python tools/launch_probe.py 2>&1 | tail -1 | tee -a gpurun_out/launch_probe.jsonl
done
