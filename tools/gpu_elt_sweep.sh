#!/bin/bash
for v in 1 2; do echo "POOL_BLOCKS=$v"; TEETHRT_POOL_BLOCKS=$v ELT_ONLY=pool_act timeout 200 python tools/elt_probe.py 2>&1 | tail -1; done
for v in 1 2 4; do echo "SEBR_BLOCKS=$v"; TEETHRT_SEBR_BLOCKS=$v ELT_ONLY=se_bwd_reduce timeout 200 python tools/elt_probe.py 2>&1 | tail -12; done
