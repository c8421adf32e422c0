#!/bin/bash
# grid-size multiplier sweeps of the per-image streaming kernels (isolated, HBM-cold, inside a CUDA graph)
for v in 3 4 5 8; do echo "POOL_BLOCKS=$v"; TEETHRT_POOL_BLOCKS=$v ELT_ONLY=pool_act timeout 200 python tools/elt_probe.py 2>&1 | tail -1; done
for v in 2 4 6 8 12; do echo "GATE_BLOCKS=$v"; TEETHRT_GATE_BLOCKS=$v ELT_ONLY=gate_apply timeout 200 python tools/elt_probe.py 2>&1 | tail -1; done
for v in 2 4 6 8 12; do echo "ABA_BLOCKS=$v"; TEETHRT_ABA_BLOCKS=$v ELT_ONLY=act_bwd_apply timeout 200 python tools/elt_probe.py 2>&1 | tail -1; done
for v in 4 5; do echo "SEBR_BLOCKS=$v"; TEETHRT_SEBR_BLOCKS=$v ELT_ONLY=se_bwd_reduce timeout 200 python tools/elt_probe.py 2>&1 | tail -1; done
