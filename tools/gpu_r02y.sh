#!/bin/bash
# Round-2 late A/B: 1024-thread one-block tab MLP kernels + warp-parallel entry search in pack_w_batch_kernel (new library)
# against the previous library (tools/ab/libteethrt_base.so, built from the previous commit), interleaved on one box.
# The base library is the previous commit built the same way: git stash (or checkout) -> python -c "import __graft_entry__ as g; g.build()" -> cp the .so to tools/ab/libteethrt_base.so -> restore and rebuild.
mkdir -p gpurun_out
LIB=multimodal-teeth-restoration-selection_b200/libteethrt.so
cp $LIB /tmp/new.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "pack or tab or heads" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2 3; do
for which in new base; do
  if [ $which = new ]; then cp /tmp/new.so $LIB; else cp tools/ab/libteethrt_base.so $LIB; fi
  timeout 300 python bench.py --steps 40 --warmup 5 --no-infer --no-u8 --no-cpu-baseline --sustain-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02y_tmp.log
  python - "$which" gpurun_out/r02y_tmp.log <<'PY' | tee -a gpurun_out/r02y_ab.jsonl
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); print(json.dumps({"lib": sys.argv[1], "ms_per_step": round(d['ms_per_step'],3), "images_per_s": round(d['value'],1)}))
PY
done; done
cp /tmp/new.so $LIB
