#!/bin/bash
# A/B of compile-time variants on the GPU box.  Variants are pre-built by tools/build_variants.sh into
# tiddit_b200/_variants/ (the box only runs them): VARIANTS="name1 name2" bash tools/gpu_ab.sh
# each variant: the sort + cluster parity tests, a short device bench; stage times to gpurun_out/ab_<name>.json
mkdir -p gpurun_out
for name in ${VARIANTS:-default}; do
  lib=tiddit_b200/_variants/libtdt_b200_$name.so
  [ "$name" = default ] && lib=tiddit_b200/libtdt_b200.so
  echo "== variant $name ($lib)"
  export TDT_B200_LIB=$PWD/$lib
  timeout 600 python -m pytest tests/test_gpu_segsort.py tests/test_gpu_cluster.py -m gpu -q -x --timeout 300 2>&1 | tail -2
  timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu --no-coverage --no-extra ${BENCH_ARGS} > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$name.json"))
    print("$name", "ms_per_step=%.4f e2e_ms=%.3f verified=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified")), d["roofline"]["stages_ms"])
except Exception as e:
    print("$name: no result", e)
PY
done
unset TDT_B200_LIB
