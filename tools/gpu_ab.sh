#!/bin/bash
# A/B of compile-time variants on the GPU box: VARIANTS="name1:-DX=1 name2:-DX=0,-DY=2" bash tools/gpu_ab.sh
# each variant: forced rebuild, the sort + cluster parity tests, a short device bench; stage times to gpurun_out/ab_<name>.json
mkdir -p gpurun_out
for v in ${VARIANTS:-base:}; do
  name=${v%%:*}; defs=${v#*:}; defs=${defs//,/ }
  echo "== variant $name  defs: $defs"
  TDT_NVCC_DEFS="$defs" python -m tiddit_b200.build --force > gpurun_out/ab_build_$name.log 2>&1 || { echo build failed; tail -5 gpurun_out/ab_build_$name.log; continue; }
  timeout 600 python -m pytest tests/test_gpu_segsort.py tests/test_gpu_cluster.py -m gpu -q -x --timeout 300 2>&1 | tail -2
  timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-coverage --no-extra ${BENCH_ARGS} > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$name.json"))
    print("$name", "ms_per_step=%.4f e2e_ms=%.3f verified=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified")), d["roofline"]["stages_ms"])
except Exception as e:
    print("$name: no result", e)
PY
done
# leave the default build in place
python -m tiddit_b200.build --force > /dev/null 2>&1
