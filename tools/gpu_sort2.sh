#!/bin/bash
# sort generation A/B: parity tests with the second-generation sort (default), then a device bench of both
mkdir -p gpurun_out
echo "== segsort tests (v2)"; timeout 900 python -m pytest tests/test_gpu_segsort.py -m gpu -q -x --timeout 600 2>&1 | tail -15
echo "== cluster + aggregate tests (v2)"; timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_aggregate.py -m gpu -q -x --timeout 600 2>&1 | tail -15
for gen in 2 1; do
  [ $gen = 1 ] && export TDT_SEGSORT_V1=1 || unset TDT_SEGSORT_V1
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/sort_gen$gen.json 2> gpurun_out/sort_gen$gen.err
  echo "gen $gen rc=$?"; tail -3 gpurun_out/sort_gen$gen.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sort_gen$gen.json"))
    print("gen$gen ms_per_step=%.4f e2e=%.3f verified=%s launches=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified"), d["gpu_launches_per_step"]), d["roofline"]["stages_ms"])
except Exception as e:
    print("gen$gen: no result", e)
PY
done
unset TDT_SEGSORT_V1
TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py > gpurun_out/kernel_times_sort2.txt 2>&1; head -70 gpurun_out/kernel_times_sort2.txt
