#!/bin/bash
# generation-3 sort (MSD rounds + shared-memory finish): parity, then A/B against the LSD chain
mkdir -p gpurun_out
echo "== segsort tests (msd)"; timeout 900 python -m pytest tests/test_gpu_segsort.py -m gpu -q -x --timeout 300 -k "msd" 2>&1 | tail -15
echo "== segsort tests (lsd, samplesort)"; timeout 900 python -m pytest tests/test_gpu_segsort.py -m gpu -q --timeout 300 -k "not msd" 2>&1 | tail -8
echo "== cluster + aggregate tests (msd default)"; timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_aggregate.py -m gpu -q -x --timeout 600 2>&1 | tail -8
for gen in msd lsd; do
  export TDT_SEGSORT=$gen
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/sort3_$gen.json 2> gpurun_out/sort3_$gen.err
  echo "gen $gen rc=$?"; tail -3 gpurun_out/sort3_$gen.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sort3_$gen.json"))
    print("$gen ms_per_step=%.4f e2e=%.3f verified=%s launches=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified"), d["gpu_launches_per_step"]), d["roofline"]["stages_ms"])
except Exception as e:
    print("$gen: no result", e)
PY
done
unset TDT_SEGSORT
TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py > gpurun_out/kernel_times_sort3.txt 2>&1; head -90 gpurun_out/kernel_times_sort3.txt
