#!/bin/bash
# generation-3 sort, fourth session: by-value mode (unordered passes), full parity, serial kernel times, tiny variants
mkdir -p gpurun_out
echo "== segsort tests (all generations)"; timeout 900 python -m pytest tests/test_gpu_segsort.py -m gpu -q -x --timeout 300 2>&1 | tail -6
echo "== cluster + aggregate"; timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_aggregate.py -m gpu -q -x --timeout 600 2>&1 | tail -3
echo "== debug counters"; TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/sort_target.py 2>&1 | tail -3
echo "== serialised kernel times"; TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py > gpurun_out/kernel_times_sort3_serial.txt 2>&1; head -34 gpurun_out/kernel_times_sort3_serial.txt
for name in ${VARIANTS:-default}; do
  lib=tiddit_b200/_variants/libtdt_b200_$name.so
  [ "$name" = default ] && lib=tiddit_b200/libtdt_b200.so
  export TDT_B200_LIB=$PWD/$lib
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/sort3_$name.json 2> gpurun_out/sort3_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sort3_$name.json"))
    print("$name ms_per_step=%.4f e2e=%.3f verified=%s launches=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified"), d["gpu_launches_per_step"]), d["roofline"]["stages_ms"])
except Exception as e:
    print("$name: no result", e)
PY
done
unset TDT_B200_LIB
