#!/bin/bash
mkdir -p gpurun_out
echo "== segsort tests (auto, msd)"; timeout 900 python -m pytest tests/test_gpu_segsort.py -m gpu -q -x --timeout 300 -k "auto or msd" 2>&1 | tail -3
echo "== cluster + aggregate"; timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_aggregate.py -m gpu -q -x --timeout 600 2>&1 | tail -2
echo "== debug counters"; TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/sort_target.py 2>&1 | tail -3
echo "== serialised kernel times"; TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py > gpurun_out/kernel_times_sort3_serial.txt 2>&1; head -20 gpurun_out/kernel_times_sort3_serial.txt
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/sort3_default.json 2> gpurun_out/sort3_default.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sort3_default.json"))
    print("default ms_per_step=%.4f e2e=%.3f verified=%s launches=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified"), d["gpu_launches_per_step"]), d["roofline"]["stages_ms"])
except Exception as e:
    print("default: no result", e)
PY
done
