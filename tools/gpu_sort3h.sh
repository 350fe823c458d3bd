#!/bin/bash
mkdir -p gpurun_out
echo "== segsort tests (auto, msd)"; timeout 900 python -m pytest tests/test_gpu_segsort.py -m gpu -q -x --timeout 300 -k "auto or msd" 2>&1 | tail -3
echo "== cluster + aggregate"; timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_aggregate.py -m gpu -q -x --timeout 600 2>&1 | tail -2
echo "== serialised kernel times"; TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py > gpurun_out/kernel_times_sort3_serial.txt 2>&1; head -20 gpurun_out/kernel_times_sort3_serial.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/sort3_default.json 2> gpurun_out/sort3_default.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sort3_default.json"))
    print("default ms_per_step=%.4f e2e=%.3f verified=%s launches=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified"), d["gpu_launches_per_step"]), d["roofline"]["stages_ms"])
except Exception as e:
    print("default: no result", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'m3_finish' -s 3 -c 1 -o gpurun_out/src_m3fin_${R:-r02_v5} -f python tools/sort_target.py > gpurun_out/src_m3fin.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/src_m3fin_*
