#!/bin/bash
mkdir -p gpurun_out
echo "== segsort + cluster tests"; timeout 900 python -m pytest tests/test_gpu_segsort.py tests/test_gpu_cluster.py -m gpu -q -x --timeout 300 -k "auto or msd or cluster" 2>&1 | tail -2
echo "== debug counters (tumor)"; TDT_KT_WORKLOAD=tumor TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/kernel_times.py 50000000 2>&1 | grep "^m3" | head -2
for gen in auto lsd; do
  [ $gen = lsd ] && export TDT_SEGSORT=lsd || unset TDT_SEGSORT
  echo "== tumor serial kernel times ($gen)"; TDT_KT_WORKLOAD=tumor TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py 50000000 2>&1 | head -22
done
unset TDT_SEGSORT
echo "== 30X serial kernel times"; TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py 2>&1 | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/sort3_default.json 2> gpurun_out/sort3_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/sort3_default.json"))
print("default ms_per_step=%.4f e2e=%.3f verified=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified")), d["roofline"]["stages_ms"])
PY
