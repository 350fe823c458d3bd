#!/bin/bash
# generation-3 sort, third session: finish-kernel debug counters, ncu capture of the finish kernels, tiny-kernel variants
mkdir -p gpurun_out
echo "== debug counters"; TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/sort_target.py 2>&1 | tail -12
for name in ${VARIANTS:-default}; do
  lib=tiddit_b200/_variants/libtdt_b200_$name.so
  [ "$name" = default ] && lib=tiddit_b200/libtdt_b200.so
  export TDT_B200_LIB=$PWD/$lib
  timeout 300 python -m pytest tests/test_gpu_segsort.py -m gpu -q -x --timeout 300 -k "msd" 2>&1 | tail -1
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'m3_finish' -s 4 -c 2 -o gpurun_out/src_m3fin_${R:-r02_v4} -f python tools/sort_target.py > gpurun_out/src_m3fin.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/src_m3fin.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'m3_pass' -s 8 -c 1 -o gpurun_out/src_m3pass_${R:-r02_v4} -f python tools/sort_target.py > gpurun_out/src_m3pass.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/src_m3pass.log
