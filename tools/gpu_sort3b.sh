#!/bin/bash
# generation-3 sort, second session: parity, serialised per-kernel times, variants A/B, ncu capture of the new kernels
mkdir -p gpurun_out
echo "== segsort tests (msd)"; timeout 600 python -m pytest tests/test_gpu_segsort.py tests/test_gpu_cluster.py -m gpu -q -x --timeout 300 -k "msd or cluster" 2>&1 | tail -4
echo "== serialised kernel times"; TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py > gpurun_out/kernel_times_sort3_serial.txt 2>&1; head -46 gpurun_out/kernel_times_sort3_serial.txt
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'m3_finish|m3_pass|m3_hist' -s 12 -c 6 -o gpurun_out/src_m3_${R:-r02_v4} -f python tools/sort_target.py > gpurun_out/src_m3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/src_m3.log; ls -la gpurun_out/src_m3_*.ncu-rep
