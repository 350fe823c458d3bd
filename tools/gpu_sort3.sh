#!/bin/bash
# generation-3 sort (tdt_segsort3.cuh): parity of every generation, serialised per-kernel times (TDT_M3_SERIAL=1 puts the
# later rounds on the caller's stream so that event times do not overlap), debug counters of the finish kernel
# (-DTDT_M3_DEBUG variant), A/B of pre-built variants (tools/build_variants.sh), optional ncu capture of the sort kernels.
#   VARIANTS="default hot16 ..." WORKLOADS="wgs30x tumor60x" NCU=1 bash tools/gpu_sort3.sh
mkdir -p gpurun_out
echo "== segsort tests (all generations)"; timeout 900 python -m pytest tests/test_gpu_segsort.py -m gpu -q -x --timeout 300 2>&1 | tail -3
echo "== cluster + aggregate"; timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_aggregate.py -m gpu -q -x --timeout 600 2>&1 | tail -2
if [ -f tiddit_b200/_variants/libtdt_b200_dbg.so ]; then
  echo "== finish-kernel counters (30X, tumour)"
  TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/sort_target.py 2>&1 | grep "^m3" | head -2
  TDT_KT_WORKLOAD=tumor TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/kernel_times.py 50000000 2>&1 | grep "^m3" | head -2
fi
echo "== serialised kernel times (30X)"; TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py > gpurun_out/kernel_times_sort3_serial.txt 2>&1; head -40 gpurun_out/kernel_times_sort3_serial.txt
for name in ${VARIANTS:-default}; do
  lib=tiddit_b200/_variants/libtdt_b200_$name.so
  [ "$name" = default ] && lib=tiddit_b200/libtdt_b200.so
  export TDT_B200_LIB=$PWD/$lib
  for wl in ${WORKLOADS:-wgs30x}; do
    timeout 600 python bench.py --workload $wl --steps ${STEPS:-20} --warmup 3 --no-cpu --no-extra > gpurun_out/sort3_${name}_$wl.json 2> gpurun_out/sort3_${name}_$wl.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sort3_${name}_$wl.json"))
    print("$name $wl ms_per_step=%.4f e2e=%.3f verified=%s launches=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified"), d["gpu_launches_per_step"]), d["roofline"]["stages_ms"])
except Exception as e:
    print("$name $wl: no result", e)
PY
  done
done
unset TDT_B200_LIB
if [ "${NCU:-0}" = 1 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'m3_finish|m3_pass|m3_hist|m3_plan' -s 13 -c 13 -o gpurun_out/src_m3_${R:-r02_v4} -f python tools/sort_target.py > gpurun_out/src_m3.log 2>&1
  echo "ncu rc=$?"; ls -la gpurun_out/src_m3_*.ncu-rep
  ncu -i gpurun_out/src_m3_${R:-r02_v4}.ncu-rep --page raw --csv > gpurun_out/src_m3_${R:-r02_v4}_raw.csv 2>/dev/null
fi
