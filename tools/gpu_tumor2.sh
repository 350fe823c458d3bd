#!/bin/bash
mkdir -p gpurun_out
echo "== segsort + cluster tests"; timeout 900 python -m pytest tests/test_gpu_segsort.py tests/test_gpu_cluster.py -m gpu -q -x --timeout 300 -k "auto or msd or cluster" 2>&1 | tail -2
echo "== debug counters (tumor)"; TDT_KT_WORKLOAD=tumor TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/kernel_times.py 50000000 2>&1 | grep "^m3" | head -2
echo "== debug counters (30X)"; TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_dbg.so TDT_M3_SERIAL=1 timeout 300 python tools/sort_target.py 2>&1 | grep "^m3" | head -2
for name in default hot32 hot8; do
  lib=tiddit_b200/_variants/libtdt_b200_$name.so
  [ "$name" = default ] && lib=tiddit_b200/libtdt_b200.so
  export TDT_B200_LIB=$PWD/$lib
  echo "== $name: tumor serial kernel times"; TDT_KT_WORKLOAD=tumor TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py 50000000 2>&1 | sed -n '4,7p;16,18p'
  echo "== $name: 30X serial"; TDT_M3_SERIAL=1 TDT_PROF_DETAIL=1 timeout 600 python tools/kernel_times.py 2>&1 | sed -n '4,7p;16,18p'
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/sort3_$name.json 2> gpurun_out/sort3_$name.err
  timeout 600 python bench.py --workload tumor60x --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/sort3t_$name.json 2> gpurun_out/sort3t_$name.err
  python - <<PY
import json
for f in ("sort3_$name", "sort3t_$name"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms_per_step=%.4f e2e=%.3f verified=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified")), d["roofline"]["stages_ms"])
    except Exception as e:
        print(f, "no result", e)
PY
done
unset TDT_B200_LIB
