#!/bin/bash
mkdir -p gpurun_out
echo "== segsort + cluster tests"; timeout 900 python -m pytest tests/test_gpu_segsort.py tests/test_gpu_cluster.py -m gpu -q -x --timeout 300 -k "auto or cluster" 2>&1 | tail -2
run() {  # name, args...
  name=$1; shift
  timeout 600 python bench.py "$@" --no-cpu --no-extra > gpurun_out/t3_$name.json 2> gpurun_out/t3_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/t3_$name.json"))
    print("$name ms_per_step=%.4f e2e=%.3f verified=%s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("verified")), d["roofline"]["stages_ms"])
except Exception as e:
    print("$name: no result", e)
PY
}
run default30 --steps 20 --warmup 3
run small30 --steps 20 --warmup 3 --signals 2500000
run tumor --workload tumor60x --steps 10 --warmup 3
for v in t256 t512; do
  export TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_$v.so
  run tumor_$v --workload tumor60x --steps 10 --warmup 3
  run default30_$v --steps 20 --warmup 3
done
unset TDT_B200_LIB
