#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_segsort.py tests/test_gpu_cluster.py -m gpu -q -x --timeout 600 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-coverage --no-tumor > gpurun_out/agg2.json 2> gpurun_out/agg2.err; echo "bench rc=$?"; tail -3 gpurun_out/agg2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/agg2.json"))
print("step", d["ms_per_step"], d["verified"], d["roofline"]["stages_ms"])
a = d["aggregate"]; print("aggregate", a["ms_per_step"], a["verified"], a["stages_ms"], a["e2e"])
print("cluster_main", d["cluster_main"])
print("gc", d["gc"]["ms_per_step"], d["gc"]["roofline"]["frac"])
PY
