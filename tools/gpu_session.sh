#!/bin/bash
# One gpurun call: check (smoke, gpu tests, bench), ncu evidence, compute-sanitizer, e2e chunk sweep.
bash tools/gpu_check.sh
ROUND=${ROUND:-r01} bash tools/gpu_profile.sh
bash tools/gpu_sanitize.sh
for k in 4 8 12; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-coverage --no-extra --chunks $k > gpurun_out/chunks_$k.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/chunks_$k.json')); print('chunks $k e2e ms', d['e2e']['ms_per_step'])"
done
du -sh gpurun_out
