#!/bin/bash
# One gpurun call: check (smoke, gpu tests, bench), ncu evidence, compute-sanitizer, the other BASELINE workloads.
bash tools/gpu_check.sh
ROUND=${ROUND:-r01} bash tools/gpu_profile.sh
bash tools/gpu_sanitize.sh
for w in config2 tumor60x; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-coverage --no-extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json'))
print('$w', d['config']['signals'], 'signals', 'ms', round(d['ms_per_step'],4), 'G/s', round(d['value']/1e9,2), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'verified', d.get('verified'), 'wr frac', round(d['roofline']['frac'],3))"
done
du -sh gpurun_out
