#!/bin/bash
# ncu evidence: (1) launch list of a short bench run, (2) full captures of our top kernels.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-coverage > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'window_runs|coverage_kernel|gc_small|final_labels|pack_keys' \
    -c 14 -o gpurun_out/prof_${R} -f python tools/profile_target.py > gpurun_out/prof.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/prof.log
ls -la gpurun_out
