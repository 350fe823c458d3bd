#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run, (2) per-kernel metrics of ONE clustering pass,
# (3) full captures of our top kernels.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
SKIP=58 COUNT=29 bash tools/gpu_launches.sh
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'window_runs_small|coverage_kernel|gc_small|segsort_pass|segsort_local|final_labels|pack_y' \
    -s 29 -c 22 -o gpurun_out/prof_${R} -f python tools/profile_target.py > gpurun_out/prof.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/prof.log
