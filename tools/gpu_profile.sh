#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run, (2) per-kernel metrics of ONE clustering pass and ONE
# aggregation pass, (3) full-set metrics of our kernels exported as CSV on the box (the report itself would exceed the
# 64 MiB gpurun_out/ limit), (4) a small --import-source capture of the top kernels, (5) in-stream per-kernel times.
mkdir -p gpurun_out
R=${ROUND:-r02_v4}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-tumor --main-lines 100000 --cov-reads 61765396 --bam-reads 200000 > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
SKIP=${STEP_SKIP:-72} COUNT=${STEP_COUNT:-36} bash tools/gpu_launches.sh
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'agg_|segsort' -s 119 -c 50 --csv --log-file gpurun_out/agg_launches.csv python tools/agg_target.py > gpurun_out/agg_launches.log 2>&1
echo "agg step rc=$?"
timeout 1200 ncu --set full --clock-control none \
    -k regex:'window_runs_small|wr_|coverage_kernel|gc_small|m3_|segsort_pass|segsort_local|segsort_tiny|final_labels|pack_y|agg_|med_pass' \
    -s ${FULL_SKIP:-56} -c ${FULL_COUNT:-84} -o /tmp/prof_${R} -f python tools/profile_target.py > gpurun_out/prof.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/prof.log
ncu -i /tmp/prof_${R}.ncu-rep --page raw --csv > gpurun_out/prof_${R}_raw.csv 2> gpurun_out/prof_export.err
echo "raw export rc=$? $(wc -c < gpurun_out/prof_${R}_raw.csv) bytes"
# source-level capture of the three kernels the bench reports on: one steady-state launch each
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'window_runs_small' \
    -s 4 -c 1 -o gpurun_out/src_wr_${R} -f python tools/profile_target.py > gpurun_out/src_wr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'m3_finish' \
    -s 5 -c 1 -o gpurun_out/src_fin_${R} -f python tools/profile_target.py > gpurun_out/src_fin.log 2>&1
echo "source captures done"; ls -la gpurun_out/*.ncu-rep
timeout 300 python tools/kernel_times.py > gpurun_out/kernel_times_${R}.txt 2>&1
echo "kernel times rc=$?"
du -sh gpurun_out
