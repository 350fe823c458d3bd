#!/bin/bash
# per-kernel durations of ONE clustering pass (after 2 warm-up passes) of tools/profile_target.py
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'segsort|m3_|window_runs|wr_|pack_y|heads_from|set_dims|pair_first|group_|final_labels' -s ${SKIP:-64} -c ${COUNT:-32} --csv --log-file gpurun_out/step_launches.csv python tools/profile_target.py ${NSIG:-20000000} > gpurun_out/step_launches.log 2>&1
echo rc=$?
