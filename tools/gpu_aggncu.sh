#!/bin/bash
# ncu --set full of the aggregation's gather / direct / run kernels (third pass of tools/agg_target.py), exported as CSV on the box
mkdir -p gpurun_out
timeout 110 ncu --set full --clock-control none -k regex:'agg_direct|agg_gather|agg_runs' -s 10 -c 5 -o /tmp/prof_agg -f python tools/agg_target.py > gpurun_out/prof_agg.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/prof_agg.log
ncu -i /tmp/prof_agg.ncu-rep --page raw --csv > gpurun_out/prof_agg_raw.csv 2> gpurun_out/prof_agg_export.err
echo "export rc=$? $(wc -c < gpurun_out/prof_agg_raw.csv) bytes"
