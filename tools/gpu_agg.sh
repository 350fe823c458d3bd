#!/bin/bash
# aggregation: tests, stage timing, per-kernel ncu list of ONE pass (the third)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aggregate.py -x -q --timeout 300 2>&1 | tail -3
timeout 300 python tools/agg_probe.py 2>&1 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'agg_|segsort' --csv --log-file gpurun_out/agg_launches.csv python tools/agg_target.py ${NSIG:-20000000} > gpurun_out/agg_launches.log 2>&1
echo rc=$?
