#!/bin/bash
# direct modes / names kernel of the aggregation: parity on every threshold, the A/B (tools/agg_direct_ab.py), in-stream kernel times
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_aggregate.py -m gpu -q -x --timeout 150 2>&1 | tail -5
TDT_AB_SETTINGS="${TDT_AB_SETTINGS:-0 256:0 256:1 128:1}" timeout 170 python tools/agg_direct_ab.py 20000000 20000000 3 2>&1 | tail -10
TDT_KT_ONLY=aggregate timeout 100 python tools/kernel_times.py > gpurun_out/agg_kernel_times.txt 2>&1; tail -45 gpurun_out/agg_kernel_times.txt
