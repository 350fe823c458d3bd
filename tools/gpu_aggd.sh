#!/bin/bash
# direct modes / names kernel of the aggregation: parity on every threshold, then the A/B (tools/agg_direct_ab.py)
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_aggregate.py -m gpu -q -x --timeout 150 2>&1 | tail -5
timeout 170 python tools/agg_direct_ab.py 20000000 20000000 3 2>&1 | tail -14
