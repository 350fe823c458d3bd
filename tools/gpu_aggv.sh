#!/bin/bash
# compile-time variants of the aggregation scans (tools/build_variants.sh: TDT_AG_ITEMS) next to the default build
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_aggregate.py -m gpu -q -x --timeout 90 2>&1 | tail -3
for v in default items16 items32; do
  lib=""; [ "$v" != default ] && lib="$PWD/tiddit_b200/_variants/libtdt_b200_$v.so"
  echo "== $v"
  TDT_B200_LIB=$lib TDT_AB_SETTINGS="256" timeout 60 python tools/agg_direct_ab.py 20000000 0 5 2>&1 | tail -1
  cp gpurun_out/agg_direct_ab.json gpurun_out/agg_items_$v.json
done
for v in items16 items32; do
  TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_$v.so timeout 60 python -m pytest tests/test_gpu_aggregate.py -m gpu -q -x --timeout 50 -k "default" 2>&1 | tail -2
done
