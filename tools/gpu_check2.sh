#!/bin/bash
# N=2 box: full gpu test-suite on GPU 0, then the 2-GPU bench (p2p exchange)
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/tests.log
EXCH=p2p NGPU=2 bash tools/gpu_multi.sh 2>&1 | grep -v "^== engine" 
