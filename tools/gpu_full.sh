#!/bin/bash
# ncu --set full of selected kernels of tools/profile_target.py: KREGEX, SKIP, COUNT, OUT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s ${SKIP:-0} -c ${COUNT:-3} -o gpurun_out/${OUT:-full} -f python tools/profile_target.py ${NSIG:-20000000} > gpurun_out/full.log 2>&1
echo rc=$?; tail -2 gpurun_out/full.log
