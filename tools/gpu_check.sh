#!/bin/bash
# One GPU session: smoke, the gpu test-suite, a short bench.  Everything is wrapped in `timeout` so a
# hung kernel cannot hold the box; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ${PYTEST_ARGS} > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/tests.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
