#!/bin/bash
# One gpurun call: check (smoke, gpu tests, bench), ncu evidence, compute-sanitizer.
bash tools/gpu_check.sh
ROUND=${ROUND:-r01_v3} bash tools/gpu_profile.sh
bash tools/gpu_sanitize.sh
du -sh gpurun_out
