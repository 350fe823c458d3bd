#!/bin/bash
# Pre-build compile-time variants HERE (nvcc cross-compiles) so that the GPU box only runs them:
#   VARIANTS="name1:-DX=1 name2:-DX=0,-DY=2" bash tools/build_variants.sh   -> tiddit_b200/_variants/libtdt_b200_<name>.so
mkdir -p tiddit_b200/_variants
for v in ${VARIANTS}; do
  name=${v%%:*}; defs=${v#*:}; defs=${defs//,/ }
  TDT_NVCC_DEFS="$defs" python -c "
from tiddit_b200 import build
print(build.build(force=True, out='tiddit_b200/_variants/libtdt_b200_$name.so', build_dir='tiddit_b200/_build/var_$name'))" 2>&1 | tail -1 &
done
wait
