#!/bin/bash
mkdir -p gpurun_out
export TDT_B200_LIB=$PWD/tiddit_b200/_variants/libtdt_b200_s2prof.so
timeout 300 python - > gpurun_out/s2prof.txt 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
from tiddit_b200 import device_ops, synth
a, b, off, L = synth.wgs30x_signals(20_000_000)
A, B, O = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(off).cuda()
for _ in range(2):
    device_ops.cluster_labels_device(A, B, O, len(off)-1, 500, 3, L)
    torch.cuda.synchronize()
PY
cat gpurun_out/s2prof.txt
