#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ploidy.py -m gpu -q -x --timeout 600 2>&1 | tail -3
TDT_PROF_DETAIL=1 timeout 600 python - <<'PY'
import sys, json, os
sys.path.insert(0, '.')
import torch, argparse, numpy as np
import bench
from tiddit_b200 import device_ops, synth, _lib
args = argparse.Namespace(steps=20, warmup=3, no_cpu=True)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = bench.medians_leg(torch, args, bench.peaks()[0], lambda: flush_buf.add_(1))
print(out["ms_per_step"], out["roofline"]["frac"])
lens = np.array([ln for _, ln in synth.GRCH38], dtype=np.int64); nb = (lens + 49) // 50
off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64); n = int(off[-1])
g = torch.Generator(device="cuda"); g.manual_seed(5)
cov = torch.round(torch.rand(n, device="cuda", generator=g, dtype=torch.float64) * 3000) / 50.0
gc = torch.randint(-1, 80, (n,), device="cuda", generator=g, dtype=torch.int8)
off_d = torch.from_numpy(off).cuda()
for _ in range(2): device_ops.coverage_medians_device(cov, gc, off_d, len(nb))
torch.cuda.synchronize(); _lib.profile_begin(); device_ops.coverage_medians_device(cov, gc, off_d, len(nb))
for k, v in _lib.profile_end(): print("%-40s %8.1f us" % (k[:40], v * 1e3))
PY
