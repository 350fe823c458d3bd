"""ncu target: 3 candidate-aggregation passes over the 30X signal set (labels computed once)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiddit_b200 import device_ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
a, b, off, L = synth.wgs30x_signals(n)
rec = synth.signal_records(a, b, off)
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
A, B, O = d(a), d(b), d(off)
span, name, flags, same = d(rec["span"]), d(rec["name_id"]), d(rec["flags"]), d(rec["same_chrom"])
P = len(off) - 1
labels = device_ops.cluster_labels_device(A, B, O, P, 500, 3, L)
rows = torch.empty((n, 16), dtype=torch.int32, device="cuda")
mem = torch.empty(n, dtype=torch.int32, device="cuda")
counts = torch.zeros(4, dtype=torch.int64, device="cuda")
for _ in range(3):
    device_ops.cluster_aggregate_device(labels, A, B, span, name, flags, O, same, P, 5000, False, 3, L, n, rows, mem, counts)
torch.cuda.synchronize()
print("done", counts.tolist())
