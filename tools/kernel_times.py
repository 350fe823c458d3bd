"""In-stream per-kernel durations (CUDA events around every launch, TDT_PROF_DETAIL=1) of one clustering call and one
aggregation call on the 30X set -- unlike ncu's serialised cold launches these are the times the step really pays.
    TDT_PROF_DETAIL=1 python tools/kernel_times.py [n]"""
import os, sys
os.environ["TDT_PROF_DETAIL"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tiddit_b200 import device_ops, synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
tumor = os.environ.get("TDT_KT_WORKLOAD") == "tumor"   # BASELINE configs[4]: eps 1000, m 5
EPS, M = (1000, 5) if tumor else (500, 3)
a, b, off, L = (synth.tumor60x_signals if tumor else synth.wgs30x_signals)(n)
rec = synth.signal_records(a, b, off)
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
A, B, O = d(a), d(b), d(off)
span, name, flags, same = d(rec["span"]), d(rec["name_id"]), d(rec["flags"]), d(rec["same_chrom"])
P = len(off) - 1
labels = device_ops.cluster_labels_device(A, B, O, P, EPS, M, L)
rows = torch.empty((n, 16), dtype=torch.int32, device="cuda")
mem = torch.empty(n, dtype=torch.int32, device="cuda")
counts = torch.zeros(4, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
calls = {"cluster": lambda: device_ops.cluster_labels_device(A, B, O, P, EPS, M, L, labels_out=labels),
         "aggregate": lambda: device_ops.cluster_aggregate_device(labels, A, B, span, name, flags, O, same, P, 5000, False, M, L, n, rows, mem, counts)}
only = os.environ.get("TDT_KT_ONLY")   # "cluster" / "aggregate": just that call
for what, fn in calls.items():
    if only and what != only:
        continue
    fn(); fn()
    acc, reps = None, 5
    for _ in range(reps):
        flush.add_(1); torch.cuda.synchronize()
        _lib.profile_begin(); fn()
        got = _lib.profile_end()
        if acc is None:
            acc = [[k, 0.0] for k, _ in got]
        for i, (k, ms) in enumerate(got):
            acc[i][1] += ms / reps
    print("== %s: %d launches, %.1f us in kernels" % (what, len(acc), 1e3 * sum(v for _, v in acc)))
    for i, (k, v) in enumerate(acc):
        print("%3d %-44s %8.1f us" % (i, k[:44], v * 1e3))
