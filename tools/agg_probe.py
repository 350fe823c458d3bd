"""Stage timing of tdt_cluster_aggregate on the 30X-shaped set (tuning aid):  python tools/agg_probe.py [n] [reps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tiddit_b200 import device_ops, synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
run = lambda: device_ops.cluster_aggregate_device(labels, A, B, span, name, flags, O, same, P, 5000, False, 3, L, n, rows, mem, counts)
for _ in range(2):
    run()
torch.cuda.synchronize()
print("counts", counts.tolist())
tot = {}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
whole = []
for _ in range(reps):
    flush.add_(1)
    torch.cuda.synchronize()
    ev[0].record(); run(); ev[1].record(); torch.cuda.synchronize()
    whole.append(ev[0].elapsed_time(ev[1]))
    flush.add_(1)
    torch.cuda.synchronize()
    _lib.profile_begin()
    run()
    for k, ms in _lib.profile_end():
        tot[k] = tot.get(k, 0) + ms / reps
print("whole ms", np.mean(whole), "=> %.2f G signals/s" % (n / np.mean(whole) / 1e6))
for k, v in tot.items():
    print("  %-18s %.3f ms" % (k, v))
