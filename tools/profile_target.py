"""A short, fixed sequence of launches for ncu: 3 clustering passes over the 30X signal set, 2 coverage passes
over 3X-worth of reads, 1 GC pass over 250 Mbp, 1 candidate-aggregation pass over the same signals, 1 ploidy-median
pass over 61.8 M bins.  No timing, no CPU work."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiddit_b200 import device_ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
a, b, off, L = synth.wgs30x_signals(n)
a_d, b_d, off_d = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(off).cuda()
for _ in range(3):
    labels = device_ops.cluster_labels_device(a_d, b_d, off_d, len(off) - 1, 500, 3, L)
torch.cuda.synchronize()

lens = np.array([ln for _, ln in synth.GRCH38], dtype=np.int64)
n_reads = 61_765_396
per = np.floor(n_reads * lens / lens.sum()).astype(np.int64)
starts = [synth.sorted_starts_device(torch, int(k), int(ln)) for ln, k in zip(lens, per)]
start = torch.cat(starts)
end = torch.cat([torch.clamp(s + 150, max=int(ln)) for s, ln in zip(starts, lens)])
nb = np.ceil(lens / 500.0).astype(np.int64)
read_off = torch.from_numpy(np.concatenate([[0], np.cumsum(per)]).astype(np.int64)).cuda()
bin_off = torch.from_numpy(np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)).cuda()
ebs = torch.from_numpy((lens - (nb - 1) * 500).astype(np.int32)).cuda()
bins = torch.zeros(int(nb.sum()), dtype=torch.float64, device="cuda")
bad = device_ops.new_first_bad(torch)
for _ in range(2):
    device_ops.coverage_accumulate_contigs_device(start, end, read_off, bin_off, ebs, 500, bins, bad)
torch.cuda.synchronize()

seq, ln = device_ops.padded_sequence_device(synth.fasta_sequence(250_000_000))
device_ops.gc_bins_device(seq, ln, 50, 0.5)
torch.cuda.synchronize()
del seq, start, end, bins
torch.cuda.empty_cache()

rec = synth.signal_records(a, b, off)
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
device_ops.cluster_aggregate_device(labels, a_d, b_d, d(rec["span"]), d(rec["name_id"]), d(rec["flags"]), off_d,
                                    d(rec["same_chrom"]), len(off) - 1, 5000, False, 3, L, n)
torch.cuda.synchronize()

nb50 = (lens + 49) // 50
off50 = torch.from_numpy(np.concatenate([[0], np.cumsum(nb50)]).astype(np.int64)).cuda()
nbin = int(nb50.sum())
cov = torch.round(torch.rand(nbin, device="cuda", dtype=torch.float64) * 3000) / 50.0
gcb = torch.randint(-1, 80, (nbin,), device="cuda", dtype=torch.int8)
device_ops.coverage_medians_device(cov, gcb, off50, len(nb50))
torch.cuda.synchronize()
print("done")
