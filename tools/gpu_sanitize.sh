#!/bin/bash
# compute-sanitizer over small invocations of every kernel family: memcheck (out-of-bounds, misaligned), racecheck
# (shared-memory hazards), synccheck.  Small inputs only: the tools slow kernels by 10-100x.
mkdir -p gpurun_out
cat > /tmp/san_target.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import oracle
from tiddit_b200 import device_ops, synth, tiddit_coverage, tiddit_gc
a, b, off, L = synth.wgs30x_signals(60_000)
lab = device_ops.cluster_labels(a, b, off, 500, 3, L)
assert np.array_equal(lab, oracle.cluster_segments(a, b, off, 500, 3))
rec = synth.signal_records(a, b, off)
args = (lab, a, b, rec["span"], rec["name_id"], rec["flags"], off, rec["same_chrom"], 5000, False, 3)
rows, mem = device_ops.cluster_aggregate(*args, max_pos=L, n_names=rec["n_names"])
wr, wm = oracle.cluster_aggregate(*args)
keep = [c for c in range(16) if c != 3]
assert np.array_equal(rows[:, keep], wr[:, keep])
rng = np.random.default_rng(1)
sizes = rng.integers(0, 5000, 40); boff = np.concatenate([[0], np.cumsum(sizes)])
cov = np.float32(rng.integers(0, 400, boff[-1])).astype(np.float64) / np.float32(50)
gc = rng.integers(-1, 60, boff[-1]).astype(np.int8)
m1, c1 = device_ops.coverage_medians(cov, gc, boff); m2, c2 = oracle.coverage_medians(cov, gc, boff)
assert np.array_equal(m1.view(np.uint64), m2.view(np.uint64))
s, e, roff, lens = synth.coverage_reads(50_000, contigs=synth.GRCH38[20:22])
header = {"SQ": [{"SN": n, "LN": l} for n, l in synth.GRCH38[20:22]]}
cv, ebs = tiddit_coverage.create_coverage(header, 500, "chr21")
tiddit_coverage.update_coverage_batch(s[:roff[1]], e[:roff[1]], 500, cv, ebs)
seq = synth.fasta_sequence(200_003)
assert np.array_equal(tiddit_gc.gc_bins(seq, 50, 0.5), oracle.gc_bins(seq, 50, 0.5))
print("sanitizer target ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer target ok|hazard" gpurun_out/sanitize_$tool.log | tail -4
done
