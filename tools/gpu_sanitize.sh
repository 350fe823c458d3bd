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
assert np.array_equal(tiddit_gc.gc_bins(seq, 7, 0.5), oracle.gc_bins(seq, 7, 0.5))
assert np.array_equal(tiddit_gc.gc_bins(seq, 500, 0.5), oracle.gc_bins(seq, 500, 0.5))
# the queued per-read path
cq, eq = tiddit_coverage.create_coverage(header, 500, "chr21")
for x, y in zip(s[:2000].tolist(), e[:2000].tolist()):
    cq = tiddit_coverage.update_coverage(x, y, 500, cq, eq)
want = np.zeros(len(cq)); oracle.update_coverage_batch(s[:2000], e[:2000], 500, want, eq)
assert np.array_equal(np.asarray(cq).view(np.uint64), want.view(np.uint64))
# one large pair (large-segment chain: generation 3 for posA, LSD for posB), then everything through the LSD chain
a2, b2, off2, L2 = synth.config2_signals(40_000, n_clusters=800)
want2 = oracle.cluster_segments(a2, b2, off2, 500, 3)
assert np.array_equal(device_ops.cluster_labels(a2, b2, off2, 500, 3, L2), want2)
os.environ["TDT_SEGSORT"] = "lsd"
assert np.array_equal(device_ops.cluster_labels(a2, b2, off2, 500, 3, L2), want2)
assert np.array_equal(device_ops.cluster_labels(a, b, off, 500, 3, L), lab)
del os.environ["TDT_SEGSORT"]
# generation 3 of the segmented sort (tdt_segsort3.cuh): by-value mode (value = element index) and, forced, the stable mode;
# clustered keys (equi-depth path), pile-ups (later rounds), > 8192 equal keys (compaction path / copy batches), narrow keys
import torch
def sort_case(keys, sizes, key_bits, with_vals):
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    vals = np.random.default_rng(3).permutation(n).astype(np.int32) if with_vals else np.arange(n, dtype=np.int32)
    k = torch.from_numpy(keys.astype(np.int64).astype(np.uint32).view(np.int32)).cuda()
    v = torch.from_numpy(vals).cuda() if with_vals else None
    ko, vo = device_ops.segsort_device(k, v, torch.from_numpy(off).cuda(), key_bits, None)
    ko, vo = ko.cpu().numpy().view(np.uint32), vo.cpu().numpy()
    for s in range(len(sizes)):
        lo, hi = off[s], off[s + 1]
        order = np.argsort(keys[lo:hi], kind="stable")
        assert np.array_equal(ko[lo:hi], keys[lo:hi][order].astype(np.uint32)) and np.array_equal(vo[lo:hi], vals[lo:hi][order])
r3 = np.random.default_rng(11)
sizes3 = [90_000, 9000, 30_000, 8193, 5000]
n3 = sum(sizes3)
for key_bits in (28, 20):
    hi = 1 << key_bits
    centres = r3.integers(0, hi, 300)
    clustered = np.clip(centres[r3.integers(0, 300, n3)] + r3.integers(-100, 100, n3), 0, hi - 1)
    hot = r3.integers(0, hi, n3); m = r3.random(n3) < 0.7; hot[m] = hi // 2 + r3.integers(0, 3, int(m.sum()))
    equal = np.full(n3, hi - 5)
    narrow = hi // 3 + r3.integers(0, 40, n3)
    for keys in (clustered, hot, equal, narrow):
        for mode, with_vals in (("", False), ("msd", True), ("", True)):
            if mode:
                os.environ["TDT_SEGSORT"] = mode
            sort_case(keys, sizes3, key_bits, with_vals)
            os.environ.pop("TDT_SEGSORT", None)
print("sanitizer target ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer target ok|hazard" gpurun_out/sanitize_$tool.log | tail -4
done

# the peer-memory label exchange (needs 2 GPUs): memcheck over both ranks
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
cat > /tmp/san_peer.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from tiddit_b200 import engine, synth
from oracle import oracle
os.environ["TDT_LABEL_EXCHANGE"] = "p2p"
a, b, off, L = synth.wgs30x_signals(80_000)
for _ in range(3):
    got = engine.sharded_labels(a, b, off, 500, 3, L)
assert np.array_equal(got, oracle.cluster_segments(a, b, off, 500, 3))
dist.barrier(); dist.destroy_process_group()
print("peer target ok", rank)
PY
  timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 /tmp/san_peer.py > gpurun_out/sanitize_peer_memcheck.log 2>&1
  echo "peer memcheck rc=$?"; grep -E "ERROR SUMMARY|peer target ok" gpurun_out/sanitize_peer_memcheck.log | tail -6
fi
