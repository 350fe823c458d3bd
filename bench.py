#!/usr/bin/env python
"""bench.py -- signals clustered / s (+ coverage bins / s) of the TIDDIT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload wgs30x|config2|tumor60x]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the clustering hot path (all (chrA,chrB) pairs: key packing, both radix sorts, both
eps-range-query/run-labelling kernels, id assignment, scatter) over the synthetic 30X-WGS-shaped signal set of
BASELINE.json configs[2] (20 M signals, 300 pairs, eps=500, m=3).  With N > 1 the pairs are sharded over the
ranks (LPT by signal count, no data-path collective) and the labels are all-gathered once per step
(BASELINE.json configs[3]; total work fixed => "scaling": "strong").  Rank 0 prints ONE JSON line.

value        device-resident inputs, CUDA events, max over ranks
e2e          the same through the host front end: pinned host arrays -> H2D -> kernels -> D2H labels, every step
roofline     the eps-range-query kernel (window_runs<X>): 8 B/signal / its event-timed duration
cpu_baseline the REAL reference (oracle/_ref, compiled from the upstream sources) or the C port, 1 core,
             on a bounded sample of whole pairs of the same workload
coverage     the coverage kernel on 30X-shaped reads (reads/s, bins/s, its own roofline and CPU baseline)

--impl reference times the reference's own CPU path (all host cores, a bounded sample per step).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "wgs30x": dict(gen="wgs30x_signals", n=20_000_000, eps=500, m=3,
                   desc="30X-WGS-shaped synthetic set: 20M signals over 300 chrom-pairs (BASELINE configs[2])"),
    "config2": dict(gen="config2_signals", n=1_000_000, eps=500, m=3,
                    desc="1M synthetic signals, single chrom-pair (BASELINE configs[1])"),
    "tumor60x": dict(gen="tumor60x_signals", n=50_000_000, eps=1000, m=5,
                     desc="60X tumor-like synthetic set: 50M signals (BASELINE configs[4])"),
}
METRIC = "signals_clustered_per_sec"
UNIT = "signals/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML, sampled from a thread)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# the reference on the host cores
# ---------------------------------------------------------------------------------------------------
def _ref_modules():
    from oracle import ref
    return ref.load()


def _ref_pair_seconds(job):
    """One (chrA,chrB) list through the reference: sorted(key=posA) + DBSCAN.main (tiddit_cluster.pyx:152-154)."""
    a, b, eps, m, use_ref = job
    rows = [[int(x), int(y), i] for i, (x, y) in enumerate(zip(a, b))]
    t0 = time.perf_counter()
    if use_ref:
        R = _ref_modules()
        arr = np.array(sorted(rows, key=lambda l: l[0]))
        R.DBSCAN.main(arr, eps, m)
    else:
        from oracle import oracle
        arr = np.array(sorted(rows, key=lambda l: l[0]))
        oracle.main(arr, eps, m)
    return time.perf_counter() - t0


def sample_jobs(posA, posB, seg_off, n_crops, crop=50_000, sparse_share=0.3):
    """A bounded, workload-shaped sample for the CPU reference: `n_crops` posA-window crops of ~`crop` signals
    cut out of the largest pairs (local density and clusters intact -- the dense, expensive 70 % of the
    workload) plus whole small pairs adding `sparse_share` of the signals (the sparse inter-chromosomal 30 %).
    Whole large pairs cost the reference minutes to hours each (O(#clusters * n) y-pass), so per signal this
    sample is CHEAPER for the CPU than the full workload."""
    sizes = np.diff(seg_off)
    big = np.argsort(-sizes, kind="stable")
    jobs, total = [], 0
    for i in range(n_crops):
        p = int(big[i % min(len(big), 24)])
        lo, hi = int(seg_off[p]), int(seg_off[p + 1])
        k = min(crop, hi - lo)
        order = np.argsort(posA[lo:hi], kind="stable")
        s0 = ((i // 24) * 7 + 1) * k % max(1, (hi - lo) - k + 1)
        pick = np.sort(order[s0:s0 + k])                      # insertion order kept
        jobs.append((posA[lo:hi][pick], posB[lo:hi][pick]))
        total += k
    want_sparse = int(total * sparse_share / (1 - sparse_share)) if len(sizes) > 1 else 0
    got = 0
    for p in np.argsort(sizes, kind="stable"):
        if got >= want_sparse:
            break
        if sizes[p] < 1000 or sizes[p] > crop:
            continue
        jobs.append((posA[seg_off[p]:seg_off[p + 1]], posB[seg_off[p]:seg_off[p + 1]]))
        got += int(sizes[p])
    return jobs, total + got


def cpu_cluster_baseline(posA, posB, seg_off, eps, m, n_crops=4):
    use_ref = _ref_modules() is not None
    jobs, total = sample_jobs(posA, posB, seg_off, n_crops)
    secs = sum(_ref_pair_seconds((a, b, eps, m, use_ref)) for a, b in jobs)
    return {"value": total / secs, "unit": UNIT, "cores": 1, "kind": "reference" if use_ref else "port",
            "sample": "%d signals: %d posA-window crops of 50k signals from the largest pairs + whole small pairs "
                      "for the sparse 30%%; sorted(key=posA) + DBSCAN.main each (tiddit_cluster.pyx:152-154), "
                      "%.1f s on 1 core; whole large pairs are far slower per signal (O(#clusters*n) y-pass), so "
                      "this flatters the CPU" % (total, n_crops, secs)}


def run_reference(args):
    """--impl reference: the reference's CPU clustering on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    from tiddit_b200 import synth
    w = WORKLOADS[args.workload]
    posA, posB, seg_off, _ = getattr(synth, w["gen"])(args.signals or w["n"])
    use_ref = _ref_modules() is not None
    cores = os.cpu_count() or 1
    jobs, total = sample_jobs(posA, posB, seg_off, cores * args.ref_crops_per_core)
    jobs = [(a, b, w["eps"], w["m"], use_ref) for a, b in jobs]
    times = []
    with ProcessPoolExecutor(max_workers=cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(pool.map(_ref_pair_seconds, jobs, chunksize=1))
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    sec = float(np.mean(times))
    value = total / sec
    sample = ("%d signals per step: %d posA-window crops of 50k signals from the largest pairs + whole small pairs, "
              "over a %d-process pool (one pair per task, as the reference clusters a pair serially)"
              % (total, cores * args.ref_crops_per_core, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": w["desc"], "eps": w["eps"], "min_pts": w["m"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# coverage leg (N = 1)
# ---------------------------------------------------------------------------------------------------
def ncu_traffic(kernel, note_only=False):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full summary (profiles/*_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None
    table = json.load(open(files[-1]))
    rec = table.get(kernel)
    if rec is None:       # template arguments are printed differently across captures: match on the prefix
        hits = [v for k, v in table.items() if k.startswith(kernel)]
        rec = hits[0] if hits else None
    if rec is None:
        return None
    if note_only:
        return "%s: %.1f MB per launch in the shape tools/profile_target.py runs" % (os.path.basename(files[-1]),
                                                                             rec["first_launch_bytes"] / 1e6)
    return rec["first_launch_bytes"]


def coverage_leg(torch, args, hbm_peak, flush):
    from tiddit_b200 import device_ops, synth, _lib
    lens = np.array([ln for _, ln in synth.GRCH38], dtype=np.int64)
    z = 500
    n_reads = args.cov_reads
    per = np.floor(n_reads * lens / lens.sum()).astype(np.int64)
    per[0] += n_reads - per.sum()
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    starts, ends = [], []
    for ln, k in zip(lens, per):                       # coordinate-sorted per contig, like a BAM
        starts.append(synth.sorted_starts_device(torch, int(k), int(ln), g))
        ends.append(torch.clamp(starts[-1] + 150, max=int(ln)))
    start, end = torch.cat(starts), torch.cat(ends)
    del starts, ends
    nb = np.ceil(lens / float(z)).astype(np.int64)
    read_off = torch.from_numpy(np.concatenate([[0], np.cumsum(per)]).astype(np.int64)).cuda()
    bin_off_h = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    bin_off = torch.from_numpy(bin_off_h).cuda()
    ebs = torch.from_numpy((lens - (nb - 1) * z).astype(np.int32)).cuda()
    n_bins = int(bin_off_h[-1])
    bins = torch.zeros(n_bins, dtype=torch.float64, device="cuda")
    bad = device_ops.new_first_bad(torch)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(args.warmup + args.steps):
        bins.zero_()
        flush()
        ev[0].record()
        device_ops.coverage_accumulate_contigs_device(start, end, read_off, bin_off, ebs, z, bins, bad)
        ev[1].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(ev[0].elapsed_time(ev[1]))
    ms = float(np.mean(times))
    mean_cov = float(bins.mean().item())
    alg_bytes = 8.0 * n_reads + 8.0 * n_bins
    out = {"reads": n_reads, "bins": n_bins, "bin_size": z, "ms_per_step": ms, "reads_per_sec": n_reads / ms * 1e3,
           "bins_per_sec": n_bins / ms * 1e3, "mean_coverage": mean_cov,
           "roofline": {"bound": "hbm", "achieved": alg_bytes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / ms / 1e6 / hbm_peak, "traffic": None, "kernel": "coverage_kernel",
                        "traffic_note": ncu_traffic("coverage_kernel<1>", note_only=True),
                        "algorithmic_bytes": "8 B/read + 8 B/bin"}}
    if args.no_cpu:
        return out
    # CPU: the reference's update_coverage per read (1 core), bounded sample
    k = min(n_reads, 2_000_000)
    hs, he = start[:k].cpu().numpy(), end[:k].cpu().numpy()
    R = _ref_modules()
    ln0 = int(lens[0])
    t0 = time.perf_counter()
    if R is not None:
        cov, e0 = R.tiddit_coverage.create_coverage({"SQ": [{"SN": "chr1", "LN": ln0}]}, z, "chr1")
        upd = R.tiddit_coverage.update_coverage
        for a, b in zip(hs.tolist(), he.tolist()):
            upd(a, b, z, cov, e0)
        kind = "reference"
    else:
        from oracle import oracle
        cov, e0 = oracle.create_coverage({"SQ": [{"SN": "chr1", "LN": ln0}]}, z, "chr1")
        oracle.update_coverage_batch(hs, he, z, cov, e0)
        kind = "port"
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": k / dt, "unit": "reads/s", "cores": 1, "kind": kind,
                           "sample": "first %d reads of chr1, one update_coverage call per read, %.1f s" % (k, dt)}
    # parity spot check on the same sample (test infrastructure as checker)
    from oracle import oracle
    chk = np.zeros(int(nb[0]))
    oracle.update_coverage_batch(hs, he, z, chk, int(lens[0] - (nb[0] - 1) * z))
    dchk = torch.zeros(int(nb[0]), dtype=torch.float64, device="cuda")
    device_ops.coverage_accumulate_device(start[:k], end[:k], z, int(lens[0] - (nb[0] - 1) * z), dchk,
                                          device_ops.new_first_bad(torch))
    out["verified"] = bool(np.array_equal(dchk.cpu().numpy().view(np.uint64), chk.view(np.uint64)))
    del start, end, bins
    torch.cuda.empty_cache()
    return out


def gc_leg(torch, args, hbm_peak, flush):
    """GC bins (tiddit_gc.pyx:6-33) of a chr1-sized contig resident in HBM, bin 50 like `--sv`: bases/s and bins/s."""
    from tiddit_b200 import device_ops, synth
    n_bases = args.gc_bases
    seq_h = synth.fasta_sequence(n_bases)
    seq, ln = device_ops.padded_sequence_device(seq_h)
    z = 50
    n_bins = (n_bases + z - 1) // z
    out_bins = torch.zeros(n_bins, dtype=torch.int8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(args.warmup + args.steps):
        flush()
        ev[0].record()
        device_ops.gc_bins_device(seq, ln, z, 0.5, out=out_bins)
        ev[1].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(ev[0].elapsed_time(ev[1]))
    ms = float(np.mean(times))
    alg = float(n_bases + n_bins)
    out = {"bases": n_bases, "bins": n_bins, "bin_size": z, "ms_per_step": ms, "bases_per_sec": n_bases / ms * 1e3,
           "bins_per_sec": n_bins / ms * 1e3,
           "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg / ms / 1e6 / hbm_peak, "kernel": "gc_small_kernel",
                        "traffic_note": ncu_traffic("gc_small_kernel", note_only=True),
                        "algorithmic_bytes": "1 B/base + 1 B/bin"}}
    from oracle import oracle
    k = min(n_bases, 5_000_000) // z * z
    out["verified"] = bool(np.array_equal(out_bins[:k // z].cpu().numpy(), oracle.gc_bins(seq_h[:k], z, 0.5)))
    if not args.no_cpu:
        R = _ref_modules()
        if R is not None and getattr(R, "tiddit_gc", None) is not None:
            import tempfile
            kk = min(k, 3_000_000)
            tmp = tempfile.mkdtemp(prefix="tdt_bench_gc_")
            fa = os.path.join(tmp, "ref.fa")
            with open(fa, "w") as f:
                f.write(">c\n")
                text = bytes(seq_h[:kk]).decode("ascii")
                f.write("\n".join(text[i:i + 60] for i in range(0, kk, 60)) + "\n")
            R.tiddit_gc.binned_gc(fa, "c", z, 0.5)        # warms the stand-in FastaFile's cache: parsing is not timed
            t0 = time.perf_counter()
            R.tiddit_gc.binned_gc(fa, "c", z, 0.5)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": kk / dt, "unit": "bases/s", "cores": 1, "kind": "reference",
                                   "sample": "binned_gc on the first %d bases (FASTA already in memory), %.1f s" % (kk, dt)}
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
    del seq, out_bins
    torch.cuda.empty_cache()
    return out


def bam_leg(torch, args):
    """`--cov` from a BAM FILE: libtdt_bam.so (BGZF inflated on the host cores, records as columns)
    + the coverage kernel batch by batch, host->device copies and the final device->host read of the bins included.
    The synthetic BAM (150-bp reads with bases and qualities, coordinate-sorted) is written once, untimed."""
    import tempfile
    from tiddit_b200 import bamio, synth, __main__ as cli
    contigs = synth.GRCH38[20:22]
    s, e, roff, lens = synth.coverage_reads(args.bam_reads, contigs=contigs)
    n = len(s)
    rng = np.random.default_rng(17)
    rid = np.repeat(np.arange(len(contigs)), np.diff(roff)).astype(np.int32)
    s = np.minimum(s, (lens[rid] - 150).astype(np.int32))      # whole 150M reads inside the contig (order is kept)
    flag = np.where(rng.random(n) < 0.02, 0x400, 0).astype(np.uint16)
    mapq = rng.integers(0, 61, n).astype(np.uint8)
    tmp = tempfile.mkdtemp(prefix="tdt_bench_bam_")
    path = os.path.join(tmp, "reads.bam")
    bamio.write_bam_columns(path, contigs, rid, s, flag, mapq)
    z, q = 500, 20
    times = []
    for it in range(4):
        t0 = time.perf_counter()
        cov, header = cli.coverage_from_bam(path, z, q)
        torch.cuda.synchronize()
        if it:
            times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    out = {"reads": n, "file_bytes": os.path.getsize(path), "bin_size": z, "min_q": q, "s_per_pass": dt,
           "reads_per_sec": n / dt, "file_MB_per_sec": os.path.getsize(path) / dt / 1e6, "host_threads": os.cpu_count(),
           "path": "tiddit_b200.__main__.coverage_from_bam: BGZF inflate (zlib) on all host cores + record decode -> numpy "
                   "columns -> filters -> H2D -> tdt_coverage_accumulate_contigs per 1M-read batch -> bins D2H",
           "bound": "host: zlib inflate of the BGZF blocks (the coverage kernel takes microseconds per batch)"}
    from oracle import oracle
    keep = ((flag & 0x400) == 0) & (mapq >= q)
    ok = True
    for ci, (name, ln) in enumerate(contigs):
        sel = keep & (rid == ci)
        want, ebs = oracle.create_coverage({"SQ": [{"SN": name, "LN": ln}]}, z, name)
        oracle.update_coverage_batch(s[sel], np.minimum(s[sel] + 150, 2 ** 31 - 1).astype(np.int32), z, want, ebs)
        ok = ok and np.array_equal(cov[name].view(np.uint64), want.view(np.uint64))
    out["verified"] = bool(ok)
    if not args.no_cpu:
        R = _ref_modules()
        k = min(n, 1_000_000)
        name, ln = contigs[0]
        sel = np.flatnonzero(keep & (rid == 0))[:k]
        hs, he = s[sel].tolist(), (s[sel] + 150).tolist()
        if R is not None:
            c0, e0 = R.tiddit_coverage.create_coverage({"SQ": [{"SN": name, "LN": ln}]}, z, name)
            upd = R.tiddit_coverage.update_coverage
            t0 = time.perf_counter()
            for a, b in zip(hs, he):
                upd(a, b, z, c0, e0)
            dtc = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": len(hs) / dtc, "unit": "reads/s", "cores": 1, "kind": "reference",
                                   "sample": "%d reads: the reference's update_coverage call per read only (%.1f s); its BAM "
                                             "iteration through pysam, absent here, comes on top, so this flatters the CPU" % (len(hs), dtc)}
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return out


# ---------------------------------------------------------------------------------------------------
# candidate aggregation + ploidy medians legs (N = 1; the SURVEY 8(f) rows built so far)
# ---------------------------------------------------------------------------------------------------
def aggregate_leg(torch, args, posA, posB, seg_off, L, eps, m, hbm_peak, flush):
    """tdt_cluster_aggregate on the step's own labels: candidates/s, stage times, CPU port on a bounded sample."""
    from tiddit_b200 import device_ops, synth, _lib
    n, P = len(posA), len(seg_off) - 1
    rec = synth.signal_records(posA, posB, seg_off)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    A, B, O = d(posA), d(posB), d(seg_off)
    span, name, flags, same = d(rec["span"]), d(rec["name_id"]), d(rec["flags"]), d(rec["same_chrom"])
    labels = device_ops.cluster_labels_device(A, B, O, P, eps, m, L)
    rows = torch.empty((n, 16), dtype=torch.int32, device="cuda")
    mem = torch.empty(n, dtype=torch.int32, device="cuda")
    counts = torch.zeros(4, dtype=torch.int64, device="cuda")
    max_ins, is_mp, min_reads = 5000, False, 3
    run = lambda: device_ops.cluster_aggregate_device(labels, A, B, span, name, flags, O, same, P, max_ins, is_mp,
                                                      min_reads, L, n, rows, mem, counts)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times, stages = [], {}
    c0 = _lib.launch_count()
    run()
    launches = _lib.launch_count() - c0
    for it in range(args.warmup + args.steps):
        flush()
        ev[0].record()
        run()
        ev[1].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(ev[0].elapsed_time(ev[1]))
    for _ in range(3):
        flush()
        torch.cuda.synchronize()
        _lib.profile_begin()
        run()
        for k, v in _lib.profile_end():
            stages[k] = stages.get(k, 0.0) + v / 3
    ms = float(np.mean(times))
    C, M, err, _ = (int(v) for v in counts.cpu().tolist())
    alg = 33.0 * n + 4.0 * M + 64.0 * C
    out = {"signals": n, "candidates": C, "members": M, "ms_per_step": ms, "signals_per_sec": n / ms * 1e3,
           "candidates_per_sec": C / ms * 1e3, "gpu_launches_per_step": int(launches), "data_error": err,
           "stages_ms": {k: round(v, 4) for k, v in stages.items()},
           "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg / ms / 1e6 / hbm_peak,
                        "algorithmic_bytes": "33 B/signal in (label, posA, posB, 4 x span, name id, flags) + 4 B/member "
                                             "+ 64 B/candidate out; the four sorts are implementation traffic"}}
    if not args.no_cpu:
        from oracle import oracle
        k = int(np.searchsorted(seg_off, 2_000_000, side="right"))          # whole pairs, about 2M signals
        k = max(k, 1)
        hi = int(seg_off[k])
        lab_h = labels[:hi].cpu().numpy()
        a = (lab_h, posA[:hi], posB[:hi], rec["span"][:hi], rec["name_id"][:hi], rec["flags"][:hi], seg_off[:k + 1],
             rec["same_chrom"][:k], max_ins, is_mp, min_reads)
        t0 = time.perf_counter()
        want_rows, want_mem = oracle.cluster_aggregate(*a)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": hi / dt, "unit": "signals/s", "cores": 1, "kind": "port",
                               "sample": "the first %d pairs (%d signals) through the C restatement of "
                                         "tiddit_cluster.pyx:156-336, %.2f s; the reference itself does this fold in "
                                         "interpreted Python per signal" % (k, hi, dt)}
        got_rows = rows[:C].cpu().numpy()
        sel = got_rows[:, 0] < k
        keep = [c for c in range(16) if c != 3]
        out["verified"] = bool(np.array_equal(got_rows[sel][:, keep], want_rows[:, keep]))
    return out


def medians_leg(torch, args, hbm_peak, flush):
    """tdt_coverage_medians on 61.8 M bins (GRCh38, bin size 50: the arrays determine_ploidy walks in --sv runs)."""
    from tiddit_b200 import device_ops, synth
    lens = np.array([ln for _, ln in synth.GRCH38], dtype=np.int64)
    nb = (lens + 49) // 50
    off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    n = int(off[-1])
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    cov = torch.round(torch.rand(n, device="cuda", generator=g, dtype=torch.float64) * 3000) / 50.0
    cov[torch.rand(n, device="cuda", generator=g) < 0.08] = 0.0
    gc = torch.randint(-1, 80, (n,), device="cuda", generator=g, dtype=torch.int8)
    off_d = torch.from_numpy(off).cuda()
    med = torch.empty(len(nb) + 1, dtype=torch.float64, device="cuda")
    cnt = torch.empty(len(nb) + 1, dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(args.warmup + args.steps):
        flush()
        ev[0].record()
        device_ops.coverage_medians_device(cov, gc, off_d, len(nb), med, cnt)
        ev[1].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(ev[0].elapsed_time(ev[1]))
    ms = float(np.mean(times))
    out = {"bins": n, "contigs": len(nb), "ms_per_step": ms, "bins_per_sec": n / ms * 1e3, "gpu_launches_per_step": 18,
           "roofline": {"bound": "hbm", "achieved": 9.0 * n / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": 9.0 * n / ms / 1e6 / hbm_peak,
                        "algorithmic_bytes": "9 B/bin (float64 coverage + int8 GC), read once; the radix select "
                                             "streams them 6 times (digit passes of 11/11/11/11/11/9 bits; the upper median comes out of the last one)"}}
    if not args.no_cpu:
        R = _ref_modules()
        lo, hi = int(off[20]), int(off[22])                                  # chr21 + chr22: 1.95 M bins
        cov_h, gc_h = cov[lo:hi].cpu().numpy(), gc[lo:hi].cpu().numpy()
        names = ["chr21", "chr22"]
        cd = {nm: cov_h[int(off[20 + i]) - lo:int(off[21 + i]) - lo] for i, nm in enumerate(names)}
        gd = {nm: gc_h[int(off[20 + i]) - lo:int(off[21 + i]) - lo] for i, nm in enumerate(names)}
        import tempfile
        t0 = time.perf_counter()
        if R is not None and hasattr(R, "tiddit_coverage_analysis"):
            lib = R.tiddit_coverage_analysis.determine_ploidy(cd, names, {}, 2, os.path.join(tempfile.mkdtemp(), "p"), 0,
                                                              "", 50, {"SQ": []}, gd)
            want = [lib["avg_coverage_chr21"], lib["avg_coverage_chr22"]]
            kind = "reference"
        else:
            from oracle import oracle
            m_, _ = oracle.coverage_medians(cov_h, gc_h, off[20:23] - off[20])
            want = m_[:2].tolist()
            kind = "port"
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": (hi - lo) / dt, "unit": "bins/s", "cores": 1, "kind": kind,
                               "sample": "determine_ploidy on chr21 + chr22 (%d bins), %.1f s" % (hi - lo, dt)}
        got = med.cpu().numpy()
        out["verified"] = bool(got[20] == want[0] and got[21] == want[1])
    del cov, gc
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------
# main (our arm)
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wgs30x", choices=sorted(WORKLOADS))
    ap.add_argument("--signals", type=int, default=0, help="override the workload's signal count")
    ap.add_argument("--cov-reads", type=int, default=617_653_966, help="reads for the coverage leg (30X = 617653966)")
    ap.add_argument("--no-coverage", action="store_true")
    ap.add_argument("--gc-bases", type=int, default=248_956_422, help="bases of the GC leg (chr1 of GRCh38)")
    ap.add_argument("--bam-reads", type=int, default=2_000_000, help="reads of the synthetic BAM of the bam_coverage leg")
    ap.add_argument("--no-graph", action="store_true", help="issue the step's kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--chunks", type=int, default=6, help="pair chunks of the pipelined host path (e2e, N=1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    ap.add_argument("--no-extra", action="store_true", help="skip the candidate-aggregation and ploidy-median legs")
    ap.add_argument("--ref-crops-per-core", type=int, default=1, help="--impl reference: 50k-signal crops per core per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from tiddit_b200 import build
    build.build()
    from tiddit_b200 import device_ops, engine, synth, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = peaks()
    w = WORKLOADS[args.workload]
    eps, m = w["eps"], w["m"]
    posA, posB, seg_off, L = getattr(synth, w["gen"])(args.signals or w["n"])
    n_total = len(posA)
    P_total = len(seg_off) - 1

    plan = engine.ShardPlan(seg_off, world)
    if world > 1:
        idx_h = plan.shard_index(rank)
        a_h, b_h, off_h = np.ascontiguousarray(posA[idx_h]), np.ascontiguousarray(posB[idx_h]), plan.shard_seg_off(rank)
    else:
        a_h, b_h, off_h = posA, posB, seg_off
    n_mine, P_mine = len(a_h), len(off_h) - 1
    pad = plan.pad

    a_pin = torch.from_numpy(a_h).pin_memory()
    b_pin = torch.from_numpy(b_h).pin_memory()
    off_pin = torch.from_numpy(off_h).pin_memory()
    a_d, b_d, off_d = a_pin.cuda(), b_pin.cuda(), off_pin.cuda()
    labels_d = torch.empty(pad, dtype=torch.int32, device="cuda")
    gathered = torch.empty(world * pad, dtype=torch.int32, device="cuda") if world > 1 else None
    out_pin = torch.empty(world * pad if world > 1 else n_total, dtype=torch.int32).pin_memory()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        flush_buf.add_(1)          # 256 MB read+write > the 126 MB L2

    runner = None if args.no_graph else engine.GraphRunner(a_d, b_d, off_d, P_mine, eps, m, L, labels_d[:n_mine])

    def step_device():
        if runner is not None:   # the call's kernels replayed from a CUDA graph (same launches, no host latency)
            runner.replay()
        else:
            device_ops.cluster_labels_device(a_d, b_d, off_d, P_mine, eps, m, L, labels_out=labels_d[:n_mine])
        if world > 1:
            dist.all_gather_into_tensor(gathered, labels_d)

    # e2e: every rank pipelines ITS shard (pinned host arrays -> H2D | kernels | D2H of the shard's labels into the
    # rank's own pinned buffer, chunked over three streams); with N > 1 the labels are then all-gathered on the device
    # like in the device-resident step, so every GPU ends up with every label.
    pipe = engine.HostPipeline(max(pad, n_mine, 1), n_chunks=args.chunks if world == 1 else max(2, args.chunks // 2))
    shard_pin = out_pin if world == 1 else torch.empty(max(n_mine, 1), dtype=torch.int32).pin_memory()

    def step_e2e():
        pipe.run(a_pin, b_pin, off_h, eps, m, L, shard_pin)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pipe.lab_d[:pad])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            flush()
            if world > 1:
                dist.barrier()
            s.record()
            fn()
            e.record()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in ev)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    c0 = _lib.launch_count()
    device_ops.cluster_labels_device(a_d, b_d, off_d, P_mine, eps, m, L, labels_out=labels_d[:n_mine])
    launches_per_call = _lib.launch_count() - c0
    launches0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        ms_step = timed(step_device, args.steps, args.warmup)
        launches = (_lib.launch_count() - launches0) // (args.steps + args.warmup)
        if runner is not None:
            runner.check()
            launches = launches_per_call   # replayed launches are not re-issued through the library's counter
        ms_e2e = timed(step_e2e, max(3, args.steps // 2), 3)
    clocks = clk.summary()

    # per-stage device times of one more (untimed) pass -> roofline of the eps-range-query kernel
    stage_ms = {}
    reps = 5
    for _ in range(reps):
        flush()
        torch.cuda.synchronize()
        _lib.profile_begin()
        device_ops.cluster_labels_device(a_d, b_d, off_d, P_mine, eps, m, L, labels_out=labels_d[:n_mine])
        for name, ms in _lib.profile_end():
            stage_ms[name] = stage_ms.get(name, 0.0) + ms / reps
    tot_stage = sum(stage_ms.values()) or 1.0
    k_ms = stage_ms.get("window_runs_x", float("nan"))
    alg = 8.0 * n_mine
    n_pass = (max(int(L), 1).bit_length() + 7) // 8
    sx_ms = stage_ms.get("sort_x", float("nan"))
    sort_bytes = (4.0 + 16.0 * n_pass) * n_mine
    same_shape = world == 1 and args.workload == "wgs30x" and not args.signals   # the shape the ncu capture was taken on
    roofline = {"bound": "hbm", "kernel": "window_runs_small_kernel<X, two-phase> (eps-range query + run labelling, posA axis; "
                                            "its tile sums are scanned by wr_tile_scan_kernel = stage tile_scan_x)",
                "achieved": alg / k_ms / 1e6, "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": alg / k_ms / 1e6 / hbm_peak,
                "traffic": ncu_traffic("window_runs_small_kernel<0") if same_shape else None,
                "traffic_source": "profiles/*_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of this kernel's "
                                  "launch in the committed ncu --set full capture of the same 20M-signal set",
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": k_ms,
                "share_of_step": k_ms / tot_stage,
                "stages_ms": {k: round(v, 4) for k, v in stage_ms.items()},
                "largest_stage": {"name": "sort_x (segmented radix sort of posA: histogram + %d passes)" % n_pass,
                                  "ms": sx_ms, "share_of_step": sx_ms / tot_stage,
                                  "bytes_moved": sort_bytes, "achieved": sort_bytes / sx_ms / 1e6,
                                  "frac": sort_bytes / sx_ms / 1e6 / hbm_peak,
                                  "note": "implementation traffic (4 B/key histogram read + 16 B/element per pass), "
                                          "not algorithmic bytes"},
                "pipeline": {"algorithmic_bytes": 12.0 * n_mine, "achieved": 12.0 * n_mine / ms_step / 1e6,
                             "frac": 12.0 * n_mine / ms_step / 1e6 / hbm_peak}}

    line = {"metric": METRIC, "value": n_total / ms_step * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": w["desc"], "signals": n_total, "pairs": P_total, "eps": eps, "min_pts": m,
                       "sharding": "pairs LPT over %d ranks, one all-gather of int32 labels" % world if world > 1
                       else "single GPU", "l2": "256 MB flush between timed steps (and inputs > L2 at N=1)",
                       "launch": "eager" if args.no_graph else "CUDA-graph replay of the ABI call's kernels"},
            "e2e": {"value": n_total / ms_e2e * 1e3, "unit": UNIT, "ms_per_step": ms_e2e,
                    "path": "engine.HostPipeline: %d tapered pair chunks, H2D | CUDA-graph replay of the chunk kernels | D2H on 3 streams, one host sync" % args.chunks
                    if world == 1 else "per rank: engine.HostPipeline over its shard (H2D | kernels | D2H of the shard's "
                    "labels to the rank's pinned buffer), then the device all-gather of all labels",
                    "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes + off_h.nbytes) * (1 if world == 1 else 1),
                    "d2h_bytes_per_step": int(n_mine * 4),
                    "bytes_note": "per rank (rank 0 shown)" if world > 1 else "whole job"},
            "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
            "clocks": clocks, "roofline": roofline}

    # N > 1, secondary figure: fixed work PER GPU (every rank clusters a whole 20M-signal set of its own -- N samples
    # in flight, no collective), next to the headline strong-scaling number of BASELINE configs[3]
    if world > 1:
        a_full = torch.from_numpy(posA).cuda()
        b_full = torch.from_numpy(posB).cuda()
        off_full = torch.from_numpy(seg_off).cuda()
        lab_full = torch.empty(n_total, dtype=torch.int32, device="cuda")
        full_runner = None if args.no_graph else engine.GraphRunner(a_full, b_full, off_full, P_total, eps, m, L, lab_full)

        def step_full():
            if full_runner is not None:
                full_runner.replay()
            else:
                device_ops.cluster_labels_device(a_full, b_full, off_full, P_total, eps, m, L, labels_out=lab_full)

        ms_full = timed(step_full, args.steps, args.warmup)
        if full_runner is not None:
            full_runner.check()
        line["weak_scaling"] = {"value": world * n_total / ms_full * 1e3, "unit": UNIT, "ms_per_step": ms_full,
                                "work": "every rank clusters its own copy-sized set of %d signals (N sets in flight), "
                                        "no collective; max over ranks" % n_total}
        del a_full, b_full, off_full, lab_full, full_runner

    # result check of what the timed paths produced (the oracle as checker, bounded to the sample pairs' cost)
    torch.cuda.synchronize()
    if world > 1:
        shard_ok = bool(torch.equal(shard_pin[:n_mine], labels_d[:n_mine].cpu()))   # e2e shard == device-resident shard
        out_pin.copy_(gathered)          # the all-gathered labels of the last e2e step
        torch.cuda.synchronize()
    if rank == 0:
        from oracle import oracle
        got_all = out_pin.numpy()[plan.gather_index()] if world > 1 else out_pin.numpy()
        want_all = oracle.cluster_segments(posA, posB, seg_off, eps, m)
        line["verified"] = bool(np.array_equal(got_all, want_all)) and (world == 1 or shard_ok)
    if world == 1 and rank == 0:
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_cluster_baseline(posA, posB, seg_off, eps, m)
        if not args.no_coverage:
            torch.cuda.empty_cache()
            try:
                line["coverage"] = coverage_leg(torch, args, hbm_peak, flush)
            except torch.cuda.OutOfMemoryError as exc:
                line["coverage"] = {"error": "out of memory: %s" % exc}
        if not args.no_extra:
            _lib.release_workspaces()
            torch.cuda.empty_cache()
            line["aggregate"] = aggregate_leg(torch, args, posA, posB, seg_off, L, eps, m, hbm_peak, flush)
            _lib.release_workspaces()
            torch.cuda.empty_cache()
            line["ploidy_medians"] = medians_leg(torch, args, hbm_peak, flush)
            line["gc"] = gc_leg(torch, args, hbm_peak, flush)
            line["bam_coverage"] = bam_leg(torch, args)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
