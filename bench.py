#!/usr/bin/env python
"""bench.py -- signals clustered / s (+ coverage bins / s) of the TIDDIT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload wgs30x|config2|tumor60x]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the clustering hot path (all (chrA,chrB) pairs: both segmented sorts, both eps-range-query /
run-labelling kernels, id assignment, scatter) over the synthetic 30X-WGS-shaped signal set of BASELINE.json
configs[2] (20 M signals, 300 pairs, eps=500, m=3).  With N > 1 the pairs are sharded over the ranks (LPT by signal
count, no data-path collective) and the labels are exchanged once per step so that every GPU holds every label
(BASELINE.json configs[3]; total work fixed => "scaling": "strong").  Rank 0 prints ONE JSON line.

value        device-resident inputs, CUDA events, max over ranks
e2e          the same through the host front end: pinned host arrays -> H2D -> kernels -> D2H labels, every step
roofline     the eps-range-query kernel (window_runs<X>): 8 B/signal / its event-timed duration
cpu_baseline the REAL reference (oracle/_ref, compiled from the upstream sources) or the C port, 1 core,
             on a bounded sample of whole pairs of the same workload
legs         N = 1: coverage, gc, tumor60x (BASELINE configs[4]), aggregate, ploidy_medians, cluster_main (tab files ->
             tiddit_cluster.main -> dict next to the reference), bam_coverage -- each with device-resident and
             end-to-end figures, its roofline, CPU baselines (1 core and all cores) and an oracle check of the TIMED
             output.  N > 1: tumor60x sharded like the headline, sharded_coverage / sharded_gc parity.

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
                     desc="60X tumor-like synthetic set: 50M signals, eps=1000, m=5 (BASELINE configs[4])"),
}
METRIC = "signals_clustered_per_sec"
UNIT = "signals/s"
REFERENCE_SAMPLE = ("bounded sample (posA-window crops of 50k signals from the largest pairs + whole small pairs): "
                    "same_config false -- whole large pairs cost the reference minutes to hours (O(#clusters*n) y-pass), "
                    "so its signals/s on the full set is LOWER than on this sample")


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
                      "%.1f s on 1 core; flatters the CPU (see config.reference_sample)" % (total, n_crops, secs)}


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
            "config": {"workload": w["desc"], "eps": w["eps"], "min_pts": w["m"], "same_config": False,
                       "reference_sample": REFERENCE_SAMPLE},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


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


def _event_ms(torch, fn, steps, warmup, flush, pre=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(warmup + steps):
        if pre is not None:
            pre()
        flush()
        ev[0].record()
        fn()
        ev[1].record()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(ev[0].elapsed_time(ev[1]))
    return float(np.mean(times))


# ---------------------------------------------------------------------------------------------------
# coverage / GC CPU baselines on all host cores (per-contig pools like tiddit_signal.pyx:259, tiddit_gc.pyx:36)
# ---------------------------------------------------------------------------------------------------
def _cov_worker(job):
    ln, z, k, seed = job
    rng = np.random.default_rng(seed)
    s = np.sort(rng.integers(0, ln - 150, k))
    hs, he = s.tolist(), (s + 150).tolist()
    R = _ref_modules()
    t0 = time.perf_counter()
    if R is not None:
        cov, e0 = R.tiddit_coverage.create_coverage({"SQ": [{"SN": "c", "LN": int(ln)}]}, z, "c")
        upd = R.tiddit_coverage.update_coverage
        for a, b in zip(hs, he):
            upd(a, b, z, cov, e0)
    else:
        from oracle import oracle
        cov, e0 = oracle.create_coverage({"SQ": [{"SN": "c", "LN": int(ln)}]}, z, "c")
        oracle.update_coverage_batch(s, s + 150, z, cov, e0)
    return time.perf_counter() - t0


def _gc_worker(job):
    path, z = job
    R = _ref_modules()
    R.tiddit_gc.binned_gc(path, "c", z, 0.5)         # warms the stand-in FastaFile's cache: parsing is not timed
    t0 = time.perf_counter()
    R.tiddit_gc.binned_gc(path, "c", z, 0.5)
    return time.perf_counter() - t0


def _pool_rate(worker, jobs, units_per_job):
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores) as pool:
        list(pool.map(worker, jobs, chunksize=1))
    dt = time.perf_counter() - t0
    return units_per_job * len(jobs) / dt, cores, dt


# ---------------------------------------------------------------------------------------------------
# coverage leg (N = 1)
# ---------------------------------------------------------------------------------------------------
def coverage_leg(torch, args, hbm_peak, flush):
    from tiddit_b200 import device_ops, engine, synth
    from oracle import oracle
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
    read_off_h = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    read_off = torch.from_numpy(read_off_h).cuda()
    bin_off_h = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    bin_off = torch.from_numpy(bin_off_h).cuda()
    ebs = torch.from_numpy((lens - (nb - 1) * z).astype(np.int32)).cuda()
    n_bins = int(bin_off_h[-1])
    bins = torch.zeros(n_bins, dtype=torch.float64, device="cuda")
    bad = device_ops.new_first_bad(torch)
    ms = _event_ms(torch, lambda: device_ops.coverage_accumulate_contigs_device(start, end, read_off, bin_off, ebs, z, bins, bad),
                   args.steps, args.warmup, flush, pre=bins.zero_)
    alg_bytes = 8.0 * n_reads + 8.0 * n_bins
    out = {"reads": n_reads, "bins": n_bins, "bin_size": z, "ms_per_step": ms, "reads_per_sec": n_reads / ms * 1e3,
           "bins_per_sec": n_bins / ms * 1e3,
           "roofline": {"bound": "hbm", "achieved": alg_bytes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / ms / 1e6 / hbm_peak, "kernel": "coverage_kernel",
                        "traffic_note": ncu_traffic("coverage_kernel<1>", note_only=True),
                        "algorithmic_bytes": "8 B/read + 8 B/bin"}}
    timed_bins = bins.cpu().numpy()          # the output of the last TIMED call (zeroed before it)
    out["mean_coverage"] = float(timed_bins.mean())
    # parity of the timed output: whole contigs through the C oracle, smallest first, ~20 s of CPU at most
    t_start, ok, checked, checked_reads = time.perf_counter(), True, [], 0
    for c in np.argsort(lens, kind="stable"):
        if time.perf_counter() - t_start > args.verify_seconds and checked:
            break
        lo, hi = int(read_off_h[c]), int(read_off_h[c + 1])
        want = np.zeros(int(nb[c]))
        oracle.update_coverage_batch(start[lo:hi].cpu().numpy(), end[lo:hi].cpu().numpy(), z, want,
                                     int(lens[c] - (nb[c] - 1) * z))
        ok = ok and np.array_equal(timed_bins[bin_off_h[c]:bin_off_h[c + 1]].view(np.uint64), want.view(np.uint64))
        checked.append(synth.GRCH38[c][0])
        checked_reads += hi - lo
    out["verified"] = bool(ok)
    out["verified_on"] = "the timed call's bins of %d whole contigs (%d reads) vs oracle.update_coverage_batch, bit for bit" % (
        len(checked), checked_reads)
    # e2e: pinned host (start, end) -> chunked H2D || kernel -> bins D2H (engine.coverage_host)
    try:
        s_pin = torch.empty(n_reads, dtype=torch.int32).pin_memory()
        e_pin = torch.empty(n_reads, dtype=torch.int32).pin_memory()
        s_pin.copy_(start)
        e_pin.copy_(end)
        out_pin = torch.empty(n_bins, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        del start, end
        torch.cuda.empty_cache()
        e2e_times = []
        for it in range(3):
            t0 = time.perf_counter()
            engine.coverage_host(s_pin, e_pin, read_off_h, lens, z, out=out_pin)
            torch.cuda.synchronize()
            if it:
                e2e_times.append(time.perf_counter() - t0)
        e2e_ms = float(np.mean(e2e_times)) * 1e3
        out["e2e"] = {"ms_per_step": e2e_ms, "reads_per_sec": n_reads / e2e_ms * 1e3, "bins_per_sec": n_bins / e2e_ms * 1e3,
                      "h2d_bytes_per_step": int(8 * n_reads), "d2h_bytes_per_step": int(8 * n_bins),
                      "h2d_GBps": 8.0 * n_reads / e2e_ms / 1e6,
                      "verified": bool(np.array_equal(out_pin.numpy().view(np.uint64), timed_bins.view(np.uint64))),
                      "path": "engine.coverage_host: pinned (start,end) -> 32M-read chunks, H2D of chunk k+1 || "
                              "tdt_coverage_accumulate_contigs on chunk k -> bins D2H; PCIe-bound"}
        del s_pin, e_pin, out_pin
    except (RuntimeError, MemoryError) as exc:
        out["e2e"] = {"error": str(exc)[:200]}
    if not args.no_cpu:
        ln0, k = int(lens[0]), 2_000_000
        dt = _cov_worker((ln0, z, k, 7))
        R = _ref_modules()
        out["cpu_baseline"] = {"value": k / dt, "unit": "reads/s", "cores": 1, "kind": "reference" if R is not None else "port",
                               "sample": "%d reads of a chr1-sized contig, one update_coverage call per read, %.1f s" % (k, dt)}
        cores = os.cpu_count() or 1
        kk = 1_000_000
        rate, cores, dtp = _pool_rate(_cov_worker, [(int(lens[i % len(lens)]), z, kk, 100 + i) for i in range(cores)], kk)
        out["cpu_baseline_all_cores"] = {"value": rate, "unit": "reads/s", "cores": cores,
                                         "kind": "reference" if R is not None else "port",
                                         "sample": "%d contigs x %d reads, one process per contig like the joblib fan-out of "
                                                   "tiddit_signal.pyx:259, %.1f s wall" % (cores, kk, dtp)}
    del bins
    torch.cuda.empty_cache()
    return out


def gc_leg(torch, args, hbm_peak, flush):
    """GC bins (tiddit_gc.pyx:6-33) of a chr1-sized contig, bin 50 like `--sv`: bases/s and bins/s, device-resident
    and from pinned host bytes."""
    from tiddit_b200 import device_ops, engine, synth
    n_bases = args.gc_bases
    seq_h = synth.fasta_sequence(n_bases)
    seq, ln = device_ops.padded_sequence_device(seq_h)
    z = 50
    n_bins = (n_bases + z - 1) // z
    out_bins = torch.zeros(n_bins, dtype=torch.int8, device="cuda")
    ms = _event_ms(torch, lambda: device_ops.gc_bins_device(seq, ln, z, 0.5, out=out_bins), args.steps, args.warmup, flush)
    alg = float(n_bases + n_bins)
    out = {"bases": n_bases, "bins": n_bins, "bin_size": z, "ms_per_step": ms, "bases_per_sec": n_bases / ms * 1e3,
           "bins_per_sec": n_bins / ms * 1e3,
           "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg / ms / 1e6 / hbm_peak, "kernel": "gc_small_kernel",
                        "algorithmic_bytes": "1 B/base + 1 B/bin"}}
    from oracle import oracle
    want = oracle.gc_bins(seq_h, z, 0.5)                       # the whole contig: the C oracle takes ~1 s
    out["verified"] = bool(np.array_equal(out_bins.cpu().numpy(), want))
    seq_pin = torch.from_numpy(seq_h).pin_memory()
    out_pin = torch.empty(n_bins, dtype=torch.int8).pin_memory()
    e2e_times = []
    for it in range(4):
        t0 = time.perf_counter()
        engine.gc_host(seq_pin, z, 0.5, out=out_pin)
        torch.cuda.synchronize()
        if it:
            e2e_times.append(time.perf_counter() - t0)
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    out["e2e"] = {"ms_per_step": e2e_ms, "bases_per_sec": n_bases / e2e_ms * 1e3, "h2d_bytes_per_step": int(n_bases),
                  "d2h_bytes_per_step": int(n_bins), "h2d_GBps": n_bases / e2e_ms / 1e6,
                  "verified": bool(np.array_equal(out_pin.numpy(), want)),
                  "path": "engine.gc_host: pinned FASTA bytes -> 32 MB chunks, H2D of chunk k+1 || tdt_gc_bins on chunk k -> int8 bins D2H"}
    if not args.no_cpu:
        R = _ref_modules()
        if R is not None and getattr(R, "tiddit_gc", None) is not None:
            import shutil
            import tempfile
            kk = 3_000_000
            tmp = tempfile.mkdtemp(prefix="tdt_bench_gc_")
            fa = os.path.join(tmp, "ref.fa")
            with open(fa, "w") as f:
                f.write(">c\n")
                text = bytes(seq_h[20_000_000:20_000_000 + kk]).decode("ascii")
                f.write("\n".join(text[i:i + 60] for i in range(0, kk, 60)) + "\n")
            dt = _gc_worker((fa, z))
            out["cpu_baseline"] = {"value": kk / dt, "unit": "bases/s", "cores": 1, "kind": "reference",
                                   "sample": "binned_gc on %d bases (FASTA already in memory), %.1f s" % (kk, dt)}
            rate, cores, dtp = _pool_rate(_gc_worker, [(fa, z)] * (os.cpu_count() or 1), kk)
            out["cpu_baseline_all_cores"] = {"value": rate, "unit": "bases/s", "cores": cores, "kind": "reference",
                                             "sample": "%d processes x binned_gc on %d bases each (per-contig pool of "
                                                       "tiddit_gc.pyx:36; includes each process' FASTA load), %.1f s wall" % (cores, kk, dtp)}
            shutil.rmtree(tmp, ignore_errors=True)
    del seq, out_bins
    torch.cuda.empty_cache()
    return out


def bam_leg(torch, args):
    """`--cov` from a BAM FILE: libtdt_bam.so (BGZF inflated on the host cores, records as columns)
    + the coverage kernel batch by batch, host->device copies and the final device->host read of the bins included.
    The synthetic BAM (150-bp reads with bases and qualities, coordinate-sorted) is written once, untimed."""
    import shutil
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
           "path": "coverage_from_bam: BGZF inflate (csrc/tdt_inflate.h block decoder + zlib CRC32, all host cores, 16 MiB "
                   "windows) -> columns -> H2D -> coverage kernel per window batch -> bins D2H",
           "bound": "host: inflate of the BGZF blocks + the sequential record walk"}
    from oracle import oracle
    keep = ((flag & 0x400) == 0) & (mapq >= q)
    ok = True
    for ci, (name, ln) in enumerate(contigs):
        sel = keep & (rid == ci)
        want, ebs = oracle.create_coverage({"SQ": [{"SN": name, "LN": ln}]}, z, name)
        oracle.update_coverage_batch(s[sel], np.minimum(s[sel] + 150, 2 ** 31 - 1).astype(np.int32), z, want, ebs)
        ok = ok and np.array_equal(cov[name].view(np.uint64), want.view(np.uint64))
    out["verified"] = bool(ok)
    shutil.rmtree(tmp, ignore_errors=True)
    return out


# ---------------------------------------------------------------------------------------------------
# candidate aggregation + ploidy medians legs (N = 1)
# ---------------------------------------------------------------------------------------------------
def aggregate_leg(torch, args, posA, posB, seg_off, L, eps, m, hbm_peak, flush):
    """tdt_cluster_aggregate on the step's own labels: candidates/s, stage times, CPU port on a bounded sample; e2e =
    tiddit_cluster.cluster_packed (host arrays -> labels -> candidate table back on the host)."""
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
    stages = {}
    c0 = _lib.launch_count()
    run()
    launches = _lib.launch_count() - c0
    ms = _event_ms(torch, run, args.steps, args.warmup, flush)
    for _ in range(3):
        flush()
        torch.cuda.synchronize()
        _lib.profile_begin()
        run()
        for k, v in _lib.profile_end():
            stages[k] = stages.get(k, 0.0) + v / 3
    C, M, err, _ = (int(v) for v in counts.cpu().tolist())
    alg = 33.0 * n + 4.0 * M + 64.0 * C
    out = {"signals": n, "candidates": C, "members": M, "ms_per_step": ms, "signals_per_sec": n / ms * 1e3,
           "candidates_per_sec": C / ms * 1e3, "gpu_launches_per_step": int(launches), "data_error": err,
           "stages_ms": {k: round(v, 4) for k, v in stages.items()},
           "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg / ms / 1e6 / hbm_peak,
                        "algorithmic_bytes": "33 B/signal in + 4 B/member + 64 B/candidate out; the sorts are implementation traffic"}}
    # e2e through the package API: packed host arrays -> cluster_packed -> (labels, CandidateTable) on the host
    from tiddit_b200 import tiddit_cluster
    from tiddit_b200.signals import PackedSignals
    try:
        packed = PackedSignals.from_arrays(posA, posB, seg_off, rec, synth.GRCH38)
        ts = []
        for it in range(3):
            t0 = time.perf_counter()
            _, table = tiddit_cluster.cluster_packed(packed, eps, m, max_ins, is_mp, min_reads)
            if it:
                ts.append(time.perf_counter() - t0)
        e2e_ms = float(np.mean(ts)) * 1e3
        in_bytes = int(posA.nbytes + posB.nbytes + rec["span"].nbytes + rec["name_id"].nbytes + rec["flags"].nbytes)
        out["e2e"] = {"ms_per_step": e2e_ms, "signals_per_sec": n / e2e_ms * 1e3, "h2d_bytes_per_step": in_bytes,
                      "d2h_bytes_per_step": int(len(table.rows) * 64 + len(table.member_idx) * 4),
                      "path": "tiddit_cluster.cluster_packed: pageable host arrays -> tdt_cluster_labels -> "
                              "tdt_cluster_aggregate (labels stay in HBM) -> rows + member index to the host",
                      "candidates": int(len(table))}
    except Exception as exc:
        out["e2e"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
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
                                         "tiddit_cluster.pyx:156-336, %.2f s; the reference does this fold in "
                                         "interpreted Python per signal (see the cluster_main leg)" % (k, hi, dt)}
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
    ms = _event_ms(torch, lambda: device_ops.coverage_medians_device(cov, gc, off_d, len(nb), med, cnt), args.steps,
                   args.warmup, flush)
    out = {"bins": n, "contigs": len(nb), "ms_per_step": ms, "bins_per_sec": n / ms * 1e3,
           "roofline": {"bound": "hbm", "achieved": 9.0 * n / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": 9.0 * n / ms / 1e6 / hbm_peak,
                        "algorithmic_bytes": "9 B/bin (float64 coverage + int8 GC), read once; the radix select streams them 3 times (2 digit passes + 1 compaction of the prefix buckets)"}}
    # e2e: the arrays on the host (as tiddit_signal.main / tiddit_gc.main return them) -> medians on the host
    cov_h, gc_h = cov.cpu().numpy(), gc.cpu().numpy()
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        device_ops.coverage_medians(cov_h, gc_h, off)
        if it:
            ts.append(time.perf_counter() - t0)
    e2e_ms = float(np.mean(ts)) * 1e3
    out["e2e"] = {"ms_per_step": e2e_ms, "bins_per_sec": n / e2e_ms * 1e3, "h2d_bytes_per_step": int(9 * n),
                  "d2h_bytes_per_step": int(16 * (len(nb) + 1)),
                  "path": "device_ops.coverage_medians: pageable host bins + GC -> H2D -> tdt_coverage_medians -> medians D2H"}
    if not args.no_cpu:
        R = _ref_modules()
        lo, hi = int(off[20]), int(off[22])                                  # chr21 + chr22: 1.95 M bins
        cd = {nm: cov_h[int(off[20 + i]):int(off[21 + i])] for i, nm in enumerate(["chr21", "chr22"])}
        gd = {nm: gc_h[int(off[20 + i]):int(off[21 + i])] for i, nm in enumerate(["chr21", "chr22"])}
        import tempfile
        t0 = time.perf_counter()
        if R is not None and hasattr(R, "tiddit_coverage_analysis"):
            lib = R.tiddit_coverage_analysis.determine_ploidy(cd, ["chr21", "chr22"], {}, 2, os.path.join(tempfile.mkdtemp(), "p"), 0,
                                                              "", 50, {"SQ": []}, gd)
            want = [lib["avg_coverage_chr21"], lib["avg_coverage_chr22"]]
            kind = "reference"
        else:
            from oracle import oracle
            m_, _ = oracle.coverage_medians(cov_h[lo:hi], gc_h[lo:hi], off[20:23] - off[20])
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


def cluster_main_leg(torch, args):
    """The reference's own entry point end to end: tab files on disk -> tiddit_cluster.main(prefix, ...)
    -> candidates dict, ours next to the compiled reference's on the SAME files."""
    import shutil
    import tempfile
    from tiddit_b200 import synth, tiddit_cluster
    n = args.main_lines
    posA, posB, seg_off, L = synth.wgs30x_signals(n)
    rec = synth.signal_records(posA, posB, seg_off)
    tmp = tempfile.mkdtemp(prefix="tdt_bench_main_")
    prefix = os.path.join(tmp, "s")
    synth.write_tab_files(prefix, "S", posA, posB, seg_off, rec)
    chrom, clen = [c for c, _ in synth.GRCH38], dict(synth.GRCH38)
    call = (prefix, chrom, clen, ["S"], False, 500, 3, 5000, 10000, True, 3)
    tiddit_cluster.main(*call)                        # warm-up: library load, pinned pools
    t0 = time.perf_counter()
    got = tiddit_cluster.main(*call)
    dt = time.perf_counter() - t0
    n_cand = sum(len(got[a][b]) for a in got for b in got[a])
    out = {"lines": n, "candidates": n_cand, "s_per_call": dt, "lines_per_sec": n / dt,
           "path": "tiddit_cluster.main(prefix, ...): native tab scanner (libtdt_tab.so, all host cores) -> packed arrays -> 2 GPU calls -> candidates dict (Python)"}
    # where the time goes
    from tiddit_b200.signals import PackedSignals
    t0 = time.perf_counter()
    packed = PackedSignals.from_tab(prefix, chrom, clen, ["S"], False, 10000, True)
    t1 = time.perf_counter()
    _, table = tiddit_cluster.cluster_packed(packed, 500, 3, 5000, False, 3)
    t2 = time.perf_counter()
    tiddit_cluster.candidates_from_table(packed, table)
    t3 = time.perf_counter()
    out["stages_s"] = {"parse_tab": round(t1 - t0, 3), "gpu_cluster_packed": round(t2 - t1, 3), "build_dict": round(t3 - t2, 3)}
    if not args.no_cpu:
        R = _ref_modules()
        if R is not None:
            t0 = time.perf_counter()
            want = R.tiddit_cluster.main(*call)
            dtr = time.perf_counter() - t0
            same = list(want) == list(got) and all(list(want[a]) == list(got[a]) and all(
                list(want[a][b]) == list(got[a][b]) for b in want[a]) for a in want)
            if same:                                   # numbers of every candidate
                for a in want:
                    for b in want[a]:
                        for cid, c in want[a][b].items():
                            g = got[a][b][cid]
                            if any(g[k] != c[k] for k in ("N_discordants", "N_splits", "N_contigs", "posA", "posB", "startA",
                                                          "endA", "startB", "endB")) or g["discordants"] != c["discordants"]:
                                same = False
            out["cpu_baseline"] = {"value": n / dtr, "unit": "lines/s", "cores": 1, "kind": "reference",
                                   "sample": "the compiled reference's tiddit_cluster.main on the same %d-line tab files "
                                             "(same_config true), %.1f s" % (n, dtr)}
            out["verified"] = bool(same)
            out["speedup_vs_reference"] = dtr / dt
    shutil.rmtree(tmp, ignore_errors=True)
    return out


# ---------------------------------------------------------------------------------------------------
# one clustering workload: device-resident + e2e (+ the label exchange at N > 1) + parity of what was timed
# ---------------------------------------------------------------------------------------------------
def cluster_workload(ctx, wname, steps, warmup, detail):
    torch, dist, args = ctx["torch"], ctx["dist"], ctx["args"]
    world, rank, flush = ctx["world"], ctx["rank"], ctx["flush"]
    from tiddit_b200 import device_ops, engine, synth, _lib
    w = WORKLOADS[wname]
    eps, m = w["eps"], w["m"]
    posA, posB, seg_off, L = getattr(synth, w["gen"])((args.signals if wname == args.workload else 0) or w["n"])
    n_total, P_total = len(posA), len(seg_off) - 1
    plan = engine.ShardPlan(seg_off, world)
    if world > 1:
        idx_h = plan.shard_index(rank)
        a_h, b_h, off_h = np.ascontiguousarray(posA[idx_h]), np.ascontiguousarray(posB[idx_h]), plan.shard_seg_off(rank)
    else:
        a_h, b_h, off_h = posA, posB, seg_off
    n_mine, P_mine = len(a_h), len(off_h) - 1
    pad = plan.pad
    a_pin, b_pin = torch.from_numpy(a_h).pin_memory(), torch.from_numpy(b_h).pin_memory()
    a_d, b_d, off_d = a_pin.cuda(), b_pin.cuda(), torch.from_numpy(off_h).cuda()
    exch = engine.LabelExchange(pad, world, rank) if world > 1 else None
    labels_d = exch.shard if world > 1 else torch.empty(pad, dtype=torch.int32, device="cuda")
    out_pin = torch.empty(world * pad if world > 1 else n_total, dtype=torch.int32).pin_memory()
    runner = None if args.no_graph else engine.GraphRunner(a_d, b_d, off_d, P_mine, eps, m, L, labels_d[:n_mine])

    def compute():
        if runner is not None:   # the call's kernels replayed from a CUDA graph (same launches, no host latency)
            runner.replay()
        else:
            device_ops.cluster_labels_device(a_d, b_d, off_d, P_mine, eps, m, L, labels_out=labels_d[:n_mine])

    def step_device():
        compute()
        if world > 1:
            exch.run()

    # e2e: every rank pipelines ITS shard (pinned host arrays -> H2D | kernels | D2H of the shard's labels into the
    # rank's own pinned buffer -- one CUDA graph with memcpy nodes); with N > 1 the labels are then exchanged on the
    # device like in the device-resident step, so every GPU ends up with every label.
    pipe = engine.HostPipeline(max(pad, n_mine, 1), n_chunks=args.chunks)
    shard_pin = out_pin if world == 1 else torch.empty(max(n_mine, 1), dtype=torch.int32).pin_memory()

    def step_e2e():
        pipe.run(a_pin, b_pin, off_h, eps, m, L, shard_pin)
        if world > 1:
            exch.shard[:n_mine].copy_(pipe.lab_d[:n_mine], non_blocking=True)
            exch.run()

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
    ms_step = timed(step_device, steps, warmup)
    if runner is not None:
        runner.check()
    # parity of what the DEVICE-RESIDENT timed path left behind (graph-replayed buffer / exchanged labels)
    torch.cuda.synchronize()
    if world > 1:
        dev_all = exch.gathered.cpu().numpy()
    else:
        dev_all = labels_d[:n_mine].cpu().numpy()
    ms_compute = timed(compute, max(3, steps // 2), 2) if world > 1 else ms_step
    ms_exch = timed(exch.run, max(3, steps // 2), 2) if world > 1 else 0.0
    ms_e2e = timed(step_e2e, max(3, steps // 2), 3)
    torch.cuda.synchronize()
    res = {"workload": w["desc"], "signals": n_total, "pairs": P_total, "eps": eps, "min_pts": m,
           "value": n_total / ms_step * 1e3, "ms_per_step": ms_step, "launches_per_step": int(launches_per_call) + (exch.launches if world > 1 else 0),
           "e2e": {"value": n_total / ms_e2e * 1e3, "unit": UNIT, "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes + off_h.nbytes), "d2h_bytes_per_step": int(n_mine * 4),
                   "bytes_note": "per rank (rank 0 shown)" if world > 1 else "whole job"},
           "n_mine": n_mine}
    if world > 1:
        res["exchange"] = {"kind": exch.kind, "ms": ms_exch, "compute_ms": ms_compute,
                           "bytes_out_per_rank": int(pad * 4 * (world - 1)), "bytes_in_per_rank": int(pad * 4 * (world - 1))}
        e2e_all = exch.gathered.cpu().numpy()
        shard_ok = bool(np.array_equal(shard_pin[:n_mine].numpy(), dev_all[rank * pad:rank * pad + n_mine]))
    if rank == 0:
        from oracle import oracle
        want_all = oracle.cluster_segments(posA, posB, seg_off, eps, m)
        if world > 1:
            gi = plan.gather_index()
            res["verified"] = bool(np.array_equal(dev_all[gi], want_all)) and bool(np.array_equal(e2e_all[gi], want_all)) and shard_ok
        else:
            res["verified"] = bool(np.array_equal(dev_all, want_all)) and bool(np.array_equal(out_pin.numpy(), want_all))
        res["verified_on"] = "labels left by the timed device-resident path AND by the timed e2e path vs oracle.cluster_segments, all %d signals" % n_total
    if detail:
        # per-stage device times of a few more (untimed) passes -> roofline of the eps-range-query kernel
        stage_ms, reps = {}, 5
        for _ in range(reps):
            flush()
            torch.cuda.synchronize()
            _lib.profile_begin()
            device_ops.cluster_labels_device(a_d, b_d, off_d, P_mine, eps, m, L, labels_out=labels_d[:n_mine])
            for name, ms in _lib.profile_end():
                stage_ms[name] = stage_ms.get(name, 0.0) + ms / reps
        if world > 1:
            stage_ms["label_exchange"] = ms_exch
        res["stage_ms"] = stage_ms
        res["L"] = L
    if world == 1 and rank == 0 and detail and not args.no_cpu:
        res["cpu_baseline"] = cpu_cluster_baseline(posA, posB, seg_off, eps, m)
    res["arrays"] = (posA, posB, seg_off, L) if detail else None
    del runner, pipe
    if exch is not None:
        exch.close()
    _lib.release_workspaces()
    torch.cuda.empty_cache()
    return res


def sharded_parity(ctx):
    """N > 1: engine.sharded_coverage / sharded_gc on a bounded set through NCCL, against the oracle (driver-visible
    evidence for SURVEY 8(e) rows 2-3 even when the test box has one GPU)."""
    torch, dist, rank = ctx["torch"], ctx["dist"], ctx["rank"]
    from tiddit_b200 import engine, synth
    contigs = synth.GRCH38[16:24]
    s, e, roff, lens = synth.coverage_reads(8_000_000, contigs=contigs)
    t0 = time.perf_counter()
    bins, bin_off = engine.sharded_coverage(s, e, roff, lens, 500)
    t_cov = time.perf_counter() - t0
    seqs = {name: synth.fasta_sequence(min(ln, 6_000_000) + 7 * i, seed=20 + i) for i, (name, ln) in enumerate(contigs)}
    t0 = time.perf_counter()
    gcs = engine.sharded_gc(seqs, 50, 0.5)
    t_gc = time.perf_counter() - t0
    out = {"coverage": {"reads": int(len(s)), "contigs": len(contigs), "s": round(t_cov, 3),
                        "how": "reads sliced over ranks, tdt_coverage_accumulate_contigs per rank, one all-reduce(sum) of float64 bins"},
           "gc": {"bases": int(sum(len(v) for v in seqs.values())), "contigs": len(seqs), "s": round(t_gc, 3),
                  "how": "contigs LPT over ranks, tdt_gc_bins per rank, one all-gather of int8 bins"}}
    if rank == 0:
        from oracle import oracle
        ok = True
        for c, (name, ln) in enumerate(contigs):
            want, ebs = oracle.create_coverage({"SQ": [{"SN": name, "LN": ln}]}, 500, name)
            oracle.update_coverage_batch(s[roff[c]:roff[c + 1]], e[roff[c]:roff[c + 1]], 500, want, ebs)
            ok = ok and np.array_equal(bins[bin_off[c]:bin_off[c + 1]].view(np.uint64), want.view(np.uint64))
        out["coverage"]["verified"] = bool(ok)
        out["gc"]["verified"] = bool(all(np.array_equal(gcs[k], oracle.gc_bins(v, 50, 0.5)) for k, v in seqs.items()))
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
    ap.add_argument("--main-lines", type=int, default=2_000_000, help="tab-file lines of the cluster_main leg")
    ap.add_argument("--verify-seconds", type=float, default=20.0, help="CPU budget of the coverage leg's oracle check")
    ap.add_argument("--no-graph", action="store_true", help="issue the step's kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--chunks", type=int, default=6, help="pair chunks of the pipelined host path (e2e)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    ap.add_argument("--no-extra", action="store_true", help="skip the legs next to the headline workload")
    ap.add_argument("--no-tumor", action="store_true", help="skip the 60X tumour leg (BASELINE configs[4])")
    ap.add_argument("--ref-crops-per-core", type=int, default=1, help="--impl reference: 50k-signal crops per core per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from tiddit_b200 import build
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from tiddit_b200 import _lib
    hbm_peak, peak_src = peaks()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        flush_buf.add_(1)          # 256 MB read+write > the 126 MB L2

    ctx = {"torch": torch, "dist": dist, "args": args, "world": world, "rank": rank, "flush": flush}
    with ClockSampler(local) as clk:
        main_res = cluster_workload(ctx, args.workload, args.steps, args.warmup, detail=True)
    clocks = clk.summary()
    w = WORKLOADS[args.workload]
    posA, posB, seg_off, L = main_res.pop("arrays")
    n_total, n_mine = main_res["signals"], main_res["n_mine"]
    stage_ms = main_res["stage_ms"]
    tot_stage = sum(v for k, v in stage_ms.items()) or 1.0
    k_ms = stage_ms.get("window_runs_x", float("nan"))
    alg = 8.0 * n_mine
    sx_ms = stage_ms.get("sort_x", float("nan"))
    same_shape = world == 1 and args.workload == "wgs30x" and not args.signals   # the shape the ncu capture was taken on
    roofline = {"bound": "hbm", "kernel": "window_runs_small_kernel<X, two-phase> (eps-range query + run labelling, posA axis)",
                "achieved": alg / k_ms / 1e6, "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": alg / k_ms / 1e6 / hbm_peak,
                "traffic": ncu_traffic("window_runs_small_kernel<0") if same_shape else None,
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": k_ms, "share_of_step": k_ms / tot_stage,
                "stages_ms": {k: round(v, 4) for k, v in stage_ms.items()},
                "largest_stage": {"name": "sort_x (segmented sort of posA inside every pair)", "ms": sx_ms,
                                  "share_of_step": sx_ms / tot_stage,
                                  "algorithmic_bytes": 12.0 * n_mine, "achieved": 12.0 * n_mine / sx_ms / 1e6,
                                  "frac": 12.0 * n_mine / sx_ms / 1e6 / hbm_peak,
                                  "note": "4 B key in + 8 B (key, index) out per signal over the stage time"},
                "pipeline": {"algorithmic_bytes": 12.0 * n_mine, "achieved": 12.0 * n_mine / main_res["ms_per_step"] / 1e6,
                             "frac": 12.0 * n_mine / main_res["ms_per_step"] / 1e6 / hbm_peak}}
    line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": w["desc"], "signals": n_total, "pairs": main_res["pairs"], "eps": w["eps"], "min_pts": w["m"],
                       "sharding": "pairs LPT over %d ranks, one exchange of the int32 labels (%s)" % (
                           world, main_res["exchange"]["kind"]) if world > 1 else "single GPU",
                       "l2": "256 MB flush between timed steps (and inputs > L2 at N=1)",
                       "launch": "eager" if args.no_graph else "CUDA-graph replay of the ABI call's kernels",
                       "same_config": False, "reference_sample": REFERENCE_SAMPLE},
            "e2e": dict(main_res["e2e"], path="engine.HostPipeline: tapered pair chunks, H2D | kernels | D2H on 3 streams, "
                                             "the whole call one CUDA graph with memcpy nodes, one host sync" +
                                             ("; then the device label exchange" if world > 1 else "")),
            "gpu_launches": int(main_res["launches_per_step"]) * args.steps,
            "gpu_launches_per_step": int(main_res["launches_per_step"]),
            "verified": main_res.get("verified"), "verified_on": main_res.get("verified_on"),
            "clocks": clocks}
    if world > 1:
        line["exchange"] = main_res["exchange"]
    if "cpu_baseline" in main_res:
        line["cpu_baseline"] = main_res["cpu_baseline"]
    line["roofline"] = roofline

    if world == 1 and rank == 0 and not args.no_extra:
        if not args.no_coverage:
            torch.cuda.empty_cache()
            try:
                line["coverage"] = coverage_leg(torch, args, hbm_peak, flush)
            except torch.cuda.OutOfMemoryError as exc:
                line["coverage"] = {"error": "out of memory: %s" % exc}
        line["gc"] = gc_leg(torch, args, hbm_peak, flush)
    if not args.no_tumor and not args.no_extra and args.workload == "wgs30x" and not args.signals:
        t = cluster_workload(ctx, "tumor60x", max(5, args.steps // 2), 3, detail=False)
        t.pop("arrays", None)
        t.pop("n_mine", None)
        if rank == 0:
            line["tumor60x"] = t
    if world > 1 and not args.no_extra:
        sp = sharded_parity(ctx)
        if rank == 0:
            line["sharded"] = sp
    if world == 1 and rank == 0 and not args.no_extra:
        _lib.release_workspaces()
        torch.cuda.empty_cache()
        line["aggregate"] = aggregate_leg(torch, args, posA, posB, seg_off, L, w["eps"], w["m"], hbm_peak, flush)
        _lib.release_workspaces()
        torch.cuda.empty_cache()
        line["ploidy_medians"] = medians_leg(torch, args, hbm_peak, flush)
        line["cluster_main"] = cluster_main_leg(torch, args)
        line["bam_coverage"] = bam_leg(torch, args)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
