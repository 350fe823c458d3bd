#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REAL reference (compiled from
/root/reference into oracle/_ref/ by oracle/build_ref.py).  Run in the build container only; the
outputs are committed so that the tests need neither /root/reference nor oracle/_ref.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, ref  # noqa: E402

build_ref.build()
R = ref.load()
assert R is not None, "the reference could not be compiled"


def dbscan_cases():
    """App. A.4 known-answer vectors + random small cases, labels from the real DBSCAN.py."""
    cases = []
    kats = [
        ([0, 1, 5, 9, 12, 14, 14], [0] * 7, 10, 3),
        ([1, 1, 1, 10], [2, 2, 2, 11], 0.1, 2),
        ([0, 1, 2, 3, 100, 101, 102, 103], [0] * 8, 10, 3),
        (list(range(8)), [0, 0, 0, 0, 1000, 1000, 1000, 1000], 10, 3),
        (list(range(8)), [0, 1000] * 4, 10, 3),
        (list(range(9)), [0, 0, 0, 500, 500, 500, 900, 900, 900], 10, 3),
        ([5, 5, 5], [7, 7, 7], 10, 3),
        ([5, 5, 5, 5], [7, 7, 7, 7], 10, 3),
    ]
    rng = np.random.default_rng(20261017)
    for t in range(400):
        n = int(rng.integers(0, 70))
        m = int(rng.integers(2, 7))
        eps = int(rng.integers(1, 40))
        span = int(rng.integers(5, 300))
        x = np.sort(rng.integers(0, span, n)).tolist()
        if t % 5 == 0 and n:   # unsorted x: DBSCAN.main does not sort, window max of |dx|
            x = rng.integers(0, span, n).tolist()
        y = rng.integers(0, span, n).tolist()
        kats.append((x, y, eps, m))
    for x, y, eps, m in kats:
        data = np.array([[a, b, i] for i, (a, b) in enumerate(zip(x, y))], dtype=np.int64).reshape(len(x), 3)
        if len(x) == 0:
            xl, cid, fl = [], -1, []
        else:
            xl, cid = R.DBSCAN.x_coordinate_clustering(data, eps, m)
            fl = R.DBSCAN.main(data, eps, m)
            xl, fl = [int(v) for v in xl], [int(v) for v in fl]
        cases.append({"x": x, "y": y, "eps": eps, "m": m, "x_labels": xl, "x_last_id": int(cid), "labels": fl})
    with open(os.path.join(HERE, "dbscan_small.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    # medium cases in the regimes that matter (sparse / at-threshold / dense / hotspot), stored as npz
    med = {}
    for k, (n, span, eps, m) in enumerate([(6000, 4_000_000, 500, 3), (6000, 1_000_000, 500, 3),
                                            (6000, 300_000, 500, 3), (5000, 2_000_000, 1000, 5),
                                            (4000, 50_000, 50, 2)]):
        r = np.random.default_rng(100 + k)
        nc = n // 20
        sizes = 2 + r.geometric(1 / 8, nc)
        cx, cy = r.integers(0, span, nc), r.integers(0, span, nc)
        ax = np.repeat(cx, sizes) + np.rint(r.normal(0, 120, sizes.sum())).astype(np.int64)
        ay = np.repeat(cy, sizes) + np.rint(r.normal(0, 120, sizes.sum())).astype(np.int64)
        hot = r.integers(0, span - 3000)
        ax = np.concatenate([ax, r.integers(0, span, n // 2), hot + r.integers(0, 2000, n // 8)])
        ay = np.concatenate([ay, r.integers(0, span, n // 2), hot + r.integers(0, 2000, n // 8)])
        ax, ay = np.clip(ax, 0, span), np.clip(ay, 0, span)
        perm = r.permutation(len(ax))
        ax, ay = ax[perm], ay[perm]
        order = np.argsort(ax, kind="stable")
        data = np.stack([ax[order], ay[order], np.arange(len(ax))], 1).astype(np.int64)
        lab = R.DBSCAN.main(data, eps, m)
        out = np.empty(len(ax), dtype=np.int32)
        out[order] = lab.astype(np.int32)
        med["posA_%d" % k], med["posB_%d" % k], med["labels_%d" % k] = ax.astype(np.int32), ay.astype(np.int32), out
        med["param_%d" % k] = np.array([eps, m])
    np.savez_compressed(os.path.join(HERE, "dbscan_medium.npz"), **med)


def coverage_cases():
    cov = R.tiddit_coverage
    out = []
    # App. B known-answer vector
    specs = [("c1", 1234, 500, [(0, 150), (400, 550), (990, 1234), (1100, 1234), (499, 501), (0, 1234)])]
    rng = np.random.default_rng(7)
    for t in range(40):
        z = int(rng.choice([1, 7, 37, 50, 100, 500, 1000, 4096]))
        ln = int(rng.integers(1, 30000 if z >= 37 else 1500))
        k = int(rng.integers(0, 400))
        s = np.sort(rng.integers(0, ln, k))
        length = rng.integers(1, 400, k) if t % 3 else rng.integers(1, 6000, k)
        e = np.minimum(s + length, ln)
        specs.append(("ctg%d" % t, ln, z, [(int(a), int(b)) for a, b in zip(s, e)]))
    for name, ln, z, reads in specs:
        header = {"SQ": [{"SN": name, "LN": ln}]}
        data, ebs = cov.create_coverage(header, z, name)
        for a, b in reads:
            data = cov.update_coverage(a, b, z, data, ebs)
        out.append({"name": name, "LN": ln, "bin": z, "reads": reads, "end_bin_size": int(ebs),
                    "bins_hex": [float(v).hex() for v in data]})
    with open(os.path.join(HERE, "coverage_cases.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    # text output of print_coverage (bed + wig) for a 3-contig header
    header = {"SQ": [{"SN": "c1", "LN": 1234}, {"SN": "c2", "LN": 500}, {"SN": "c3", "LN": 2001}]}
    data, ebs = cov.create_coverage(header, 500)
    r = np.random.default_rng(11)
    reads = {}
    for c in header["SQ"]:
        s = np.sort(r.integers(0, c["LN"], 60))
        e = np.minimum(s + r.integers(1, 300, 60), c["LN"])
        reads[c["SN"]] = [(int(a), int(b)) for a, b in zip(s, e)]
        for a, b in reads[c["SN"]]:
            cov.update_coverage(a, b, 500, data[c["SN"]], ebs[c["SN"]])
    cov.print_coverage(data, header, 500, "bed", os.path.join(HERE, "print_coverage.bed"))
    cov.print_coverage(data, header, 500, "wig", os.path.join(HERE, "print_coverage.wig"))
    with open(os.path.join(HERE, "print_coverage_input.json"), "w") as f:
        json.dump({"header": header, "bin": 500, "reads": reads}, f, separators=(",", ":"))


def gc_cases():
    out = []
    seqs = [("ACGTACGTACNNNNNNacgtNNNNNGGGGGATTTACAT", 10, 0.5), ("GGGATTTT", 50, 0.5)]
    rng = np.random.default_rng(9)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM", dtype=np.uint8)
    for t in range(30):
        ln = int(rng.integers(1, 4000))
        z = int(rng.choice([1, 3, 10, 16, 50, 64, 100, 191, 192, 193, 500, 1000]))
        p = np.array([6, 6, 6, 6, 3, 3, 3, 3, 4, 1, .2, .2, .2, .2])
        s = rng.choice(alphabet, ln, p=p / p.sum())
        k = int(rng.integers(0, ln))
        s[k:k + int(rng.integers(0, 300))] = ord("N")
        seqs.append((bytes(s).decode(), z, float(rng.choice([0.5, 0.0, 0.2, 1.0]))))
    tmp = os.path.join(HERE, "_tmp.fa")
    for k, (s, z, cut) in enumerate(seqs):
        with open(tmp, "w") as f:
            f.write(">ctg\n")
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
        import pysam  # the stand-in next to the compiled reference
        pysam.FastaFile._cache.clear()
        name, bins = R.tiddit_gc.binned_gc(tmp, "ctg", z, cut)
        out.append({"seq": s, "bin": z, "n_cutoff": cut, "gc": [int(v) for v in bins]})
    os.remove(tmp)
    with open(os.path.join(HERE, "gc_cases.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (set, frozenset)):
        return {"__set__": sorted(_jsonable(v) for v in o)}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    return o


def cluster_cases():
    """Synthetic <prefix>_tiddit/*.tab files (App. D formats) + the reference's candidates dict."""
    contigs = {"chr1": 60000, "chr2": 45000, "chr10": 30000, "tiny": 900}
    names = list(contigs)
    rng = np.random.default_rng(5)

    def tf():
        return "True" if rng.random() < 0.5 else "False"

    for case, (samples, is_mp, skip_assembly, eps, m, min_reads) in enumerate(
            [(["S1"], False, False, 300, 3, 3), (["S1", "S2"], True, False, 200, 2, 2), (["S1"], False, True, 300, 4, 5)]):
        prefix = os.path.join(HERE, "cluster_case%d" % case)
        os.makedirs(prefix + "_tiddit", exist_ok=True)
        centres = [(names[a], names[b], int(rng.integers(100, contigs[names[a]])), int(rng.integers(100, contigs[names[b]])))
                   for a, b in [(0, 0), (0, 0), (0, 1), (1, 1), (0, 2), (2, 2), (0, 0), (1, 2), (0, 3)]]
        for sample in samples:
            with open("%s_tiddit/discordants_%s.tab" % (prefix, sample), "w") as f:
                for k in range(260):
                    ca, cb, xa, xb = centres[int(rng.integers(0, len(centres)))]
                    if rng.random() < 0.25:
                        xa, xb = int(rng.integers(1, contigs[ca])), int(rng.integers(1, contigs[cb]))
                    sa = xa + int(rng.normal(0, 80)); sb = xb + int(rng.normal(0, 80))
                    sa, sb = max(1, sa), max(1, sb)
                    if rng.random() < 0.03:
                        sa = contigs[ca] + 50   # beyond the contig end: exercises the clamp
                    f.write("\t".join(["read%d_%s" % (k // 2, sample), ca, cb, str(sa), str(sa + 100), tf(),
                                       str(sb), str(sb + 100), tf()]) + "\n")
            with open("%s_tiddit/splits_%s.tab" % (prefix, sample), "w") as f:
                for k in range(140):
                    ca, cb, xa, xb = centres[int(rng.integers(0, len(centres)))]
                    pa = max(1, xa + int(rng.integers(-3, 4))); pb = max(1, xb + int(rng.integers(-3, 4)))
                    if rng.random() < 0.03:
                        pb = contigs[cb] + 7
                    f.write("\t".join(["split%d_%s" % (k, sample), ca, cb, str(pa), tf(), str(pb), tf(),
                                       str(pa - 40), str(pa), str(pb), str(pb + 40)]) + "\n")
            with open("%s_tiddit/contigs_%s.tab" % (prefix, sample), "w") as f:
                for k in range(40):
                    ca, cb, xa, xb = centres[int(rng.integers(0, len(centres)))]
                    if rng.random() < 0.5:
                        cb = ca
                        xa = int(rng.integers(1, contigs[ca] - 500)); xb = xa + int(rng.integers(10, 900))
                    f.write("\t".join(["ctg%d_%s" % (k, sample), ca, cb, str(xa), tf(), str(xb), tf(),
                                       str(xa - 200), str(xa), str(xb), str(xb + 200)]) + "\n")
        chromosomes = ["chr1", "chr2", "chr10", "tiny"]
        args = dict(samples=samples, is_mp=is_mp, epsilon=eps, m=m, max_ins_len=400, min_contig=1000,
                    skip_assembly=skip_assembly, min_reads=min_reads, chromosomes=chromosomes, contig_length=contigs)
        cand = R.tiddit_cluster.main(prefix, chromosomes, contigs, samples, is_mp, eps, m, 400, 1000, skip_assembly,
                                     min_reads)
        order = {a: {b: [int(c) for c in cand[a][b]] for b in cand[a]} for a in cand}
        with open(prefix + "_expected.json", "w") as f:
            json.dump({"args": args, "candidates": _jsonable(cand), "order": order}, f, separators=(",", ":"))


def config1_cov():
    """BASELINE config 1: `tiddit --cov` on a 1-contig 10k-read synthetic BAM, -z 500.  The BAM comes from our
    writer; the expected bed / wig come from the REFERENCE's coverage functions driven by the loop of
    tiddit/__main__.py:225-247 (the reference CLI itself cannot start here: it imports pysam)."""
    from tiddit_b200 import bamio
    contigs = [("chrS", 1_000_000)]
    bam = os.path.join(HERE, "config1.bam")
    bamio.write_bam(bam, contigs, bamio.synthetic_reads(contigs, 10_000, seed=1))
    cov = R.tiddit_coverage
    for z, q in ((500, 20), (50, 5)):
        samfile = bamio.AlignmentFile(bam)
        header = samfile.header
        coverage_data, end_bin_size = cov.create_coverage(header, z)
        for read in samfile.fetch(until_eof=True):
            if read.is_unmapped or read.is_duplicate:
                continue
            if read.mapq >= q:
                name = read.reference_name
                coverage_data[name] = cov.update_coverage(read.reference_start, read.reference_end, z,
                                                          coverage_data[name], end_bin_size[name])
        cov.print_coverage(coverage_data, header, z, "bed", os.path.join(HERE, "config1_z%d_q%d.bed" % (z, q)))
        if z == 500:
            cov.print_coverage(coverage_data, header, z, "wig", os.path.join(HERE, "config1_z%d_q%d.wig" % (z, q)))


def aggregate_cases():
    """Candidate aggregation (tiddit_cluster.pyx:156-336): random tab files written to a scratch directory, the REAL
    tiddit_cluster.main run on them, and per case the packed arrays (our reader, itself pinned by cluster_cases) +
    the reference's labels + one expected numeric row per candidate in dict order.  Scenarios are built to reach every
    branch of :265-330 (splits >= min_reads, contigs, few splits, discordants by orientation in PE and MP libraries,
    discordants by mode) with ties in the modes and repeated read names."""
    import tempfile
    from tiddit_b200.signals import PackedSignals
    contigs = {"chr1": 90000, "chr2": 70000, "chr3": 40000, "chrX": 52000, "small": 800}
    names = list(contigs)
    rng = np.random.default_rng(77)
    out = {}
    scenarios = []
    for k in range(14):
        scenarios.append(dict(samples=["S1"] if k % 3 else ["S1", "S2"], is_mp=bool(k % 2), skip_assembly=(k % 5 == 4),
                              eps=int(rng.integers(100, 600)), m=int(rng.integers(2, 6)), min_reads=int(rng.integers(1, 7)),
                              n_disc=int(rng.integers(200, 900)), n_split=int(rng.integers(0, 400)) if k % 4 else 0,
                              n_ctg=int(rng.integers(0, 60)) if k % 3 == 0 else 0, jitter=int(rng.choice([2, 30, 120])),
                              max_ins_len=int(rng.integers(200, 600))))
    tmp = tempfile.mkdtemp(prefix="tdt_agg_")
    for case, sc in enumerate(scenarios):
        prefix = os.path.join(tmp, "case%d" % case)
        os.makedirs(prefix + "_tiddit")
        centres = []
        for _ in range(int(rng.integers(6, 25))):
            a, b = sorted(rng.integers(0, 4, 2).tolist())
            centres.append((names[a], names[b], int(rng.integers(200, contigs[names[a]] - 200)),
                            int(rng.integers(200, contigs[names[b]] - 200)), float(rng.choice([0.02, 0.5, 0.97])),
                            float(rng.choice([0.02, 0.5, 0.97]))))
        for sample in sc["samples"]:
            with open("%s_tiddit/discordants_%s.tab" % (prefix, sample), "w") as f:
                for k in range(sc["n_disc"]):
                    ca, cb, xa, xb, pa, pb = centres[int(rng.integers(0, len(centres)))]
                    if rng.random() < 0.2:
                        xa, xb = int(rng.integers(1, contigs[ca])), int(rng.integers(1, contigs[cb]))
                    sa = max(1, xa + int(rng.integers(-sc["jitter"], sc["jitter"] + 1)))
                    sb = max(1, xb + int(rng.integers(-sc["jitter"], sc["jitter"] + 1)))
                    if rng.random() < 0.02:
                        sa = contigs[ca] + 30
                    if rng.random() < 0.02:
                        sb = contigs[cb] + 30
                    f.write("\t".join(["read%d" % int(rng.integers(0, sc["n_disc"] // 2 + 1)), ca, cb, str(sa), str(sa + 100),
                                       str(rng.random() < pa), str(sb), str(sb + 100), str(rng.random() < pb)]) + "\n")
            with open("%s_tiddit/splits_%s.tab" % (prefix, sample), "w") as f:
                for k in range(sc["n_split"]):
                    ca, cb, xa, xb, pa, pb = centres[int(rng.integers(0, len(centres)))]
                    qa = max(1, xa + int(rng.integers(-2, 3))); qb = max(1, xb + int(rng.integers(-2, 3)))
                    if rng.random() < 0.02:
                        qb = contigs[cb] + 5
                    extra = ["x"] * 8 if rng.random() < 0.1 else []   # readers only use the first 11 fields
                    f.write("\t".join(["split%d" % int(rng.integers(0, sc["n_split"] // 2 + 1)), ca, cb, str(qa),
                                       str(rng.random() < pa), str(qb), str(rng.random() < pb), str(qa - 40), str(qa),
                                       str(qb), str(qb + 40)] + extra) + "\n")
            with open("%s_tiddit/contigs_%s.tab" % (prefix, sample), "w") as f:
                for k in range(sc["n_ctg"]):
                    ca, cb, xa, xb, pa, pb = centres[int(rng.integers(0, len(centres)))]
                    if rng.random() < 0.5:
                        cb = ca
                        xa = int(rng.integers(1, contigs[ca] - 1500)); xb = xa + int(rng.integers(10, 1400))
                    f.write("\t".join(["ctg%d" % k, ca, cb, str(xa), str(rng.random() < 0.5), str(xb),
                                       str(rng.random() < 0.5), str(xa - 200), str(xa), str(xb), str(xb + 200)]) + "\n")
        chromosomes = ["chr1", "chr2", "chrX", "chr3", "small"]
        cand = R.tiddit_cluster.main(prefix, chromosomes, contigs, sc["samples"], sc["is_mp"], sc["eps"], sc["m"],
                                     sc["max_ins_len"], 1000, sc["skip_assembly"], sc["min_reads"])
        pk = PackedSignals.from_tab(prefix, chromosomes, contigs, sc["samples"], sc["is_mp"], 1000, sc["skip_assembly"])
        labels = np.full(len(pk), -1, dtype=np.int32)
        for p in range(pk.n_pairs):
            lo, hi = int(pk.seg_off[p]), int(pk.seg_off[p + 1])
            rows = sorted([[int(pk.posA[i]), int(pk.posB[i]), i - lo] for i in range(lo, hi)], key=lambda l: l[0])
            arr = np.array(rows)
            lab = R.DBSCAN.main(arr, sc["eps"], sc["m"])
            labels[lo + arr[:, 2]] = lab.astype(np.int32)
        rows = []
        pair_index = {pr: i for i, pr in enumerate(pk.pairs)}
        for a in cand:
            for b in cand[a]:
                for cid, c in cand[a][b].items():
                    rows.append([pair_index[(a, b)], int(cid), c["N_discordants"], c["N_splits"], c["N_contigs"],
                                 int(c["posA"]), int(c["posB"]), c["startA"], c["endA"], c["startB"], c["endB"],
                                 len(c["positions_A"]["start"])])
        key = "c%d_" % case
        for f in ("seg_off", "posA", "posB", "span", "name_id", "flags", "same_chrom"):
            out[key + f] = getattr(pk, f)
        out[key + "labels"] = labels
        out[key + "expected"] = np.array(rows, dtype=np.int64).reshape(-1, 12)
        out[key + "params"] = np.array([sc["max_ins_len"], int(sc["is_mp"]), sc["min_reads"], sc["eps"], sc["m"]])
    out["n_cases"] = np.array(len(scenarios))
    np.savez_compressed(os.path.join(HERE, "aggregate_cases.npz"), **out)


def signal_cases():
    """tiddit_signal.main (the REAL one, compiled against the pysam stand-in of oracle/ref_shims -- pysam itself is
    not installed here) on synthetic BAMs from our writer that reach every branch of the worker loop: expected
    discordants / splits tab files, clips fasta and the 50-bp coverage of every contig."""
    import shutil
    import tempfile
    from tiddit_b200 import bamio, synth
    signal = R.tiddit_signal
    assert signal is not None, "tiddit_signal was not compiled"
    cases = [dict(contigs=[("chr2", 400000), ("chr10", 300000), ("chr1", 250000), ("tiny", 4000), ("chrX", 120000)],
                  n=2500, seed=13, min_q=10, max_ins=600, min_contig=5000, min_anchor_len=40, min_clip_len=20),
             dict(contigs=[("b", 90000), ("a", 150000), ("c", 60000)],
                  n=1500, seed=14, min_q=30, max_ins=450, min_contig=0, min_anchor_len=60, min_clip_len=4)]
    tmp = tempfile.mkdtemp(prefix="tdt_sig_")
    for k, c in enumerate(cases):
        bam = os.path.join(HERE, "signal_case%d.bam" % k)
        bamio.write_bam(bam, c["contigs"], synth.sv_bam_reads(c["contigs"], c["n"], seed=c["seed"], max_ins=c["max_ins"]))
        prefix = os.path.join(tmp, "case%d" % k)
        os.makedirs(prefix + "_tiddit/clips")
        cov = signal.main(bam, "", prefix, c["min_q"], c["max_ins"], "S", 1, c["min_contig"], True,
                          c["min_anchor_len"], c["min_clip_len"])
        out = os.path.join(HERE, "signal_case%d_expected" % k)
        os.makedirs(out, exist_ok=True)
        for name in ("discordants_S.tab", "splits_S.tab", "clips_S.fa"):
            shutil.copy(os.path.join(prefix + "_tiddit", name), os.path.join(out, name))
        np.savez_compressed(os.path.join(out, "coverage.npz"), order=np.array(list(cov), dtype=str),
                            **{"cov_" + n: v for n, v in cov.items()})
        with open(os.path.join(out, "args.json"), "w") as f:
            json.dump({kk: v for kk, v in c.items() if kk not in ("contigs", "n", "seed")}, f)


def ploidy_cases():
    """tiddit_coverage_analysis.determine_ploidy (the REAL one) on synthetic coverage / GC bins: library values and
    the ploidies.tab text."""
    import tempfile
    rng = np.random.default_rng(31)
    cases = []
    tmp = tempfile.mkdtemp(prefix="tdt_ploidy_")
    for k in range(6):
        names = ["chr%d" % (i + 1) for i in range(int(rng.integers(2, 7)))]
        cov, gc = {}, {}
        for i, nm in enumerate(names):
            nb = int(rng.integers(1, 1500))
            base = float(rng.choice([10, 30, 15, 60]))
            v = np.round(rng.gamma(20, base / 20, nb) * 50) / 50.0          # multiples of 1/50 with many ties
            v[rng.random(nb) < 0.1] = 0.0
            g = rng.integers(20, 70, nb).astype(np.int8)
            g[rng.random(nb) < 0.05] = -1
            if k == 2 and i == 1:
                v[:] = 0.0                                                 # nothing qualifies -> nan -> 0
            if k == 3 and i == 0:
                nb2 = 2 * (nb // 2) + 2                                    # even count, all distinct
                v = rng.permutation(nb2).astype(np.float64) + 1.5
                g = np.full(nb2, 40, dtype=np.int8)
            cov[nm], gc[nm] = v, g
        c = 0 if k != 4 else 27.5
        contigs = names[::-1] + ["not_there"]
        prefix = os.path.join(tmp, "case%d" % k)
        lib = R.tiddit_coverage_analysis.determine_ploidy(dict(cov), contigs, {"x": 1}, 2, prefix, c, "ref.fa", 50,
                                                          {"SQ": []}, dict(gc))
        cases.append({"names": names, "contigs": contigs, "c": c, "ploidy": 2,
                      "cov": {n: cov[n].tolist() for n in names}, "gc": {n: gc[n].tolist() for n in names},
                      "library": {kk: (float(v) if not isinstance(v, int) else v) for kk, v in lib.items()},
                      "library_types": {kk: type(v).__name__ for kk, v in lib.items()},
                      "tab": open(prefix + ".ploidies.tab").read()})
    with open(os.path.join(HERE, "ploidy_cases.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))


if __name__ == "__main__":
    steps = [config1_cov, dbscan_cases, coverage_cases, gc_cases, cluster_cases, aggregate_cases, signal_cases, ploidy_cases]
    only = set(sys.argv[1:])          # e.g. `make_golden.py signal_cases` regenerates one family
    for step in steps:
        if not only or step.__name__ in only:
            step()
    print("golden vectors written to", HERE)
