"""Synthetic workloads shaped like TIDDIT's inputs (SURVEY.md App. E / BASELINE.json configs).

Pure numpy generators, deterministic per seed; used by bench.py and the tests.  Signals come out
grouped by (chrA,chrB) pair -- seg_off delimits the pairs -- and shuffled inside each pair, which is
the insertion order the cluster driver sees."""
import numpy as np

GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
          ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
          ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
          ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
          ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
          ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


def config2_signals(n=1_000_000, seed=42, length=250_000_000, n_clusters=20_000, mean_size=15, sigma=120):
    """BASELINE config 2: one pair, planted clusters of size 2+Geom(1/mean_size) + uniform noise up to n."""
    rng = np.random.default_rng(seed)
    sizes = 2 + rng.geometric(1.0 / mean_size, n_clusters)
    keep = np.cumsum(sizes) <= n
    sizes = sizes[keep]
    cx = rng.integers(1, length, len(sizes))
    cy = rng.integers(1, length, len(sizes))
    ax = np.repeat(cx, sizes) + np.rint(rng.normal(0, sigma, sizes.sum())).astype(np.int64)
    ay = np.repeat(cy, sizes) + np.rint(rng.normal(0, sigma, sizes.sum())).astype(np.int64)
    k = n - len(ax)
    ax = np.concatenate([ax, rng.integers(1, length, k)])
    ay = np.concatenate([ay, rng.integers(1, length, k)])
    perm = rng.permutation(n)
    posA = np.clip(ax[perm], 1, length).astype(np.int32)
    posB = np.clip(ay[perm], 1, length).astype(np.int32)
    return posA, posB, np.array([0, n], dtype=np.int64), length


def populated_pairs(contigs=GRCH38):
    """(ia, ib) with nameA <= nameB in string order, as tiddit_signal.pyx:214-219 orders inter-chromosomal pairs;
    listed in the cluster driver's visiting order (header order of chrA, then chrB)."""
    out = []
    for ia, (na, _) in enumerate(contigs):
        for ib, (nb, _) in enumerate(contigs):
            if na <= nb:
                out.append((ia, ib))
    return out


def _wgs_signals(n, seed, bg, cl, hs, cluster_mean, sigma, n_hot, hot_lo, hot_hi, contigs):
    rng = np.random.default_rng(seed)
    lens = np.array([ln for _, ln in contigs], dtype=np.float64)
    pairs = populated_pairs(contigs)
    intra = np.array([ia == ib for ia, ib in pairs])
    w = np.where(intra, 0.7 * np.array([lens[ia] for ia, _ in pairs]) / lens.sum(), 0.0)
    wi = np.array([lens[ia] * lens[ib] for ia, ib in pairs]) * (~intra)
    w = w + 0.3 * wi / wi.sum()
    w /= w.sum()
    n_hs = int(n * hs)
    n_pair = rng.multinomial(n - n_hs, w)
    hot_sizes = rng.integers(hot_lo, hot_hi, n_hot).astype(np.float64)
    hot_sizes = np.maximum(1, np.floor(hot_sizes * n_hs / hot_sizes.sum())).astype(np.int64)
    hot_sizes[0] += n_hs - hot_sizes.sum()
    hot_pair = rng.choice(len(pairs), n_hot, p=w)
    A, B, off = [], [], [0]
    for k, (ia, ib) in enumerate(pairs):
        la, lb = int(lens[ia]), int(lens[ib])
        nk = int(n_pair[k])
        n_bg = int(round(nk * bg / (bg + cl)))
        n_cl = nk - n_bg
        xa = [rng.integers(1, la + 1, n_bg)]
        if ia == ib:
            xb = [np.minimum(xa[0] + np.floor(rng.lognormal(9, 2, n_bg)).astype(np.int64), lb)]
        else:
            xb = [rng.integers(1, lb + 1, n_bg)]
        if n_cl > 0:
            sizes = 2 + rng.geometric(1.0 / cluster_mean, max(1, n_cl // (cluster_mean + 1) + 8))
            sizes = sizes[np.cumsum(sizes) <= n_cl]
            rest = n_cl - int(sizes.sum())
            if rest > 0:
                sizes = np.append(sizes, rest)
            cx = rng.integers(1, la + 1, len(sizes))
            cy = rng.integers(1, lb + 1, len(sizes))
            xa.append(np.repeat(cx, sizes) + np.rint(rng.normal(0, sigma, n_cl)).astype(np.int64))
            xb.append(np.repeat(cy, sizes) + np.rint(rng.normal(0, sigma, n_cl)).astype(np.int64))
        for h in np.flatnonzero(hot_pair == k):
            hx, hy = rng.integers(1, la + 1), rng.integers(1, lb + 1)
            xa.append(hx + rng.integers(0, 2000, hot_sizes[h]))
            xb.append(hy + rng.integers(0, 2000, hot_sizes[h]))
        a = np.clip(np.concatenate(xa), 1, la)
        b = np.clip(np.concatenate(xb), 1, lb)
        perm = rng.permutation(len(a))
        A.append(a[perm].astype(np.int32))
        B.append(b[perm].astype(np.int32))
        off.append(off[-1] + len(a))
    return np.concatenate(A), np.concatenate(B), np.array(off, dtype=np.int64), int(lens.max())


def wgs30x_signals(n=20_000_000, seed=3, contigs=GRCH38):
    """BASELINE config 3: 30X-WGS-shaped signals over the 300 populated chromosome pairs (eps=500, m=3)."""
    return _wgs_signals(n, seed, 0.55, 0.40, 0.05, 12, 120, 50, 5_000, 50_000, contigs)


def tumor60x_signals(n=50_000_000, seed=5, contigs=GRCH38):
    """BASELINE config 5: 60X tumour-like signals, dense clusters + 100 hotspots (eps=1000, m=5)."""
    return _wgs_signals(n, seed, 0.25, 0.70, 0.05, 60, 250, 100, 10_000, 100_000, contigs)


def coverage_reads(n_reads, seed=7, contigs=GRCH38, read_len=150):
    """Coordinate-sorted reads per contig (BAM order): starts uniform, end = min(start + read_len, LN).
    -> start, end (int32), read_off, bin-independent contig lengths."""
    rng = np.random.default_rng(seed)
    lens = np.array([ln for _, ln in contigs], dtype=np.int64)
    per = np.floor(n_reads * lens / lens.sum()).astype(np.int64)
    per[0] += n_reads - per.sum()
    starts, ends, off = [], [], [0]
    for ln, k in zip(lens, per):
        s = np.sort(rng.integers(0, ln, int(k)))
        starts.append(s.astype(np.int32))
        ends.append(np.minimum(s + read_len, ln).astype(np.int32))
        off.append(off[-1] + int(k))
    return np.concatenate(starts), np.concatenate(ends), np.array(off, dtype=np.int64), lens


def sorted_starts_device(torch, k, ln, generator=None):
    """k sorted uniform read starts in [0, ln) generated ON the device without a sort: normalised running sums of
    exponential gaps are distributed exactly like the order statistics of k uniforms (bench / profiling input only)."""
    if k <= 0:
        return torch.empty(0, dtype=torch.int32, device="cuda")
    gaps = -torch.log1p(-torch.rand(k + 1, device="cuda", generator=generator, dtype=torch.float64))
    c = torch.cumsum(gaps, 0)
    s = torch.floor(c[:k] * (float(ln) / float(c[k]))).clamp_(0, ln - 1)
    return s.to(torch.int32)


def fasta_sequence(length, seed=9, gc=0.41):
    """uint8 bases: i.i.d. with `gc` GC content, half soft-masked in blocks of 300-3000 bp, one N block in the
    middle (3 % of the contig, <= 3 Mbp) and 10 kb (<= 1 %) N telomeres."""
    rng = np.random.default_rng(seed)
    r = rng.random(length)
    seq = np.where(r < gc / 2, ord("C"), np.where(r < gc, ord("G"), np.where(r < gc + (1 - gc) / 2, ord("A"), ord("T"))))
    seq = seq.astype(np.uint8)
    pos = 0
    lower = rng.random() < 0.5
    while pos < length:
        blk = int(rng.integers(300, 3001))
        if lower:
            seq[pos:pos + blk] |= 0x20
        lower = not lower
        pos += blk
    tel = min(10_000, max(1, length // 100))
    seq[:tel] = ord("N")
    seq[length - tel:] = ord("N")
    cen = min(3_000_000, max(1, length * 3 // 100))
    mid = length // 2
    seq[mid:mid + cen] = ord("N")
    return seq


def signal_records(posA, posB, seg_off, seed=11, split_frac=0.18, contig_frac=0.02, dup_names=0.3, contigs=GRCH38):
    """The remaining record fields (tiddit_cluster.pyx:72,101,134) for a signal set: kind (D / S / A), read-name ids
    (a share of the names occurs on several records, like reads with supplementary alignments), orientation bits
    drawn with a per-region bias (so that both the orientation rule and the mode rule of :284-329 occur), and
    (startA, endA, startB, endB) around the breakpoints.  Assembly contigs on intra-chromosomal pairs get
    posB close to posA now and then, so that noise contigs survive as singleton candidates (:163-168).
    -> dict(span int32 [n,4], name_id int32, flags uint8, same_chrom uint8 [P], n_names)."""
    rng = np.random.default_rng(seed)
    n = len(posA)
    pairs = populated_pairs(contigs)
    P = len(seg_off) - 1
    same = np.array([ia == ib for ia, ib in pairs[:P]], dtype=np.uint8) if P == len(pairs) else np.ones(P, dtype=np.uint8)
    r = rng.random(n)
    kind = np.where(r < contig_frac, 2, np.where(r < contig_frac + split_frac, 1, 0)).astype(np.uint8)
    # orientation bias by 4 kb region of posA / posB: 0.03, 0.5 or 0.97
    bias = np.array([0.03, 0.5, 0.97])
    pa = bias[(posA.astype(np.int64) >> 12) * 2654435761 % 3]
    pb = bias[(posB.astype(np.int64) >> 12) * 40503 % 3]
    ta, tb = rng.random(n) < pa, rng.random(n) < pb
    flags = kind | np.where(ta, 0x04, 0x08).astype(np.uint8) | np.where(tb, 0x10, 0x20).astype(np.uint8)
    name_id = np.arange(n, dtype=np.int64)
    dup = rng.random(n) < dup_names
    name_id[dup] = np.maximum(name_id[dup] - rng.integers(1, 40, int(dup.sum())), 0)   # shares a name with an earlier record
    la, lb = rng.integers(30, 151, n), rng.integers(30, 151, n)
    span = np.stack([posA - la, posA.astype(np.int64), posB.astype(np.int64), posB + lb], axis=1)
    return {"span": np.ascontiguousarray(span, dtype=np.int32), "name_id": name_id.astype(np.int32), "flags": flags,
            "same_chrom": same, "n_names": n}


def write_tab_files(prefix, sample, posA, posB, seg_off, rec, contigs=GRCH38):
    """discordants_<sample>.tab / splits_<sample>.tab (writer tiddit_signal.pyx:298-326) for a synthetic
    signal set, laid out so that the reference's reader (tiddit_cluster.pyx:47-137, paired-end library) parses back
    exactly (posA, posB): discordant rows pick endA / startA and startB / endB by orientation (find_discordant_pos),
    split rows carry the positions directly.  Assembly contigs (kind A) are written as splits.  -> number of lines."""
    import os
    import pandas as pd
    os.makedirs(prefix + "_tiddit", exist_ok=True)
    pairs = populated_pairs(contigs)
    names = np.array([c for c, _ in contigs])
    pid = np.repeat(np.arange(len(seg_off) - 1), np.diff(seg_off))
    chrA = names[np.array([ia for ia, _ in pairs])[pid]]
    chrB = names[np.array([ib for _, ib in pairs])[pid]]
    flags = rec["flags"]
    revA = np.where(flags & 0x04, "True", "False")
    revB = np.where(flags & 0x10, "True", "False")
    a, b = posA.astype(np.int64), posB.astype(np.int64)
    read = np.char.add("r", rec["name_id"].astype(str))
    disc = (flags & 3) == 0
    startA = np.where(revA == "True", a, a - 100)
    endA = np.where(revA == "True", a + 100, a)
    startB = np.where(revB == "True", b, b - 100)
    endB = np.where(revB == "True", b + 100, b)
    d = pd.DataFrame({"n": read[disc], "ca": chrA[disc], "cb": chrB[disc], "sa": startA[disc], "ea": endA[disc],
                      "ra": revA[disc], "sb": startB[disc], "eb": endB[disc], "rb": revB[disc]})
    d.to_csv("%s_tiddit/discordants_%s.tab" % (prefix, sample), sep="\t", header=False, index=False)
    sp = ~disc
    span = rec["span"]
    t = pd.DataFrame({"n": read[sp], "ca": chrA[sp], "cb": chrB[sp], "pa": a[sp], "ra": revA[sp], "pb": b[sp], "rb": revB[sp],
                      "sa": span[sp, 0], "ea": span[sp, 1], "sb": span[sp, 2], "eb": span[sp, 3]})
    t.to_csv("%s_tiddit/splits_%s.tab" % (prefix, sample), sep="\t", header=False, index=False)
    return int(len(posA))


def sv_bam_reads(contigs, n_fragments=4000, seed=13, read_len=100, max_ins=600, n_events=12):
    """Reads for bamio.write_bam that exercise every branch of the signal worker (tiddit_signal.pyx:169-221):
    proper pairs, discordant pairs (far apart on one contig, across contigs, every orientation, mates that are
    unmapped / filtered / low quality), split reads with SA tags (one or several entries, either strand, soft and
    hard clips, SA on a smaller- or larger-named contig, quality below the threshold), soft-clipped reads,
    duplicates, secondary and supplementary records.  Coordinate-sorted per contig like a real BAM."""
    rng = np.random.default_rng(seed)
    nc = len(contigs)
    out = [[] for _ in contigs]
    bases = np.array(list("ACGT"))

    def seq(n):
        return "".join(bases[rng.integers(0, 4, n)])

    def place(ci, n):
        return int(rng.integers(0, max(1, contigs[ci][1] - n - 1)))

    def emit(ci, pos, name, flag, mapq, cigar, mate=(-1, -1), tlen=0, tags=None, with_seq=False):
        rd = {"name": name, "flag": int(flag), "ref": ci, "pos": int(pos), "mapq": int(mapq), "cigar": cigar,
              "next_ref": int(mate[0]), "next_pos": int(mate[1]), "tlen": int(tlen)}
        if tags:
            rd["tags"] = tags
        if with_seq:
            rd["seq"] = seq(sum(n for op, n in cigar if op in (0, 1, 4, 7, 8)))
        out[ci].append(rd)

    def mapq():
        return 0 if rng.random() < 0.08 else int(rng.integers(5, 61))

    for f in range(n_fragments):
        name = "frag%d" % f
        kind = rng.random()
        ca = int(rng.integers(0, nc))
        pa = place(ca, 2 * read_len + max_ins)
        ra, rb = (int(rng.random() < 0.5) * 0x10), (int(rng.random() < 0.5) * 0x20)
        extra = 0x400 if rng.random() < 0.03 else 0
        if kind < 0.45:                                   # proper pair
            ins = int(rng.integers(read_len, max_ins - 1))
            emit(ca, pa, name, 0x1 | 0x2 | 0x40 | 0x20 | extra, mapq(), [(0, read_len)], (ca, pa + ins - read_len), ins)
            emit(ca, pa + ins - read_len, name, 0x1 | 0x2 | 0x80 | 0x10 | extra, mapq(), [(0, read_len)], (ca, pa), -ins)
        elif kind < 0.70:                                 # discordant pair
            if rng.random() < 0.5:
                cb, pb = ca, pa + int(rng.integers(max_ins, 30 * max_ins))
                pb = min(pb, contigs[cb][1] - read_len - 1)
            else:
                cb = int(rng.integers(0, nc))
                pb = place(cb, read_len)
            tl = (pb + read_len - pa) if cb == ca else 0
            fa = 0x1 | 0x40 | ra | (0x20 if rb else 0) | extra
            fb = 0x1 | 0x80 | (0x10 if rb else 0) | (0x20 if ra else 0) | extra
            u = rng.random()
            if u < 0.08:
                fa |= 0x8                                  # mate unmapped: never paired up
                emit(ca, pa, name, fa, mapq(), [(0, read_len)], (ca, pa), 0)
                out[ca].append({"name": name, "flag": 0x1 | 0x4 | 0x80, "ref": ca, "pos": pa, "mapq": 0, "cigar": [],
                                "seq_len": read_len, "next_ref": ca, "next_pos": pa, "tlen": 0})
                continue
            emit(ca, pa, name, fa, mapq(), [(0, read_len)], (cb, pb), tl)
            if u < 0.16:
                continue                                   # the mate is missing from the file
            emit(cb, pb, name, fb, mapq(), [(0, read_len - 10), (4, 10)] if u < 0.3 else [(0, read_len)], (ca, pa), -tl)
            if u > 0.92:                                   # a third record under the same name (supplementary-free duplicate name)
                emit(cb, min(pb + 50, contigs[cb][1] - read_len - 1), name, fb & ~0x400, mapq(), [(0, read_len)], (ca, pa), -tl)
        elif kind < 0.88:                                 # split read: primary with SA (+ its supplementary record)
            cb = ca if rng.random() < 0.5 else int(rng.integers(0, nc))
            pb = place(cb, read_len)
            k = int(rng.integers(20, read_len - 20))
            before = rng.random() < 0.5                    # is the clipped part before the aligned part?
            cigar = [(4, k), (0, read_len - k)] if before else [(0, read_len - k), (4, k)]
            clipop = "H" if rng.random() < 0.3 else "S"
            sa_cig = ("%dM%d%s" % (k, read_len - k, clipop)) if before else ("%d%s%dM" % (read_len - k, clipop, k))
            if rng.random() < 0.2:
                sa_cig = sa_cig.replace("M", "M2D3I", 1) if rng.random() < 0.5 else sa_cig
            strand = "-" if rng.random() < 0.4 else "+"
            sa_q = 0 if rng.random() < 0.1 else int(rng.integers(10, 61))
            sa = "%s,%d,%s,%s,%d,%d;" % (contigs[cb][0], pb + 1, strand, sa_cig, sa_q, int(rng.integers(0, 4)))
            if rng.random() < 0.25:                        # more SA entries: the reference only ever uses the first
                cc = int(rng.integers(0, nc))
                sa += "%s,%d,%s,%dM%dS,%d,0;" % (contigs[cc][0], place(cc, read_len) + 1, "+", 30, read_len - 30,
                                                 int(rng.integers(0, 61)))
            paired = rng.random() < 0.7
            flag = (0x1 | 0x2 | 0x40 | 0x20 if paired else 0) | ra | extra
            ins = int(rng.integers(read_len, max_ins - 1))
            emit(ca, pa, name, flag, mapq(), cigar, (ca, pa + ins) if paired else (-1, -1), ins if paired else 0,
                 {"NM": int(rng.integers(0, 5)), "SA": sa}, with_seq=True)
            if rng.random() < 0.8:                         # the supplementary record itself (skipped by the worker)
                back = "%s,%d,%s,%s,%d,0;" % (contigs[ca][0], pa + 1, "-" if ra else "+",
                                               "".join("%d%s" % (n, "MIDNSHP=X"[op]) for op, n in cigar), 60)
                emit(cb, pb, name, 0x800 | (0x10 if strand == "-" else 0), sa_q, [(5, read_len - k), (0, k)], tags={"SA": back})
            if paired:
                emit(ca, pa + ins, name, 0x1 | 0x2 | 0x80 | 0x10, mapq(), [(0, read_len)], (ca, pa), -ins)
        else:                                             # soft-clipped reads (clips fasta), secondary noise
            k = int(rng.integers(1, 60))
            cigar = [(4, k), (0, read_len - k)] if rng.random() < 0.5 else [(0, read_len - k), (4, k)]
            if rng.random() < 0.15:
                cigar = [(4, k), (0, read_len - k - 5), (4, 5)]
            ins = int(rng.integers(read_len, max_ins - 1))
            emit(ca, pa, name, 0x1 | 0x2 | 0x40 | 0x20 | ra | extra, mapq(), cigar, (ca, pa + ins), ins, with_seq=True)
            emit(ca, pa + ins, name, 0x1 | 0x2 | 0x80 | 0x10, mapq(), [(0, read_len)], (ca, pa), -ins)
            if rng.random() < 0.2:
                emit(ca, place(ca, read_len), name, 0x100 | ra, mapq(), [(0, read_len)])
    # planted structural variants: fragments that agree on a pair of breakpoints, so that the cluster stage has
    # candidates to find (discordant pairs scattered around the breakpoints + split reads exactly on them)
    for ev in range(n_events):
        ca, cb = sorted(rng.integers(0, nc, 2).tolist()) if rng.random() < 0.4 else [int(rng.integers(0, nc))] * 2
        bpa = place(ca, 4 * read_len + max_ins) + 2 * read_len
        bpb = place(cb, 4 * read_len + max_ins) + 2 * read_len
        if ca == cb and abs(bpb - bpa) < 3 * max_ins:
            bpb = min(bpa + 5 * max_ins, contigs[cb][1] - 3 * read_len)
        rev_b = rng.random() < 0.5
        for k in range(int(rng.integers(4, 14))):
            name = "sv%d_%d" % (ev, k)
            q = int(rng.integers(25, 61))
            if rng.random() < 0.65:
                pa = max(0, bpa - read_len - int(rng.integers(0, 150)))
                pb = min(bpb + int(rng.integers(0, 150)), contigs[cb][1] - read_len - 1)
                tl = (pb + read_len - pa) if ca == cb else 0
                emit(ca, pa, name, 0x1 | 0x40 | (0x20 if rev_b else 0), q, [(0, read_len)], (cb, pb), tl)
                emit(cb, pb, name, 0x1 | 0x80 | (0x10 if rev_b else 0), q, [(0, read_len)], (ca, pa), -tl)
            else:
                keep = int(rng.integers(30, read_len - 30))
                sa = "%s,%d,%s,%dS%dM,%d,0;" % (contigs[cb][0], bpb + 1, "-" if rev_b else "+", keep, read_len - keep, q)
                emit(ca, bpa - keep, name, 0, q, [(0, keep), (4, read_len - keep)], tags={"SA": sa}, with_seq=True)
    reads = []
    for ci in range(nc):
        reads.extend(sorted(out[ci], key=lambda r: r["pos"]))   # stable: equal positions keep emission order
    return reads
