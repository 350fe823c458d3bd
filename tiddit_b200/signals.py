"""Packed signal arrays: the reference's per-pair record lists as struct-of-arrays.

The reference carries every discordant pair / split read / assembly contig from `tiddit_signal` to
`tiddit_cluster` through three tab files per sample (writer tiddit_signal.pyx:298-326, reader
tiddit_cluster.pyx:47-137) and keeps them as Python lists of 12-field records
`[name, sample, "D"|"S"|"A", posA, oriA, posB, oriB, idx, startA, endA, startB, endB]` (:72,101,134).
`PackedSignals` holds the same records, grouped by (chrA,chrB) pair in the reference's visiting order
(:140-150), as int32 / uint8 arrays that go to HBM unchanged -- the layout `tdt_cluster_labels` and
`tdt_cluster_aggregate` (include/tdt_b200.h) take -- plus the string side tables (read names, samples,
orientation strings).  `from_tab` reads the reference's files (compatibility path); `save` / `load` keep the
arrays in one .npz so that the 20M-line text round trip disappears between the two stages.
"""
import numpy as np

KIND_D, KIND_S, KIND_A = 0, 1, 2
KIND_CHAR = ("D", "S", "A")
SIG_A_TRUE, SIG_A_FALSE, SIG_B_TRUE, SIG_B_FALSE = 0x04, 0x08, 0x10, 0x20

# which of (startA, endA) / (startB, endB) is the breakpoint, by (reverse A, reverse B); index into the
# tab fields 3/4 (A) and 6/7 (B).  Mate-pair libraries point the other way (tiddit_cluster.pyx:8-35).
_PE_CHOICE = {("False", "True"): (4, 6), ("False", "False"): (4, 7), ("True", "True"): (3, 6)}
_MP_CHOICE = {("False", "True"): (3, 7), ("False", "False"): (3, 6), ("True", "True"): (4, 7)}


def find_discordant_pos(fragment, is_mp):
    """tiddit_cluster.pyx:7-37 -> (posA, posB) as the strings found in the tab line."""
    if is_mp:
        a, b = _MP_CHOICE.get((fragment[5], fragment[8]), (4, 6))
    else:
        a, b = _PE_CHOICE.get((fragment[5], fragment[8]), (3, 7))
    return fragment[a], fragment[b]


def orientation_flags(kind, oriA, oriB):
    """TDT_SIG_* byte of one record."""
    f = kind
    if oriA == "True":
        f |= SIG_A_TRUE
    elif oriA == "False":
        f |= SIG_A_FALSE
    if oriB == "True":
        f |= SIG_B_TRUE
    elif oriB == "False":
        f |= SIG_B_FALSE
    return f


class PackedSignals:
    """Signals of all (chrA,chrB) pairs, pair by pair, insertion order inside a pair."""

    FIELDS = ("seg_off", "posA", "posB", "span", "name_id", "flags", "sample_id", "oriA_id", "oriB_id", "same_chrom")

    def __init__(self, pairs, seg_off, posA, posB, span, name_id, flags, sample_id, oriA_id, oriB_id, names, samples,
                 ori_table, chrA_present=None):
        self.pairs = [tuple(p) for p in pairs]                       # [(chrA, chrB)] in visiting order
        # every chrA that has a signal at all: the reference opens candidates[chrA] for them (:141-144)
        self.chrA_present = list(chrA_present) if chrA_present is not None else list(dict.fromkeys(a for a, _ in self.pairs))
        self.seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)  # [P+1]
        self.posA = np.ascontiguousarray(posA, dtype=np.int32)       # int(rec[3])
        self.posB = np.ascontiguousarray(posB, dtype=np.int32)       # int(rec[5])
        self.span = np.ascontiguousarray(span, dtype=np.int32).reshape(-1, 4)   # rec[8..11]
        self.name_id = np.ascontiguousarray(name_id, dtype=np.int32)  # index into names
        self.flags = np.ascontiguousarray(flags, dtype=np.uint8)     # TDT_SIG_* bits
        self.sample_id = np.ascontiguousarray(sample_id, dtype=np.int32)
        self.oriA_id = np.ascontiguousarray(oriA_id, dtype=np.int32)  # index into ori_table (rec[4] verbatim)
        self.oriB_id = np.ascontiguousarray(oriB_id, dtype=np.int32)
        self.names = list(names)
        self.samples = list(samples)
        self.ori_table = list(ori_table)
        self.same_chrom = np.array([a == b for a, b in self.pairs], dtype=np.uint8)

    def __len__(self):
        return int(self.seg_off[-1]) if len(self.seg_off) else 0

    @property
    def n_pairs(self):
        return len(self.pairs)

    def max_pos(self):
        return int(max(self.posA.max(), self.posB.max())) if len(self) else 0

    # ---- the reference's tab files --------------------------------------------------------------------
    @classmethod
    def from_tab(cls, prefix, chromosomes, contig_length, samples, is_mp, min_contig, skip_assembly):
        """tiddit_cluster.pyx:47-137.  Quirks kept: positions are clamped to the contig length; for discordants
        the posB test is nested inside the posA test and overwrites posA (:67-70)."""
        def lines(stem, sample):
            with open("{}_tiddit/{}_{}.tab".format(prefix, stem, sample)) as handle:
                yield from handle

        sources = [(lines("discordants", sample), lines("splits", sample),
                    None if skip_assembly else lines("contigs", sample)) for sample in samples]
        return cls.from_lines(sources, chromosomes, contig_length, samples, is_mp, min_contig)

    @classmethod
    def from_lines(cls, sources, chromosomes, contig_length, samples, is_mp, min_contig):
        """The same records from in-memory lines: sources[k] = (discordant lines, split lines, contig lines or None)
        of samples[k], in the tab-file format -- e.g. tiddit_signal.discordant_lines / split_lines, so that the
        text files between the signal and the cluster stage need not be read back."""
        chrA_l, chrB_l, posA_l, posB_l, span_l, name_l, flag_l, samp_l, oa_l, ob_l = ([] for _ in range(10))
        name_ids, ori_ids = {}, {}

        def intern(table, key):
            v = table.get(key)
            if v is None:
                v = table[key] = len(table)
            return v

        def add(sample_k, kind, name, chrA, chrB, posA, oriA, posB, oriB, sA, eA, sB, eB):
            chrA_l.append(chrA)
            chrB_l.append(chrB)
            posA_l.append(int(posA))
            posB_l.append(int(posB))
            span_l.append((int(sA), int(eA), int(sB), int(eB)))
            name_l.append(intern(name_ids, name))
            flag_l.append(orientation_flags(kind, oriA, oriB))
            samp_l.append(sample_k)
            oa_l.append(intern(ori_ids, oriA))
            ob_l.append(intern(ori_ids, oriB))

        for k, (disc, splits, contigs) in enumerate(sources):
            for line in disc:
                f = line.rstrip().split("\t")
                chrA, chrB = f[1], f[2]
                if contig_length[chrA] < min_contig or contig_length[chrB] < min_contig:
                    continue
                posA, posB = find_discordant_pos(f, is_mp)
                if int(posA) > contig_length[chrA]:
                    posA = contig_length[chrA]
                    if int(posB) > contig_length[chrB]:
                        posA = contig_length[chrB]
                add(k, KIND_D, f[0], chrA, chrB, posA, f[5], posB, f[8], f[3], f[4], f[6], f[7])
            for handle, kind in ((splits, KIND_S), (contigs, KIND_A)):
                if handle is None:
                    continue
                for line in handle:
                    f = line.rstrip().split("\t")
                    chrA, chrB = f[1], f[2]
                    if contig_length[chrA] < min_contig or contig_length[chrB] < min_contig:
                        continue
                    posA, posB = f[3], f[5]
                    if int(posA) > contig_length[chrA]:
                        posA = contig_length[chrA]
                    if int(posB) > contig_length[chrB]:
                        posB = contig_length[chrB]
                    add(k, kind, f[0], chrA, chrB, posA, f[4], posB, f[6], f[7], f[8], f[9], f[10])

        # group by pair in the reference's visiting order (:140-150); a pair whose chrA or chrB is not listed in
        # `chromosomes` is never visited
        seen = {}
        for a, b in zip(chrA_l, chrB_l):
            seen.setdefault((a, b), len(seen))
        pairs = [(a, b) for a in chromosomes for b in chromosomes if (a, b) in seen]
        pairs = list(dict.fromkeys(pairs))
        rank = {p: r for r, p in enumerate(pairs)}
        pair_rank = np.fromiter((rank.get((a, b), -1) for a, b in zip(chrA_l, chrB_l)), dtype=np.int64,
                                count=len(chrA_l))
        order = np.argsort(pair_rank, kind="stable")
        order = order[pair_rank[order] >= 0]
        counts = np.bincount(pair_rank[order], minlength=len(pairs)) if len(pairs) else np.zeros(0, dtype=np.int64)
        seg_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)

        def col(lst, dtype):
            return np.asarray(lst, dtype=dtype)[order] if len(lst) else np.zeros(0, dtype=dtype)

        posA, posB = col(posA_l, np.int64), col(posB_l, np.int64)
        span = np.asarray(span_l, dtype=np.int64).reshape(-1, 4)[order] if len(span_l) else np.zeros((0, 4), np.int64)
        for arr in (posA, posB, span):
            if arr.size and (arr.min() < -2 ** 31 or arr.max() >= 2 ** 31 - 1):
                raise OverflowError("signal coordinates must fit int32")
        names = [None] * len(name_ids)
        for s, i in name_ids.items():
            names[i] = s
        ori_table = [None] * len(ori_ids)
        for s, i in ori_ids.items():
            ori_table[i] = s
        return cls(pairs, seg_off, posA, posB, span, col(name_l, np.int32), col(flag_l, np.uint8), col(samp_l, np.int32),
                   col(oa_l, np.int32), col(ob_l, np.int32), names, samples, ori_table,
                   chrA_present=[a for a in dict.fromkeys(chromosomes) if a in set(chrA_l)])

    # ---- .npz round trip --------------------------------------------------------------------------------
    def save(self, path):
        np.savez(path, pairs=np.array(self.pairs, dtype=object).reshape(-1, 2).astype(str),
                 names=np.array(self.names, dtype=str), samples=np.array(self.samples, dtype=str),
                 ori_table=np.array(self.ori_table, dtype=str), chrA_present=np.array(self.chrA_present, dtype=str),
                 **{f: getattr(self, f) for f in self.FIELDS if f != "same_chrom"})

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        return cls([tuple(p) for p in z["pairs"].tolist()], z["seg_off"], z["posA"], z["posB"], z["span"], z["name_id"],
                   z["flags"], z["sample_id"], z["oriA_id"], z["oriB_id"], z["names"].tolist(), z["samples"].tolist(),
                   z["ori_table"].tolist(), z["chrA_present"].tolist())


class CandidateTable:
    """Result of the device aggregation: one int32 row per candidate (columns TDT_CAND_* of include/tdt_b200.h) in
    the reference's dict insertion order + the member lists."""
    COLS = ("pair", "id", "first", "member_off", "size", "N_discordants", "N_splits", "N_contigs", "posA", "posB",
            "startA", "endA", "startB", "endB", "rule")

    def __init__(self, rows, member_idx):
        self.rows = rows              # int32 [C, 16]
        self.member_idx = member_idx  # int32 [M]

    def __len__(self):
        return len(self.rows)

    def column(self, name):
        return self.rows[:, self.COLS.index(name)]

    def members(self, c):
        off, size = int(self.rows[c, 3]), int(self.rows[c, 4])
        return self.member_idx[off:off + size]
