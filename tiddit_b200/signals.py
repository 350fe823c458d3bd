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
import os

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


def _part_from_lines(handle, sample_k, kind, contig_length, is_mp, min_contig):
    """One tab file, line by line, exactly as tiddit_cluster.pyx:47-137 reads it -> columns of the records kept."""
    name, chrA_l, chrB_l, posA_l, posB_l, oa, ob, span = [], [], [], [], [], [], [], []
    for line in handle:
        f = line.rstrip().split("\t")
        chrA, chrB = f[1], f[2]
        if contig_length[chrA] < min_contig or contig_length[chrB] < min_contig:
            continue
        if kind == KIND_D:
            posA, posB = find_discordant_pos(f, is_mp)
            if int(posA) > contig_length[chrA]:
                posA = contig_length[chrA]
                if int(posB) > contig_length[chrB]:
                    posA = contig_length[chrB]
            oriA, oriB, sp = f[5], f[8], (f[3], f[4], f[6], f[7])
        else:
            posA, posB = f[3], f[5]
            if int(posA) > contig_length[chrA]:
                posA = contig_length[chrA]
            if int(posB) > contig_length[chrB]:
                posB = contig_length[chrB]
            oriA, oriB, sp = f[4], f[6], (f[7], f[8], f[9], f[10])
        name.append(f[0])
        chrA_l.append(chrA)
        chrB_l.append(chrB)
        posA_l.append(int(posA))
        posB_l.append(int(posB))
        oa.append(oriA)
        ob.append(oriB)
        span.append((int(sp[0]), int(sp[1]), int(sp[2]), int(sp[3])))
    return dict(kind=kind, sample=sample_k, name=name, chrA=chrA_l, chrB=chrB_l, posA=posA_l, posB=posB_l, oriA=oa, oriB=ob,
                span=span)


def _part_from_columns(path, sample_k, kind, contig_length, is_mp, min_contig):
    """The same columns through the pandas C parser, the rules applied to whole columns.  Only a perfectly regular file
    passes (same number of fields on every line, integer coordinates, no blank-padded or empty text field); anything
    else raises _IrregularTab and the caller reads the file line by line."""
    import pandas as pd
    dtypes = "scciiciic" if kind == KIND_D else "sccicic" + "iiii"
    kinds = {"s": str, "c": "category", "i": np.int64}
    try:
        df = pd.read_csv(path, sep="\t", header=None, usecols=list(range(len(dtypes))),
                         dtype={i: kinds[d] for i, d in enumerate(dtypes)}, quoting=3, na_filter=False, engine="c",
                         skip_blank_lines=False)
    except pd.errors.EmptyDataError:
        return dict(kind=kind, sample=sample_k, name=[], chrA=[], chrB=[], posA=[], posB=[], oriA=[], oriB=[], span=[])
    except (pd.errors.ParserError, ValueError, TypeError, UnicodeDecodeError, OverflowError) as exc:
        raise _IrregularTab(str(exc))
    for i, d in enumerate(dtypes):
        if d == "c":
            if any((not isinstance(c, str)) or c == "" or c != c.strip() for c in df[i].cat.categories):
                raise _IrregularTab("empty or blank-padded field")
            df[i] = df[i].astype(object)
    if len(df) and df[0].str.contains("^$|^\\s|\\s$", regex=True).any():
        raise _IrregularTab("empty or blank-padded read name")
    chrA, chrB = df[1], df[2]
    length_of = pd.Series(contig_length, dtype=object)
    la, lb = chrA.map(length_of), chrB.map(length_of)
    if la.isna().any() or lb.isna().any():
        raise KeyError(chrA[la.isna()].iloc[0] if la.isna().any() else chrB[lb.isna()].iloc[0])
    la, lb = la.to_numpy().astype(np.int64), lb.to_numpy().astype(np.int64)
    col = lambda i: df[i].to_numpy().astype(np.int64)
    if kind == KIND_D:
        sA, eA, sB, eB = col(3), col(4), col(6), col(7)
        oriA, oriB = df[5], df[8]
        fa, ta = (oriA == "False").to_numpy(), (oriA == "True").to_numpy()
        fb, tb = (oriB == "False").to_numpy(), (oriB == "True").to_numpy()
        ft, ff, tt = fa & tb, fa & fb, ta & tb
        if is_mp:    # tiddit_cluster.pyx:8-35 (_MP_CHOICE / _PE_CHOICE above)
            posA, posB = np.where(ft | ff, sA, eA), np.where(ft | tt, eB, sB)
        else:
            posA, posB = np.where(ft | ff, eA, sA), np.where(ft | tt, sB, eB)
        posA = np.where(posA > la, np.where(posB > lb, lb, la), posA)          # :67-70: the nested test overwrites posA
        span = np.stack([sA, eA, sB, eB], 1)
    else:
        posA, posB = np.minimum(col(3), la), np.minimum(col(5), lb)
        oriA, oriB = df[4], df[6]
        span = np.stack([col(7), col(8), col(9), col(10)], 1)
    keep = (la >= min_contig) & (lb >= min_contig)
    take = lambda x: np.asarray(x)[keep]
    return dict(kind=kind, sample=sample_k, name=take(df[0]), chrA=take(chrA), chrB=take(chrB), posA=posA[keep],
                posB=posB[keep], oriA=take(oriA), oriB=take(oriB), span=span[keep])


class _IrregularTab(Exception):
    """A tab file the column parser does not take as is; the line-by-line reader handles (or rejects) it."""


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
        self.names = names if hasattr(names, "_blob") or isinstance(names, _LazyNames) else list(names)
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
    def from_tab(cls, prefix, chromosomes, contig_length, samples, is_mp, min_contig, skip_assembly, fast=True):
        """tiddit_cluster.pyx:47-137.  Quirks kept: positions are clamped to the contig length; for discordants
        the posB test is nested inside the posA test and overwrites posA (:67-70).

        fast=True: the files go through the native scanner libtdt_tab.so (mapped, parsed on all host cores, strings
        interned natively; 2 M lines in ~0.5 s instead of 3.6 s) when every file is perfectly regular; otherwise, and
        with fast="columns": regular files go through the pandas C parser column by column (no Python loop per line; measured 3.0 s vs
        6.0 s per million lines, the rest is string handling); a file that is not perfectly regular -- split lines grow by eight fields for every further
        record of the same read name -- is read line by line like the reference does.  Same result either way."""
        if fast is True:
            native = cls._from_tab_native(prefix, chromosomes, contig_length, samples, is_mp, min_contig, skip_assembly)
            if native is not None:
                return native
        parts = []
        for k, sample in enumerate(samples):
            stems = [("discordants", KIND_D), ("splits", KIND_S)] + ([] if skip_assembly else [("contigs", KIND_A)])
            for stem, kind in stems:
                path = "{}_tiddit/{}_{}.tab".format(prefix, stem, sample)
                part = None
                if fast:
                    try:
                        part = _part_from_columns(path, k, kind, contig_length, is_mp, min_contig)
                    except _IrregularTab:
                        part = None
                if part is None:
                    with open(path) as handle:
                        part = _part_from_lines(handle, k, kind, contig_length, is_mp, min_contig)
                parts.append(part)
        return cls._assemble(parts, chromosomes, samples)

    @classmethod
    def _from_tab_native(cls, prefix, chromosomes, contig_length, samples, is_mp, min_contig, skip_assembly):
        """The files through libtdt_tab.so (include/tdt_tab.h): mapped, split at line ends over the host cores, fields
        parsed into columns, strings interned natively; the reference's per-record rules (:52-72, :80-101) are then
        applied to whole columns.  -> PackedSignals, or None when a file is not perfectly regular (the caller reads
        line by line like the reference) or the scanner is not built."""
        from . import tabio
        try:
            ts = tabio.TabSet()
        except (RuntimeError, OSError):
            return None
        try:
            ranges = []
            for k, sample in enumerate(samples):
                stems = [("discordants", KIND_D), ("splits", KIND_S)] + ([] if skip_assembly else [("contigs", KIND_A)])
                for stem, kind in stems:
                    path = "{}_tiddit/{}_{}.tab".format(prefix, stem, sample)
                    try:
                        lo, hi = ts.parse(path, kind)
                    except tabio.IrregularTab:
                        return None
                    except OSError:
                        if not os.path.exists(path):
                            raise FileNotFoundError(2, "No such file or directory", path)
                        raise
                    ranges.append((kind, k, lo, hi))
            n = len(ts)
            contig_tab = ts.table(1)
            missing = [c for c in contig_tab if c not in contig_length]
            cA, cB = ts.col_i32(1), ts.col_i32(2)
            if missing:      # the reference fails on the first line that names an unknown contig (chrA looked up first)
                bad = np.isin(cA, [contig_tab.index(c) for c in missing]) | np.isin(cB, [contig_tab.index(c) for c in missing])
                i = int(np.flatnonzero(bad)[0])
                raise KeyError(contig_tab[cA[i]] if contig_tab[cA[i]] in missing else contig_tab[cB[i]])
            clen = np.array([contig_length[c] for c in contig_tab] or [0], dtype=np.int64)
            la, lb = (clen[cA], clen[cB]) if n else (np.zeros(0, np.int64), np.zeros(0, np.int64))
            ori_tab = ts.table(2)
            oA, oB = ts.col_i32(3), ts.col_i32(4)
            is_true = np.array([o == "True" for o in ori_tab] or [False])
            is_false = np.array([o == "False" for o in ori_tab] or [False])
            num = [ts.col_i64(j) for j in range(6)]
            posA, posB = np.zeros(n, np.int64), np.zeros(n, np.int64)
            span = [np.zeros(n, np.int64) for _ in range(4)]          # rec[8..11] as four columns until the very end
            kind_col, sample_col = np.zeros(n, np.uint8), np.zeros(n, np.int32)
            for kind, k, lo, hi in ranges:
                sl = slice(lo, hi)
                kind_col[sl], sample_col[sl] = kind, k
                if kind == KIND_D:
                    sA, eA, sB, eB = (num[j][sl] for j in range(4))
                    fa, ta, fb, tb = is_false[oA[sl]], is_true[oA[sl]], is_false[oB[sl]], is_true[oB[sl]]
                    ft, ff, tt = fa & tb, fa & fb, ta & tb
                    if is_mp:    # tiddit_cluster.pyx:8-35
                        pa, pb = np.where(ft | ff, sA, eA), np.where(ft | tt, eB, sB)
                    else:
                        pa, pb = np.where(ft | ff, eA, sA), np.where(ft | tt, sB, eB)
                    posA[sl] = np.where(pa > la[sl], np.where(pb > lb[sl], lb[sl], la[sl]), pa)   # :67-70 nested test
                    posB[sl] = pb
                    for j, col in enumerate((sA, eA, sB, eB)):
                        span[j][sl] = col
                else:
                    posA[sl], posB[sl] = np.minimum(num[0][sl], la[sl]), np.minimum(num[1][sl], lb[sl])
                    for j in range(4):
                        span[j][sl] = num[j + 2][sl]
            keep = (la >= min_contig) & (lb >= min_contig)
            all_kept = bool(keep.all())
            sel = (lambda arr: arr) if all_kept else (lambda arr: arr[keep])   # (no copies when nothing is dropped)
            for arr in [sel(posA), sel(posB)] + [sel(c) for c in span]:
                if arr.size and (arr.min() < -2 ** 31 or arr.max() >= 2 ** 31 - 1):
                    raise OverflowError("signal coordinates must fit int32")
            # (checked on the kept records: dropped ones may hold anything) -- from here on 32 bits per coordinate
            with np.errstate(over="ignore"):
                posA, posB, span = posA.astype(np.int32), posB.astype(np.int32), [c.astype(np.int32) for c in span]
            name_id = ts.col_i32(0)
            names = ts.table(0, lazy=True)
            if not all_kept:
                # ids in order of first appearance among the KEPT records, like the line reader interns them
                name_id, names = _reintern(name_id[keep], names)
                ori = np.empty(2 * int(keep.sum()), dtype=np.int32)
                ori[0::2], ori[1::2] = oA[keep], oB[keep]
                ori, ori_tab = _reintern(ori, ori_tab)
                oA_k, oB_k = ori[0::2], ori[1::2]
                is_true = np.array([o == "True" for o in ori_tab] or [False])
                is_false = np.array([o == "False" for o in ori_tab] or [False])
            else:
                oA_k, oB_k = oA, oB
            flags = sel(kind_col).copy()
            flags |= np.where(is_true, SIG_A_TRUE, np.where(is_false, SIG_A_FALSE, 0)).astype(np.uint8)[oA_k]   # per table entry,
            flags |= np.where(is_true, SIG_B_TRUE, np.where(is_false, SIG_B_FALSE, 0)).astype(np.uint8)[oB_k]   # then one gather
            # pairs in the reference's visiting order (:140-150)
            chrom_rank = {c: i for i, c in reversed(list(enumerate(chromosomes)))}       # first listing wins
            rank_of = np.array([chrom_rank.get(c, -1) for c in contig_tab] or [-1], dtype=np.int64)
            C = max(len(chromosomes), 1)
            cA_k, cB_k = sel(cA), sel(cB)
            T = max(len(contig_tab), 1)
            if T * T <= 1 << 22:
                # the contig pairs that occur, found on the (contig id, contig id) grid instead of on the records
                grid = (np.bincount(cA_k.astype(np.int64) * T + cB_k, minlength=T * T) if len(cA_k)
                        else np.zeros(T * T, np.int64)).reshape(T, T)
                ia, ib = np.nonzero(grid)
                listed = (rank_of[ia] >= 0) & (rank_of[ib] >= 0)
                ia, ib = ia[listed], ib[listed]
                keys = rank_of[ia] * C + rank_of[ib]
                by_key = np.argsort(keys)                                  # visiting order: chrA rank, then chrB rank
                present = keys[by_key]
                rank_grid = np.full((T, T), len(present), dtype=np.int64)  # never-visited pairs sort behind the others
                rank_grid[ia[by_key], ib[by_key]] = np.arange(len(present))
                pair_rank = rank_grid[cA_k, cB_k]
                counts = grid[ia[by_key], ib[by_key]]
                seen_a = np.flatnonzero(grid.sum(axis=1)).tolist()
            else:                                                          # very many contigs: on the records
                ra, rb = rank_of[cA_k], rank_of[cB_k]
                visited = (ra >= 0) & (rb >= 0)
                key = np.where(visited, ra * C + rb, -1)
                present = np.unique(key[visited])
                pair_rank = np.where(visited, np.searchsorted(present, key), len(present))
                counts = np.bincount(pair_rank, minlength=len(present) + 1)[:len(present)]
                seen_a = np.unique(cA_k).tolist()
            pairs = [(chromosomes[int(kk) // C], chromosomes[int(kk) % C]) for kk in present]
            seg_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            # stable sort of small integers: numpy's radix sort takes 16-bit keys
            order = np.argsort(pair_rank.astype(np.uint16) if len(pairs) < 65535 else pair_rank, kind="stable")
            order = order[:int(seg_off[-1])]
            chrA_all = {contig_tab[i] for i in seen_a}
            pick = (lambda arr: arr[order]) if all_kept else (lambda arr: arr[keep][order])
            span_out = np.empty((len(order), 4), dtype=np.int32)
            for j in range(4):
                span_out[:, j] = pick(span[j])
            return cls(pairs, seg_off, pick(posA), pick(posB), span_out, name_id[order], flags[order],
                       pick(sample_col), oA_k[order], oB_k[order], names, samples, list(ori_tab),
                       chrA_present=[a for a in dict.fromkeys(chromosomes) if a in chrA_all])
        finally:
            ts.close()

    @classmethod
    def from_lines(cls, sources, chromosomes, contig_length, samples, is_mp, min_contig):
        """The same records from in-memory lines: sources[k] = (discordant lines, split lines, contig lines or None)
        of samples[k], in the tab-file format -- e.g. tiddit_signal.discordant_lines / split_lines, so that the
        text files between the signal and the cluster stage need not be read back."""
        parts = []
        for k, (disc, splits, contigs) in enumerate(sources):
            for handle, kind in ((disc, KIND_D), (splits, KIND_S), (contigs, KIND_A)):
                if handle is not None:
                    parts.append(_part_from_lines(handle, k, kind, contig_length, is_mp, min_contig))
        return cls._assemble(parts, chromosomes, samples)

    @classmethod
    def _assemble(cls, parts, chromosomes, samples):
        """Per-file column sets (processing order: per sample discordants, splits, contigs) -> PackedSignals: strings
        interned in order of first appearance, records grouped by pair in the reference's visiting order (:140-150);
        a pair whose chrA or chrB is not listed in `chromosomes` is never visited."""
        import pandas as pd

        def cat(key, dtype):
            cols = [np.asarray(pt[key], dtype=dtype) for pt in parts if len(pt["posA"])]
            return np.concatenate(cols) if cols else np.zeros(0, dtype=dtype)

        n = int(sum(len(pt["posA"]) for pt in parts))
        name_id, names = pd.factorize(cat("name", object), sort=False)
        oriA, oriB = cat("oriA", object), cat("oriB", object)
        # one interning table for both orientation columns, in the order the line-by-line reader meets them (A then B)
        inter = np.empty(2 * n, dtype=object)
        inter[0::2], inter[1::2] = oriA, oriB
        ori_id, ori_table = pd.factorize(inter, sort=False)
        kind = np.concatenate([np.full(len(pt["posA"]), pt["kind"], dtype=np.uint8) for pt in parts]) if parts else np.zeros(0, np.uint8)
        sample_id = np.concatenate([np.full(len(pt["posA"]), pt["sample"], dtype=np.int32) for pt in parts]) if parts else np.zeros(0, np.int32)
        flags = kind.copy()
        flags |= np.where(oriA == "True", SIG_A_TRUE, np.where(oriA == "False", SIG_A_FALSE, 0)).astype(np.uint8)
        flags |= np.where(oriB == "True", SIG_B_TRUE, np.where(oriB == "False", SIG_B_FALSE, 0)).astype(np.uint8)
        posA, posB = cat("posA", np.int64), cat("posB", np.int64)
        spans = [np.asarray(pt["span"], dtype=np.int64).reshape(-1, 4) for pt in parts if len(pt["posA"])]
        span = np.concatenate(spans) if spans else np.zeros((0, 4), np.int64)
        for arr in (posA, posB, span):
            if arr.size and (arr.min() < -2 ** 31 or arr.max() >= 2 ** 31 - 1):
                raise OverflowError("signal coordinates must fit int32")
        chrA, chrB = cat("chrA", object), cat("chrB", object)
        chrom_rank = {c: i for i, c in reversed(list(enumerate(chromosomes)))}       # first listing wins
        C = max(len(chromosomes), 1)
        ra = pd.Series(chrA, dtype=object).map(chrom_rank).fillna(-1).to_numpy().astype(np.int64)
        rb = pd.Series(chrB, dtype=object).map(chrom_rank).fillna(-1).to_numpy().astype(np.int64)
        visited = (ra >= 0) & (rb >= 0)
        key = np.where(visited, ra * C + rb, -1)
        present = np.unique(key[visited])
        pairs = [(chromosomes[int(kk) // C], chromosomes[int(kk) % C]) for kk in present]
        pair_rank = np.where(visited, np.searchsorted(present, key), -1)
        order = np.argsort(pair_rank, kind="stable")
        order = order[pair_rank[order] >= 0]
        counts = np.bincount(pair_rank[order], minlength=len(pairs)) if len(pairs) else np.zeros(0, dtype=np.int64)
        seg_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        chrA_all = set(pd.unique(chrA).tolist()) if n else set()
        return cls(pairs, seg_off, posA[order], posB[order], span[order], name_id[order].astype(np.int32), flags[order],
                   sample_id[order], ori_id[0::2][order].astype(np.int32), ori_id[1::2][order].astype(np.int32),
                   [str(x) for x in names], samples, [str(x) for x in ori_table],
                   chrA_present=[a for a in dict.fromkeys(chromosomes) if a in chrA_all])

    @classmethod
    def from_arrays(cls, posA, posB, seg_off, rec, contigs, sample="S"):
        """Packed arrays as a producer holds them (synth.signal_records layout: span, name_id, flags over the populated
        (chrA,chrB) pairs of `contigs` in visiting order) -> PackedSignals, names interned lazily as "r<id>"."""
        from .synth import populated_pairs
        names_c = [c for c, _ in contigs]
        pairs = [(names_c[ia], names_c[ib]) for ia, ib in populated_pairs(contigs)][:len(seg_off) - 1]
        n = len(posA)
        flags = np.asarray(rec["flags"], dtype=np.uint8)
        ori_table = ["True", "False"]
        oriA = np.where(flags & SIG_A_TRUE, 0, 1).astype(np.int32)
        oriB = np.where(flags & SIG_B_TRUE, 0, 1).astype(np.int32)
        n_names = int(rec.get("n_names", n))
        return cls(pairs, seg_off, posA, posB, rec["span"], rec["name_id"], flags, np.zeros(n, dtype=np.int32), oriA, oriB,
                   _LazyNames(n_names), [sample], ori_table)

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


def _reintern(ids, table):
    """ids into `table` -> (ids renumbered in order of first appearance, the sub-table in that order)."""
    if len(ids) == 0:
        return ids.astype(np.int32), []
    uniq, first, inv = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    new_of = np.empty(len(uniq), dtype=np.int32)
    new_of[order] = np.arange(len(uniq), dtype=np.int32)
    return new_of[inv].astype(np.int32), [table[int(u)] for u in uniq[order]]


class _LazyNames:
    """names[i] = "r<i>" without materialising millions of strings (PackedSignals.from_arrays)."""

    def __init__(self, n):
        self.n = int(n)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return "r%d" % i

    def __iter__(self):
        return ("r%d" % i for i in range(self.n))


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
