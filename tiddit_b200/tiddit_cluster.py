"""Drop-in for tiddit/tiddit_cluster.pyx: `find_discordant_pos` (:7-37) and `main` (:39-338).

`main` keeps the reference's signature and returns the same nested `candidates` dict (schema in
SURVEY.md App. D), but instead of one sorted() + DBSCAN.main call per (chrA,chrB) list it packs
every pair's (posA, posB) into two int32 arrays, labels ALL pairs with one GPU call
(device_ops.cluster_labels -> tdt_cluster_labels) and folds the labels back on the host.
"""
from collections import Counter

import numpy as np

from . import device_ops

__all__ = ["find_discordant_pos", "main", "read_signals", "label_signals"]

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


class _Pair:
    """The signals of one (chrA, chrB) in insertion order (the reference's discordants[chrA][chrB] rows)."""
    __slots__ = ("records", "posA", "posB")

    def __init__(self):
        self.records = []   # [name, sample, type, posA, oriA, posB, oriB, idx, startA, endA, startB, endB]
        self.posA = []
        self.posB = []


def read_signals(prefix, contig_length, samples, is_mp, min_contig, skip_assembly):
    """tiddit_cluster.pyx:47-137: parse discordants_/splits_/contigs_<sample>.tab -> {chrA: {chrB: _Pair}}.

    Quirks kept: positions are clamped to the contig length; for discordants the posB test is nested
    inside the posA test and overwrites posA (:67-70); the record keeps the position as found (a
    string) unless it was clamped (then the int length)."""
    pairs = {}
    serial = 0

    def slot(chrA, chrB):
        row = pairs.get(chrA)
        if row is None:
            row = pairs[chrA] = {}
        cell = row.get(chrB)
        if cell is None:
            cell = row[chrB] = _Pair()
        return cell

    for sample in samples:
        with open("{}_tiddit/discordants_{}.tab".format(prefix, sample)) as handle:
            for line in handle:
                f = line.rstrip().split("\t")
                chrA, chrB = f[1], f[2]
                if contig_length[chrA] < min_contig or contig_length[chrB] < min_contig:
                    continue
                cell = slot(chrA, chrB)
                posA, posB = find_discordant_pos(f, is_mp)
                if int(posA) > contig_length[chrA]:
                    posA = contig_length[chrA]
                    if int(posB) > contig_length[chrB]:
                        posA = contig_length[chrB]
                cell.records.append([f[0], sample, "D", posA, f[5], posB, f[8], serial,
                                     int(f[3]), int(f[4]), int(f[6]), int(f[7])])
                cell.posA.append(int(posA))
                cell.posB.append(int(posB))
                serial += 1

        sources = [("splits", "S")] + ([] if skip_assembly else [("contigs", "A")])
        for stem, kind in sources:
            with open("{}_tiddit/{}_{}.tab".format(prefix, stem, sample)) as handle:
                for line in handle:
                    f = line.rstrip().split("\t")
                    chrA, chrB = f[1], f[2]
                    if contig_length[chrA] < min_contig or contig_length[chrB] < min_contig:
                        continue
                    cell = slot(chrA, chrB)
                    posA, posB = f[3], f[5]
                    if int(posA) > contig_length[chrA]:
                        posA = contig_length[chrA]
                    if int(posB) > contig_length[chrB]:
                        posB = contig_length[chrB]
                    cell.records.append([f[0], sample, kind, posA, f[4], posB, f[6], serial,
                                         int(f[7]), int(f[8]), int(f[9]), int(f[10])])
                    cell.posA.append(int(posA))
                    cell.posB.append(int(posB))
                    serial += 1
    return pairs


def label_signals(pair_list, epsilon, m):
    """One GPU call for every pair (tiddit_cluster.pyx:152-154 for all of them): list of _Pair ->
    list of int32 label arrays in insertion order."""
    if not pair_list:
        return []
    sizes = [len(p.posA) for p in pair_list]
    seg_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    posA = np.fromiter((v for p in pair_list for v in p.posA), dtype=np.int64, count=int(seg_off[-1]))
    posB = np.fromiter((v for p in pair_list for v in p.posB), dtype=np.int64, count=int(seg_off[-1]))
    if len(posA) and (min(posA.min(), posB.min()) < 0 or max(posA.max(), posB.max()) >= device_ops.INT32_MAX):
        raise OverflowError("signal positions must lie in [0, 2^31-1)")
    max_pos = int(max(posA.max(), posB.max())) if len(posA) else 0
    labels = device_ops.cluster_labels(posA.astype(np.int32), posB.astype(np.int32), seg_off, epsilon, m, max_pos)
    return [labels[seg_off[k]:seg_off[k + 1]] for k in range(len(pair_list))]


_KIND_KEY = {"D": "discordants", "S": "splits", "A": "contigs"}


def _new_candidate():
    side = lambda: {"contigs": [], "splits": [], "discordants": [], "orientation_contigs": [],
                    "orientation_splits": [], "orientation_discordants": [], "start": [], "end": []}
    return {"signal_type": {}, "samples": set(), "sample_discordants": {}, "sample_splits": {}, "sample_contigs": {},
            "N_discordants": 0, "discordants": set(), "N_splits": 0, "splits": set(), "N_contigs": 0,
            "contigs": set(), "n_signals": 0,
            "posA": 0, "positions_A": side(), "start_A": 0, "end_A": 0,
            "posB": 0, "positions_B": side(), "start_B": 0, "end_B": 0}


def _fold_pair(cell, labels, same_chrom, max_ins_len):
    """tiddit_cluster.pyx:161-254: signals in insertion order -> {cluster id: candidate}."""
    out = {}
    n = len(cell.records)
    extra_contigs = 0
    for rec, label in zip(cell.records, labels):
        cid = int(label)
        if cid == -1:
            # noise survives only as a short intra-chromosomal assembly contig, with a fresh id (:164-168)
            if not (same_chrom and rec[2] == "A" and (int(rec[5]) - int(rec[3])) < max_ins_len * 2):
                continue
            cid = n + extra_contigs
            extra_contigs += 1
        cand = out.get(cid)
        if cand is None:
            cand = out[cid] = _new_candidate()
        name, sample, kind = rec[0], rec[1], rec[2]
        if sample not in cand["samples"]:
            cand["sample_discordants"][sample] = set()
            cand["sample_splits"][sample] = set()
            cand["sample_contigs"][sample] = set()
            cand["samples"].add(sample)
        A, B = cand["positions_A"], cand["positions_B"]
        A["start"].append(rec[8])
        A["end"].append(rec[9])
        B["start"].append(rec[10])
        B["end"].append(rec[11])
        key = _KIND_KEY[kind]
        cand[key].add(name)
        A[key].append(int(rec[3]))
        A["orientation_" + key].append(rec[4])
        B[key].append(int(rec[5]))
        B["orientation_" + key].append(rec[6])
        cand["sample_" + key][sample].add(name)
    return out


def _mode(values):
    return Counter(values).most_common(1)[0][0]   # first-inserted wins ties, like the reference


def _breakpoints(cand, is_mp, min_reads):
    """tiddit_cluster.pyx:265-330: representative posA / posB of one candidate."""
    A, B = cand["positions_A"], cand["positions_B"]
    if cand["N_splits"] and min_reads <= cand["N_splits"]:
        return _mode(A["splits"]), _mode(B["splits"])
    if cand["N_contigs"]:
        return _mode(A["contigs"]), _mode(B["contigs"])
    if cand["N_splits"]:
        return _mode(A["splits"]), _mode(B["splits"])
    rev_a, fwd_a = A["orientation_discordants"].count("True"), A["orientation_discordants"].count("False")
    rev_b, fwd_b = B["orientation_discordants"].count("True"), B["orientation_discordants"].count("False")
    consistent_a = rev_a >= 5 * fwd_a or rev_a * 5 <= fwd_a
    consistent_b = rev_b >= 5 * fwd_b or rev_b * 5 <= fwd_b
    if not (consistent_a and consistent_b):
        return _mode(A["discordants"]), _mode(B["discordants"])
    a_rev, b_rev = rev_a > fwd_a, rev_b > fwd_b
    # a reverse read points left (breakpoint = smallest position) in a paired-end library; mate-pair
    # libraries are mirrored (:289-322)
    take_max_a = a_rev if is_mp else not a_rev
    take_max_b = b_rev if is_mp else not b_rev
    posA = max(A["discordants"]) if take_max_a else min(A["discordants"])
    posB = max(B["discordants"]) if take_max_b else min(B["discordants"])
    return posA, posB


def main(prefix, chromosomes, contig_length, samples, is_mp, epsilon, m, max_ins_len, min_contig, skip_assembly,
         min_reads):
    """tiddit_cluster.pyx:39-338 -> candidates[chrA][chrB][cluster id] (same keys, same insertion order)."""
    pairs = read_signals(prefix, contig_length, samples, is_mp, min_contig, skip_assembly)

    order = []       # (chrA, chrB) in the reference's visiting order (:140-150)
    candidates = {}
    for chrA in chromosomes:
        if chrA not in pairs:
            continue
        candidates.setdefault(chrA, {})
        for chrB in chromosomes:
            if chrB in pairs[chrA] and (chrA, chrB) not in order:
                order.append((chrA, chrB))
    labels = label_signals([pairs[a][b] for a, b in order], epsilon, m)
    for (chrA, chrB), lab in zip(order, labels):
        candidates[chrA][chrB] = _fold_pair(pairs[chrA][chrB], lab, chrA == chrB, max_ins_len)

    for row in candidates.values():
        for cell in row.values():
            for cand in cell.values():
                cand["N_discordants"] = len(cand["discordants"])
                cand["N_splits"] = len(cand["splits"])
                cand["N_contigs"] = len(cand["contigs"])
                cand["posA"], cand["posB"] = _breakpoints(cand, is_mp, min_reads)
                cand["startB"] = min(cand["positions_B"]["start"])
                cand["endB"] = max(cand["positions_B"]["end"])
                cand["startA"] = min(cand["positions_A"]["start"])
                cand["endA"] = max(cand["positions_A"]["end"])
    return candidates
