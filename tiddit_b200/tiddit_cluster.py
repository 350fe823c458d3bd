"""Drop-in for tiddit/tiddit_cluster.pyx: `find_discordant_pos` (:7-37) and `main` (:39-338).

`main` keeps the reference's signature and returns the same nested `candidates` dict (schema in SURVEY.md
App. D).  Instead of one sorted() + DBSCAN.main call per (chrA,chrB) list and a Python fold per signal it

  1. packs every pair's records into int32 / uint8 arrays (`signals.PackedSignals`),
  2. labels ALL pairs with one GPU call (device_ops.cluster_labels -> tdt_cluster_labels, replaces :140-160),
  3. aggregates ALL candidates with one GPU call (device_ops.cluster_aggregate -> tdt_cluster_aggregate,
     replaces the per-signal fold :156-254 and the per-candidate statistics :258-336), and
  4. only unpacks the result into the reference's dicts (name sets and per-signal lists, which are strings and
     Python containers by contract) on the host.

`main_packed` / `cluster_packed` are the same on packed arrays and return the `CandidateTable` without building
dicts -- the form a packed downstream consumes.
"""
import numpy as np

from . import device_ops
from .signals import CandidateTable, PackedSignals, find_discordant_pos, KIND_CHAR

__all__ = ["find_discordant_pos", "main", "main_packed", "cluster_packed", "candidates_from_table"]

_KIND_KEY = ("discordants", "splits", "contigs")


def cluster_packed(packed, epsilon, m, max_ins_len, is_mp, min_reads):
    """PackedSignals -> (labels int32 [n] in insertion order, CandidateTable).  Two GPU calls, no per-signal host
    work (tiddit_cluster.pyx:140-336 minus the dict building)."""
    n = len(packed)
    if n == 0:
        return np.zeros(0, dtype=np.int32), CandidateTable(np.zeros((0, 16), dtype=np.int32), np.zeros(0, dtype=np.int32))
    if min(int(packed.posA.min()), int(packed.posB.min())) < 0:
        raise OverflowError("signal positions must lie in [0, 2^30)")
    max_pos = packed.max_pos()
    if max_pos >= 2 ** 30:
        raise OverflowError("signal positions must lie in [0, 2^30)")
    labels, rows, members = device_ops.cluster_and_aggregate(packed.posA, packed.posB, packed.seg_off, packed.span,
                                                             packed.name_id, packed.flags, packed.same_chrom, epsilon, m,
                                                             max_ins_len, is_mp, min_reads, max_pos,
                                                             max(len(packed.names), 1))
    return labels, CandidateTable(rows, members)


def _new_candidate():
    side = lambda: {"contigs": [], "splits": [], "discordants": [], "orientation_contigs": [],
                    "orientation_splits": [], "orientation_discordants": [], "start": [], "end": []}
    return {"signal_type": {}, "samples": set(), "sample_discordants": {}, "sample_splits": {}, "sample_contigs": {},
            "N_discordants": 0, "discordants": set(), "N_splits": 0, "splits": set(), "N_contigs": 0,
            "contigs": set(), "n_signals": 0,
            "posA": 0, "positions_A": side(), "start_A": 0, "end_A": 0,
            "posB": 0, "positions_B": side(), "start_B": 0, "end_B": 0}


def candidates_from_table(packed, table):
    """CandidateTable -> candidates[chrA][chrB][id] exactly as tiddit_cluster.pyx:156-336 leaves it: the numbers
    (N_*, posA/posB, start/end) come from the device rows, the name sets and per-signal lists are unpacked from the
    member lists (members are in insertion order, rows in dict insertion order)."""
    candidates = {a: {} for a in packed.chrA_present}
    for a, b in packed.pairs:
        candidates.setdefault(a, {})[b] = {}
    if len(table) == 0:
        return candidates
    # Millions of small containers are created below and none of them is garbage: with the cyclic collector on,
    # its full passes over the growing heap took two thirds of the time (4.8 of 7.7 s for 61 k candidates).
    import gc
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        return _fill_candidates(candidates, packed, table)
    finally:
        if was_enabled:
            gc.enable()


def _names_of(packed, ids):
    """Read names of `ids` (numpy) as Python strings; one bulk decode when the table is a native blob."""
    names = packed.names
    blob = getattr(names, "_blob", None)
    if blob is not None and blob.isascii():
        text = blob.decode("ascii")                       # byte offsets are character offsets
        off = names._off
        return [text[a:b] for a, b in zip(off[ids].tolist(), off[ids + 1].tolist())]
    return [names[i] for i in ids.tolist()]


def _fill_candidates(candidates, packed, table):
    """The reference's fold (tiddit_cluster.pyx:156-254) without a Python step per signal: members are grouped by
    (candidate, kind) with one stable sort, every per-signal column becomes ONE Python list, and a candidate's lists and
    name sets are slices of those (list slicing and set() run at C speed).  Candidates whose members come from several
    samples take the per-member loop (`_fill_one_multi_sample`); for one sample -- every single-sample run -- the loop
    below does ~25 slice operations per candidate."""
    rows_np, mem = table.rows, table.member_idx
    C, M = len(rows_np), len(mem)
    lo_np, size_np = rows_np[:, 3].astype(np.int64), rows_np[:, 4].astype(np.int64)
    by_off = np.argsort(lo_np, kind="stable")
    if M == 0 or not (np.array_equal(lo_np[by_off], np.concatenate([[0], np.cumsum(size_np[by_off])[:-1]]))
                      and int(size_np.sum()) == M):
        raise ValueError("candidate member ranges do not tile the member index")
    cand_of = np.repeat(by_off, size_np[by_off])                      # candidate (row) of every member
    kind_np = (packed.flags[mem] & 3).astype(np.int64)
    key2 = cand_of * 3 + kind_np
    order2 = np.argsort(key2, kind="stable")                          # members grouped by (candidate, kind), order kept
    off2 = np.concatenate([[0], np.cumsum(np.bincount(key2, minlength=3 * C))]).tolist()
    mem2 = mem[order2]
    posA, posB = packed.posA[mem2].tolist(), packed.posB[mem2].tolist()
    ori_table = packed.ori_table
    oriA = [ori_table[i] for i in packed.oriA_id[mem2].tolist()]
    oriB = [ori_table[i] for i in packed.oriB_id[mem2].tolist()]
    names = _names_of(packed, packed.name_id[mem2].astype(np.int64))
    span = packed.span[mem]
    sA, eA, sB, eB = span[:, 0].tolist(), span[:, 1].tolist(), span[:, 2].tolist(), span[:, 3].tolist()
    # one sample per candidate?  (min == max of the members' sample ids)
    smp = packed.sample_id[mem]
    starts = lo_np[by_off]
    smin = np.empty(C, dtype=np.int64)
    smax = np.empty(C, dtype=np.int64)
    smin[by_off] = np.minimum.reduceat(smp, starts)
    smax[by_off] = np.maximum.reduceat(smp, starts)
    single = (smin == smax).tolist()
    smin = smin.tolist()
    samples = packed.samples
    pairs = packed.pairs
    for r, row in enumerate(rows_np.tolist()):
        chrA, chrB = pairs[row[0]]
        lo, hi = row[3], row[3] + row[4]
        if not single[r]:
            cand = candidates[chrA][chrB][row[1]] = _new_candidate()
            _fill_one_multi_sample(cand, packed, mem[lo:hi])
        else:
            sample = samples[smin[r]]
            d0, d1, s1, c1 = off2[3 * r], off2[3 * r + 1], off2[3 * r + 2], off2[3 * r + 3]
            nd, ns, nc = set(names[d0:d1]), set(names[d1:s1]), set(names[s1:c1])
            cand = candidates[chrA][chrB][row[1]] = {
                "signal_type": {}, "samples": {sample},
                "sample_discordants": {sample: nd.copy()}, "sample_splits": {sample: ns.copy()},
                "sample_contigs": {sample: nc.copy()},
                "N_discordants": 0, "discordants": nd, "N_splits": 0, "splits": ns, "N_contigs": 0,
                "contigs": nc, "n_signals": 0,
                "posA": 0,
                "positions_A": {"contigs": posA[s1:c1], "splits": posA[d1:s1], "discordants": posA[d0:d1],
                                "orientation_contigs": oriA[s1:c1], "orientation_splits": oriA[d1:s1],
                                "orientation_discordants": oriA[d0:d1], "start": sA[lo:hi], "end": eA[lo:hi]},
                "start_A": 0, "end_A": 0,
                "posB": 0,
                "positions_B": {"contigs": posB[s1:c1], "splits": posB[d1:s1], "discordants": posB[d0:d1],
                                "orientation_contigs": oriB[s1:c1], "orientation_splits": oriB[d1:s1],
                                "orientation_discordants": oriB[d0:d1], "start": sB[lo:hi], "end": eB[lo:hi]},
                "start_B": 0, "end_B": 0}
        cand["N_discordants"], cand["N_splits"], cand["N_contigs"] = row[5], row[6], row[7]
        cand["posA"], cand["posB"] = row[8], row[9]
        cand["startB"], cand["endB"] = row[12], row[13]
        cand["startA"], cand["endA"] = row[10], row[11]
    return candidates


def _fill_one_multi_sample(cand, packed, members):
    """One candidate, member by member, exactly in the reference's order of operations (several samples)."""
    A, B = cand["positions_A"], cand["positions_B"]
    span = packed.span[members]
    A["start"], A["end"], B["start"], B["end"] = span[:, 0].tolist(), span[:, 1].tolist(), span[:, 2].tolist(), span[:, 3].tolist()
    kind = (packed.flags[members] & 3).tolist()
    posA, posB = packed.posA[members].tolist(), packed.posB[members].tolist()
    names = [packed.names[i] for i in packed.name_id[members].tolist()]
    samples = [packed.samples[i] for i in packed.sample_id[members].tolist()]
    oriA = [packed.ori_table[i] for i in packed.oriA_id[members].tolist()]
    oriB = [packed.ori_table[i] for i in packed.oriB_id[members].tolist()]
    for j in range(len(kind)):
        sample, name, key = samples[j], names[j], _KIND_KEY[kind[j]]
        if sample not in cand["samples"]:
            cand["sample_discordants"][sample] = set()
            cand["sample_splits"][sample] = set()
            cand["sample_contigs"][sample] = set()
            cand["samples"].add(sample)
        cand[key].add(name)
        A[key].append(posA[j])
        A["orientation_" + key].append(oriA[j])
        B[key].append(posB[j])
        B["orientation_" + key].append(oriB[j])
        cand["sample_" + key][sample].add(name)


def main_packed(packed, epsilon, m, max_ins_len, is_mp, min_reads):
    """`main` on packed arrays -> the reference's candidates dict."""
    _, table = cluster_packed(packed, epsilon, m, max_ins_len, is_mp, min_reads)
    return candidates_from_table(packed, table)


def main(prefix, chromosomes, contig_length, samples, is_mp, epsilon, m, max_ins_len, min_contig, skip_assembly,
         min_reads):
    """tiddit_cluster.pyx:39-338 -> candidates[chrA][chrB][cluster id] (same keys, same insertion order)."""
    packed = PackedSignals.from_tab(prefix, chromosomes, contig_length, samples, is_mp, min_contig, skip_assembly)
    return main_packed(packed, epsilon, m, max_ins_len, is_mp, min_reads)
