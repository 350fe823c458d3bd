"""Drop-in for tiddit/tiddit_signal.pyx: signal extraction with the coverage on the GPU.

The reference runs one joblib worker per contig (tiddit_signal.pyx:147-228, :259): each walks its reads through
pysam, calls `update_coverage` once per read (:181-182) and collects clipped reads (:191-197), split reads
(`SA_analysis`, :31-145) and discordant pairs (:204-221); `main` (:230-334) pairs the mates and writes
`discordants_<sample>.tab`, `splits_<sample>.tab`, `clips_<sample>.fa` and returns the per-contig coverage.

Here the BAM is scanned once by libtdt_bam.so (include/tdt_bam.h: BGZF inflated on all host cores, records decoded
into columns).  The (start, end) columns of every batch go to the GPU coverage kernel in one
`tdt_coverage_accumulate_contigs` call (tiddit_coverage.DeviceCoverage); the filters of the worker loop are
evaluated on the columns with numpy, and only the reads that carry a signal -- a few per thousand -- are looked at
record by record.  Same call signature, same three files byte for byte, same returned dict.

`collect` is the array-level entry point (no files); `main` keeps the reference's signature and side effects.
"""
import itertools
import os

import numpy as np

from . import bamio, tiddit_coverage

__all__ = ["main", "collect", "packed_signals", "discordant_lines", "split_lines", "SA_analysis", "find_SA_query_range"]

_SA_OPS = {"M": 0, "S": 4, "H": 5, "D": 2, "I": 1}   # the operations the reference knows (tiddit_signal.pyx:24)


class _SASegment:
    """What tiddit_signal.pyx:12-29 builds as a pysam.AlignedSegment from one SA entry: reference_start is the SA
    position as written (1-based, not shifted), the CIGAR is parsed with the five-letter table above (any other
    operation letter raises KeyError there, and here)."""
    __slots__ = ("reference_start", "reference_end", "query_alignment_start", "query_alignment_end", "flag", "cigar")

    def __init__(self, SA):
        self.reference_start = int(SA[1])
        self.flag = 64 if SA[2] == "+" else 80
        parts = ["".join(g) for _, g in itertools.groupby(SA[3], key=str.isdigit)]
        self.cigar = tuple((_SA_OPS[parts[2 * i + 1]], int(parts[2 * i])) for i in range(int(len(parts) / 2)))
        # pysam semantics for a segment without sequence (see bamio.AlignedRead)
        self.reference_end = (self.reference_start + sum(n for op, n in self.cigar if op in (0, 2))) if self.cigar else None
        start = 0
        for op, n in self.cigar:
            if op == 5:
                continue
            if op != 4:
                break
            start += n
        self.query_alignment_start = start
        end = 0
        for op, n in self.cigar:
            if op in (0, 1) or (op == 4 and end == 0):
                end += n
        self.query_alignment_end = end


def find_SA_query_range(SA):
    """tiddit_signal.pyx:12-29."""
    return _SASegment(SA)


def SA_analysis(read, min_q, tag, reference_name):
    """tiddit_signal.pyx:31-145 -> [chrA, chrB, name, split_pos, is_reverse, SA_split_pos, SA_is_reverse, startA, endA,
    startB, endB] or () when the supplementary alignment fails the quality threshold.

    With several SA entries the reference's selection loop re-reads entry 0 on every iteration (:39-45), so either
    every entry "passes" or none does, all candidate lengths are equal and entry 0 stays selected (:50-61): only the
    first entry is ever used."""
    entries = read.get_tag(tag).rstrip(";").split(";")
    SA = entries[0].split(",")
    if len(entries) > 1 and int(SA[4]) >= min_q:
        _SASegment(SA)              # the reference parses it inside the loop (same KeyError on an unknown CIGAR letter)
    if int(SA[4]) < min_q:
        return ()
    seg = _SASegment(SA)
    clip_before = seg.query_alignment_start < read.query_alignment_start
    sa_reverse = SA[2] == "-"
    read_span = (read.reference_start + 1, read.reference_end + 1)
    sa_span = (seg.reference_start, seg.reference_end)
    split_pos = read_span[0] if bool(read.is_reverse) != clip_before else read_span[1]      # :77-90
    sa_split_pos = sa_span[0] if sa_reverse == clip_before else sa_span[1]                   # :100-111
    sa_chr = SA[0]
    swap = sa_chr < reference_name or (sa_chr == reference_name and sa_split_pos < split_pos)    # :114-143
    if sa_chr < reference_name:
        chrA, chrB = sa_chr, reference_name
    else:
        chrA, chrB = reference_name, sa_chr
    if swap:
        split_pos, sa_split_pos = sa_split_pos, split_pos
        spanA, spanB = sa_span, read_span
    else:
        spanA, spanB = read_span, sa_span
    return [chrA, chrB, read.query_name, split_pos, read.is_reverse, sa_split_pos, sa_reverse,
            spanA[0], spanA[1], spanB[0], spanB[1]]


class Signals:
    """Result of `collect`: per-contig record lists in the reference worker's layout + the device coverage."""

    def __init__(self, chromosomes, header, coverage):
        self.chromosomes = chromosomes                  # contigs with LN >= min_contig, header order
        self.header = header
        self.coverage = coverage                        # tiddit_coverage.DeviceCoverage over `chromosomes`
        self.data = {c: [] for c in chromosomes}        # [chrA, chrB, name, start+1, end+1, is_reverse, read contig]
        self.splits = {c: [] for c in chromosomes}      # SA_analysis rows
        self.clips = {c: [] for c in chromosomes}       # (fasta header line, sequence line)
        self.n_reads = 0
        self.n_covered = 0
        self.n_inspected = 0


def collect(bam_file_name, min_q, max_ins, min_contig, min_anchor_len, min_clip_len, bin_size=50, threads=0,
            batch_reads=1 << 20):
    """One pass over the BAM: coverage of every read with mapq >= min_q into HBM, signals into per-contig lists.

    Equivalent to running tiddit_signal.pyx:147-228 for every contig with LN >= min_contig."""
    with bamio.ColumnReader(bam_file_name, threads=threads, batch_reads=batch_reads) as reader:
        header = reader.header
        refs = reader.references
        keep_contig = np.array([ln >= min_contig for ln in reader.lengths], dtype=bool)
        chromosomes = [n for n, k in zip(refs, keep_contig) if k]
        sub_header = {"SQ": [sq for sq, k in zip(header["SQ"], keep_contig) if k]}
        out = Signals(chromosomes, header, tiddit_coverage.DeviceCoverage(sub_header, bin_size))
        for b in reader.batches():
            out.n_reads += len(b)
            flag, ref_id = b.flag, b.ref_id
            on_kept = (ref_id >= 0) & keep_contig[np.maximum(ref_id, 0)]
            # :171-172, :181-182: unmapped / duplicate reads are skipped, the rest counts when mapq >= min_q
            counted = on_kept & ((flag & (0x4 | 0x400)) == 0) & (b.mapq >= min_q)
            idx = np.flatnonzero(counted)
            if len(idx):
                if np.any(b.end[idx] < 0):
                    raise TypeError("an integer is required")      # update_coverage(long None): a mapped read without CIGAR
                rid = ref_id[idx]
                if np.any(rid[1:] < rid[:-1]):
                    idx = idx[np.argsort(rid, kind="stable")]
                    rid = ref_id[idx]
                cuts = np.flatnonzero(rid[1:] != rid[:-1]) + 1
                for lo, hi in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [len(idx)]])):
                    sel = idx[lo:hi]
                    out.coverage.add_reads(refs[rid[lo]], b.pos[sel], b.end[sel])
                out.n_covered += len(idx)
            # :186-221 on the columns: which reads need a record-level look?
            primary = counted & ((flag & (0x100 | 0x800)) == 0)
            same = b.mate_ref == ref_id
            # numpy abs of INT32_MIN stays negative; no template is that long, the int64 view keeps it exact anyway
            tl = np.abs(b.tlen.astype(np.int64))
            f_op, f_len = b.cig_first & 15, b.cig_first >> 4
            l_op, l_len = b.cig_last & 15, b.cig_last >> 4
            clipped = primary & (tl < max_ins) & same & (
                ((f_op == 4) & (f_len > min_clip_len) & (l_op == 0) & (l_len > min_anchor_len)) |
                ((l_op == 4) & (l_len > min_clip_len) & (f_op == 0) & (f_len > min_anchor_len)))
            split = primary & (b.has_sa != 0)
            disc = primary & ((flag & 0x8) == 0) & ((flag & 0x1) != 0) & ((tl > max_ins) | ~same)
            look = np.flatnonzero(clipped | split | disc)
            out.n_inspected += len(look)
            for k in look:
                read = b.record(int(k))
                contig = refs[ref_id[k]]
                if clipped[k]:
                    out.clips[contig].append((">{}|{}|{}\n".format(read.query_name, contig, read.reference_start + 1),
                                              read.query_sequence + "\n"))
                if split[k]:
                    row = SA_analysis(read, min_q, "SA", contig)
                    if row:
                        out.splits[contig].append(row)
                if disc[k]:
                    mate = read.next_reference_name
                    if mate < contig:                  # contig NAMES compared as strings (:210)
                        chrA, chrB = mate, contig
                    else:
                        chrA, chrB = contig, mate
                    out.data[contig].append([chrA, chrB, read.query_name, read.reference_start + 1, read.reference_end + 1,
                                             read.is_reverse, contig])
        out.coverage.flush()
    return out


def _merge(chromosomes, header, per_contig, extend):
    """tiddit_signal.pyx:247-286: [chrA][chrB][read name] -> fields, dict insertion order = the order the files are
    written in.  `extend` False: a list of per-read records (discordants); True: one flat list (splits)."""
    merged = {a: {sq["SN"]: {} for sq in header["SQ"]} for a in chromosomes}
    for contig in chromosomes:
        for rec in per_contig[contig]:
            if rec[0] not in merged:
                continue
            slot = merged[rec[0]][rec[1]]
            if extend:
                slot.setdefault(rec[2], []).extend(rec[3:])
            else:
                slot.setdefault(rec[2], []).append(rec[3:])
    return merged


def discordant_lines(signals):
    """The lines of discordants_<sample>.tab (tiddit_signal.pyx:298-318): fragments seen at least twice; the first
    two records, the one on chrA first.  For chrA == chrB the reference compares the two records' contig names
    (equal strings), so they stay in file order."""
    data = _merge(signals.chromosomes, signals.header, signals.data, extend=False)
    for chrA, row in data.items():
        for chrB, fragments in row.items():
            for name, recs in fragments.items():
                if len(recs) < 2:
                    continue
                first, second = recs[0], recs[1]
                if chrA != chrB and first[-1] != chrA:
                    first, second = second, first
                yield "{}\t{}\t{}\t{}\n".format(name, chrA, chrB, "\t".join(map(str, first[:-1] + second[:-1])))


def split_lines(signals):
    """The lines of splits_<sample>.tab (tiddit_signal.pyx:320-326)."""
    splits = _merge(signals.chromosomes, signals.header, signals.splits, extend=True)
    for chrA, row in splits.items():
        for chrB, fragments in row.items():
            for name, fields in fragments.items():
                yield "{}\t{}\t{}\t{}\n".format(name, chrA, chrB, "\t".join(map(str, fields)))


def _record_parts(signals, sample_k, contig_length, is_mp, min_contig):
    """The discordant and split records as the column sets PackedSignals assembles -- what reading the two tab files
    back (tiddit_cluster.pyx:47-107) yields, without formatting and re-parsing the text."""
    from .signals import KIND_D, KIND_S, find_discordant_pos

    def new_part(kind):
        return dict(kind=kind, sample=sample_k, name=[], chrA=[], chrB=[], posA=[], posB=[], oriA=[], oriB=[], span=[])

    disc, split = new_part(KIND_D), new_part(KIND_S)
    data = _merge(signals.chromosomes, signals.header, signals.data, extend=False)
    for chrA, row in data.items():
        for chrB, fragments in row.items():
            if contig_length[chrA] < min_contig or contig_length[chrB] < min_contig:
                continue
            la, lb = contig_length[chrA], contig_length[chrB]
            for name, recs in fragments.items():
                if len(recs) < 2:
                    continue
                first, second = recs[0], recs[1]
                if chrA != chrB and first[-1] != chrA:
                    first, second = second, first
                # the tab line's fields 3..8: startA, endA, reverseA, startB, endB, reverseB
                f = [name, chrA, chrB, first[0], first[1], str(first[2]), second[0], second[1], str(second[2])]
                posA, posB = find_discordant_pos(f, is_mp)
                if posA > la:                      # tiddit_cluster.pyx:67-70, the nested test overwrites posA
                    posA = la
                    if posB > lb:
                        posA = lb
                disc["name"].append(name)
                disc["chrA"].append(chrA)
                disc["chrB"].append(chrB)
                disc["posA"].append(posA)
                disc["posB"].append(posB)
                disc["oriA"].append(f[5])
                disc["oriB"].append(f[8])
                disc["span"].append((f[3], f[4], f[6], f[7]))
    splits = _merge(signals.chromosomes, signals.header, signals.splits, extend=True)
    for chrA, row in splits.items():
        for chrB, fragments in row.items():
            if contig_length[chrA] < min_contig or contig_length[chrB] < min_contig:
                continue
            la, lb = contig_length[chrA], contig_length[chrB]
            for name, f in fragments.items():      # f: split_pos, reverse, SA_split_pos, SA reverse, startA, endA, startB, endB, ...
                if None in f[:8]:
                    raise ValueError("invalid literal for int() with base 10: 'None'")   # what reading the file back does
                split["name"].append(name)
                split["chrA"].append(chrA)
                split["chrB"].append(chrB)
                split["posA"].append(min(f[0], la))
                split["posB"].append(min(f[2], lb))
                split["oriA"].append(str(f[1]))
                split["oriB"].append(str(f[3]))
                split["span"].append((f[4], f[5], f[6], f[7]))
    return disc, split


def packed_signals(signals, sample_id, is_mp, min_contig, contig_lines=None):
    """Signals -> signals.PackedSignals, the arrays tiddit_cluster.cluster_packed takes: the records
    `main` writes to the tab files, handed over in memory without the text round trip.  contig_lines: the lines of
    contigs_<sample>.tab when the assembly stage ran, None for --skip_assembly."""
    from .signals import KIND_A, PackedSignals, _part_from_lines
    contig_length = {sq["SN"]: sq["LN"] for sq in signals.header["SQ"]}
    chromosomes = [sq["SN"] for sq in signals.header["SQ"]]     # tiddit/__main__.py:119-123: every @SQ, header order
    parts = list(_record_parts(signals, 0, contig_length, is_mp, min_contig))
    if contig_lines is not None:
        parts.append(_part_from_lines(contig_lines, 0, KIND_A, contig_length, is_mp, min_contig))
    return PackedSignals._assemble(parts, chromosomes, [sample_id])


def main(bam_file_name, ref, prefix, min_q, max_ins, sample_id, threads, min_contig, skip_index, min_anchor_len,
         min_clip_len):
    """tiddit_signal.pyx:230-334: writes <prefix>_tiddit/{discordants,splits}_<sample>.tab, clips/<contig>.fa,
    clips_<sample>.fa and returns {contig: float64 coverage per 50-bp bin} for the contigs with LN >= min_contig.
    `skip_index` is accepted for compatibility (the scan is sequential, no index is read); `ref` only matters for
    CRAM input, which this reader rejects with an explicit error (bamio.require_bam)."""
    from . import bamio
    bamio.require_bam(bam_file_name, ref)
    signals = collect(bam_file_name, int(min_q), int(max_ins), int(min_contig), int(min_anchor_len), int(min_clip_len),
                      bin_size=50, threads=int(threads))
    print("Writing signals to file")
    out_dir = "{}_tiddit".format(prefix)
    os.makedirs(os.path.join(out_dir, "clips"), exist_ok=True)
    with open("{}/discordants_{}.tab".format(out_dir, sample_id), "w") as f:
        f.writelines(discordant_lines(signals))
    with open("{}/splits_{}.tab".format(out_dir, sample_id), "w") as f:
        f.writelines(split_lines(signals))
    with open("{}/clips_{}.fa".format(out_dir, sample_id), "w") as merged:
        for contig in signals.chromosomes:
            with open("{}/clips/{}.fa".format(out_dir, contig), "w") as f:
                for head, seq in signals.clips[contig]:
                    f.write(head + seq)
                    merged.write(head + seq)
    coverage_data, _ = signals.coverage.to_host()
    return coverage_data
