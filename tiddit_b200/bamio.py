"""Minimal BAM access for the coverage path.

The reference reads alignments through pysam (`AlignmentFile(...).fetch(until_eof=True)`,
`__main__.py:225-242`).  pysam / htslib are not part of this image, so `open_alignment_file` returns
the real `pysam.AlignmentFile` when it is importable and otherwise this pure-Python reader, which offers
the attributes the two coverage loops touch: header["SQ"], fetch(until_eof=True), and per read
is_unmapped, is_duplicate, is_secondary, is_supplementary, mapq / mapping_quality, reference_start,
reference_end, reference_name, query_name, flag.  `write_bam` produces small BGZF-compressed BAM files for
tests and benchmarks.
"""
import gzip
import struct
import zlib

import numpy as np

_CIGAR_REF = (1, 0, 1, 1, 0, 0, 0, 1, 1)   # M I D N S H P = X : consumes reference?
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


class AlignedRead:
    __slots__ = ("query_name", "flag", "reference_id", "reference_start", "reference_end", "mapping_quality",
                 "reference_name", "cigartuples", "next_reference_id", "next_reference_start", "template_length")

    @property
    def mapq(self):
        return self.mapping_quality

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_duplicate(self):
        return bool(self.flag & 0x400)

    @property
    def is_secondary(self):
        return bool(self.flag & 0x100)

    @property
    def is_supplementary(self):
        return bool(self.flag & 0x800)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_paired(self):
        return bool(self.flag & 0x1)


class AlignmentFile:
    """Sequential BAM reader (no index, no CRAM)."""

    def __init__(self, path, mode="r", reference_filename=None, **_):
        self._fh = gzip.open(path, "rb")      # BGZF is a series of gzip members
        head = self._fh.read(8)
        if head[:4] != b"BAM\x01":
            raise ValueError("%s is not a BAM file" % path)
        l_text, = struct.unpack_from("<i", head, 4)
        text = self._fh.read(l_text).split(b"\0", 1)[0].decode("ascii", "replace")
        n_ref, = struct.unpack("<i", self._fh.read(4))
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name, = struct.unpack("<i", self._fh.read(4))
            self.references.append(self._fh.read(l_name)[:-1].decode("ascii"))
            self.lengths.append(struct.unpack("<i", self._fh.read(4))[0])
        self.text = text
        self.header = {"HD": {}, "SQ": [{"SN": n, "LN": l} for n, l in zip(self.references, self.lengths)], "RG": []}
        for line in text.splitlines():
            if line.startswith("@RG"):
                self.header["RG"].append(dict(f.split(":", 1) for f in line.split("\t")[1:] if ":" in f))

    def fetch(self, contig=None, until_eof=False, **_):
        if contig is not None:
            raise NotImplementedError("the pure-Python BAM reader is sequential: use fetch(until_eof=True)")
        read_block = self._fh.read
        unpack = struct.Struct("<iiBBHHHiiii").unpack_from
        while True:
            raw = read_block(4)
            if len(raw) < 4:
                return
            size, = struct.unpack("<i", raw)
            rec = read_block(size)
            ref_id, pos, l_name, mapq, _bin, n_cigar, flag, l_seq, next_ref, next_pos, tlen = unpack(rec, 0)
            r = AlignedRead()
            r.query_name = rec[32:32 + l_name - 1].decode("ascii")
            r.flag, r.reference_id, r.reference_start, r.mapping_quality = flag, ref_id, pos, mapq
            r.next_reference_id, r.next_reference_start, r.template_length = next_ref, next_pos, tlen
            r.reference_name = self.references[ref_id] if ref_id >= 0 else None
            ops = struct.unpack_from("<%dI" % n_cigar, rec, 32 + l_name) if n_cigar else ()
            r.cigartuples = [(o & 0xF, o >> 4) for o in ops]
            if flag & 0x4 or not ops:
                r.reference_end = None
            else:
                r.reference_end = pos + sum(n for op, n in r.cigartuples if op < 9 and _CIGAR_REF[op])
            yield r

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def open_alignment_file(path, reference_filename=None):
    """pysam.AlignmentFile when pysam is installed, the pure-Python reader otherwise."""
    try:
        import pysam
        return pysam.AlignmentFile(path, "r", reference_filename=reference_filename)
    except ImportError:
        return AlignmentFile(path, "r", reference_filename=reference_filename)


def _bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bam(path, contigs, reads, header_text=None):
    """contigs: [(name, length)]; reads: iterable of dicts with keys name, flag, ref (index or -1), pos (0-based),
    mapq, cigar [(op, len)] (op as in BAM: 0=M 1=I 2=D 3=N 4=S ...), optional seq_len."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in contigs)
    out = bytearray(b"BAM\x01")
    text = header_text.encode("ascii")
    out += struct.pack("<i", len(text)) + text + struct.pack("<i", len(contigs))
    for name, length in contigs:
        nm = name.encode("ascii") + b"\0"
        out += struct.pack("<i", len(nm)) + nm + struct.pack("<i", length)
    for rd in reads:
        nm = rd["name"].encode("ascii") + b"\0"
        cigar = rd.get("cigar") or []
        l_seq = rd.get("seq_len", sum(n for op, n in cigar if op in (0, 1, 4, 7, 8)))
        ref_len = sum(n for op, n in cigar if _CIGAR_REF[op])
        end = rd["pos"] + (ref_len or 1)
        # UCSC binning scheme (SAM spec 5.3)
        b, e = rd["pos"], end - 1
        if b >> 14 == e >> 14: bin_ = ((1 << 15) - 1) // 7 + (b >> 14)
        elif b >> 17 == e >> 17: bin_ = ((1 << 12) - 1) // 7 + (b >> 17)
        elif b >> 20 == e >> 20: bin_ = ((1 << 9) - 1) // 7 + (b >> 20)
        elif b >> 23 == e >> 23: bin_ = ((1 << 6) - 1) // 7 + (b >> 23)
        elif b >> 26 == e >> 26: bin_ = ((1 << 3) - 1) // 7 + (b >> 26)
        else: bin_ = 0
        body = struct.pack("<iiBBHHHiiii", rd["ref"], rd["pos"], len(nm), rd["mapq"], bin_, len(cigar), rd["flag"],
                           l_seq, -1, -1, 0)
        body += nm + b"".join(struct.pack("<I", (n << 4) | op) for op, n in cigar)
        body += b"\xff" * ((l_seq + 1) // 2) + b"\xff" * l_seq
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        for i in range(0, len(out), 0xff00):
            f.write(_bgzf_block(bytes(out[i:i + 0xff00])))
        f.write(_BGZF_EOF)


def synthetic_reads(contigs, n_reads, seed=1, read_len=150):
    """BASELINE config 1 shape: coordinate-sorted 150-bp reads, uniform starts, mapq uniform 0..60, 2 % duplicates,
    1 % unmapped, a few soft-clipped / spliced CIGARs."""
    rng = np.random.default_rng(seed)
    lens = np.array([l for _, l in contigs], dtype=np.int64)
    per = np.floor(n_reads * lens / lens.sum()).astype(int)
    per[0] += n_reads - per.sum()
    reads = []
    for ci, (name, ln) in enumerate(contigs):
        starts = np.sort(rng.integers(0, ln, per[ci]))
        for k, s in enumerate(starts):
            flag = 0
            u = rng.random()
            if u < 0.02: flag |= 0x400
            elif u < 0.03: flag |= 0x4
            if rng.random() < 0.05: flag |= 0x100 if rng.random() < 0.5 else 0x800
            length = int(min(read_len, ln - s))
            v = rng.random()
            if v < 0.1 and length > 40:
                cigar = [(4, 20), (0, length - 20)]
            elif v < 0.15 and length > 60:
                cigar = [(0, 30), (2, 5), (0, length - 35)] if s + length + 5 <= ln else [(0, length)]
            elif v < 0.2 and length > 60:
                cigar = [(0, 40), (1, 7), (0, length - 47)]
            else:
                cigar = [(0, length)]
            reads.append({"name": "r%d_%d" % (ci, k), "flag": flag, "ref": ci, "pos": int(s),
                          "mapq": int(rng.integers(0, 61)), "cigar": cigar})
    return reads
