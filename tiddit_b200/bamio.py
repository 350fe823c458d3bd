"""BAM access for the coverage / signal path.

The reference reads alignments through pysam (`AlignmentFile(...).fetch(...)`, `__main__.py:225-242`,
`tiddit_signal.pyx:156-221`).  pysam / htslib are not part of this image.  Two readers replace it:

  * `ColumnReader` -- the product path: libtdt_bam.so (include/tdt_bam.h, csrc/tdt_bam.cpp) inflates BGZF
    blocks on all host cores and decodes whole batches of records into numpy columns (reference id, start,
    end, flag, mapq, mate, template length, first / last CIGAR word, "has SA"); the start / end columns go to
    the GPU coverage kernel batch by batch and only signal-carrying reads are materialised (`record(k)`).
  * `AlignmentFile` -- a small pure-Python sequential reader with pysam's attribute names, used by tests as an
    independent cross-check of the scanner and by `open_alignment_file` when pysam is absent.

`AlignedRead` offers the pysam attributes the two reference loops touch: is_unmapped, is_duplicate,
is_secondary, is_supplementary, is_paired, is_reverse, mate_is_unmapped, mapq / mapping_quality,
reference_start, reference_end, reference_name, next_reference_name, isize / template_length, cigartuples,
query_name, query_sequence, query_alignment_start / _end, has_tag, get_tag.  `write_bam` produces small
BGZF-compressed BAM files for tests and benchmarks.
"""
import ctypes
import gzip
import os
import struct
import zlib

import numpy as np

_CIGAR_REF = (1, 0, 1, 1, 0, 0, 0, 1, 1)   # M I D N S H P = X : consumes reference?
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_SEQ_CODE = "=ACMGRSVTWYHKDBN"
_CORE = struct.Struct("<iiBBHHHiiii")
_AUX_FIXED = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
_AUX_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}


class AlignedRead:
    """One BAM record (the bytes after its block_size word), decoded on demand, pysam attribute names."""
    __slots__ = ("_rec", "_refs", "reference_id", "reference_start", "mapping_quality", "flag", "next_reference_id",
                 "next_reference_start", "template_length", "_l_name", "_n_cigar", "_l_seq", "_cigar", "_tags")

    def __init__(self, rec, references):
        self._rec, self._refs = rec, references
        (self.reference_id, self.reference_start, self._l_name, self.mapping_quality, _bin, self._n_cigar, self.flag,
         self._l_seq, self.next_reference_id, self.next_reference_start, self.template_length) = _CORE.unpack_from(rec, 0)
        self._cigar = None
        self._tags = None

    # ---- names / flags ---------------------------------------------------------------------------------
    @property
    def query_name(self):
        return bytes(self._rec[32:32 + self._l_name - 1]).decode("ascii")

    @property
    def mapq(self):
        return self.mapping_quality

    @property
    def isize(self):
        return self.template_length

    @property
    def reference_name(self):
        return self._refs[self.reference_id] if self.reference_id >= 0 else None

    @property
    def next_reference_name(self):
        return self._refs[self.next_reference_id] if self.next_reference_id >= 0 else None

    is_paired = property(lambda self: bool(self.flag & 0x1))
    is_unmapped = property(lambda self: bool(self.flag & 0x4))
    mate_is_unmapped = property(lambda self: bool(self.flag & 0x8))
    is_reverse = property(lambda self: bool(self.flag & 0x10))
    is_secondary = property(lambda self: bool(self.flag & 0x100))
    is_duplicate = property(lambda self: bool(self.flag & 0x400))
    is_supplementary = property(lambda self: bool(self.flag & 0x800))

    # ---- alignment -------------------------------------------------------------------------------------
    @property
    def cigartuples(self):
        """[(op, len)], or None for a record without CIGAR (pysam)."""
        if self._cigar is None:
            ops = struct.unpack_from("<%dI" % self._n_cigar, self._rec, 32 + self._l_name) if self._n_cigar else ()
            self._cigar = [(o & 0xF, o >> 4) for o in ops]
        return self._cigar or None

    @property
    def reference_end(self):
        cig = self.cigartuples
        if self.flag & 0x4 or not cig:
            return None
        return self.reference_start + sum(n for op, n in cig if op < 9 and _CIGAR_REF[op])

    @property
    def query_alignment_start(self):
        """Leading soft clip (hard clips skipped), like pysam's getQueryStart."""
        start = 0
        for op, n in self.cigartuples or ():
            if op == 5:
                continue
            if op == 4:
                start += n
            else:
                break
        return start

    @property
    def query_alignment_end(self):
        """pysam's getQueryEnd: sequence length minus the trailing soft clip."""
        end = self._l_seq
        cig = self.cigartuples or ()
        if end == 0:
            for op, n in cig:
                if op in (0, 1, 7, 8) or (op == 4 and end == 0):
                    end += n
            return end
        for op, n in reversed(cig[1:]):
            if op == 5:
                continue
            if op == 4:
                end -= n
            else:
                break
        return end

    @property
    def query_sequence(self):
        if self._l_seq == 0:
            return None
        at = 32 + self._l_name + 4 * self._n_cigar
        packed = np.frombuffer(self._rec, dtype=np.uint8, count=(self._l_seq + 1) // 2, offset=at)
        codes = np.empty(2 * len(packed), dtype=np.uint8)
        codes[0::2], codes[1::2] = packed >> 4, packed & 15
        return "".join(_SEQ_CODE[c] for c in codes[:self._l_seq])

    # ---- aux fields ------------------------------------------------------------------------------------
    def _aux(self):
        if self._tags is None:
            rec, tags = self._rec, {}
            p = 32 + self._l_name + 4 * self._n_cigar + (self._l_seq + 1) // 2 + self._l_seq
            end = len(rec)
            while p + 3 <= end:
                tag, typ = bytes(rec[p:p + 2]).decode("ascii"), chr(rec[p + 2])
                p += 3
                if typ == "A":
                    val = chr(rec[p])
                    p += 1
                elif typ in _AUX_FMT:
                    val = struct.unpack_from(_AUX_FMT[typ], rec, p)[0]
                    p += _AUX_FIXED[typ]
                elif typ in "ZH":
                    z = p
                    while rec[z]:
                        z += 1
                    val = bytes(rec[p:z]).decode("ascii")
                    p = z + 1
                elif typ == "B":
                    sub, cnt = chr(rec[p]), struct.unpack_from("<I", rec, p + 1)[0]
                    val = list(struct.unpack_from("<%d%s" % (cnt, _AUX_FMT[sub][1]), rec, p + 5))
                    p += 5 + cnt * _AUX_FIXED[sub]
                else:
                    raise ValueError("unknown aux type %r" % typ)
                tags.setdefault(tag, val)
            self._tags = tags
        return self._tags

    def has_tag(self, tag):
        return tag in self._aux()

    def get_tag(self, tag):
        try:
            return self._aux()[tag]
        except KeyError:
            raise KeyError("tag '%s' not present" % tag)


def _parse_header_text(text, references, lengths):
    header = {"HD": {}, "SQ": [{"SN": n, "LN": l} for n, l in zip(references, lengths)], "RG": []}
    for line in text.splitlines():
        if line.startswith("@RG"):
            header["RG"].append(dict(f.split(":", 1) for f in line.split("\t")[1:] if ":" in f))
    return header


class AlignmentFile:
    """Sequential pure-Python BAM reader (no index, no CRAM)."""

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
        self.header = _parse_header_text(text, self.references, self.lengths)

    def fetch(self, contig=None, until_eof=False, **_):
        """All records in file order; with `contig`, the records placed on it (a linear scan: there is no index)."""
        want = None if contig is None else self.references.index(contig)
        read_block = self._fh.read
        while True:
            raw = read_block(4)
            if len(raw) < 4:
                return
            size, = struct.unpack("<i", raw)
            r = AlignedRead(read_block(size), self.references)
            if want is None or r.reference_id == want:
                yield r

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def require_bam(path, ref=None):
    """The native scanner reads sequential BAM only.  A CRAM input (which the reference hands to pysam together with
    --ref) is detected up front and reported as such instead of failing deep inside the BGZF block hopper."""
    with open(path, "rb") as f:
        magic = f.read(4)
    if magic == b"CRAM":
        raise NotImplementedError(
            "%s is a CRAM file: libtdt_bam.so reads BAM only (CRAM needs htslib/pysam and the --ref FASTA%s); "
            "convert it with `samtools view -b` first" % (path, "" if ref else ", which was not given"))
    if magic[:2] != b"\x1f\x8b":
        raise ValueError("%s is not a BGZF-compressed BAM file (magic %r)" % (path, magic))


def open_alignment_file(path, reference_filename=None):
    """pysam.AlignmentFile when pysam is installed, the pure-Python reader otherwise."""
    try:
        import pysam
        return pysam.AlignmentFile(path, "r", reference_filename=reference_filename)
    except ImportError:
        return AlignmentFile(path, "r", reference_filename=reference_filename)


# ---------------------------------------------------------------------------------------------------------
# the native scanner
# ---------------------------------------------------------------------------------------------------------
_BAM_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtdt_bam.so")
_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
BAM_SIGNATURES = {
    "tdt_bam_last_error": (ctypes.c_char_p, []),
    "tdt_bam_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_vp)]),
    "tdt_bam_close": (None, [_vp]),
    "tdt_bam_header_text": (_vp, [_vp, ctypes.POINTER(_i64)]),
    "tdt_bam_n_ref": (_i32, [_vp]),
    "tdt_bam_ref_name": (ctypes.c_char_p, [_vp, _i32]),
    "tdt_bam_ref_len": (_i32, [_vp, _i32]),
    "tdt_bam_read_columns": (_i64, [_vp, _i64] + [_vp] * 12),
    "tdt_bam_batch_data": (_vp, [_vp, ctypes.POINTER(_i64)]),
    "tdt_bam_inflate_raw": (ctypes.c_int, [ctypes.c_char_p, _i64, _vp, _i64]),
    "tdt_bam_crc32": (ctypes.c_uint32, [ctypes.c_char_p, _i64]),
}
_bam_lib = None


def bam_lib():
    global _bam_lib
    if _bam_lib is None:
        if not os.path.exists(_BAM_LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m tiddit_b200.build`" % _BAM_LIB_PATH)
        L = ctypes.CDLL(_BAM_LIB_PATH)
        for name, (res, args) in BAM_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bam_lib = L
    return _bam_lib


class ColumnBatch:
    """Columns of one batch of records (numpy views, valid until the reader's next batch)."""
    COLUMNS = (("ref_id", np.int32), ("pos", np.int32), ("end", np.int32), ("mate_ref", np.int32),
               ("mate_pos", np.int32), ("tlen", np.int32), ("flag", np.uint16), ("mapq", np.uint8),
               ("cig_first", np.uint32), ("cig_last", np.uint32), ("has_sa", np.uint8), ("rec_off", np.int64))

    def __init__(self, n, cols, data, references):
        self.n = n
        for (name, _), col in zip(self.COLUMNS, cols):
            setattr(self, name, None if col is None else col[:n])    # None: a column the reader was told to skip
        self._data, self._refs = data, references

    def __len__(self):
        return self.n

    def record(self, k):
        """The k-th record of the batch as an AlignedRead (copies its bytes: it outlives the batch)."""
        off = int(self.rec_off[k])
        size = int(self._data[off:off + 4].view(np.int32)[0])
        return AlignedRead(self._data[off + 4:off + 4 + size].tobytes(), self._refs)


class ColumnReader:
    """`pysam.AlignmentFile(path).fetch(until_eof=True)` as batches of columns (libtdt_bam.so)."""

    def __init__(self, path, threads=0, batch_reads=1 << 20, columns=None):
        """columns: names of the ColumnBatch.COLUMNS to fill (default: all).  The scanner skips the work of the others --
        `has_sa` walks the aux fields at the far end of every record -- and the batch carries None for them."""
        L = bam_lib()
        h = _vp()
        rc = L.tdt_bam_open(os.fsencode(path), int(threads), ctypes.byref(h))
        if rc != 0:
            msg = L.tdt_bam_last_error().decode("utf-8", "replace")
            raise (FileNotFoundError if rc == -2 else ValueError)(msg)
        self._h, self._L = h, L
        n = _i64()
        p = L.tdt_bam_header_text(h, ctypes.byref(n))
        self.text = ctypes.string_at(p, n.value).decode("ascii", "replace")
        self.references = [L.tdt_bam_ref_name(h, i).decode("ascii") for i in range(L.tdt_bam_n_ref(h))]
        self.lengths = [L.tdt_bam_ref_len(h, i) for i in range(len(self.references))]
        self.header = _parse_header_text(self.text, self.references, self.lengths)
        self.batch_reads = int(batch_reads)
        known = [name for name, _ in ColumnBatch.COLUMNS]
        if columns is not None and any(c not in known for c in columns):
            raise ValueError("unknown column in %r (known: %s)" % (columns, ", ".join(known)))
        self._cols = [np.empty(self.batch_reads, dtype=dt) if columns is None or name in columns else None
                      for name, dt in ColumnBatch.COLUMNS]

    def batches(self):
        L, h = self._L, self._h
        ptrs = [None if c is None else c.ctypes.data_as(_vp) for c in self._cols]
        while True:
            n = L.tdt_bam_read_columns(h, self.batch_reads, *ptrs)
            if n < 0:
                raise ValueError(L.tdt_bam_last_error().decode("utf-8", "replace"))
            if n == 0:
                return
            size = _i64()
            base = L.tdt_bam_batch_data(h, ctypes.byref(size))
            data = np.ctypeslib.as_array(ctypes.cast(base, ctypes.POINTER(ctypes.c_uint8)), shape=(size.value,))
            yield ColumnBatch(int(n), self._cols, data, self.references)

    def close(self):
        if self._h:
            self._L.tdt_bam_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _bgzf_block(data, level=6):
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bam(path, contigs, reads, header_text=None):
    """contigs: [(name, length)]; reads: iterable of dicts with keys name, flag, ref (index or -1), pos (0-based),
    mapq, cigar [(op, len)] (op as in BAM: 0=M 1=I 2=D 3=N 4=S ...); optional seq_len or seq (ACGTN string),
    next_ref / next_pos / tlen (mate), tags {"SA": "chr,pos,strand,CIGAR,mapq,NM;", "NM": 3} (str -> Z, int -> i)."""
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
        seq = rd.get("seq")
        l_seq = len(seq) if seq is not None else rd.get("seq_len", sum(n for op, n in cigar if op in (0, 1, 4, 7, 8)))
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
                           l_seq, rd.get("next_ref", -1), rd.get("next_pos", -1), rd.get("tlen", 0))
        body += nm + b"".join(struct.pack("<I", (n << 4) | op) for op, n in cigar)
        if seq is None:
            body += b"\xff" * ((l_seq + 1) // 2)
        else:
            codes = [_SEQ_CODE.index(ch) for ch in seq.upper()] + [0]
            body += bytes((codes[i] << 4) | codes[i + 1] for i in range(0, l_seq, 2))
        body += b"\xff" * l_seq
        for tag, val in (rd.get("tags") or {}).items():
            if isinstance(val, str):
                body += tag.encode("ascii") + b"Z" + val.encode("ascii") + b"\0"
            else:
                body += tag.encode("ascii") + b"i" + struct.pack("<i", int(val))
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


def write_bam_columns(path, contigs, ref_id, pos, flag, mapq, read_len=150, seed=0, level=1):
    """Vectorised writer for benchmarks: one fixed-size record per read (a `read_len`M CIGAR, random bases,
    constant qualities, 10-character names), coordinate order as given.  Millions of reads per second of
    numpy work + zlib; the result is an ordinary BGZF BAM."""
    n = len(pos)
    rng = np.random.default_rng(seed)
    l_name, seq_b = 11, (read_len + 1) // 2
    size = 32 + l_name + 4 + seq_b + read_len
    rec = np.zeros((n, 4 + size), dtype=np.uint8)

    def put(col, values, dtype):
        rec[:, col:col + np.dtype(dtype).itemsize] = np.ascontiguousarray(values, dtype=dtype).reshape(n, 1).view(np.uint8)

    pos = np.asarray(pos, dtype=np.int64)
    put(0, np.full(n, size), "<i4")
    put(4, ref_id, "<i4")
    put(8, pos, "<i4")
    rec[:, 12] = l_name
    rec[:, 13] = np.asarray(mapq, dtype=np.uint8)
    put(14, np.full(n, 4680), "<u2")               # bin: recomputed by readers that care; htslib ignores it on read
    put(16, np.full(n, 1), "<u2")
    put(18, flag, "<u2")
    put(20, np.full(n, read_len), "<i4")
    put(24, np.full(n, -1), "<i4")
    put(28, np.full(n, -1), "<i4")
    put(32, np.zeros(n), "<i4")
    digits = (np.arange(n, dtype=np.int64)[:, None] // 10 ** np.arange(8, -1, -1)) % 10
    rec[:, 36] = ord("r")
    rec[:, 37:46] = (digits + ord("0")).astype(np.uint8)
    put(36 + l_name, np.full(n, (read_len << 4) | 0), "<u4")
    at = 36 + l_name + 4
    codes = np.array([1, 2, 4, 8], dtype=np.uint8)[rng.integers(0, 4, (n, 2 * seq_b), dtype=np.uint8)]
    rec[:, at:at + seq_b] = (codes[:, 0::2] << 4) | codes[:, 1::2]
    rec[:, at + seq_b:at + seq_b + read_len] = 30
    text = ("@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in contigs)).encode("ascii")
    head = bytearray(b"BAM\x01") + struct.pack("<i", len(text)) + text + struct.pack("<i", len(contigs))
    for name, length in contigs:
        nm = name.encode("ascii") + b"\0"
        head += struct.pack("<i", len(nm)) + nm + struct.pack("<i", length)
    body = memoryview(rec.reshape(-1))
    with open(path, "wb") as f:
        f.write(_bgzf_block(bytes(head), level))
        for i in range(0, len(body), 0xff00):
            f.write(_bgzf_block(bytes(body[i:i + 0xff00]), level))
        f.write(_BGZF_EOF)
