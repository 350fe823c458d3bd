# cython: language_level=3
"""Stand-in for the part of pysam the reference's signal worker uses (tiddit/tiddit_signal.pyx).

TEST INFRASTRUCTURE ONLY -- our own code, nothing from the reference or from pysam.  The real pysam (htslib
bindings; the reference does not pin a version: pyproject.toml:2, setup.py:41) is not installed in this image, so
`oracle/build_ref.py` compiles this module next to the unmodified `tiddit_signal.pyx`; the `cimport` at
tiddit_signal.pyx:7 and the typed variables (`cdef AlignmentFile samfile`, `cdef AlignedSegment read`) then resolve
here.  It restates pysam's documented behaviour for exactly what the worker touches:

  AlignmentFile(path, "r", reference_filename=, index_filename=)   .header["SQ"][i]["SN"/"LN"]   .close()
  .fetch(contig, until_eof=True)  -> the records placed on `contig`, in file order (what an index fetch of the
                                     whole contig returns for a coordinate-sorted file)
  AlignedSegment(): settable reference_start / flag / cigar (tiddit_signal.pyx:13-28); per record is_unmapped,
  is_duplicate, is_supplementary, is_secondary, is_paired, is_reverse, mate_is_unmapped, mapq, reference_start,
  reference_end (None when unmapped or without CIGAR, else start + reference-consuming CIGAR lengths),
  reference_name, next_reference_name, isize, cigartuples, query_name, query_sequence, has_tag, get_tag,
  query_alignment_start (leading soft clip) and query_alignment_end (sequence length minus trailing soft clip; for a
  record without sequence: leading soft clip + M/I/=/X lengths -- pysam's getQueryEnd).

The BAM decoding below is deliberately independent of tiddit_b200/bamio.py so that the two cross-check each other.
"""
import gzip
import struct

cdef tuple _REF_OPS = (0, 2, 3, 7, 8)
cdef str _SEQ = "=ACMGRSVTWYHKDBN"


cdef class AlignedSegment:
    def __init__(self, header=None):
        self.query_name = None
        self.query_sequence = None
        self.flag = 0
        self.reference_id = -1
        self.reference_start = -1
        self.mapping_quality = 0
        self.next_reference_id = -1
        self.next_reference_start = -1
        self.template_length = 0
        self._cigar = []
        self._tags = {}
        self._refs = []

    property cigar:
        def __get__(self):
            return list(self._cigar)
        def __set__(self, value):
            self._cigar = [(int(op), int(n)) for op, n in (value or ())]

    property cigartuples:
        def __get__(self):
            return list(self._cigar) if self._cigar else None
        def __set__(self, value):
            self._cigar = [(int(op), int(n)) for op, n in (value or ())]

    property mapq:
        def __get__(self):
            return self.mapping_quality

    property isize:
        def __get__(self):
            return self.template_length

    property is_paired:
        def __get__(self):
            return (self.flag & 1) != 0

    property is_unmapped:
        def __get__(self):
            return (self.flag & 4) != 0

    property mate_is_unmapped:
        def __get__(self):
            return (self.flag & 8) != 0

    property is_reverse:
        def __get__(self):
            return (self.flag & 16) != 0

    property is_secondary:
        def __get__(self):
            return (self.flag & 256) != 0

    property is_duplicate:
        def __get__(self):
            return (self.flag & 1024) != 0

    property is_supplementary:
        def __get__(self):
            return (self.flag & 2048) != 0

    property reference_name:
        def __get__(self):
            if self.reference_id < 0 or self.reference_id >= len(self._refs):
                return None
            return self._refs[self.reference_id]

    property next_reference_name:
        def __get__(self):
            if self.next_reference_id < 0 or self.next_reference_id >= len(self._refs):
                return None
            return self._refs[self.next_reference_id]

    property reference_end:
        def __get__(self):
            if (self.flag & 4) or not self._cigar:
                return None
            cdef long e = self.reference_start
            for op, n in self._cigar:
                if op in _REF_OPS:
                    e += n
            return e

    property query_alignment_start:
        def __get__(self):
            cdef long start = 0
            for op, n in self._cigar:
                if op == 5:
                    continue
                elif op == 4:
                    start += n
                else:
                    break
            return start

    property query_alignment_end:
        def __get__(self):
            cdef long end = len(self.query_sequence) if self.query_sequence else 0
            cdef long k
            if end == 0:
                for op, n in self._cigar:
                    if op in (0, 1, 7, 8) or (op == 4 and end == 0):
                        end += n
                return end
            for k in range(len(self._cigar) - 1, 0, -1):
                op, n = self._cigar[k]
                if op == 5:
                    continue
                elif op == 4:
                    end -= n
                else:
                    break
            return end

    def has_tag(self, tag):
        return tag in self._tags

    def get_tag(self, tag):
        if tag not in self._tags:
            raise KeyError("tag '%s' not present" % tag)
        return self._tags[tag]


cdef object _decode(bytes rec, list refs):
    cdef AlignedSegment a = AlignedSegment()
    (ref_id, pos, l_name, mapq, _bin, n_cigar, flag, l_seq, next_ref, next_pos, tlen) = struct.unpack_from("<iiBBHHHiiii", rec, 0)
    a.reference_id, a.reference_start, a.mapping_quality, a.flag = ref_id, pos, mapq, flag
    a.next_reference_id, a.next_reference_start, a.template_length = next_ref, next_pos, tlen
    a._refs = refs
    p = 32
    a.query_name = rec[p:p + l_name - 1].decode("ascii")
    p += l_name
    a._cigar = [(w & 15, w >> 4) for w in struct.unpack_from("<%dI" % n_cigar, rec, p)]
    p += 4 * n_cigar
    if l_seq:
        chars = []
        for i in range(l_seq):
            b = rec[p + (i >> 1)]
            chars.append(_SEQ[(b >> 4) if (i & 1) == 0 else (b & 15)])
        a.query_sequence = "".join(chars)
    p += (l_seq + 1) // 2 + l_seq
    tags = {}
    while p + 3 <= len(rec):
        tag = rec[p:p + 2].decode("ascii")
        t = chr(rec[p + 2])
        p += 3
        if t == "A":
            val = chr(rec[p]); p += 1
        elif t in "cC":
            val = struct.unpack_from("<b" if t == "c" else "<B", rec, p)[0]; p += 1
        elif t in "sS":
            val = struct.unpack_from("<h" if t == "s" else "<H", rec, p)[0]; p += 2
        elif t in "iIf":
            val = struct.unpack_from({"i": "<i", "I": "<I", "f": "<f"}[t], rec, p)[0]; p += 4
        elif t in "ZH":
            z = rec.index(b"\0", p)
            val = rec[p:z].decode("ascii"); p = z + 1
        elif t == "B":
            sub = chr(rec[p]); cnt = struct.unpack_from("<I", rec, p + 1)[0]
            w = 1 if sub in "cC" else 2 if sub in "sS" else 4
            val = None; p += 5 + cnt * w
        else:
            break
        if tag not in tags:
            tags[tag] = val
    a._tags = tags
    return a


cdef class AlignmentFile:
    def __init__(self, filename, mode="r", reference_filename=None, index_filename=None, **kwargs):
        self.filename = filename
        self._fh = None
        fh = gzip.open(filename, "rb")
        if fh.read(4) != b"BAM\x01":
            raise ValueError("not a BAM file: %s" % filename)
        l_text = struct.unpack("<i", fh.read(4))[0]
        fh.read(l_text)
        n_ref = struct.unpack("<i", fh.read(4))[0]
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack("<i", fh.read(4))[0]
            self.references.append(fh.read(l_name)[:-1].decode("ascii"))
            self.lengths.append(struct.unpack("<i", fh.read(4))[0])
        self._fh = fh
        self.header = {"SQ": [{"SN": n, "LN": l} for n, l in zip(self.references, self.lengths)]}

    def fetch(self, contig=None, until_eof=False, **kwargs):
        want = None if contig is None else self.references.index(contig)
        out = []
        fh = self._fh
        while True:
            raw = fh.read(4)
            if len(raw) < 4:
                break
            rec = fh.read(struct.unpack("<i", raw)[0])
            if want is not None and struct.unpack_from("<i", rec, 0)[0] != want:
                continue
            out.append(_decode(rec, self.references))
        return iter(out)

    def close(self):
        if self._fh is not None:
            self._fh.close()
            self._fh = None
