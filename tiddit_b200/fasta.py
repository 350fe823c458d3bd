"""Minimal FASTA access for the GC path: the two pysam.FastaFile calls tiddit_gc.pyx:7-8,15 makes
(get_reference_length, fetch), on a numpy reader that loads the file once and strips line breaks.
Plain and gzip / bgzip-compressed files are read (pysam.FastaFile accepts bgzip); the whole genome is held in
host memory as one uint8 array per contig (3.1 GB for GRCh38), which is what the GC kernel's H2D copies read."""
import gzip
import os

import numpy as np

_cache = {}


def _read_bytes(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":              # gzip / bgzip (concatenated members): inflate once
        with gzip.open(path, "rb") as f:
            return np.frombuffer(f.read(), dtype=np.uint8)
    return np.fromfile(path, dtype=np.uint8)


class NumpyFasta:
    def __init__(self, path):
        raw = _read_bytes(path)
        if len(raw) and raw[0] not in (ord(">"), ord(";"), 10, 13):
            raise ValueError("%s does not look like a FASTA file (first byte %r)" % (path, bytes(raw[:1])))
        gt = np.flatnonzero(raw == ord(">"))
        if len(gt):
            keep = (gt == 0) | (raw[np.maximum(gt - 1, 0)] == 10)
            gt = gt[keep]
        self._seqs = {}
        self.references = []
        nl = np.flatnonzero(raw == 10)
        for k, h in enumerate(gt):
            pos = np.searchsorted(nl, h)
            line_end = nl[pos] if pos < len(nl) else len(raw)
            name = bytes(raw[h + 1:line_end]).decode("ascii").split()[0] if line_end > h + 1 else ""
            stop = gt[k + 1] if k + 1 < len(gt) else len(raw)
            body = raw[min(line_end + 1, stop):stop]
            seq = body[(body != 10) & (body != 13)]
            self._seqs[name] = np.ascontiguousarray(seq)
            self.references.append(name)

    def get_reference_length(self, contig):
        return int(len(self._seqs[contig]))

    def fetch_bytes(self, contig):
        return self._seqs[contig]

    def fetch(self, contig, start=None, end=None):
        return bytes(self._seqs[contig][start:end]).decode("ascii")


def open_fasta(path):
    """Cached per (path, mtime, size): replacing the file invalidates the entry; one genome is kept at a time."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    fa = _cache.get(key)
    if fa is None:
        fa = NumpyFasta(path)
        _cache.clear()
        _cache[key] = fa
    return fa
