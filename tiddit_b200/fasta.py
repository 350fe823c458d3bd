"""Minimal FASTA access for the GC path: the two pysam.FastaFile calls tiddit_gc.pyx:7-8,15 makes
(get_reference_length, fetch).  Uses the real pysam when it is importable; otherwise a numpy
reader that loads the file once and strips line breaks."""
import numpy as np

_cache = {}


class NumpyFasta:
    def __init__(self, path):
        raw = np.fromfile(path, dtype=np.uint8)
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
    fa = _cache.get(path)
    if fa is None:
        fa = NumpyFasta(path)
        _cache.clear()      # one genome at a time
        _cache[path] = fa
    return fa
