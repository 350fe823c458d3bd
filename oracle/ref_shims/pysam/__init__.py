"""Stand-in `pysam` package for the compiled reference modules: `FastaFile` (two methods, GC module) here,
`AlignmentFile` / `AlignedSegment` (signal worker) in libcalignmentfile.pyx.

TEST INFRASTRUCTURE ONLY.  The real pysam (htslib bindings) is not installed in this
image; `tiddit/tiddit_gc.pyx:1,7-8,15` only needs `FastaFile(path)`,
`.get_reference_length(contig)` and `.fetch(contig, start, end) -> str`.  This file is
our own code (nothing from the reference); `oracle/build_ref.py` copies it next to the
compiled reference modules so that `import pysam` inside `tiddit_gc` resolves here.
"""
from .libcalignmentfile import AlignedSegment, AlignmentFile  # noqa: F401  (compiled by oracle/build_ref.py)


class FastaFile:
    _cache = {}

    def __init__(self, path):
        self.path = path
        seqs = FastaFile._cache.get(path)
        if seqs is None:
            seqs = {}
            name = None
            chunks = []
            with open(path) as handle:
                for line in handle:
                    if line.startswith(">"):
                        if name is not None:
                            seqs[name] = "".join(chunks)
                        name = line[1:].split()[0]
                        chunks = []
                    else:
                        chunks.append(line.strip())
            if name is not None:
                seqs[name] = "".join(chunks)
            FastaFile._cache[path] = seqs
        self._seqs = seqs

    @property
    def references(self):
        return list(self._seqs)

    def get_reference_length(self, contig):
        return len(self._seqs[contig])

    def fetch(self, contig, start=None, end=None):
        seq = self._seqs[contig]
        if start is None:
            start = 0
        if end is None:
            end = len(seq)
        return seq[start:end]

    def close(self):
        pass
