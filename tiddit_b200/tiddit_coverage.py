"""Drop-in for tiddit/tiddit_coverage.pyx (create_coverage :10-21, print_coverage :22-45,
update_coverage :48-74), with the accumulation done on the GPU.

The reference adds one read per Python call; the GPU path wants batches, so next to the per-read
`update_coverage` (same signature, same result, one tiny launch per call) this module offers
`update_coverage_batch` (arrays of starts / ends for one contig) and `DeviceCoverage`, which keeps
every contig's bins resident in HBM while a BAM is streamed and hands the
float64 arrays back at the end.  Bins are bit-identical to the reference's in every case
(all partial sums are exact).
"""
import math

import numpy as np

from . import _lib, device_ops

__all__ = ["create_coverage", "update_coverage", "update_coverage_batch", "print_coverage", "DeviceCoverage"]


def create_coverage(bam_header, bin_size, c="all"):
    """tiddit_coverage.pyx:10-21.  c == "all" -> (dict, dict); a contig name -> (ndarray, int)."""
    coverage_data, end_bin_size = {}, {}
    for contig in bam_header["SQ"]:
        name, length = contig["SN"], contig["LN"]
        if c != "all" and name != c:
            continue
        bins = int(math.ceil(length / float(bin_size)))
        coverage_data[name] = np.zeros(bins)
        end_bin_size[name] = length - (bins - 1) * bin_size
        if c != "all":
            return coverage_data[name], end_bin_size[name]
    return coverage_data, end_bin_size


def _as_int32_positions(values, what):
    arr = np.asarray(values)
    if arr.dtype.kind not in "iu":
        arr = arr.astype(np.int64)
    if arr.size and (arr.min() < -2 ** 31 or arr.max() > 2 ** 31 - 1):
        # a C long beyond 2^31 lands far outside any bin array: the reference's bounds check fires
        raise IndexError("Out of bounds on buffer access (axis 0)")
    return np.ascontiguousarray(arr, dtype=np.int32)


def update_coverage_batch(ref_start, ref_end, bin_size, coverage_data, end_bin_size):
    """tiddit_coverage.pyx:48-74 for many reads of one contig: coverage_data (float64 ndarray) += reads.

    Raises IndexError like the reference when a read touches a bin outside the array (the state of
    coverage_data is then unspecified; the reference leaves the partial sums of the earlier reads)."""
    torch = _lib.torch_cuda()
    if not isinstance(coverage_data, np.ndarray) or coverage_data.dtype != np.float64 or coverage_data.ndim != 1:
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got something else")
    if int(bin_size) == 0:
        raise ZeroDivisionError("integer division or modulo by zero")
    if int(bin_size) < 0:
        raise ValueError("bin_size must be positive")
    s = _as_int32_positions(ref_start, "ref_start")
    e = _as_int32_positions(ref_end, "ref_end")
    if s.shape != e.shape:
        raise ValueError("ref_start and ref_end differ in length")
    if s.size == 0:
        return coverage_data
    bins = torch.from_numpy(np.ascontiguousarray(coverage_data)).cuda()
    first_bad = device_ops.new_first_bad(torch)
    device_ops.coverage_accumulate_device(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda(), int(bin_size),
                                          int(end_bin_size), bins, first_bad)
    bad = int(first_bad.item())
    coverage_data[...] = bins.cpu().numpy()
    if bad != device_ops.FIRST_BAD_NONE:
        raise IndexError("Out of bounds on buffer access (axis 0)")
    return coverage_data


def update_coverage(ref_start, ref_end, bin_size, coverage_data, end_bin_size):
    """tiddit_coverage.pyx:48-74, one read (signature kept for tiddit_signal.pyx:182 / __main__.py:242)."""
    return update_coverage_batch([int(ref_start)], [int(ref_end)], bin_size, coverage_data, end_bin_size)


class DeviceCoverage:
    """All contigs' bins resident in HBM; reads are appended per contig and flushed in batches.

    Stands in for the dict pair create_coverage(header, bin_size) returns plus the per-read
    update_coverage calls of __main__.py:227-242 / tiddit_signal.pyx:156-182."""

    def __init__(self, bam_header, bin_size, flush_reads=1 << 22):
        torch = _lib.torch_cuda()
        if int(bin_size) <= 0:
            raise ZeroDivisionError("integer division or modulo by zero")
        self.bin_size = int(bin_size)
        self.names = [c["SN"] for c in bam_header["SQ"]]
        self.lengths = [int(c["LN"]) for c in bam_header["SQ"]]
        nb = [int(math.ceil(ln / float(bin_size))) for ln in self.lengths]
        self.index = {n: i for i, n in enumerate(self.names)}
        self.bin_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        self.end_bin_size = np.array([ln - (b - 1) * self.bin_size for ln, b in zip(self.lengths, nb)], dtype=np.int32)
        self.bins = torch.zeros(int(self.bin_off[-1]), dtype=torch.float64, device="cuda")
        self._bin_off_d = torch.from_numpy(self.bin_off).cuda()
        self._ebs_d = torch.from_numpy(self.end_bin_size).cuda()
        self.first_bad = device_ops.new_first_bad(torch)
        self.flush_reads = int(flush_reads)
        self._pending = [[] for _ in self.names]   # per contig: list of (starts, ends) int32 arrays
        self._n_pending = 0

    def add_reads(self, contig, ref_start, ref_end):
        s = _as_int32_positions(np.atleast_1d(ref_start), "ref_start")
        e = _as_int32_positions(np.atleast_1d(ref_end), "ref_end")
        self._pending[self.index[contig]].append((s, e))
        self._n_pending += len(s)
        if self._n_pending >= self.flush_reads:
            self.flush()

    def flush(self):
        if self._n_pending == 0:
            return
        torch = _lib.torch_cuda()
        starts, ends, counts = [], [], []
        for chunks in self._pending:
            counts.append(sum(len(s) for s, _ in chunks))
            starts.extend(s for s, _ in chunks)
            ends.extend(e for _, e in chunks)
        read_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        s = torch.from_numpy(np.concatenate(starts)).cuda()
        e = torch.from_numpy(np.concatenate(ends)).cuda()
        device_ops.coverage_accumulate_contigs_device(s, e, torch.from_numpy(read_off).cuda(), self._bin_off_d,
                                                      self._ebs_d, self.bin_size, self.bins, self.first_bad)
        self._pending = [[] for _ in self.names]
        self._n_pending = 0

    def to_host(self):
        """-> ({contig: float64 ndarray}, {contig: end_bin_size}) like create_coverage + all updates."""
        self.flush()
        if int(self.first_bad.item()) != device_ops.FIRST_BAD_NONE:
            raise IndexError("Out of bounds on buffer access (axis 0)")
        host = self.bins.cpu().numpy()
        cov = {n: host[self.bin_off[i]:self.bin_off[i + 1]].copy() for i, n in enumerate(self.names)}
        ebs = {n: int(self.end_bin_size[i]) for i, n in enumerate(self.names)}
        return cov, ebs


def print_coverage(coverage_data, bam_header, bin_size, file_type, outfile):
    """tiddit_coverage.pyx:22-45: bed rows `contig, 1+i*bin, (i+1)*bin+1 (last: LN), value` or wig values,
    byte-identical to the reference (values printed like str(numpy.float64))."""
    with open(outfile, "w", buffering=1 << 20) as out:
        if file_type == "bed":
            out.write("#chromosome\tstart\tend\tcoverage\n")
        elif file_type == "wig":
            out.write('track type=wiggle_0 name="Coverage" description="Per bin average coverage"\n')
        for contig in bam_header["SQ"]:
            name = contig["SN"]
            values = [str(v) for v in np.asarray(coverage_data[name], dtype=np.float64)]
            if file_type == "wig":
                out.write("fixedStep chrom={} start=1 step={}\n".format(name, bin_size))
                if values:
                    out.write("\n".join(values))
                    out.write("\n")
            elif file_type == "bed":
                n = len(values)
                if n == 0:
                    continue
                starts = range(1, 1 + n * bin_size, bin_size)
                ends = list(range(bin_size + 1, (n + 1) * bin_size + 1, bin_size))
                ends[-1] = contig["LN"]
                out.write("".join("%s\t%d\t%s\t%s\n" % (name, s, e, v) for s, e, v in zip(starts, ends, values)))
