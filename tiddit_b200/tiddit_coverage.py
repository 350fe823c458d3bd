"""Drop-in for tiddit/tiddit_coverage.pyx (create_coverage :10-21, print_coverage :22-45,
update_coverage :48-74), with the accumulation done on the GPU.

The reference adds one read per Python call; the GPU path wants batches.  `create_coverage` therefore returns
`CoverageArray`s -- float64 ndarrays that queue the per-read `update_coverage` calls (same signature) and apply
them with the kernel every 2^22 reads and before any access to the data, so the reference's own read loops run
unchanged without per-call device traffic.  Next to it: `update_coverage_batch` (arrays of starts / ends for one
contig) and `DeviceCoverage`, which keeps
every contig's bins resident in HBM while a BAM is streamed and hands the
float64 arrays back at the end.  Bins are bit-identical to the reference's in every case
(all partial sums are exact).
"""
import math

import numpy as np

from . import _lib, device_ops

__all__ = ["create_coverage", "update_coverage", "update_coverage_batch", "print_coverage", "DeviceCoverage",
           "CoverageArray"]


def create_coverage(bam_header, bin_size, c="all"):
    """tiddit_coverage.pyx:10-21.  c == "all" -> (dict, dict); a contig name -> (ndarray, int)."""
    coverage_data, end_bin_size = {}, {}
    for contig in bam_header["SQ"]:
        name, length = contig["SN"], contig["LN"]
        if c != "all" and name != c:
            continue
        bins = int(math.ceil(length / float(bin_size)))
        coverage_data[name] = CoverageArray(bins)
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
    if type(coverage_data) is CoverageArray:
        update_coverage_batch(ref_start, ref_end, bin_size, coverage_data._plain(), end_bin_size)   # queued reads first
        return coverage_data
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


class CoverageArray(np.lib.mixins.NDArrayOperatorsMixin):
    """What create_coverage hands out: a float64 array whose per-read updates are QUEUED and applied by the GPU kernel
    in batches.

    The reference's callers (tiddit/__main__.py:229-242, tiddit_signal.pyx:181-182) call update_coverage once per read
    -- 618 M times on a 30X BAM -- and only look at the bins when the BAM is exhausted.  update_coverage on a
    CoverageArray appends (start, end) to the array's queue (two list appends, no device work); the queue is flushed
    to tdt_coverage_accumulate every FLUSH_READS reads and before ANY access to the data.  Between flushes the bins
    stay resident in HBM; the host copy is refreshed on access.  No coverage arithmetic happens on the CPU.

    It is an array-like that WRAPS its ndarray instead of subclassing it, on purpose: numpy hands out base-class views
    of a subclass instance (np.asarray, np.array(copy=False), C extensions) without calling any Python hook, which
    would read the bins before the queue has been applied.  As a wrapper every consumer has to come through
    __array__ / __buffer__ / __getitem__ / __array_ufunc__ / __array_function__ / attribute access, and each of them
    flushes first: indexing, slicing, len, iteration, numpy functions and operators, printing, memoryview and Cython
    typed memoryviews (PEP 688), copying and pickling (joblib returns the arrays from worker processes) all see the
    finished bins -- which covers every use the reference makes of them (tiddit_coverage_analysis.pyx:14-20,
    tiddit_variant.pyx:267-309, tiddit_contig_analysis.pyx:192, print_coverage).

    A read that touches a bin outside the array raises the reference's IndexError inside update_coverage, like the
    reference (two integer comparisons on the fast path); the kernel's own bounds report is a second line of defence
    at flush time."""

    FLUSH_READS = 1 << 22
    __array_priority__ = 100.0
    _META = frozenset(("shape", "dtype", "ndim", "size", "nbytes", "itemsize"))

    def __init__(self, n_bins):
        self._host = np.zeros(int(n_bins), dtype=np.float64)
        self._qs, self._qe = [], []       # queued starts / ends (Python ints)
        self._q_bin = None                # (bin_size, end_bin_size) of the queued reads
        self._q_end = -1                  # n_bins * bin_size: reads inside [0, _q_end] cannot leave the array
        self._dev = None                  # device copy of the bins while updates are streaming
        self._host_stale = False          # the device copy is ahead of the host memory

    # ---- the queue ------------------------------------------------------------------------------
    def _queue(self, ref_start, ref_end, bin_size, end_bin_size):
        key = (bin_size, end_bin_size)
        if self._q_bin != key:
            if self._qs:
                self._flush_reads()
            self._q_bin = key
            self._q_end = len(self._host) * bin_size if bin_size > 0 else -1
        if not (0 <= ref_start and ref_end <= self._q_end):
            # off the fast path: the reference's bounds check (tiddit_coverage.pyx:48-74 indexes with wrap-around,
            # so bins -n_bins .. n_bins-1 are legal) -- raise where the reference raises, inside the call
            nb = len(self._host)
            fb, eb = ref_start // bin_size, (ref_end - 1) // bin_size
            if not (-nb <= fb < nb) or (eb != fb and not (-nb <= eb < nb)):
                raise IndexError("Out of bounds on buffer access (axis 0)")
        self._qs.append(ref_start)
        self._qe.append(ref_end)
        if len(self._qs) >= self.FLUSH_READS:
            self._flush_reads()

    def _flush_reads(self):
        """Queued reads -> the kernel; the bins stay on the device."""
        if not self._qs:
            return
        qs, qe = self._qs, self._qe
        self._qs, self._qe = [], []
        s = _as_int32_positions(qs, "ref_start")
        e = _as_int32_positions(qe, "ref_end")
        bin_size, end_bin_size = self._q_bin
        self._dev, bad = _accumulate_resident(s, e, int(bin_size), int(end_bin_size), self._host, self._dev)
        self._host_stale = True
        if bad:
            self._plain()
            raise IndexError("Out of bounds on buffer access (axis 0)")

    def _plain(self):
        """Everything queued is applied and the host ndarray holds the result; the device copy is dropped, so host-side
        writes that follow (cov[i] = x) are seen by the next batch.  -> the ndarray."""
        if self._qs:
            self._flush_reads()
        if self._host_stale:
            self._host[...] = _download(self._dev)
            self._host_stale = False
        self._dev = None
        return self._host

    def flush(self):
        """Apply everything queued now (errors of queued reads surface here)."""
        self._plain()
        return self

    # ---- every way of looking at the data goes through _plain -------------------------------------
    def __len__(self):
        return len(self._host)

    def __getitem__(self, key):
        return self._plain()[key]

    def __setitem__(self, key, value):
        self._plain()[key] = value

    def __iter__(self):
        return iter(self._plain())

    def __array__(self, dtype=None, copy=None):
        base = self._plain()
        if dtype is not None and np.dtype(dtype) != base.dtype:
            return base.astype(dtype)
        return base.copy() if copy else base

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = [x._plain() if isinstance(x, CoverageArray) else x for x in inputs]
        if out is not None:
            kwargs["out"] = tuple(x._plain() if isinstance(x, CoverageArray) else x for x in out)
        res = getattr(ufunc, method)(*args, **kwargs)
        if out is not None and len(out) == 1 and isinstance(out[0], CoverageArray):
            return out[0]                                        # cov += x keeps the queueing array
        return res

    def __array_function__(self, func, types, args, kwargs):
        def strip(x):
            if isinstance(x, CoverageArray):
                return x._plain()
            if isinstance(x, (list, tuple)):
                return type(x)(strip(v) for v in x)
            if isinstance(x, dict):
                return {k: strip(v) for k, v in x.items()}
            return x
        return func(*strip(args), **strip(kwargs))

    def __buffer__(self, flags):                                 # memoryview(), Cython `double[:]` (PEP 688)
        return self._plain().__buffer__(flags)

    def __getattr__(self, name):                                 # ndarray attributes / methods (sum, mean, tolist, ...)
        if name.startswith("_"):
            raise AttributeError(name)
        host = self.__dict__.get("_host")
        if host is None:
            raise AttributeError(name)
        return getattr(host if name in self._META else self._plain(), name)

    def __reduce__(self):                                        # pickled as the finished plain array
        return (np.array, (self._plain().copy(),))

    def __copy__(self):
        return self._plain().copy()

    def __deepcopy__(self, memo):
        return self._plain().copy()

    def __repr__(self):
        return repr(self._plain())

    def __str__(self):
        return str(self._plain())

    def __bool__(self):
        return bool(self._plain())

    def __contains__(self, x):
        return x in self._plain()


def _accumulate_resident(s, e, bin_size, end_bin_size, host_bins, dev_bins):
    """One batch of one contig's reads through tdt_coverage_accumulate; the bins live on the device between batches.
    -> (device bins, any read out of bounds).  (tests/test_cli.py swaps this for the oracle on CPU-only machines.)"""
    torch = _lib.torch_cuda()
    if dev_bins is None:
        dev_bins = torch.from_numpy(host_bins).cuda()
    first_bad = device_ops.new_first_bad(torch)
    device_ops.coverage_accumulate_device(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda(), bin_size, end_bin_size,
                                          dev_bins, first_bad)
    return dev_bins, int(first_bad.item()) != device_ops.FIRST_BAD_NONE


def _download(dev_bins):
    return dev_bins.cpu().numpy()


def update_coverage(ref_start, ref_end, bin_size, coverage_data, end_bin_size):
    """tiddit_coverage.pyx:48-74, one read (signature kept for tiddit_signal.pyx:182 / __main__.py:242).

    On the arrays create_coverage returns the read is queued (see CoverageArray) -- no per-call device traffic; a
    plain ndarray the caller allocated itself is updated at once through a one-read batch (O(n_bins) copies)."""
    if type(coverage_data) is CoverageArray:
        if bin_size == 0:
            raise ZeroDivisionError("integer division or modulo by zero")
        coverage_data._queue(ref_start, ref_end, bin_size, end_bin_size)
        return coverage_data
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
