"""Thin, typed wrappers over the C ABI working on torch CUDA tensors (inputs already in HBM) and
host-array front ends that add the host<->device copies.  No arithmetic happens in this file.
"""
import math

import numpy as np

from . import _lib

INT32_MAX = 2 ** 31 - 1
FIRST_BAD_NONE = 2 ** 63 - 1


def eps_to_int(epsilon):
    """Integer coordinates: d < epsilon  <=>  d < ceil(epsilon) (DBSCAN.py:51,99 compare ints to epsilon)."""
    e = math.ceil(epsilon)
    return int(min(max(e, 0), INT32_MAX))


def check_min_pts(m, n):
    if int(m) != m:
        raise TypeError("'%s' object cannot be interpreted as an integer" % type(m).__name__)
    if n > 0 and m < 2:
        # DBSCAN.py:51 / :99: max() over an empty window
        raise ValueError("max() arg is an empty sequence")


# ---------------------------------------------------------------------------------------------
# clustering
# ---------------------------------------------------------------------------------------------
def cluster_workspace_bytes(n, P):
    """Bytes of scratch tdt_cluster_labels needs for n signals in P pairs (callers that own their workspace)."""
    return int(_lib.lib().tdt_cluster_workspace_bytes(int(n), int(P)))


def cluster_labels_device(posA, posB, seg_off, P, epsilon, m, max_pos=0, labels_out=None, status=None, ws=None):
    """All (chrA,chrB) segments in one call; int32 CUDA tensors in, int32 CUDA labels (insertion order) out.

    Replaces tiddit_cluster.pyx:140-160 + DBSCAN.py:125-129.  seg_off: int64 CUDA tensor of P+1 offsets.
    status: optional one-element int32 CUDA tensor (zeroed by the caller); when given, nothing synchronises and a
    data error shows up there as a positive code (check it with `check_async_status`).
    ws: optional uint8 CUDA tensor the caller owns (>= cluster_workspace_bytes(n, P)); callers that capture the call
    into a CUDA graph MUST pass their own, because the process-wide cached workspace may be reallocated later."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    n = int(posA.numel())
    check_min_pts(m, n)
    if labels_out is None:
        labels_out = torch.empty(n, dtype=torch.int32, device=posA.device)
    if n == 0:
        return labels_out
    need = L.tdt_cluster_workspace_bytes(n, int(P))
    if ws is None:
        ws = _lib.workspace(torch, need)
    elif ws.numel() < need:
        raise ValueError("workspace of %d bytes given, %d needed" % (ws.numel(), need))
    st = _lib.stream_ptr(torch)
    if status is not None:
        rc = L.tdt_cluster_labels_async(_lib.ptr(posA), _lib.ptr(posB), _lib.ptr(seg_off), n, int(P),
                                        eps_to_int(epsilon), int(m), int(max_pos), _lib.ptr(labels_out), _lib.ptr(ws),
                                        ws.numel(), _lib.ptr(status), st)
    else:
        rc = L.tdt_cluster_labels(_lib.ptr(posA), _lib.ptr(posB), _lib.ptr(seg_off), n, int(P), eps_to_int(epsilon),
                                  int(m), int(max_pos), _lib.ptr(labels_out), _lib.ptr(ws), ws.numel(), st)
    _lib.check(rc)
    return labels_out


def check_async_status(code):
    """Raise what the synchronous call would have returned for a status word of tdt_cluster_labels_async."""
    if code:
        what = {1: "a posA / x coordinate is negative or above max_pos",
                2: "a posB / y coordinate is negative or above max_pos"}.get(int(code), "a pair / cluster id is outside its range")
        raise _lib.TdtError(_lib.TDT_E_RANGE, what)


def cluster_labels(posA, posB, seg_off, epsilon, m, max_pos=0):
    """Host front end: numpy int32 posA/posB + int64 seg_off -> numpy int32 labels (H2D, kernels, D2H)."""
    torch = _lib.torch_cuda()
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
    a = _lib.to_device(torch, posA, np.int32)
    b = _lib.to_device(torch, posB, np.int32)
    s = _lib.to_device(torch, seg_off, np.int64)
    out = cluster_labels_device(a, b, s, len(seg_off) - 1, epsilon, m, max_pos)
    return out.cpu().numpy()


def dbscan_main_device(x, y, epsilon, m, max_pos=0):
    """DBSCAN.py:125-129 on one array in the caller's order (no sort); int32 CUDA tensors."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    n = int(x.numel())
    check_min_pts(m, n)
    labels = torch.empty(n, dtype=torch.int32, device=x.device)
    if n == 0:
        return labels
    ws = _lib.workspace(torch, L.tdt_cluster_workspace_bytes(n, 1))
    rc = L.tdt_dbscan_main(_lib.ptr(x), _lib.ptr(y), n, eps_to_int(epsilon), int(m), int(max_pos), _lib.ptr(labels),
                           _lib.ptr(ws), ws.numel(), _lib.stream_ptr(torch))
    _lib.check(rc)
    return labels


def xpass_device(x, epsilon, m):
    """DBSCAN.py:33-64 -> (int32 CUDA labels, int32 CUDA scalar last id)."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    n = int(x.numel())
    check_min_pts(m, n)
    labels = torch.empty(n, dtype=torch.int32, device=x.device)
    last = torch.full((1,), -1, dtype=torch.int32, device=x.device)
    if n == 0:
        return labels, last
    ws = _lib.workspace(torch, L.tdt_cluster_workspace_bytes(n, 1))
    rc = L.tdt_xpass_labels(_lib.ptr(x), n, eps_to_int(epsilon), int(m), _lib.ptr(labels), _lib.ptr(last),
                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(torch))
    _lib.check(rc)
    return labels, last


def ypass_device(y, epsilon, m, labels_io, cluster_id_io, max_pos=0):
    """DBSCAN.py:66-123 in place on int32 CUDA labels / a one-element int32 CUDA cluster id."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    n = int(y.numel())
    check_min_pts(m, n)
    if n == 0:
        return labels_io, cluster_id_io
    ws = _lib.workspace(torch, L.tdt_cluster_workspace_bytes(n, 1))
    rc = L.tdt_ypass_labels(_lib.ptr(y), n, eps_to_int(epsilon), int(m), int(max_pos), _lib.ptr(labels_io),
                            _lib.ptr(cluster_id_io), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(torch))
    _lib.check(rc)
    return labels_io, cluster_id_io


def cluster_aggregate_device(labels, posA, posB, span, name_id, flags, seg_off, same_chrom, P, max_ins_len, is_mp,
                             min_reads, max_pos=0, n_names=0, cand_out=None, member_out=None, counts_out=None):
    """tiddit_cluster.pyx:156-336 for all pairs: CUDA tensors in -> (rows int32 [n,16], member_idx int32 [n],
    counts int64 [4] = {candidates, kept signals, data error, 0}), all on the device; nothing synchronises."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    n = int(labels.numel())
    dev = labels.device
    if cand_out is None:
        cand_out = torch.empty((max(n, 1), 16), dtype=torch.int32, device=dev)
    if member_out is None:
        member_out = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    if counts_out is None:
        counts_out = torch.zeros(4, dtype=torch.int64, device=dev)
    need = L.tdt_aggregate_workspace_bytes(n, int(P))
    ws = _lib.workspace(torch, need, tag="aggregate") if n else None
    rc = L.tdt_cluster_aggregate(_lib.ptr(labels), _lib.ptr(posA), _lib.ptr(posB), _lib.ptr(span), _lib.ptr(name_id),
                                 _lib.ptr(flags), _lib.ptr(seg_off), _lib.ptr(same_chrom), n, int(P), int(max_ins_len),
                                 int(bool(is_mp)), int(min_reads), int(max_pos), int(n_names), _lib.ptr(cand_out),
                                 _lib.ptr(member_out), _lib.ptr(counts_out), _lib.ptr(ws), ws.numel() if n else 0,
                                 _lib.stream_ptr(torch))
    _lib.check(rc)
    return cand_out, member_out, counts_out


def check_aggregate_status(code):
    if code:
        what = {1: "a sort key is out of range", 8: "a label is outside [-1, len(pair)) or a kind is not D/S/A",
                9: "a position is negative or above max_pos", 10: "a name id is negative or above n_names"}
        raise _lib.TdtError(_lib.TDT_E_RANGE, what.get(int(code), "data error %d" % int(code)))


def cluster_aggregate(labels, posA, posB, span, name_id, flags, seg_off, same_chrom, max_ins_len, is_mp, min_reads,
                      max_pos=0, n_names=0):
    """Host front end: numpy arrays in -> (rows int32 [C,16] in dict insertion order, member_idx int32 [M])."""
    torch = _lib.torch_cuda()
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
    P = len(seg_off) - 1
    d = lambda a, t: _lib.to_device(torch, a, t)
    span = np.ascontiguousarray(span, dtype=np.int32).reshape(-1, 4)
    rows, mem, counts = cluster_aggregate_device(d(labels, np.int32), d(posA, np.int32), d(posB, np.int32), d(span, np.int32),
                                                 d(name_id, np.int32), d(flags, np.uint8), d(seg_off, np.int64),
                                                 d(same_chrom, np.uint8), P, max_ins_len, is_mp, min_reads, max_pos,
                                                 n_names)
    C, M, err, _ = (int(v) for v in counts.cpu().tolist())
    check_aggregate_status(err)
    return rows[:C].cpu().numpy(), mem[:M].cpu().numpy()


def cluster_and_aggregate(posA, posB, seg_off, span, name_id, flags, same_chrom, epsilon, m, max_ins_len, is_mp, min_reads,
                          max_pos=0, n_names=0):
    """Host front end of tiddit_cluster.pyx:140-336 on packed arrays: ONE upload of every column, tdt_cluster_labels,
    tdt_cluster_aggregate on the labels where they are (they never leave HBM between the two calls), then the labels,
    the candidate rows and the member index come back once.  -> (labels int32 [n], rows int32 [C,16], member_idx int32 [M])."""
    torch = _lib.torch_cuda()
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
    P = len(seg_off) - 1
    d = lambda a, t: _lib.to_device(torch, a, t)
    A, B, O = d(posA, np.int32), d(posB, np.int32), d(seg_off, np.int64)
    span = np.ascontiguousarray(span, dtype=np.int32).reshape(-1, 4)
    labels = cluster_labels_device(A, B, O, P, epsilon, m, max_pos)
    rows, mem, counts = cluster_aggregate_device(labels, A, B, d(span, np.int32), d(name_id, np.int32), d(flags, np.uint8), O,
                                                 d(same_chrom, np.uint8), P, max_ins_len, is_mp, min_reads, max_pos, n_names)
    C, M, err, _ = (int(v) for v in counts.cpu().tolist())
    check_aggregate_status(err)
    return labels.cpu().numpy(), rows[:C].cpu().numpy(), mem[:M].cpu().numpy()


def segsort_device(keys, vals, off, key_bits, segid=None):
    """Test hook: every segment [off[s], off[s+1]) of (keys uint32-as-int32, vals int32 or None) sorted by key,
    stable -> (keys_out, vals_out) CUDA tensors."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    n, nseg = int(keys.numel()), int(off.numel()) - 1
    ko = torch.empty_like(keys)
    vo = torch.empty(n, dtype=torch.int32, device=keys.device)
    if n == 0 or nseg <= 0:
        return ko, vo
    ws = _lib.workspace(torch, L.tdt_cluster_workspace_bytes(n, nseg))
    rc = L.tdt_debug_segsort(_lib.ptr(keys), _lib.ptr(vals), _lib.ptr(off), _lib.ptr(segid), nseg, n, int(key_bits), _lib.ptr(ko),
                             _lib.ptr(vo), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(torch))
    _lib.check(rc)
    return ko, vo


# ---------------------------------------------------------------------------------------------
# coverage
# ---------------------------------------------------------------------------------------------
def new_first_bad(torch):
    return torch.full((1,), FIRST_BAD_NONE, dtype=torch.int64, device="cuda")


def coverage_accumulate_device(start, end, bin_size, end_bin_size, bins, first_bad):
    """tiddit_coverage.pyx:48-74 for a batch of reads of ONE contig; bins (float64 CUDA) updated in place."""
    torch = _lib.torch_cuda()
    if bin_size == 0:
        raise ZeroDivisionError("integer division or modulo by zero")
    rc = _lib.lib().tdt_coverage_accumulate(_lib.ptr(start), _lib.ptr(end), int(start.numel()), int(bin_size),
                                            int(end_bin_size), _lib.ptr(bins), int(bins.numel()), _lib.ptr(first_bad),
                                            _lib.stream_ptr(torch))
    _lib.check(rc)
    return bins


def coverage_accumulate_contigs_device(start, end, read_off, bin_off, end_bin_size, bin_size, bins, first_bad):
    """The same for reads grouped by contig (read_off / bin_off int64 CUDA, end_bin_size int32 CUDA)."""
    torch = _lib.torch_cuda()
    if bin_size == 0:
        raise ZeroDivisionError("integer division or modulo by zero")
    rc = _lib.lib().tdt_coverage_accumulate_contigs(
        _lib.ptr(start), _lib.ptr(end), int(start.numel()), _lib.ptr(read_off), _lib.ptr(bin_off),
        _lib.ptr(end_bin_size), int(end_bin_size.numel()), int(bin_size), _lib.ptr(bins), int(bins.numel()),
        _lib.ptr(first_bad), _lib.stream_ptr(torch))
    _lib.check(rc)
    return bins


def coverage_medians_device(bins, gc, bin_off, C, medians_out=None, counts_out=None):
    """tiddit_coverage_analysis.pyx:14-27 on CUDA tensors (float64 bins, int8 gc, int64 bin_off[C+1]) ->
    (float64 [C+1] medians: contigs then genome-wide, int64 [C+1] qualifying bins); nothing synchronises."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    if medians_out is None:
        medians_out = torch.empty(C + 1, dtype=torch.float64, device="cuda")
    if counts_out is None:
        counts_out = torch.empty(C + 1, dtype=torch.int64, device="cuda")
    ws = _lib.workspace(torch, L.tdt_coverage_medians_workspace_bytes(int(C)), tag="medians")
    rc = L.tdt_coverage_medians(_lib.ptr(bins), _lib.ptr(gc), _lib.ptr(bin_off), int(C), int(bins.numel()),
                                _lib.ptr(medians_out), _lib.ptr(counts_out), _lib.ptr(ws), ws.numel(),
                                _lib.stream_ptr(torch))
    _lib.check(rc)
    return medians_out, counts_out


def coverage_medians(bins, gc, bin_off):
    """Host front end: numpy float64 bins / int8 gc of all contigs back to back + int64 bin_off -> numpy (medians, counts)."""
    torch = _lib.torch_cuda()
    bin_off = np.ascontiguousarray(bin_off, dtype=np.int64)
    med, cnt = coverage_medians_device(_lib.to_device(torch, bins, np.float64), _lib.to_device(torch, gc, np.int8),
                                       _lib.to_device(torch, bin_off, np.int64), len(bin_off) - 1)
    return med.cpu().numpy(), cnt.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# GC
# ---------------------------------------------------------------------------------------------
def padded_sequence_device(seq_bytes):
    """uint8 host sequence -> CUDA tensor padded to a multiple of 16 bytes (the kernel's bulk copies read whole
    16-byte granules); returns (tensor, true length)."""
    torch = _lib.torch_cuda()
    arr = np.frombuffer(seq_bytes, dtype=np.uint8) if isinstance(seq_bytes, (bytes, bytearray, memoryview)) \
        else np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    n = len(arr)
    dev = torch.zeros((n + 15) // 16 * 16 + 16, dtype=torch.uint8, device="cuda")
    if n:
        dev[:n].copy_(torch.from_numpy(arr), non_blocking=False)
    return dev, n


def gc_bins_device(seq, length, bin_size, n_cutoff, out=None):
    """tiddit_gc.pyx:6-33 on a CUDA uint8 sequence (16-byte aligned, padded) -> int8 CUDA bins."""
    torch = _lib.torch_cuda()
    if bin_size == 0:
        raise ZeroDivisionError("division by zero")
    n_bins = int(math.ceil(length / bin_size))
    if out is None:
        out = torch.zeros(n_bins, dtype=torch.int8, device="cuda")
    rc = _lib.lib().tdt_gc_bins(_lib.ptr(seq), int(length), int(bin_size), float(n_cutoff), _lib.ptr(out),
                                _lib.stream_ptr(torch))
    _lib.check(rc)
    return out
