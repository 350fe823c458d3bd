"""ctypes front end for oracle/libtdt_oracle.so (the C restatement, tdt_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Function names mirror the reference
(`/root/reference/tiddit/DBSCAN.py`, `tiddit_coverage.pyx`, `tiddit_gc.pyx`); return types
mirror the reference too (float64 label arrays, float64 bins, int8 GC bins).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtdt_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "tdt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c99", "-fno-fast-math", "-shared",
                               "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.tdt_oracle_xpass.restype = ctypes.c_int64
        L.tdt_oracle_xpass.argtypes = [i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int64, i64p]
        L.tdt_oracle_ypass.restype = ctypes.c_int64
        L.tdt_oracle_ypass.argtypes = [i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int64, ctypes.c_int64, i64p]
        L.tdt_oracle_dbscan.restype = ctypes.c_int
        L.tdt_oracle_dbscan.argtypes = [i64p, i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int64, i64p]
        L.tdt_oracle_cluster_segments.restype = ctypes.c_int
        L.tdt_oracle_cluster_segments.argtypes = [i32p, i32p, i64p, ctypes.c_int64, ctypes.c_double,
                                                  ctypes.c_int64, i32p]
        L.tdt_oracle_coverage.restype = ctypes.c_int64
        L.tdt_oracle_coverage.argtypes = [i64p, i64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.POINTER(ctypes.c_double), ctypes.c_int64]
        L.tdt_oracle_gc.restype = ctypes.c_int64
        L.tdt_oracle_gc.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int64, ctypes.c_int32,
                                    ctypes.c_double, ctypes.POINTER(ctypes.c_int8)]
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.tdt_oracle_aggregate.restype = ctypes.c_int64
        L.tdt_oracle_aggregate.argtypes = [i32p, i32p, i32p, i32p, i32p, u8p, i64p, u8p, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, i32p, i32p, i64p]
        L.tdt_oracle_masked_medians.restype = None
        L.tdt_oracle_masked_medians.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int8), i64p,
                                                ctypes.c_int64, ctypes.POINTER(ctypes.c_double), i64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _col(data, c):
    data = np.asarray(data)
    if data.ndim != 2:
        data = data.reshape(len(data), -1)
    return np.ascontiguousarray(data[:, c], dtype=np.int64)


def x_coordinate_clustering(data, epsilon, m):
    """DBSCAN.py:33-64 -> (float64 labels, cluster_id)."""
    x = _col(data, 0)
    lab = np.empty(len(x), dtype=np.int64)
    cid = lib().tdt_oracle_xpass(_p(x, ctypes.c_int64), len(x), float(epsilon), int(m), _p(lab, ctypes.c_int64))
    if cid == -2:
        raise ValueError("max() arg is an empty sequence")
    return lab.astype(np.float64), int(cid)


def y_coordinate_clustering(data, epsilon, m, cluster_id, clusters):
    """DBSCAN.py:66-123 -> (float64 labels, cluster_id); `clusters` is updated in place like the reference."""
    y = _col(data, 1)
    lab = np.ascontiguousarray(np.asarray(clusters), dtype=np.int64)
    cid = lib().tdt_oracle_ypass(_p(y, ctypes.c_int64), len(y), float(epsilon), int(m), int(cluster_id),
                                 _p(lab, ctypes.c_int64))
    if cid == -2:
        raise ValueError("max() arg is an empty sequence")
    clusters[...] = lab
    return clusters, int(cid)


def main(data, epsilon, m):
    """DBSCAN.py:125-129 -> float64 labels."""
    x = _col(data, 0)
    y = _col(data, 1)
    lab = np.empty(len(x), dtype=np.int64)
    rc = lib().tdt_oracle_dbscan(_p(x, ctypes.c_int64), _p(y, ctypes.c_int64), len(x), float(epsilon), int(m),
                                 _p(lab, ctypes.c_int64))
    if rc == -2:
        raise ValueError("max() arg is an empty sequence")
    return lab.astype(np.float64)


def cluster_segments(posA, posB, seg_off, epsilon, m):
    """tiddit_cluster.pyx:140-160 over all (chrA,chrB) segments -> int32 labels in input order."""
    posA = np.ascontiguousarray(posA, dtype=np.int32)
    posB = np.ascontiguousarray(posB, dtype=np.int32)
    seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
    out = np.full(len(posA), -1, dtype=np.int32)
    rc = lib().tdt_oracle_cluster_segments(_p(posA, ctypes.c_int32), _p(posB, ctypes.c_int32),
                                           _p(seg_off, ctypes.c_int64), len(seg_off) - 1, float(epsilon), int(m),
                                           _p(out, ctypes.c_int32))
    if rc == -2:
        raise ValueError("max() arg is an empty sequence")
    return out


def cluster_aggregate(labels, posA, posB, span, name_id, flags, seg_off, same_chrom, max_ins_len, is_mp, min_reads,
                      max_pos=0, n_names=0):
    """tiddit_cluster.pyx:156-336 on packed arrays -> (rows int32 [C,16] in dict insertion order, member_idx [M]);
    same contract as tiddit_b200.device_ops.cluster_aggregate."""
    c = lambda a, t: np.ascontiguousarray(a, dtype=t)
    labels, posA, posB, name_id = c(labels, np.int32), c(posA, np.int32), c(posB, np.int32), c(name_id, np.int32)
    span = c(span, np.int32).reshape(-1, 4)
    flags, same_chrom, seg_off = c(flags, np.uint8), c(same_chrom, np.uint8), c(seg_off, np.int64)
    n = len(labels)
    rows = np.zeros((max(n, 1), 16), dtype=np.int32)
    mem = np.zeros(max(n, 1), dtype=np.int32)
    kept = np.zeros(1, dtype=np.int64)
    C = lib().tdt_oracle_aggregate(_p(labels, ctypes.c_int32), _p(posA, ctypes.c_int32), _p(posB, ctypes.c_int32),
                                   _p(span, ctypes.c_int32), _p(name_id, ctypes.c_int32), _p(flags, ctypes.c_uint8),
                                   _p(seg_off, ctypes.c_int64), _p(same_chrom, ctypes.c_uint8), len(seg_off) - 1,
                                   int(max_ins_len), int(bool(is_mp)), int(min_reads), _p(rows, ctypes.c_int32),
                                   _p(mem, ctypes.c_int32), _p(kept, ctypes.c_int64))
    return rows[:C].copy(), mem[:int(kept[0])].copy()


def create_coverage(bam_header, bin_size, c="all"):
    """tiddit_coverage.pyx:10-21."""
    coverage_data = {}
    end_bin_size = {}
    for contig in bam_header["SQ"]:
        if c == "all" or contig["SN"] == c:
            bins = int(math.ceil(contig["LN"] / float(bin_size)))
            coverage_data[contig["SN"]] = np.zeros(bins)
            end_bin_size[contig["SN"]] = contig["LN"] - (bins - 1) * bin_size
            if c != "all":
                return coverage_data[contig["SN"]], end_bin_size[contig["SN"]]
    return coverage_data, end_bin_size


def update_coverage_batch(ref_start, ref_end, bin_size, coverage_data, end_bin_size):
    """tiddit_coverage.pyx:48-74 applied to every (start, end) in order; IndexError like the reference."""
    s = np.ascontiguousarray(ref_start, dtype=np.int64)
    e = np.ascontiguousarray(ref_end, dtype=np.int64)
    assert coverage_data.dtype == np.float64 and coverage_data.flags.c_contiguous
    if bin_size == 0:
        raise ZeroDivisionError("integer division or modulo by zero")
    bad = lib().tdt_oracle_coverage(_p(s, ctypes.c_int64), _p(e, ctypes.c_int64), len(s), int(bin_size),
                                    int(end_bin_size), _p(coverage_data, ctypes.c_double), len(coverage_data))
    if bad >= 0:
        raise IndexError("Out of bounds on buffer access (axis 0)")
    return coverage_data


def update_coverage(ref_start, ref_end, bin_size, coverage_data, end_bin_size):
    return update_coverage_batch([ref_start], [ref_end], bin_size, coverage_data, end_bin_size)


def gc_bins(seq, bin_size, n_cutoff):
    """tiddit_gc.pyx:6-33 on an in-memory sequence (bytes / str / uint8 array) -> int8 bins."""
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    buf = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
    nbins = int(math.ceil(len(buf) / bin_size))
    out = np.zeros(nbins, dtype=np.int8)
    if nbins:
        lib().tdt_oracle_gc(_p(buf, ctypes.c_uint8), len(buf), int(bin_size), float(n_cutoff), _p(out, ctypes.c_int8))
    return out


def coverage_medians(bins, gc, bin_off):
    """tiddit_coverage_analysis.pyx:14-27 -> (medians float64 [C+1], counts int64 [C+1]); same contract as
    tiddit_b200.device_ops.coverage_medians."""
    bins = np.ascontiguousarray(bins, dtype=np.float64)
    gc = np.ascontiguousarray(gc, dtype=np.int8)
    bin_off = np.ascontiguousarray(bin_off, dtype=np.int64)
    C = len(bin_off) - 1
    med = np.zeros(C + 1, dtype=np.float64)
    cnt = np.zeros(C + 1, dtype=np.int64)
    lib().tdt_oracle_masked_medians(_p(bins, ctypes.c_double), _p(gc, ctypes.c_int8), _p(bin_off, ctypes.c_int64), C,
                                    _p(med, ctypes.c_double), _p(cnt, ctypes.c_int64))
    return med, cnt
