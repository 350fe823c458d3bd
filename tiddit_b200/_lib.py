"""ctypes binding of libtdt_b200.so (include/tdt_b200.h) + device-memory plumbing (torch).

The product path fails loudly: a missing library raises ImportError-like RuntimeError at first use,
a missing CUDA device raises RuntimeError; nothing here ever computes on the CPU.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TDT_B200_LIB: another build of the same library (A/B of compile-time variants, tools/gpu_ab.sh)
LIB_PATH = os.environ.get("TDT_B200_LIB") or os.path.join(_HERE, "libtdt_b200.so")

TDT_OK, TDT_E_ARG, TDT_E_WORKSPACE, TDT_E_CUDA, TDT_E_RANGE = 0, -1, -2, -3, -4

_p = ctypes.c_void_p
_i32, _i64, _sz, _dbl = ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_double

# name -> (restype, argtypes); the same list tests/test_abi.py checks against include/tdt_b200.h
SIGNATURES = {
    "tdt_version": (ctypes.c_int, []),
    "tdt_last_error": (ctypes.c_char_p, []),
    "tdt_launch_count": (_i64, []),
    "tdt_profile_begin": (None, []),
    "tdt_profile_end": (ctypes.c_int, [ctypes.c_char_p, _sz]),
    "tdt_cluster_workspace_bytes": (_sz, [_i64, _i32]),
    "tdt_cluster_labels": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "tdt_cluster_labels_async": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _sz, _p, _p]),
    "tdt_dbscan_main": (ctypes.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "tdt_xpass_labels": (ctypes.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _sz, _p]),
    "tdt_ypass_labels": (ctypes.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "tdt_aggregate_workspace_bytes": (_sz, [_i64, _i32]),
    "tdt_cluster_aggregate": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32,
                                             _p, _p, _p, _p, _sz, _p]),
    "tdt_coverage_accumulate": (ctypes.c_int, [_p, _p, _i64, _i32, _i32, _p, _i64, _p, _p]),
    "tdt_coverage_accumulate_contigs": (ctypes.c_int, [_p, _p, _i64, _p, _p, _p, _i32, _i32, _p, _i64, _p, _p]),
    "tdt_coverage_medians_workspace_bytes": (_sz, [_i32]),
    "tdt_coverage_medians": (ctypes.c_int, [_p, _p, _p, _i32, _i64, _p, _p, _p, _sz, _p]),
    "tdt_gc_bins": (ctypes.c_int, [_p, _i64, _i32, _dbl, _p, _p]),
    "tdt_peer_buffer_bytes": (_sz, [_i64, _i32]),
    "tdt_peer_alloc": (ctypes.c_int, [_sz, ctypes.POINTER(_p), ctypes.c_char_p]),
    "tdt_peer_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_p)]),
    "tdt_peer_close": (ctypes.c_int, [_p]),
    "tdt_peer_free": (ctypes.c_int, [_p]),
    "tdt_peer_allgather": (ctypes.c_int, [ctypes.POINTER(_p), _i64, _i32, _i32, _p, _p]),
    "tdt_debug_segsort": (ctypes.c_int, [_p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _sz, _p]),
}

_lib = None


class TdtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libtdt_b200 error %d: %s" % (code, message))
        self.code = code


def lib():
    """The loaded library; raises if it has not been built (python -m tiddit_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m tiddit_b200.build` "
                               "(there is no CPU fallback for this path)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != TDT_OK:
        raise TdtError(rc, lib().tdt_last_error().decode("utf-8", "replace"))


def profile_begin():
    lib().tdt_profile_begin()


def profile_end():
    """-> list of (stage, milliseconds) in call order since profile_begin()."""
    buf = ctypes.create_string_buffer(1 << 18)
    n = lib().tdt_profile_end(buf, len(buf))
    if n < 0:
        check(n)
    out = []
    for line in buf.value.decode().splitlines():
        name, ms = line.rsplit("=", 1)
        out.append((name, float(ms)))
    return out


def launch_count():
    return int(lib().tdt_launch_count())


# ---------------------------------------------------------------------------------------------
# device plumbing: torch owns HBM allocations and streams
# ---------------------------------------------------------------------------------------------
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("tiddit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def stream_ptr(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_workspaces = {}


def workspace(torch, nbytes, tag="cluster"):
    """A cached uint8 scratch tensor of at least nbytes on the current device (grown, never shrunk)."""
    key = (torch.cuda.current_device(), tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces.pop(key, None)
        ws = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device="cuda")
        _workspaces[key] = ws
    return ws


def release_workspaces():
    _workspaces.clear()


def to_device(torch, a, dtype):
    """Host array-like -> contiguous device tensor of `dtype` (numpy dtype); tensors pass through."""
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.cuda(non_blocking=True)
        want = getattr(torch, np.dtype(dtype).name)
        if t.dtype != want:
            t = t.to(want)
        return t.contiguous()
    arr = np.ascontiguousarray(a, dtype=dtype)
    return torch.from_numpy(arr).cuda(non_blocking=False)
