"""CPU: libtdt_b200.so builds for sm_100a, loads, and exports every entry point include/tdt_b200.h declares.
No compute call is made (argument validation returns before any CUDA call)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tdt_b200.h")).read()
    return sorted(set(re.findall(r"TDT_API[^;(]*?\b(tdt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(libtdt):
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(libtdt, n), n
    from tiddit_b200 import _lib
    assert sorted(_lib.SIGNATURES) == names


def test_header_compiles_as_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "tdt_b200.h")])


def test_sass_is_sm100a_with_tma(libtdt):
    from tiddit_b200 import _lib
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out          # 1-D TMA bulk copies into shared memory
    assert "RED" in out and "F64" in out   # float64 reductions of the coverage kernel


def test_argument_validation_without_gpu(libtdt):
    from tiddit_b200 import _lib
    assert libtdt.tdt_version() >= 100
    assert libtdt.tdt_cluster_workspace_bytes(1000, 3) > 1000 * 24
    rc = libtdt.tdt_cluster_labels(None, None, None, 10, 1, 500, 1, 0, None, None, 0, None)   # min_pts < 2
    assert rc == _lib.TDT_E_ARG and b"ValueError" in libtdt.tdt_last_error()
    rc = libtdt.tdt_cluster_labels(None, None, None, 10, 1, 500, 3, 0, None, None, 0, None)   # no workspace
    assert rc == _lib.TDT_E_WORKSPACE
    assert libtdt.tdt_coverage_accumulate(None, None, 5, 0, 1, None, 1, None, None) == _lib.TDT_E_ARG
    assert libtdt.tdt_gc_bins(None, 5, 0, 0.5, None, None) == _lib.TDT_E_ARG
    assert libtdt.tdt_cluster_labels(None, None, None, 0, 0, 500, 3, 0, None, None, 0, None) == _lib.TDT_OK


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tiddit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libtdt_oracle|tdt_oracle\.c", text, re.M):
                    bad.append(f)
    assert not bad, bad


def test_no_gpu_no_fallback():
    import torch
    if torch.cuda.is_available():
        return
    import numpy as np
    import pytest
    from tiddit_b200 import DBSCAN
    with pytest.raises(RuntimeError):
        DBSCAN.main(np.array([[1, 2], [1, 2], [1, 2]]), 10, 2)
