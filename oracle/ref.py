"""Loader for the REAL reference modules compiled into oracle/_ref/ (see build_ref.py).

TEST INFRASTRUCTURE ONLY.  `load()` returns a namespace with DBSCAN, tiddit_cluster,
tiddit_coverage, tiddit_gc, tiddit_coverage_analysis, tiddit_signal -- the unmodified upstream code -- or None when oracle/_ref/ has not
been built (then callers fall back to the golden fixtures / the C restatement).
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_ns = None


def available():
    from . import build_ref
    return build_ref.have_ref()


def load():
    global _ns
    if _ns is not None:
        return _ns
    if not available():
        return None
    if _REF not in sys.path:
        sys.path.insert(0, _REF)   # provides the `tiddit` package and the `pysam` FastaFile stand-in
    ns = types.SimpleNamespace()
    ns.DBSCAN = importlib.import_module("tiddit.DBSCAN")
    ns.tiddit_coverage = importlib.import_module("tiddit.tiddit_coverage")
    ns.tiddit_cluster = importlib.import_module("tiddit.tiddit_cluster")
    ns.tiddit_gc = importlib.import_module("tiddit.tiddit_gc")
    ns.tiddit_coverage_analysis = importlib.import_module("tiddit.tiddit_coverage_analysis")
    try:    # compiled against the pysam stand-in (oracle/ref_shims/pysam); needs joblib at import time
        ns.tiddit_signal = importlib.import_module("tiddit.tiddit_signal")
    except ImportError:
        ns.tiddit_signal = None
    _ns = ns
    return ns
