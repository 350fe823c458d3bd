import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) where no CUDA device exists, so plain `pytest tests` is a usable
    CPU gate."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (oracle/tdt_oracle.c), built on demand with gcc."""
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ref():
    """The real reference compiled into oracle/_ref (None when it is not there)."""
    from oracle import ref as r
    return r.load()


@pytest.fixture(scope="session")
def libtdt():
    """libtdt_b200.so, built on demand with nvcc (cross-compiles without a GPU)."""
    from tiddit_b200 import build, _lib
    build.build()
    return _lib.lib()


def unjson(o):
    """Inverse of make_golden._jsonable: {'__set__': [...]} -> set; int-like dict keys stay strings."""
    if isinstance(o, dict):
        if set(o) == {"__set__"}:
            return set(unjson(v) for v in o["__set__"])
        return {k: unjson(v) for k, v in o.items()}
    if isinstance(o, list):
        return [unjson(v) for v in o]
    return o


def data3(x, y):
    n = len(x)
    return np.stack([np.asarray(x, dtype=np.int64), np.asarray(y, dtype=np.int64), np.arange(n)], 1).reshape(n, 3)
