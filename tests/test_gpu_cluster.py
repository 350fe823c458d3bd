"""GPU parity: the CUDA clustering path (through the C ABI) against the oracle and the golden vectors
of the real reference.  Bit-exact: the reference's own cluster ids, -1 for noise."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, data3, load_json, unjson

pytestmark = pytest.mark.gpu


def assert_same(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if not np.array_equal(got, want):
        bad = np.flatnonzero(got != want)
        i = int(bad[0])
        lo, hi = max(0, i - 4), i + 5
        raise AssertionError("%s: %d of %d labels differ, first at %d: got %s want %s" %
                             (what, len(bad), len(got), i, got[lo:hi].tolist(), want[lo:hi].tolist()))


def test_golden_small_dbscan_module():
    from tiddit_b200 import DBSCAN
    for k, c in enumerate(load_json("dbscan_small.json")):
        d = data3(c["x"], c["y"])
        what = "case %d n=%d eps=%s m=%d" % (k, len(c["x"]), c["eps"], c["m"])
        if len(c["x"]) == 0:
            assert len(DBSCAN.main(d, c["eps"], c["m"])) == 0
            continue
        xl, cid = DBSCAN.x_coordinate_clustering(d, c["eps"], c["m"])
        assert xl.dtype == np.float64
        assert_same(xl.astype(int), c["x_labels"], what + " x-pass")
        assert cid == c["x_last_id"], what
        lab = DBSCAN.main(d, c["eps"], c["m"])
        assert lab.dtype == np.float64
        assert_same(lab.astype(int), c["labels"], what + " main")
        if k % 4 == 0:
            cl = np.array(c["x_labels"], dtype=np.float64)
            out, cid2 = DBSCAN.y_coordinate_clustering(d, c["eps"], c["m"], c["x_last_id"], cl)
            assert out is cl
            assert_same(cl.astype(int), c["labels"], what + " y-pass")
            assert cid2 == max([c["x_last_id"]] + c["labels"]), what


def test_golden_medium_segments():
    from tiddit_b200 import device_ops
    z = np.load(GOLDEN + "/dbscan_medium.npz")
    for k in range(5):
        a, b, want = z["posA_%d" % k], z["posB_%d" % k], z["labels_%d" % k]
        eps, m = [int(v) for v in z["param_%d" % k]]
        got = device_ops.cluster_labels(a, b, [0, len(a)], eps, m, int(max(a.max(), b.max())))
        assert_same(got, want, "medium %d" % k)


def _random_segments(rng, P, sizes, span, ties):
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    a = rng.integers(0, span, n)
    b = rng.integers(0, span, n)
    if ties:
        a = (a // ties) * ties
        b = (b // ties) * ties
    return a.astype(np.int32), b.astype(np.int32), off


@pytest.mark.parametrize("m", [2, 3, 5, 8, 31, 32, 33, 64, 100, 1000])
def test_random_segments_vs_oracle(m, oracle):
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(m)
    for trial in range(6):
        P = int(rng.choice([1, 2, 7, 300]))
        sizes = rng.integers(0, 2 * m + 50, P) if trial % 2 else rng.integers(0, 9000, P) * (rng.random(P) < 0.7)
        if trial == 4:
            sizes = np.array([4096 - m, 4096, 4097, 1, 0, m, m - 1, 8192 + m][:max(P, 1)] + [0] * max(0, P - 8))
            P = len(sizes)
        span = int(rng.choice([50, 3000, 200_000, 250_000_000]))
        eps = int(rng.choice([1, 5, 50, 500, 5000]))
        a, b, off = _random_segments(rng, P, sizes, span, int(rng.choice([0, 0, 10])))
        if len(a) == 0:
            continue
        want = oracle.cluster_segments(a, b, off, eps, m)
        got = device_ops.cluster_labels(a, b, off, eps, m, span)
        assert_same(got, want, "m=%d trial=%d P=%d n=%d span=%d eps=%d" % (m, trial, P, len(a), span, eps))


def test_dense_regime_and_hotspots(oracle):
    """Density above m/eps: chromosome-long x-runs, all separation happens in the y-pass."""
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(77)
    n = 300_000
    a = rng.integers(0, 3_000_000, n).astype(np.int32)          # 0.1 signals/bp >> 3/500
    b = rng.integers(0, 200_000_000, n).astype(np.int32)
    b[:60_000] = 5_000_000 + rng.integers(0, 2000, 60_000)       # a hotspot in y
    off = np.array([0, 100_000, 100_000, n], dtype=np.int64)
    for eps, m in [(500, 3), (1000, 5), (50, 2)]:
        assert_same(device_ops.cluster_labels(a, b, off, eps, m, 200_000_000),
                    oracle.cluster_segments(a, b, off, eps, m), "dense eps=%d m=%d" % (eps, m))


def test_config2_one_million(oracle):
    """BASELINE config 2: 1 M signals, one pair, eps=500, m=3 -- the reference's ids, bit for bit."""
    from tiddit_b200 import device_ops, synth
    a, b, off, L = synth.config2_signals()
    want = oracle.cluster_segments(a, b, off, 500, 3)
    got = device_ops.cluster_labels(a, b, off, 500, 3, L)
    assert_same(got, want, "config2")
    assert 10_000 < len(np.unique(got)) < 400_000


def test_wgs30x_full_size(oracle):
    """BASELINE config 3 at full size (20 M signals, 300 pairs) against the C oracle + invariants."""
    import torch
    from tiddit_b200 import device_ops, synth
    a, b, off, L = synth.wgs30x_signals()
    got = device_ops.cluster_labels(a, b, off, 500, 3, L)
    want = oracle.cluster_segments(a, b, off, 500, 3)
    assert_same(got, want, "wgs30x")
    # size-independent properties: idempotent (same input -> same labels); ids stay below the pair's signal
    # count (tiddit_cluster.pyx:166 hands out len(cluster_pos)+k to noise contigs); reversing the order of the
    # pairs only moves the label blocks
    again = device_ops.cluster_labels(a, b, off, 500, 3, L)
    assert np.array_equal(got, again)
    for p in (0, 1, 150, 299):
        seg = got[off[p]:off[p + 1]]
        assert seg.max() < len(seg) and seg.min() >= -1
    sizes = np.diff(off)
    rev = np.concatenate([np.arange(off[p], off[p + 1]) for p in range(len(sizes) - 1, -1, -1)])
    off_rev = np.concatenate([[0], np.cumsum(sizes[::-1])]).astype(np.int64)
    got_rev = device_ops.cluster_labels(a[rev], b[rev], off_rev, 500, 3, L)
    assert np.array_equal(got_rev, got[rev])
    del torch


def test_tumor60x_params(oracle):
    """BASELINE config 5 shape at 5 M signals: eps=1000, m=5, dense clusters + hotspots."""
    from tiddit_b200 import device_ops, synth
    a, b, off, L = synth.tumor60x_signals(5_000_000)
    assert_same(device_ops.cluster_labels(a, b, off, 1000, 5, L), oracle.cluster_segments(a, b, off, 1000, 5), "tumor")


def test_tumor60x_full_size(oracle):
    """BASELINE config 5 at FULL size: 50 M signals, eps=1000, m=5 (70 % in dense clusters, 100 hotspots of
    1e4-1e5 signals) against the C oracle, bit for bit, plus the idempotence property."""
    from tiddit_b200 import device_ops, synth, _lib
    a, b, off, L = synth.tumor60x_signals(50_000_000)
    got = device_ops.cluster_labels(a, b, off, 1000, 5, L)
    want = oracle.cluster_segments(a, b, off, 1000, 5)
    assert_same(got, want, "tumor60x 50M")
    assert 0.6 < (got >= 0).mean() < 0.9
    _lib.release_workspaces()


def test_errors():
    from tiddit_b200 import DBSCAN, device_ops, _lib
    d = data3([1, 2, 3], [1, 2, 3])
    with pytest.raises(ValueError):
        DBSCAN.main(d, 10, 1)                       # max() arg is an empty sequence
    with pytest.raises(_lib.TdtError) as e:
        device_ops.cluster_labels(np.array([5, 900], dtype=np.int32), np.array([1, 2], dtype=np.int32), [0, 2], 10, 2, 100)
    assert e.value.code == _lib.TDT_E_RANGE
    assert len(device_ops.cluster_labels(np.zeros(0, np.int32), np.zeros(0, np.int32), [0], 10, 2)) == 0
    # eps <= 0: nothing clusters; fractional eps compares like the reference (d < 0.1 <=> d == 0)
    assert DBSCAN.main(d, 0, 2).tolist() == [-1, -1, -1]
    assert DBSCAN.main(data3([1, 1, 1, 10], [2, 2, 2, 11]), 0.1, 2).tolist() == [0, 0, -1, -1]


@pytest.mark.parametrize("case", [0, 1, 2])
def test_cluster_main_end_to_end(case):
    from tiddit_b200 import tiddit_cluster
    from test_host import _same_candidates
    exp = load_json("cluster_case%d_expected.json" % case)
    a = exp["args"]
    got = tiddit_cluster.main(os.path.join(GOLDEN, "cluster_case%d" % case), a["chromosomes"], a["contig_length"],
                              a["samples"], a["is_mp"], a["epsilon"], a["m"], a["max_ins_len"], a["min_contig"],
                              a["skip_assembly"], a["min_reads"])
    _same_candidates(got, exp["candidates"], exp["order"])
