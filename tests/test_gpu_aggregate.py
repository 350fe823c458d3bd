"""GPU: candidate aggregation (tdt_cluster_aggregate) against the golden rows taken from the real
tiddit_cluster.main, and against the oracle on larger packed sets."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _agg_equal(got, want):
    """(rows, members) equal up to the placement of the member blocks inside member_idx."""
    gr, gm = got
    wr, wm = want
    assert gr.shape == wr.shape
    keep = [c for c in range(16) if c != 3]
    assert np.array_equal(gr[:, keep], wr[:, keep])
    assert len(gm) == len(wm)
    # member lists, candidate by candidate, flattened in row order
    def flat(rows, mem):
        idx = np.concatenate([np.arange(o, o + s) for o, s in zip(rows[:, 3].tolist(), rows[:, 4].tolist())]) \
            if len(rows) else np.zeros(0, dtype=np.int64)
        return mem[idx]
    assert np.array_equal(flat(gr, gm), flat(wr, wm))


def test_aggregate_golden_rows():
    from tiddit_b200 import device_ops
    z = np.load(os.path.join(GOLDEN, "aggregate_cases.npz"))
    for c in range(int(z["n_cases"])):
        k = "c%d_" % c
        mil, is_mp, mr, eps, m = z[k + "params"].tolist()
        labels = device_ops.cluster_labels(z[k + "posA"], z[k + "posB"], z[k + "seg_off"], eps, m)
        assert np.array_equal(labels, z[k + "labels"])
        rows, mem = device_ops.cluster_aggregate(labels, z[k + "posA"], z[k + "posB"], z[k + "span"], z[k + "name_id"],
                                                 z[k + "flags"], z[k + "seg_off"], z[k + "same_chrom"], mil, is_mp, mr)
        assert np.array_equal(rows[:, [0, 1, 5, 6, 7, 8, 9, 10, 11, 12, 13, 4]], z[k + "expected"]), c
        assert len(mem) == int(rows[:, 4].sum())


@pytest.mark.parametrize("n,is_mp,min_reads,eps,m", [(60_000, False, 3, 500, 3), (400_000, True, 2, 500, 3),
                                                     (1_500_000, False, 5, 1000, 5)])
def test_aggregate_vs_oracle_wgs_shaped(n, is_mp, min_reads, eps, m, oracle):
    from tiddit_b200 import device_ops, synth
    a, b, off, L = synth.wgs30x_signals(n) if m == 3 else synth.tumor60x_signals(n)
    rec = synth.signal_records(a, b, off, seed=n % 97)
    labels = device_ops.cluster_labels(a, b, off, eps, m, L)
    args = (labels, a, b, rec["span"], rec["name_id"], rec["flags"], off, rec["same_chrom"], 5000, is_mp, min_reads)
    got = device_ops.cluster_aggregate(*args, max_pos=L, n_names=rec["n_names"])
    want = oracle.cluster_aggregate(*args)
    assert len(want[0]) > 100 and len(set(want[0][:, 14].tolist())) >= 4     # the branches of :265-330 all occur
    _agg_equal(got, want)
    # unknown max_pos / n_names (30-bit keys) gives the same answer
    _agg_equal(device_ops.cluster_aggregate(*args), want)


def test_aggregate_edge_cases(oracle):
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(4)
    # empty input
    rows, mem = device_ops.cluster_aggregate(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                                             np.zeros((0, 4), np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8),
                                             np.zeros(1, np.int64), np.zeros(0, np.uint8), 100, False, 3)
    assert rows.shape == (0, 16) and len(mem) == 0
    # all noise, nothing survives; all noise contigs, everything survives; one giant candidate; empty pairs in between
    n = 5000
    posA = rng.integers(0, 1000, n).astype(np.int32)
    posB = (posA + rng.integers(0, 50, n)).astype(np.int32)
    span = np.stack([posA - 10, posA, posB, posB + 10], axis=1).astype(np.int32)
    names = rng.integers(0, 50, n).astype(np.int32)
    off = np.array([0, 0, 2000, 2000, 2000, 5000, 5000], dtype=np.int64)
    same = np.array([1, 1, 0, 1, 1, 0], dtype=np.uint8)
    for labels, kinds in [(np.full(n, -1), np.zeros(n)), (np.full(n, -1), np.full(n, 2)), (np.zeros(n), rng.integers(0, 3, n)),
                          (rng.integers(-1, 3, n), rng.integers(0, 3, n))]:
        flags = (kinds.astype(np.uint8) | rng.choice([0x04, 0x08], n).astype(np.uint8) | rng.choice([0x10, 0x20], n).astype(np.uint8))
        args = (labels.astype(np.int32), posA, posB, span, names, flags, off, same, 20, False, 3)
        _agg_equal(device_ops.cluster_aggregate(*args), oracle.cluster_aggregate(*args))
    # a label outside [-1, len(pair)) is a data error
    bad = np.zeros(n, np.int32)
    bad[7] = 4000      # pair 1 has 2000 signals
    with pytest.raises(Exception):
        device_ops.cluster_aggregate(bad, posA, posB, span, names, flags, off, same, 20, False, 3)


def test_cluster_packed_roundtrip(tmp_path, oracle):
    """PackedSignals .npz round trip + main_packed == main on a golden scenario."""
    from conftest import load_json
    from tiddit_b200 import tiddit_cluster
    from tiddit_b200.signals import PackedSignals
    exp = load_json("cluster_case1_expected.json")
    a = exp["args"]
    pk = PackedSignals.from_tab(os.path.join(GOLDEN, "cluster_case1"), a["chromosomes"], a["contig_length"], a["samples"],
                                a["is_mp"], a["min_contig"], a["skip_assembly"])
    pk.save(tmp_path / "sig.npz")
    pk2 = PackedSignals.load(tmp_path / "sig.npz")
    got = tiddit_cluster.main_packed(pk2, a["epsilon"], a["m"], a["max_ins_len"], a["is_mp"], a["min_reads"])
    want = tiddit_cluster.main(os.path.join(GOLDEN, "cluster_case1"), a["chromosomes"], a["contig_length"], a["samples"],
                               a["is_mp"], a["epsilon"], a["m"], a["max_ins_len"], a["min_contig"], a["skip_assembly"],
                               a["min_reads"])
    assert got == want
