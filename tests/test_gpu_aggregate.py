"""GPU: candidate aggregation (tdt_cluster_aggregate) against the golden rows taken from the real
tiddit_cluster.main, and against the oracle on larger packed sets."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["default", "0", "4", "1024"], autouse=True)
def direct_threshold(request, monkeypatch):
    """Every test runs with the production threshold of the direct modes / names kernel (candidates of up to 256 members),
    with the kernel off (every candidate through the three sub-sorts), with a threshold of 4 (most candidates take the
    compaction + sort path of the big candidates) and at the kernel's limit (tdt_aggregate.cu, TDT_AGG_DIRECT)."""
    monkeypatch.delenv("TDT_AGG_DIRECT", raising=False)
    if request.param != "default":
        monkeypatch.setenv("TDT_AGG_DIRECT", request.param)
    return request.param


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


def test_aggregate_ties_small_and_big_candidates(oracle):
    """Candidates of 1 ... 3000 members side by side, few distinct positions and names per candidate: every mode is
    decided by the first-inserted tie-break, members of one candidate spread over several warps / CTAs, candidates
    around every threshold of the direct kernel (4, 128, 1024) and around a warp boundary."""
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(11)
    sizes = [1, 2, 3, 4, 5, 31, 32, 33, 64, 127, 128, 129, 130, 255, 256, 257, 700, 1023, 1024, 1025, 3000] + \
        rng.integers(1, 40, 400).tolist()
    per_pair = [sizes[0::3], sizes[1::3], sizes[2::3]]
    labels, off = [], [0]
    for ps in per_pair:
        lab = np.repeat(np.arange(len(ps)), ps)
        lab = np.concatenate([lab, np.full(50, -1)])          # noise in between
        rng.shuffle(lab)                                      # members of a candidate are not neighbours in the input
        labels.append(lab)
        off.append(off[-1] + len(lab))
    labels = np.concatenate(labels).astype(np.int32)
    n = len(labels)
    posA = rng.integers(100, 104, n).astype(np.int32)         # four values: ties everywhere
    posB = (posA + rng.integers(0, 3, n)).astype(np.int32)
    span = np.stack([posA - rng.integers(0, 9, n), posA + 1, posB, posB + rng.integers(1, 9, n)], axis=1).astype(np.int32)
    names = rng.integers(0, 6, n).astype(np.int32)
    kinds = rng.integers(0, 3, n).astype(np.uint8)
    flags = kinds | rng.choice([0x04, 0x08], n).astype(np.uint8) | rng.choice([0x10, 0x20], n).astype(np.uint8)
    same = np.array([1, 0, 1], dtype=np.uint8)
    for min_reads in (1, 3, 2000):
        args = (labels, posA, posB, span, names, flags, np.asarray(off, np.int64), same, 20, False, min_reads)
        _agg_equal(device_ops.cluster_aggregate(*args), oracle.cluster_aggregate(*args))


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
