"""CPU: the oracle (oracle/tdt_oracle.c) against the golden vectors produced by the real reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the reference itself."""
import numpy as np
import pytest

from conftest import data3, load_json, GOLDEN


def test_dbscan_small_golden(oracle):
    cases = load_json("dbscan_small.json")
    assert len(cases) >= 400
    for c in cases:
        d = data3(c["x"], c["y"])
        if len(c["x"]) == 0:
            continue
        xl, cid = oracle.x_coordinate_clustering(d, c["eps"], c["m"])
        assert xl.astype(int).tolist() == c["x_labels"]
        assert cid == c["x_last_id"]
        assert oracle.main(d, c["eps"], c["m"]).astype(int).tolist() == c["labels"]
        # y-pass as a stand-alone call on the x labels
        lab = np.array(c["x_labels"], dtype=np.float64)
        out, _ = oracle.y_coordinate_clustering(d, c["eps"], c["m"], c["x_last_id"], lab)
        assert out.astype(int).tolist() == c["labels"]


def test_dbscan_medium_golden(oracle):
    z = np.load(GOLDEN + "/dbscan_medium.npz")
    for k in range(5):
        a, b, want = z["posA_%d" % k], z["posB_%d" % k], z["labels_%d" % k]
        eps, m = [int(v) for v in z["param_%d" % k]]
        got = oracle.cluster_segments(a, b, [0, len(a)], eps, m)
        assert np.array_equal(got, want)
        assert (want >= 0).sum() > 100 and (want < 0).sum() > 100


def test_coverage_golden(oracle):
    for c in load_json("coverage_cases.json"):
        header = {"SQ": [{"SN": c["name"], "LN": c["LN"]}]}
        data, ebs = oracle.create_coverage(header, c["bin"], c["name"])
        assert ebs == c["end_bin_size"]
        s = [r[0] for r in c["reads"]]
        e = [r[1] for r in c["reads"]]
        oracle.update_coverage_batch(s, e, c["bin"], data, ebs)
        assert [float(v).hex() for v in data] == c["bins_hex"]
        # order independence (the property the GPU atomics rely on)
        data2, _ = oracle.create_coverage(header, c["bin"], c["name"])
        oracle.update_coverage_batch(s[::-1], e[::-1], c["bin"], data2, ebs)
        assert np.array_equal(data, data2)


def test_coverage_out_of_range_raises(oracle):
    data, ebs = oracle.create_coverage({"SQ": [{"SN": "c", "LN": 1234}]}, 500, "c")
    with pytest.raises(IndexError):
        oracle.update_coverage(1200, 1600, 500, data, ebs)


def test_gc_golden(oracle):
    for c in load_json("gc_cases.json"):
        assert oracle.gc_bins(c["seq"], c["bin"], c["n_cutoff"]).tolist() == c["gc"]


def test_oracle_vs_compiled_reference(oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(1)
    for t in range(300):
        n = int(rng.integers(1, 120))
        m = int(rng.integers(2, 6))
        eps = int(rng.integers(1, 60))
        span = int(rng.integers(5, 600))
        d = data3(np.sort(rng.integers(0, span, n)), rng.integers(0, span, n))
        assert np.array_equal(ref.DBSCAN.main(d.copy(), eps, m), oracle.main(d, eps, m))
    cov = ref.tiddit_coverage
    for t in range(20):
        ln, z = int(rng.integers(100, 50000)), int(rng.choice([37, 50, 500]))
        header = {"SQ": [{"SN": "c", "LN": ln}]}
        a, ebs = cov.create_coverage(header, z, "c")
        b, _ = oracle.create_coverage(header, z, "c")
        s = rng.integers(0, ln, 500)
        e = np.minimum(s + rng.integers(1, 1000, 500), ln)
        for i in range(500):
            cov.update_coverage(int(s[i]), int(e[i]), z, a, ebs)
        oracle.update_coverage_batch(s, e, z, b, ebs)
        assert np.array_equal(a, b)


def test_aggregate_golden(oracle):
    """tiddit_cluster.pyx:156-336: the oracle's candidate rows against the rows of the real tiddit_cluster.main
    (14 random scenarios reaching every branch of :265-330), and its labels against the real DBSCAN.main."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "aggregate_cases.npz"))
    rules = set()
    for c in range(int(z["n_cases"])):
        k = "c%d_" % c
        mil, is_mp, mr, eps, m = z[k + "params"].tolist()
        assert np.array_equal(oracle.cluster_segments(z[k + "posA"], z[k + "posB"], z[k + "seg_off"], eps, m), z[k + "labels"])
        rows, mem = oracle.cluster_aggregate(z[k + "labels"], z[k + "posA"], z[k + "posB"], z[k + "span"], z[k + "name_id"],
                                             z[k + "flags"], z[k + "seg_off"], z[k + "same_chrom"], mil, is_mp, mr)
        assert np.array_equal(rows[:, [0, 1, 5, 6, 7, 8, 9, 10, 11, 12, 13, 4]], z[k + "expected"]), c
        rules |= set(rows[:, 14].tolist())
    assert rules == {0, 1, 2, 3, 4}


def test_masked_medians_golden(oracle):
    """tiddit_coverage_analysis.pyx:14-27: the oracle's medians against the library the real determine_ploidy built."""
    from conftest import load_json
    for c in load_json("ploidy_cases.json"):
        names = c["names"]
        cov = np.concatenate([np.array(c["cov"][n], dtype=np.float64) for n in names])
        gc = np.concatenate([np.array(c["gc"][n], dtype=np.int8) for n in names])
        off = np.concatenate([[0], np.cumsum([len(c["cov"][n]) for n in names])])
        med, cnt = oracle.coverage_medians(cov, gc, off)
        for i, n in enumerate(names):
            want = c["library"]["avg_coverage_" + n]
            assert (np.isnan(med[i]) and want == 0 and cnt[i] == 0) or med[i] == want
        if not c["c"]:
            assert med[-1] == c["library"]["avg_coverage"]
