"""GPU parity: coverage bins (float64, bit-exact) and GC bins (int8) against the golden vectors of the
real reference and the oracle."""
import numpy as np
import pytest

from conftest import load_json

pytestmark = pytest.mark.gpu


def test_coverage_golden():
    from tiddit_b200 import tiddit_coverage as cov
    for c in load_json("coverage_cases.json"):
        header = {"SQ": [{"SN": c["name"], "LN": c["LN"]}]}
        data, ebs = cov.create_coverage(header, c["bin"], c["name"])
        assert ebs == c["end_bin_size"]
        out = cov.update_coverage_batch([r[0] for r in c["reads"]], [r[1] for r in c["reads"]], c["bin"], data, ebs)
        assert out is data
        assert [float(v).hex() for v in data] == c["bins_hex"], c["name"]


def test_update_coverage_single_read_signature():
    from tiddit_b200 import tiddit_coverage as cov
    c = load_json("coverage_cases.json")[0]          # App. B known-answer vector
    data, ebs = cov.create_coverage({"SQ": [{"SN": "c1", "LN": 1234}]}, 500, "c1")
    for s, e in c["reads"]:
        data = cov.update_coverage(s, e, 500, data, ebs)
    assert data.tolist() == [1.502000014996156, 1.1179999969899654, 2.2594529390335083]
    with pytest.raises(IndexError):
        cov.update_coverage(1200, 1600, 500, data, ebs)
    with pytest.raises(ZeroDivisionError):
        cov.update_coverage(1, 2, 0, data, ebs)


@pytest.mark.parametrize("z,read_len,sorted_reads", [(500, 150, True), (50, 150, True), (37, 151, False),
                                                     (50, 12000, True), (1000, 40, False)])
def test_coverage_random_vs_oracle(z, read_len, sorted_reads, oracle):
    from tiddit_b200 import tiddit_coverage as cov
    rng = np.random.default_rng(z + read_len)
    ln = 3_000_017
    n = 400_000
    s = rng.integers(0, ln, n)
    if sorted_reads:
        s = np.sort(s)
    e = np.minimum(s + rng.integers(1, read_len + 1, n), ln)
    header = {"SQ": [{"SN": "c", "LN": ln}]}
    want, ebs = oracle.create_coverage(header, z, "c")
    oracle.update_coverage_batch(s, e, z, want, ebs)
    got, _ = cov.create_coverage(header, z, "c")
    cov.update_coverage_batch(s, e, z, got, ebs)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # accumulates into existing bins (the reference's +=), and order does not matter
    perm = rng.permutation(n)
    cov.update_coverage_batch(s[perm], e[perm], z, got, ebs)
    oracle.update_coverage_batch(s, e, z, want, ebs)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_device_coverage_all_contigs(oracle):
    from tiddit_b200 import tiddit_coverage as cov, synth
    contigs = [("a", 1_000_003), ("b", 999), ("empty", 5000), ("d", 2_500_000)]
    header = {"SQ": [{"SN": n, "LN": l} for n, l in contigs]}
    s, e, off, _ = synth.coverage_reads(300_000, contigs=[c for c in contigs if c[0] != "empty"])
    names = [c[0] for c in contigs if c[0] != "empty"]
    dc = cov.DeviceCoverage(header, 50, flush_reads=70_000)
    want, ebs = oracle.create_coverage(header, 50)
    for k, name in enumerate(names):
        lo, hi = off[k], off[k + 1]
        for c0 in range(lo, hi, 25_000):                     # streamed in chunks, like a BAM iterator
            dc.add_reads(name, s[c0:min(c0 + 25_000, hi)], e[c0:min(c0 + 25_000, hi)])
        oracle.update_coverage_batch(s[lo:hi], e[lo:hi], 50, want[name], ebs[name])
    got, gebs = dc.to_host()
    assert gebs == ebs
    for name in want:
        assert np.array_equal(got[name].view(np.uint64), want[name].view(np.uint64)), name
    assert got["empty"].sum() == 0


def test_coverage_full_density(oracle):
    """30X density on one 100 Mbp contig (20 M reads, 500-bp bins): heavy same-bin contention."""
    from tiddit_b200 import tiddit_coverage as cov
    rng = np.random.default_rng(3)
    ln, n = 100_000_000, 20_000_000
    s = np.sort(rng.integers(0, ln, n)).astype(np.int32)
    e = np.minimum(s + 150, ln).astype(np.int32)
    header = {"SQ": [{"SN": "c", "LN": ln}]}
    want, ebs = oracle.create_coverage(header, 500, "c")
    oracle.update_coverage_batch(s, e, 500, want, ebs)
    got, _ = cov.create_coverage(header, 500, "c")
    cov.update_coverage_batch(s, e, 500, got, ebs)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert abs(got.mean() - 30.0) < 0.5


def test_gc_golden():
    from tiddit_b200 import tiddit_gc
    for c in load_json("gc_cases.json"):
        got = tiddit_gc.gc_bins(c["seq"], c["bin"], c["n_cutoff"])
        assert got.dtype == np.int8
        assert got.tolist() == c["gc"], (c["bin"], c["n_cutoff"], len(c["seq"]))


@pytest.mark.parametrize("z", [1, 5, 50, 64, 191, 192, 193, 500, 4096, 100_000])
def test_gc_random_vs_oracle(z, oracle):
    from tiddit_b200 import tiddit_gc, synth
    for ln in (1, 15, 16, 17, z, z + 1, 1_000_003):
        seq = synth.fasta_sequence(ln, seed=z + ln)
        for cut in (0.5, 0.0):
            assert np.array_equal(tiddit_gc.gc_bins(seq, z, cut), oracle.gc_bins(seq, z, cut)), (z, ln, cut)


def test_gc_module_surface(tmp_path, oracle):
    from tiddit_b200 import tiddit_gc, synth
    seqs = {"c1": synth.fasta_sequence(250_007, seed=1), "c2": synth.fasta_sequence(1234, seed=2)}
    p = tmp_path / "ref.fa"
    with open(p, "w") as f:
        for name, s in seqs.items():
            f.write(">%s\n" % name)
            t = bytes(s).decode()
            for i in range(0, len(t), 70):
                f.write(t[i:i + 70] + "\n")
    got = tiddit_gc.main(str(p), ["c1", "c2"], 4, 50, 0.5)
    assert list(got) == ["c1", "c2"]
    for name in seqs:
        assert np.array_equal(got[name], oracle.gc_bins(seqs[name], 50, 0.5))
    name, bins = tiddit_gc.binned_gc(str(p), "c2", 50, 0.5)
    assert name == "c2" and bins.dtype == np.int8 and len(bins) == 25


def test_gc_chromosome_scale(oracle):
    from tiddit_b200 import tiddit_gc, synth
    seq = synth.fasta_sequence(60_000_000, seed=4)
    assert np.array_equal(tiddit_gc.gc_bins(seq, 50, 0.5), oracle.gc_bins(seq, 50, 0.5))


def test_coverage_weird_reads_and_alignment(oracle):
    """Negative starts (Python wrap-around), empty / inverted reads, unaligned device slices, bin sizes beyond the
    quotient table -- everything the reference's arithmetic does without raising."""
    import torch
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(12)
    for z in (500, 7, 5000, 100_000):
        ln = 2_000_003
        nb = -(-ln // z)
        ebs = ln - (nb - 1) * z
        s = rng.integers(0, ln, 50_000)
        e = np.minimum(s + rng.integers(1, 3 * z + 2, 50_000), ln)
        s[::97] = -rng.integers(1, min(z, 1000), len(s[::97]))      # wraps into the last bins
        e[::97] = rng.integers(1, 200, len(e[::97]))
        e[5::89] = s[5::89]                                          # empty reads
        e[7::83] = np.maximum(s[7::83] - 3, 1)                       # inverted reads
        want = np.zeros(nb)
        oracle.update_coverage_batch(s, e, z, want, ebs)
        sd = torch.from_numpy(np.concatenate([[0, 0, 0], s]).astype(np.int32)).cuda()[3:]   # 12-byte offset
        ed = torch.from_numpy(np.concatenate([[0], e]).astype(np.int32)).cuda()[1:]
        got = torch.zeros(nb, dtype=torch.float64, device="cuda")
        bad = device_ops.new_first_bad(torch)
        device_ops.coverage_accumulate_device(sd, ed, z, ebs, got, bad)
        assert int(bad.item()) == device_ops.FIRST_BAD_NONE
        assert np.array_equal(got.cpu().numpy().view(np.uint64), want.view(np.uint64)), z


@pytest.mark.parametrize("z", [500, 50, 4096, 5000])
def test_coverage_contigs_call_tile_paths(z, oracle):
    """The all-contig ABI call on a read set whose 8192-read tiles are mostly inside one contig (fast tiles), straddle
    contig boundaries, cover whole tiny contigs, and pile reads onto the last (short) bin of every contig."""
    import torch
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(z)
    lens = [1_000_003, 700, 2_500_000, 8191 * 3, 123_457, 1, 3_000_000]
    counts = [120_000, 300, 90_000, 8192, 30_000, 5, 150_000]
    ss, ee, want, ebs_l, nb_l = [], [], [], [], []
    for ln, k in zip(lens, counts):
        s = rng.integers(0, ln, k)
        s[: k // 10] = ln - 1 - rng.integers(0, min(ln, 2 * z), k // 10)      # crowd the contig end
        s = np.sort(s)
        e = np.minimum(s + rng.integers(1, 301, k), ln)
        nb = -(-ln // z)
        ebs = ln - (nb - 1) * z
        w = np.zeros(nb)
        oracle.update_coverage_batch(s, e, z, w, ebs)
        ss.append(s); ee.append(e); want.append(w); ebs_l.append(ebs); nb_l.append(nb)
    start = torch.from_numpy(np.concatenate(ss).astype(np.int32)).cuda()
    end = torch.from_numpy(np.concatenate(ee).astype(np.int32)).cuda()
    read_off = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).cuda()
    bin_off = torch.from_numpy(np.concatenate([[0], np.cumsum(nb_l)]).astype(np.int64)).cuda()
    ebs_d = torch.tensor(ebs_l, dtype=torch.int32, device="cuda")
    bins = torch.zeros(int(sum(nb_l)), dtype=torch.float64, device="cuda")
    bad = device_ops.new_first_bad(torch)
    device_ops.coverage_accumulate_contigs_device(start, end, read_off, bin_off, ebs_d, z, bins, bad)
    assert int(bad.item()) == device_ops.FIRST_BAD_NONE
    got = bins.cpu().numpy()
    assert np.array_equal(got.view(np.uint64), np.concatenate(want).view(np.uint64))
    # a read past its contig's end is reported (IndexError in the reference), everything else still lands
    end2 = end.clone()
    k_bad = int(counts[0]) - 1
    end2[k_bad] = lens[0] + 2 * z
    bins.zero_()
    device_ops.coverage_accumulate_contigs_device(start, end2, read_off, bin_off, ebs_d, z, bins, bad)
    assert int(bad.item()) == k_bad
