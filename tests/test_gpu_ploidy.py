"""GPU: masked coverage medians / determine_ploidy against the real reference's golden output and the oracle."""
import numpy as np
import pytest

from conftest import load_json

pytestmark = pytest.mark.gpu


def test_determine_ploidy_golden(tmp_path):
    from tiddit_b200 import tiddit_coverage_analysis as tca
    for k, c in enumerate(load_json("ploidy_cases.json")):
        cov = {n: np.array(c["cov"][n], dtype=np.float64) for n in c["names"]}
        gc = {n: np.array(c["gc"][n], dtype=np.int8) for n in c["names"]}
        prefix = str(tmp_path / ("case%d" % k))
        lib = tca.determine_ploidy(cov, c["contigs"], {"x": 1}, c["ploidy"], prefix, c["c"], "ref.fa", 50, {"SQ": []}, gc)
        assert list(lib) == list(c["library"])
        for key, want in c["library"].items():
            assert lib[key] == want, (k, key)
            assert type(lib[key]).__name__ == c["library_types"][key], (k, key)
        assert open(prefix + ".ploidies.tab").read() == c["tab"]          # byte-identical


@pytest.mark.parametrize("n_contigs,max_bins,seed", [(1, 1, 0), (3, 10, 1), (24, 300_000, 2), (2000, 3000, 3),
                                                      (500, 5000, 4), (40, 200_000, 7)])
def test_medians_random_vs_oracle(n_contigs, max_bins, seed, oracle):
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(seed)
    sizes = rng.integers(0, max_bins + 1, n_contigs)
    sizes[0] = max_bins
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    style = seed % 3
    if style == 0:
        cov = rng.integers(0, 4, n).astype(np.float64)                     # massive ties
    elif style == 1:
        cov = rng.random(n) * 10.0 ** rng.integers(-3, 6, n)               # spread over many exponents
    else:
        cov = np.float32(rng.integers(0, 400, n)).astype(np.float64) / np.float32(50)   # coverage-like
    cov[rng.random(n) < 0.15] = 0.0
    gc = rng.integers(-1, 60, n).astype(np.int8)
    got_m, got_c = device_ops.coverage_medians(cov, gc, off)
    want_m, want_c = oracle.coverage_medians(cov, gc, off)
    assert np.array_equal(got_c, want_c)
    assert np.array_equal(got_m.view(np.uint64), want_m.view(np.uint64))   # bit-exact, nan included


@pytest.mark.parametrize("kind", ["all_equal_over_capacity", "ties_within_capacity", "two_values", "continuous"])
def test_medians_compaction_and_fallback(kind, oracle):
    """After two digit passes the prefix buckets are copied out and the remaining passes run on the copies; when the
    buckets together exceed the reserved 4 M keys (nearly all bins equal) the streaming passes run instead -- both
    ways must give numpy's median bit for bit, including the value after the lower median."""
    from tiddit_b200 import device_ops
    rng = np.random.default_rng(11)
    sizes = [2_600_000, 1_400_001, 7, 0, 999_999]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    if kind == "all_equal_over_capacity":
        cov = np.full(n, 29.875)                                           # every bin in one bucket: 2 x 5 M keys > 4 M
    elif kind == "ties_within_capacity":
        cov = rng.integers(1, 2000, n).astype(np.float64) / 64.0           # ~2000 distinct values, exact ties
    elif kind == "two_values":
        cov = np.where(rng.random(n) < 0.5, 30.0, np.nextafter(30.0, 31.0))   # the upper median differs in the last bit
    else:
        cov = rng.gamma(9.0, 3.3, n)
    cov[rng.random(n) < 0.1] = 0.0
    gc = rng.integers(-1, 60, n).astype(np.int8)
    got_m, got_c = device_ops.coverage_medians(cov, gc, off)
    want_m, want_c = oracle.coverage_medians(cov, gc, off)
    assert np.array_equal(got_c, want_c)
    assert np.array_equal(got_m.view(np.uint64), want_m.view(np.uint64))


def test_medians_value_after_the_lower_median_known_answers(oracle):
    """Even counts: the upper median may sit in the same last digit, in a later digit of the last histogram, or in a far
    bucket (other exponent) -- the three places the kernels look for it -- and differs between a contig and the genome."""
    from tiddit_b200 import device_ops
    big, tiny = 1e300, 5e-324
    contigs = [[1.0, 1.0, big, big], [1.0, np.nextafter(1.0, 2.0)], [2.0, 2.0, 2.0, 2.5], [tiny, 3.0], [7.0],
               [0.0, 0.0], [4.0, 4.0 + 2.0 ** -44, 0.0, 9.0]]
    cov = np.array([v for c in contigs for v in c])
    off = np.concatenate([[0], np.cumsum([len(c) for c in contigs])]).astype(np.int64)
    gc = np.full(len(cov), 40, dtype=np.int8)
    got_m, got_c = device_ops.coverage_medians(cov, gc, off)
    want = [np.median([v for v in c if v > 0]) if any(v > 0 for v in c) else np.nan for c in contigs]
    want.append(np.median(cov[cov > 0]))
    assert np.array_equal(got_m.view(np.uint64), np.array(want).view(np.uint64))
    assert got_c.tolist() == [sum(v > 0 for v in c) for c in contigs] + [int((cov > 0).sum())]
    want_m, _ = oracle.coverage_medians(cov, gc, off)
    assert np.array_equal(got_m.view(np.uint64), want_m.view(np.uint64))


def test_medians_genome_scale_property(oracle):
    """61.8 M bins (bin size 50, GRCh38): the device medians satisfy the defining rank property on the full arrays,
    and equal the oracle on a 3-contig slice."""
    import torch
    from tiddit_b200 import device_ops, synth
    lens = np.array([ln for _, ln in synth.GRCH38], dtype=np.int64)
    nb = (lens + 49) // 50
    off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    n = int(off[-1])
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    cov = torch.round(torch.rand(n, device="cuda", generator=g, dtype=torch.float64) * 3000) / 50.0
    gc = torch.randint(-1, 80, (n,), device="cuda", generator=g, dtype=torch.int8)
    med, cnt = device_ops.coverage_medians_device(cov, gc, torch.from_numpy(off).cuda(), len(nb))
    med, cnt = med.cpu().numpy(), cnt.cpu().numpy()
    mask = (cov > 0) & (gc != -1)
    assert int(mask.sum().item()) == cnt[-1]
    for c in list(range(len(nb))) + [len(nb)]:
        lo, hi = (int(off[c]), int(off[c + 1])) if c < len(nb) else (0, n)
        v = cov[lo:hi][mask[lo:hi]]
        below, above = int((v < med[c]).sum().item()), int((v > med[c]).sum().item())
        assert below <= len(v) // 2 and above <= len(v) // 2 and len(v) == cnt[c]
    lo, hi = int(off[20]), int(off[23])
    want_m, want_c = oracle.coverage_medians(cov[lo:hi].cpu().numpy(), gc[lo:hi].cpu().numpy(), off[20:24] - off[20])
    assert np.array_equal(want_m[:3], med[20:23]) and np.array_equal(want_c[:3], cnt[20:23])
