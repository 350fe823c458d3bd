"""GPU: the hand-written segmented sort (tdt_segsort.cuh, tdt_segsort3.cuh) against numpy's stable sort per segment."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "msd", "lsd"], autouse=True)
def sort_generation(request, monkeypatch):
    """Every test runs on the production dispatch (generation 3 -- MSD rounds + shared-memory finish, tdt_segsort3.cuh --
    for index-valued sorts, the LSD chain otherwise) and on each chain forced for every sort (TDT_SEGSORT=msd / lsd)."""
    monkeypatch.delenv("TDT_SEGSORT", raising=False)
    if request.param in ("lsd", "msd"):
        monkeypatch.setenv("TDT_SEGSORT", request.param)
    return request.param


def _expect(keys, vals, off):
    ko, vo = keys.copy(), vals.copy()
    for s in range(len(off) - 1):
        lo, hi = off[s], off[s + 1]
        order = np.argsort(keys[lo:hi], kind="stable")
        ko[lo:hi] = keys[lo:hi][order]
        vo[lo:hi] = vals[lo:hi][order]
    return ko, vo


def _run(keys, vals, off, key_bits, with_segid=False):
    import torch
    from tiddit_b200 import device_ops
    k = torch.from_numpy(keys.astype(np.int64).astype(np.uint32).view(np.int32)).cuda()
    v = torch.from_numpy(vals).cuda() if vals is not None else None
    off = np.asarray(off, dtype=np.int64)
    o = torch.from_numpy(off).cuda()
    sid = None
    if with_segid:   # per-element segment index: switches the counting path for segments of <= 32 elements on
        sid = torch.from_numpy(np.repeat(np.arange(len(off) - 1, dtype=np.int32), np.diff(off))).cuda()
    ko, vo = device_ops.segsort_device(k, v, o, key_bits, sid)
    return ko.cpu().numpy().view(np.uint32), vo.cpu().numpy()


@pytest.mark.parametrize("sizes", [
    [10], [2048], [2049], [4096], [4097], [100_000], [0, 0, 5, 0, 3000, 1, 0, 2047, 2048, 2049, 9000, 0, 7, 0],
    [1] * 5000, [3, 4, 5] * 3000, [2048] * 7, [50_000, 3, 70_000, 2048, 2047, 1_000_001, 12, 0]])
@pytest.mark.parametrize("key_bits,ties", [(28, 0), (8, 0), (31, 0), (17, 1), (1, 0)])
def test_segsort_matches_numpy(sizes, key_bits, ties):
    rng = np.random.default_rng(len(sizes) * 31 + key_bits)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    hi = 1 << key_bits
    keys = rng.integers(0, hi, n, dtype=np.int64).astype(np.uint32)
    if ties:
        keys = (keys >> 9) << 9
    vals = rng.permutation(n).astype(np.int32)
    want_k, want_v = _expect(keys, vals, off)
    got_k, got_v = _run(keys, vals, off, key_bits)
    assert np.array_equal(got_k, want_k)
    assert np.array_equal(got_v, want_v)
    # vals = None -> element index; with a segment-id array (tiny segments take the counting path)
    got_k2, got_v2 = _run(keys, None, off, key_bits, with_segid=True)
    want_k2, want_v2 = _expect(keys, np.arange(n, dtype=np.int32), off)
    assert np.array_equal(got_k2, want_k2) and np.array_equal(got_v2, want_v2)


def test_segsort_many_tiny_and_huge_mix():
    rng = np.random.default_rng(9)
    sizes = np.concatenate([rng.integers(0, 12, 400_000), rng.integers(20, 80, 20_000), [3_000_000],
                            rng.integers(0, 5000, 300)])
    rng.shuffle(sizes)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    keys = rng.integers(0, 250_000_000, n, dtype=np.int64).astype(np.uint32)
    vals = np.arange(n, dtype=np.int32)
    got_k, got_v = _run(keys, vals, off, 28, with_segid=True)
    # check by properties (a python loop over 400k segments is slow): sorted inside segments, stable, a permutation
    seg = np.repeat(np.arange(len(sizes)), sizes)
    order = np.lexsort((vals, keys, seg))
    assert np.array_equal(got_k, keys[order])
    assert np.array_equal(got_v, vals[order])


@pytest.mark.parametrize("kind", ["all_equal", "two_values", "pileup", "hotspot", "narrow", "boundaries", "sorted", "reversed"])
def test_segsort_skewed_distributions(kind):
    """What the partition rounds and the shared-memory finish must survive: heavy exact duplicates, pile-ups inside an
    otherwise uniform segment, a 2 kb hotspot, segment sizes around the batch capacity."""
    rng = np.random.default_rng(sum(map(ord, kind)))
    if kind == "boundaries":
        sizes = [6143, 6144, 6145, 8191, 8192, 8193, 33, 32, 31, 1, 2, 12288, 12289, 0, 20000]
    else:
        sizes = [300_000, 5000, 40_000, 7, 6144]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    keys = rng.integers(0, 200_000_000, n, dtype=np.int64)
    if kind == "all_equal":
        keys[:] = 12345
    elif kind == "two_values":
        keys = np.where(rng.random(n) < 0.5, 77, 150_000_000)
    elif kind == "pileup":
        keys[rng.random(n) < 0.4] = 99_000_000          # 40 % of every segment at one position
        keys[rng.random(n) < 0.1] = 5
    elif kind == "hotspot":
        hot = rng.random(n) < 0.6
        keys[hot] = 50_000_000 + rng.integers(0, 2000, int(hot.sum()))
    elif kind == "narrow":
        keys = 1000 + rng.integers(0, 50, n)
    elif kind == "sorted":
        keys = np.sort(keys)
    elif kind == "reversed":
        keys = np.sort(keys)[::-1].copy()
    keys = keys.astype(np.uint32)
    vals = rng.permutation(n).astype(np.int32)
    want_k, want_v = _expect(keys, vals, off)
    got_k, got_v = _run(keys, vals, off, 28)
    assert np.array_equal(got_k, want_k)
    assert np.array_equal(got_v, want_v)
    # value = element index (the posA sort): the finish kernel breaks ties by value instead of by position
    got_k2, got_v2 = _run(keys, None, off, 28)
    want_k2, want_v2 = _expect(keys, np.arange(n, dtype=np.int32), off)
    assert np.array_equal(got_k2, want_k2) and np.array_equal(got_v2, want_v2)


@pytest.mark.parametrize("key_bits", [5, 12, 16, 20, 24, 28, 32])
@pytest.mark.parametrize("kind", ["uniform", "narrow", "all_equal", "hot", "clustered"])
def test_segsort_rounds_and_levels(kind, key_bits):
    """Generation 3: every number of partition rounds (1-4), ranges that end in the scratch buffers (copy batches),
    batches finished in place, crowded slots (bitonic path), sizes around the batch capacity."""
    rng = np.random.default_rng(key_bits * 7 + len(kind))
    sizes = [8191, 8192, 8193, 16384, 16385, 2049, 300_000, 0, 9000, 70_000, 1_300_000]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    hi = 1 << key_bits
    keys = rng.integers(0, hi, n, dtype=np.int64)
    if kind == "narrow":
        keys = (hi // 3) + rng.integers(0, min(hi - hi // 3, 40), n)
    elif kind == "all_equal":
        keys[:] = hi - 1
    elif kind == "hot":
        hot = rng.random(n) < 0.7
        keys[hot] = (hi // 2) + rng.integers(0, min(hi // 2, 3), int(hot.sum()))
    elif kind == "clustered":
        centres = rng.integers(0, hi, 2000, dtype=np.int64)
        keys = np.clip(centres[rng.integers(0, 2000, n)] + rng.integers(-100, 100, n), 0, hi - 1)
    keys = keys.astype(np.uint32)
    vals = rng.permutation(n).astype(np.int32)
    want_k, want_v = _expect(keys, vals, off)
    got_k, got_v = _run(keys, vals, off, key_bits)
    assert np.array_equal(got_k, want_k)
    assert np.array_equal(got_v, want_v)
    got_k2, got_v2 = _run(keys, None, off, key_bits)
    want_k2, want_v2 = _expect(keys, np.arange(n, dtype=np.int32), off)
    assert np.array_equal(got_k2, want_k2) and np.array_equal(got_v2, want_v2)


def test_segsort_key_range_error():
    from tiddit_b200 import _lib
    with pytest.raises(_lib.TdtError):
        _run(np.array([1, 2, 300], dtype=np.uint32), None, [0, 3], 8)
