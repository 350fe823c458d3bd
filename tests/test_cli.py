"""`python -m tiddit_b200 --cov` (BASELINE config 1) and the BAM reader/writer behind it."""
import filecmp
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_bam_roundtrip(tmp_path):
    from tiddit_b200 import bamio
    contigs = [("a", 5000), ("b", 70000)]
    reads = bamio.synthetic_reads(contigs, 3000, seed=3)
    path = str(tmp_path / "x.bam")
    bamio.write_bam(path, contigs, reads)
    with bamio.AlignmentFile(path) as f:
        assert f.header["SQ"] == [{"SN": "a", "LN": 5000}, {"SN": "b", "LN": 70000}]
        got = list(f.fetch(until_eof=True))
    assert len(got) == len(reads)
    for g, r in zip(got, reads):
        assert (g.query_name, g.flag, g.reference_start, g.mapq) == (r["name"], r["flag"], r["pos"], r["mapq"])
        assert g.reference_name == contigs[r["ref"]][0]
        ref_len = sum(n for op, n in r["cigar"] if op in (0, 2, 3, 7, 8))
        assert g.reference_end == (None if g.is_unmapped else r["pos"] + ref_len)
        assert g.cigartuples == r["cigar"]
    assert any(g.is_duplicate for g in got) and any(g.is_unmapped for g in got) and any(g.is_secondary for g in got)


class _OracleCoverage:
    """Stand-in for tiddit_coverage.DeviceCoverage backed by the oracle (CPU test of the CLI plumbing only)."""

    def __init__(self, header, bin_size):
        from oracle import oracle
        self.o, self.z = oracle, bin_size
        self.cov, self.ebs = oracle.create_coverage(header, bin_size)

    def add_reads(self, contig, s, e):
        self.o.update_coverage_batch(s, e, self.z, self.cov[contig], self.ebs[contig])

    def to_host(self):
        return self.cov, self.ebs


@pytest.mark.parametrize("z,q,kind", [(500, 20, "bed"), (500, 20, "wig"), (50, 5, "bed")])
def test_cov_cli_plumbing_cpu(z, q, kind, tmp_path, monkeypatch, oracle):
    from tiddit_b200 import __main__ as cli, tiddit_coverage
    monkeypatch.setattr(tiddit_coverage, "DeviceCoverage", _OracleCoverage)
    out = str(tmp_path / "out")
    argv = ["--cov", "--bam", os.path.join(GOLDEN, "config1.bam"), "-o", out, "-z", str(z), "-q", str(q)]
    assert cli.main(argv + (["-w"] if kind == "wig" else [])) == 0
    assert filecmp.cmp(out + "." + kind, os.path.join(GOLDEN, "config1_z%d_q%d.%s" % (z, q, kind)), shallow=False)


def test_cli_missing_bam_and_usage(capsys):
    from tiddit_b200 import __main__ as cli
    assert cli.main(["--cov", "--bam", "/nonexistent.bam"]) == 1
    assert "could not find the bam file" in capsys.readouterr().out
    assert cli.main([]) == 0
    assert cli.main(["--sv", "--bam", "x.bam", "--ref", "r.fa"]) == 2     # reference package absent here


@pytest.mark.gpu
@pytest.mark.parametrize("z,q,kind", [(500, 20, "bed"), (500, 20, "wig"), (50, 5, "bed")])
def test_cov_cli_gpu_byte_identical(z, q, kind, tmp_path):
    """BASELINE config 1 on the GPU path: byte-identical to the reference's output."""
    from tiddit_b200 import __main__ as cli
    out = str(tmp_path / "out")
    argv = ["--cov", "--bam", os.path.join(GOLDEN, "config1.bam"), "-o", out, "-z", str(z), "-q", str(q)]
    assert cli.main(argv + (["-w"] if kind == "wig" else [])) == 0
    assert filecmp.cmp(out + "." + kind, os.path.join(GOLDEN, "config1_z%d_q%d.%s" % (z, q, kind)), shallow=False)


# ---- the reference's own per-read loop on the queueing arrays ---------------------------------
def _reference_cov_loop(bam, z, q, out, wig=False):
    """tiddit/__main__.py:225-247 restated line by line on this package's modules (pysam -> bamio.AlignmentFile)."""
    from tiddit_b200 import bamio, tiddit_coverage
    samfile = bamio.AlignmentFile(bam, "r")
    bam_header = samfile.header
    coverage_data, end_bin_size = tiddit_coverage.create_coverage(bam_header, z)
    n_reads = 0
    for read in samfile.fetch(until_eof=True):
        if read.is_unmapped or read.is_duplicate:
            continue
        if read.mapq >= q:
            n_reads += 1
            name = read.reference_name
            coverage_data[name] = tiddit_coverage.update_coverage(read.reference_start, read.reference_end, z,
                                                                  coverage_data[name], end_bin_size[name])
    tiddit_coverage.print_coverage(coverage_data, bam_header, z, "wig" if wig else "bed", out + (".wig" if wig else ".bed"))
    return n_reads


def _oracle_backend(monkeypatch, oracle):
    from tiddit_b200 import tiddit_coverage

    def accumulate(s, e, z, ebs, host, dev):
        dev = host.copy() if dev is None else dev
        try:
            oracle.update_coverage_batch(s, e, z, dev, ebs)
            return dev, False
        except IndexError:
            return dev, True
    monkeypatch.setattr(tiddit_coverage, "_accumulate_resident", accumulate)
    monkeypatch.setattr(tiddit_coverage, "_download", lambda d: d)


def test_queued_update_coverage_host_logic_cpu(tmp_path, monkeypatch, oracle):
    """CoverageArray: queueing, flush-on-access through every access path, error timing (oracle as the flush backend)."""
    import pickle
    from tiddit_b200 import tiddit_coverage as tc
    _oracle_backend(monkeypatch, oracle)
    hdr = {"SQ": [{"SN": "c1", "LN": 1234}]}
    reads = [(0, 150), (400, 550), (990, 1234), (1100, 1234), (499, 501), (0, 1234)]
    want = [1.502000014996156, 1.1179999969899654, 2.2594529390335083]        # SURVEY App. B known answer
    for access in (lambda c: c.tolist(), lambda c: list(c), lambda c: np.asarray(c).tolist(), lambda c: [c[0], c[1], c[2]],
                   lambda c: pickle.loads(pickle.dumps(c)).tolist(), lambda c: list(memoryview(c)),
                   lambda c: (c + 0.0).tolist(), lambda c: np.sort(c)[[1, 0, 2]].tolist(), lambda c: np.array(c, copy=False).tolist(),
                   lambda c: np.asanyarray(c).tolist(), lambda c: [max(c[0:1]), np.average(c[1:2]), c[2:][0]], lambda c: c.copy().tolist(),
                   lambda c: np.frombuffer(c.tobytes()).tolist(), lambda c: [float(x) for x in str(c.astype(object))[1:-1].split()]):
        cov, ebs = tc.create_coverage(hdr, 500, "c1")
        assert cov.dtype == np.float64 and cov.shape == (3,) and len(cov) == 3 and ebs == 234
        for s, e in reads:
            cov = tc.update_coverage(s, e, 500, cov, ebs)
        assert access(cov) == want
    cov, ebs = tc.create_coverage(hdr, 500, "c1")
    cov[1] = 5.0                                            # host writes interleave with queued reads
    tc.update_coverage(400, 550, 500, cov, ebs)
    assert cov[1] == 5.0 + np.float64(np.float32(49) / np.float32(500))
    with pytest.raises(IndexError):                         # beyond the contig: the reference's IndexError, in the call
        tc.update_coverage(1200, 1600, 500, cov, ebs)
    tc.update_coverage(-100, -50, 500, cov, ebs)            # negative bins wrap around like the reference's indexing
    assert cov[2] == np.float64(np.float32(50) / np.float32(500))
    with pytest.raises(ZeroDivisionError):
        tc.update_coverage(0, 10, 0, cov, ebs)
    out = str(tmp_path / "loop")
    assert _reference_cov_loop(os.path.join(GOLDEN, "config1.bam"), 500, 20, out) > 5000
    assert filecmp.cmp(out + ".bed", os.path.join(GOLDEN, "config1_z500_q20.bed"), shallow=False)


@pytest.mark.gpu
def test_reference_cov_loop_on_gpu_byte_identical(tmp_path):
    """The reference's `--cov` read loop, unchanged, over create_coverage / update_coverage per read: byte-identical
    config-1 bed and wig; and the per-read call sustains > 1 M reads/s (it only queues; the kernel runs per batch)."""
    import time
    from tiddit_b200 import tiddit_coverage as tc
    for z, q, kind in [(500, 20, "bed"), (500, 20, "wig"), (50, 5, "bed")]:
        out = str(tmp_path / ("loop_%d_%d" % (z, q)))
        _reference_cov_loop(os.path.join(GOLDEN, "config1.bam"), z, q, out, wig=kind == "wig")
        assert filecmp.cmp(out + "." + kind, os.path.join(GOLDEN, "config1_z%d_q%d.%s" % (z, q, kind)), shallow=False)
    hdr = {"SQ": [{"SN": "chr21", "LN": 46_709_983}]}
    cov, ebs = tc.create_coverage(hdr, 50, "chr21")
    n = 3_000_000
    starts = np.sort(np.random.default_rng(1).integers(0, 46_709_983 - 150, n)).tolist()
    t0 = time.perf_counter()
    for s in starts:
        cov = tc.update_coverage(s, s + 150, 50, cov, ebs)
    total = float(cov.sum())
    dt = time.perf_counter() - t0
    assert n / dt > 1e6, "update_coverage sustains %.2f M reads/s" % (n / dt / 1e6)
    assert abs(total - n * (150 - 1) / 50.0) < n * 0.05            # ~2.98 per read (the last bin is credited one base short)
