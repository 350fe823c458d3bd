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
