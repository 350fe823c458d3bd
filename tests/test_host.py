"""CPU: host-side logic of the drop-in modules (parsing, folding, text output, generators).
Where a test needs cluster labels without a GPU it injects the ORACLE's labels (test-only stand-in)."""
import filecmp
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_json, unjson


def _same_candidates(got, want_json, order):
    want = unjson(want_json)
    assert list(got) == list(want)
    for a in want:
        assert list(got[a]) == list(want[a])
        for b in want[a]:
            assert [int(k) for k in got[a][b]] == order[a][b]          # dict insertion order drives SV numbering
            for cid, cand in got[a][b].items():
                assert cand == want[a][b][str(cid)], (a, b, cid)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_cluster_main_fold_matches_reference(case, oracle, monkeypatch):
    from tiddit_b200 import tiddit_cluster, device_ops
    exp = load_json("cluster_case%d_expected.json" % case)
    a = exp["args"]
    def cpu_cluster_and_aggregate(posA, posB, seg_off, span, name_id, flags, same_chrom, eps, m, max_ins_len, is_mp,
                                  min_reads, max_pos=0, n_names=0):
        labels = oracle.cluster_segments(posA, posB, seg_off, eps, m)
        rows, mem = oracle.cluster_aggregate(labels, posA, posB, span, name_id, flags, seg_off, same_chrom, max_ins_len,
                                             is_mp, min_reads)
        return labels, rows, mem
    monkeypatch.setattr(device_ops, "cluster_and_aggregate", cpu_cluster_and_aggregate)
    got = tiddit_cluster.main(os.path.join(GOLDEN, "cluster_case%d" % case), a["chromosomes"], a["contig_length"],
                              a["samples"], a["is_mp"], a["epsilon"], a["m"], a["max_ins_len"], a["min_contig"],
                              a["skip_assembly"], a["min_reads"])
    _same_candidates(got, exp["candidates"], exp["order"])


def test_find_discordant_pos_table(ref):
    from tiddit_b200 import tiddit_cluster
    frag = ["r", "c1", "c2", "s3", "e4", None, "s6", "e7", None]
    expected = {  # tiddit_cluster.pyx:8-35
        (True, "False", "True"): ("s3", "e7"), (True, "False", "False"): ("s3", "s6"),
        (True, "True", "True"): ("e4", "e7"), (True, "True", "False"): ("e4", "s6"),
        (False, "False", "True"): ("e4", "s6"), (False, "False", "False"): ("e4", "e7"),
        (False, "True", "True"): ("s3", "s6"), (False, "True", "False"): ("s3", "e7")}
    for (mp, ra, rb), want in expected.items():
        frag[5], frag[8] = ra, rb
        assert tiddit_cluster.find_discordant_pos(frag, mp) == want
        if ref is not None:
            assert ref.tiddit_cluster.find_discordant_pos(frag, mp) == want


def test_create_coverage_shapes():
    from tiddit_b200 import tiddit_coverage as cov
    header = {"SQ": [{"SN": "c1", "LN": 1234}, {"SN": "c2", "LN": 1000}]}
    data, ebs = cov.create_coverage(header, 500)
    assert list(data) == ["c1", "c2"] and len(data["c1"]) == 3 and len(data["c2"]) == 2
    assert ebs == {"c1": 234, "c2": 500} and data["c1"].dtype == np.float64
    one, e1 = cov.create_coverage(header, 500, "c2")
    assert np.asarray(one).dtype == np.float64 and one.shape == (2,) and len(one) == 2 and e1 == 500
    assert cov.create_coverage(header, 500, "nope") == ({}, {})


def test_print_coverage_bytes(oracle, tmp_path):
    from tiddit_b200 import tiddit_coverage as cov
    inp = load_json("print_coverage_input.json")
    header, z = inp["header"], inp["bin"]
    data, ebs = cov.create_coverage(header, z)
    for name, reads in inp["reads"].items():     # bins from the oracle: this test is about the text
        oracle.update_coverage_batch([r[0] for r in reads], [r[1] for r in reads], z, data[name], ebs[name])
    for kind in ("bed", "wig"):
        out = str(tmp_path / ("o." + kind))
        cov.print_coverage(data, header, z, kind, out)
        assert filecmp.cmp(out, os.path.join(GOLDEN, "print_coverage." + kind), shallow=False)


def test_print_coverage_number_format(ref, tmp_path):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    from tiddit_b200 import tiddit_coverage as cov
    header = {"SQ": [{"SN": "c", "LN": 12 * 50 - 3}]}
    vals = np.array([0.0, 1.0, 2.0 ** -32, 1e-5, 123456.789, 1e16, 1 / 3, 0.30000001192092896, 29.98, 1e-4, 5e-324, 7.0])
    for kind in ("bed", "wig"):
        a, b = str(tmp_path / ("a." + kind)), str(tmp_path / ("b." + kind))
        cov.print_coverage({"c": vals}, header, 50, kind, a)
        ref.tiddit_coverage.print_coverage({"c": vals}, header, 50, kind, b)
        assert filecmp.cmp(a, b, shallow=False)


def test_synth_shapes():
    from tiddit_b200 import synth
    a, b, off, L = synth.wgs30x_signals(200_000)
    assert len(a) == len(b) == 200_000 == off[-1] and len(off) == 301
    assert a.dtype == np.int32 and a.min() >= 1 and max(a.max(), b.max()) <= L
    a2, _, _, _ = synth.wgs30x_signals(200_000)
    assert np.array_equal(a, a2)
    s, e, roff, lens = synth.coverage_reads(100_000)
    assert len(s) == 100_000 and np.all(e > s) and np.all(np.diff(s[roff[0]:roff[1]]) >= 0)
    seq = synth.fasta_sequence(100_000)
    assert seq.dtype == np.uint8 and (seq == ord("N")).sum() > 1000


def test_fasta_reader(tmp_path):
    from tiddit_b200 import fasta
    p = tmp_path / "x.fa"
    p.write_text(">c1 desc\nACGT\nNNgg\n>c2\nTT\r\nA\n")
    fa = fasta.NumpyFasta(str(p))
    assert fa.references == ["c1", "c2"]
    assert fa.get_reference_length("c1") == 8 and fa.fetch("c1", 2, 6) == "GTNN" and fa.fetch("c2") == "TTA"


def test_tab_column_reader_equals_line_reader(tmp_path):
    """PackedSignals.from_tab: the pandas column reader and the line-by-line restatement of tiddit_cluster.pyx:47-137
    give the same arrays and tables on the golden scenarios; irregular files (ragged split lines, blank-padded fields,
    non-integer coordinates) are detected and read line by line."""
    import shutil
    from tiddit_b200 import signals
    from tiddit_b200.signals import PackedSignals

    def same(a, b):
        assert a.pairs == b.pairs and a.chrA_present == b.chrA_present
        assert (list(a.names), a.samples, a.ori_table) == (list(b.names), b.samples, b.ori_table)
        for f in PackedSignals.FIELDS:
            assert np.array_equal(getattr(a, f), getattr(b, f)), f

    used = []
    orig = signals._part_from_columns

    def spy(path, *a, **k):
        try:
            part = orig(path, *a, **k)
            used.append((os.path.basename(path), "columns"))
            return part
        except signals._IrregularTab:
            used.append((os.path.basename(path), "lines"))
            raise

    signals._part_from_columns = spy
    try:
        for case in range(3):
            exp = load_json("cluster_case%d_expected.json" % case)
            a = exp["args"]
            for is_mp in (False, True):
                args = (os.path.join(GOLDEN, "cluster_case%d" % case), a["chromosomes"], a["contig_length"], a["samples"],
                        is_mp, a["min_contig"], a["skip_assembly"])
                same(PackedSignals.from_tab(*args, fast="columns"), PackedSignals.from_tab(*args, fast=False))
                native = PackedSignals._from_tab_native(*args)         # libtdt_tab.so: same arrays, same tables
                assert native is not None
                same(native, PackedSignals.from_tab(*args, fast=False))
        assert used and all(how == "columns" for _, how in used)
        # irregular variants of case 0: each must fall back for that file only and still agree
        exp = load_json("cluster_case0_expected.json")
        a = exp["args"]
        src = os.path.join(GOLDEN, "cluster_case0_tiddit")
        sample = a["samples"][0]
        edits = {"splits": lambda ls: ls[:3] + [ls[3].rstrip("\n") + "\t" + "\t".join(ls[1].rstrip("\n").split("\t")[3:]) + "\n"] + ls[4:],
                 "discordants": lambda ls: [ls[0].rstrip("\n") + "  \n"] + ls[1:]}
        for stem, edit in edits.items():
            dst = str(tmp_path / ("irr_" + stem))
            shutil.copytree(src, dst + "_tiddit")
            path = os.path.join(dst + "_tiddit", "%s_%s.tab" % (stem, sample))
            lines = open(path).readlines()
            open(path, "w").writelines(edit(lines))
            del used[:]
            args = (dst, a["chromosomes"], a["contig_length"], a["samples"], a["is_mp"], a["min_contig"], a["skip_assembly"])
            same(PackedSignals.from_tab(*args, fast="columns"), PackedSignals.from_tab(*args, fast=False))
            same(PackedSignals.from_tab(*args, fast=True), PackedSignals.from_tab(*args, fast=False))
            if stem == "discordants":      # (the C parser takes longer split lines as they are: extra fields are unused)
                assert ("%s_%s.tab" % (stem, sample), "lines") in used
                assert PackedSignals._from_tab_native(*args) is None      # blank-padded field: not taken natively
            else:
                assert PackedSignals._from_tab_native(*args) is not None   # longer split lines are regular for the scanner
    finally:
        signals._part_from_columns = orig


def test_tab_scanner_irregular_and_errors(tmp_path):
    """libtdt_tab.so: what it must refuse (the line reader then decides) and what it must report like the reference."""
    from tiddit_b200 import tabio
    from tiddit_b200.signals import PackedSignals
    good_d = "r1\tc1\tc2\t10\t110\tFalse\t500\t600\tTrue\n"
    good_s = "r2\tc1\tc1\t40\tTrue\t90\tFalse\t1\t40\t90\t130\n"

    def parse(text, kind):
        path = str(tmp_path / "x.tab")
        with open(path, "w", newline="") as f:
            f.write(text)
        ts = tabio.TabSet()
        try:
            lo, hi = ts.parse(path, kind)
            return hi - lo, ts.col_i64(0).tolist(), ts.table(0), ts.table(1), ts.table(2)
        finally:
            ts.close()

    assert parse(good_d * 3, "discordants") == (3, [10, 10, 10], ["r1"], ["c1", "c2"], ["False", "True"])
    assert parse(good_d.rstrip("\n"), "discordants")[0] == 1                      # no newline at the end of the file
    assert parse("", "discordants")[0] == 0
    assert parse(good_s + good_s.rstrip("\n") + "\tx\ty\n", "splits")[0] == 2       # further fields are ignored
    for bad in (good_d + "\n", good_d.replace("\n", "\r\n"), good_d.replace("10\t", "1e1\t", 1), good_d.replace("r1", " r1"),
                good_d.replace("c2", ""), good_d.rstrip("\n") + "\textra\n", "\t".join(good_d.split("\t")[:8]) + "\n",
                good_d.replace("110", "1_10"), good_d.replace("500", "99999999999999999999")):
        with pytest.raises(tabio.IrregularTab):
            parse(bad, "discordants")
    with pytest.raises(tabio.IrregularTab):
        parse("\t".join(good_s.split("\t")[:10]) + "\n", "splits")
    with pytest.raises(OSError):
        tabio.TabSet().parse(str(tmp_path / "missing.tab"), "splits")
    # an unknown contig is the reference's KeyError (contig_length[chrA], tiddit_cluster.pyx:52)
    os.makedirs(str(tmp_path / "k_tiddit"))
    for stem, text in (("discordants", good_d), ("splits", good_s)):
        with open(str(tmp_path / "k_tiddit" / ("%s_S.tab" % stem)), "w") as f:
            f.write(text)
    with pytest.raises(KeyError):
        PackedSignals.from_tab(str(tmp_path / "k"), ["c1"], {"c1": 1000}, ["S"], False, 0, True)
    pk = PackedSignals.from_tab(str(tmp_path / "k"), ["c1", "c2"], {"c1": 1000, "c2": 550}, ["S"], False, 0, True)
    assert len(pk) == 2 and pk.pairs == [("c1", "c1"), ("c1", "c2")] and pk.posB.tolist() == [90, 500]
    with pytest.raises(FileNotFoundError):
        PackedSignals.from_tab(str(tmp_path / "nothing"), ["c1"], {"c1": 1000}, ["S"], False, 0, True)


def test_tab_scanner_parallel_interning(tmp_path):
    """libtdt_tab.so interns on all threads (per-chunk tables for contigs / orientations merged in file order, a
    hash-sharded table for the read names): ids must come out in order of FIRST APPEARANCE whatever the thread count --
    names repeated inside a file, across chunks and across the files of one set; contigs that first appear deep inside
    a later chunk -- and equal the line reader's (tiddit_cluster.pyx:47-137 restated in signals._part_from_lines)."""
    from tiddit_b200 import tabio
    from tiddit_b200.signals import PackedSignals
    rng = np.random.default_rng(5)
    contigs = ["chr%d" % i for i in range(1, 30)]
    clen = {c: 10_000_000 for c in contigs}
    nd, ns = 120_000, 60_000
    names_d = rng.integers(0, 40_000, nd)                       # every name about three times
    names_s = rng.integers(20_000, 70_000, ns)                  # half of them known from the discordants file
    ca = np.sort(rng.integers(0, 6, nd))                        # few contigs first, the others first appear late
    ca[-5000:] = rng.integers(6, len(contigs), 5000)
    cb = rng.integers(0, len(contigs), nd)
    pos = rng.integers(1, 9_000_000, (nd, 4))
    ori = np.array(["False", "True", "odd"])
    os.makedirs(str(tmp_path / "p_tiddit"))
    with open(str(tmp_path / "p_tiddit" / "discordants_S.tab"), "w") as f:
        oa, ob = rng.integers(0, 2, nd), rng.integers(0, 3, nd)
        f.write("".join("read%d\t%s\t%s\t%d\t%d\t%s\t%d\t%d\t%s\n" % (names_d[i], contigs[ca[i]], contigs[cb[i]], pos[i, 0],
                                                                     pos[i, 0] + 100, ori[oa[i]], pos[i, 1], pos[i, 1] + 100,
                                                                     ori[ob[i]]) for i in range(nd)))
    with open(str(tmp_path / "p_tiddit" / "splits_S.tab"), "w") as f:
        sa, sb = rng.integers(0, len(contigs), ns), rng.integers(0, len(contigs), ns)
        sp = rng.integers(1, 9_000_000, (ns, 2))
        f.write("".join("read%d\t%s\t%s\t%d\tTrue\t%d\tFalse\t%d\t%d\t%d\t%d\n" % (names_s[i], contigs[sa[i]], contigs[sb[i]],
                                                                                  sp[i, 0], sp[i, 1], sp[i, 0] - 5, sp[i, 0],
                                                                                  sp[i, 1], sp[i, 1] + 5) for i in range(ns)))

    def scan(threads):
        ts = tabio.TabSet()
        try:
            for stem in ("discordants", "splits"):
                ts.parse(str(tmp_path / "p_tiddit" / ("%s_S.tab" % stem)), stem, threads)
            return ([ts.col_i32(j) for j in range(5)], [ts.col_i64(j) for j in range(6)], [ts.table(t) for t in range(3)])
        finally:
            ts.close()

    one = scan(1)
    # first appearance: the id column, read in order, introduces 0, 1, 2, ... one at a time
    for j, t in ((0, 0), (1, 1), (3, 2)):
        col = one[0][j] if j == 0 else np.stack([one[0][j], one[0][j + 1]], 1).ravel()
        firsts = col[np.sort(np.unique(col, return_index=True)[1])]
        assert np.array_equal(firsts, np.arange(len(one[2][t])))
    assert one[2][0][0] == "read%d" % names_d[0] and len(one[2][0]) == len(set(names_d.tolist()) | set(names_s.tolist()))
    for threads in (2, 3, 5, 0):
        got = scan(threads)
        assert all(np.array_equal(a, b) for a, b in zip(one[0] + one[1], got[0] + got[1])) and one[2] == got[2], threads
    args = (str(tmp_path / "p"), contigs, clen, ["S"], False, 0, True)
    a, b = PackedSignals._from_tab_native(*args), PackedSignals.from_tab(*args, fast=False)
    assert a is not None and list(a.names) == list(b.names) and a.ori_table == b.ori_table and a.pairs == b.pairs
    for k in PackedSignals.FIELDS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_tab_header_symbols_exported():
    import re
    from tiddit_b200 import build, tabio
    build.build_tab()
    text = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "tdt_tab.h")).read()
    names = sorted(set(re.findall(r"TDT_TAB_API[^;(]*?\b(tdt_tab_[a-z0-9_]+)\s*\(", text)))
    assert names == sorted(tabio.TAB_SIGNATURES)
    L = tabio.tab_lib()
    for n in names:
        assert hasattr(L, n)
