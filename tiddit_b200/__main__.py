"""`python -m tiddit_b200 --cov ...`: the reference's coverage command (tiddit/__main__.py:210-247) on the GPU path.

Same flags, same output files (byte-identical bed / wig).  Reads are decoded on the host by libtdt_bam.so (BGZF inflated on all cores, records as
columns) and accumulated batch by batch by the coverage kernel.
`--sv` needs the reference's own stages (signal extraction, assembly, variant calling, all built on pysam and bwa);
when the `tiddit` package is importable it is run with this package's clustering / coverage / GC modules swapped in.
"""
import argparse
import os
import sys


def coverage_from_bam(path, bin_size, min_q, threads=0, batch_reads=1 << 20):
    """The reference's read loop (tiddit/__main__.py:225-242) on whole batches: libtdt_bam.so decodes the records into
    columns, the (start, end) columns of the reads that count (mapped, not duplicate, mapq >= min_q) go to the coverage
    kernel contig by contig.  -> ({contig: float64 bins}, header)"""
    import numpy as np
    from . import bamio, tiddit_coverage
    with bamio.ColumnReader(path, threads=threads, batch_reads=batch_reads,
                            columns=("ref_id", "pos", "end", "flag", "mapq")) as reader:
        bam_header = reader.header
        cov = tiddit_coverage.DeviceCoverage(bam_header, bin_size)
        for b in reader.batches():
            keep = np.flatnonzero(((b.flag & (0x4 | 0x400)) == 0) & (b.mapq >= min_q))
            if not len(keep):
                continue
            if np.any(b.ref_id[keep] < 0) or np.any(b.end[keep] < 0):
                raise TypeError("a mapped read without reference or CIGAR")   # the reference fails on these too
            rid = b.ref_id[keep]
            if np.any(rid[1:] < rid[:-1]):
                keep = keep[np.argsort(rid, kind="stable")]
                rid = b.ref_id[keep]
            cuts = np.flatnonzero(rid[1:] != rid[:-1]) + 1
            for lo, hi in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [len(keep)]])):
                sel = keep[lo:hi]
                cov.add_reads(reader.references[rid[lo]], b.pos[sel], b.end[sel])
    coverage_data, _ = cov.to_host()
    return coverage_data, bam_header


def run_cov(argv):
    parser = argparse.ArgumentParser("""tiddit --cov --bam inputfile [-o prefix]""")
    parser.add_argument('--cov', help="generate a coverage bed/wig file", required=False, action="store_true")
    parser.add_argument('--bam', type=str, required=True, help="coordinate sorted bam file(required)")
    parser.add_argument('-o', type=str, default="output", help="output prefix(default=output)")
    parser.add_argument('-z', type=int, default=500, help="use bins of specified size(default = 500bp) to measure the coverage of the entire bam file, set output to stdout to print to stdout")
    parser.add_argument('-w', help="generate wig instead of bed", required=False, action="store_true")
    parser.add_argument('-q', type=int, help="minimum mapping quality(default=20)", required=False, default=20)
    parser.add_argument('--ref', type=str, help="reference fasta, used for reading cram")
    args = parser.parse_args(argv)
    if not os.path.isfile(args.bam):
        print("error,  could not find the bam file")
        return 1
    from . import bamio, tiddit_coverage
    try:
        bamio.require_bam(args.bam, args.ref)
    except (NotImplementedError, ValueError) as exc:
        print("error, %s" % exc)
        return 1
    coverage_data, bam_header = coverage_from_bam(args.bam, args.z, args.q)
    if args.w:
        tiddit_coverage.print_coverage(coverage_data, bam_header, args.z, "wig", args.o + ".wig")
    else:
        tiddit_coverage.print_coverage(coverage_data, bam_header, args.z, "bed", args.o + ".bed")
    return 0


SWAPPED_MODULES = ("DBSCAN", "tiddit_cluster", "tiddit_coverage", "tiddit_gc", "tiddit_coverage_analysis",
                   "tiddit_signal")


def install_gpu_modules():
    """Put this package's modules in place of the reference's BEFORE anything of the reference pipeline is imported:
    `import tiddit.DBSCAN` / `from tiddit import tiddit_signal` inside tiddit_contig_analysis, tiddit_variant and
    tiddit/__main__ then bind to the GPU modules at their own import time (patching attributes afterwards would leave
    the bindings those modules made at import pointing at the originals).  -> the `tiddit` package."""
    import importlib
    tiddit = importlib.import_module("tiddit")        # the package itself: tiddit/__init__.py imports nothing
    pkg = importlib.import_module(__package__)
    for name in SWAPPED_MODULES:
        mod = importlib.import_module("." + name, __package__)
        sys.modules["tiddit." + name] = mod
        setattr(tiddit, name, mod)
    del pkg
    return tiddit


def run_sv(argv):
    """NEVER EXECUTED in this image (no pysam / bwa / reference package importable): only the `rc 2` branch below is
    covered by tests/test_cli.py; see INTEGRATION.md."""
    try:
        install_gpu_modules()
        import tiddit.__main__ as ref_main   # needs pysam, bwa, ... (not in this image)
    except Exception as exc:
        for name in SWAPPED_MODULES:
            sys.modules.pop("tiddit." + name, None)
        print("tiddit_b200 replaces the clustering / coverage / GC stages only; --sv also needs the reference "
              "package (pysam, bwa) for signal extraction, assembly and variant calling: %s" % exc)
        return 2
    sys.argv = ["tiddit"] + list(argv)
    ref_main.main()
    return 0


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    parser = argparse.ArgumentParser("""tiddit_b200: TIDDIT's clustering / coverage hot path on B200""", add_help=False)
    parser.add_argument('--sv', action="store_true")
    parser.add_argument('--cov', action="store_true")
    args, _ = parser.parse_known_args(argv)
    if args.cov:
        return run_cov(argv)
    if args.sv:
        return run_sv(argv)
    print("usage: python -m tiddit_b200 --cov --bam inputfile [-o prefix] [-z bin] [-q mapq] [-w]\n"
          "       python -m tiddit_b200 --sv ...   (delegates to the reference pipeline with the GPU modules swapped in)")
    return 0


if __name__ == '__main__':
    sys.exit(main())
