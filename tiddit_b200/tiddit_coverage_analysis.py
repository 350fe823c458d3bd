"""Drop-in for tiddit/tiddit_coverage_analysis.pyx (determine_ploidy :9-41) with the bin loops on the GPU.

The reference walks every coverage bin in Python (61.8 M bins at bin size 50), keeps those with coverage > 0 and
GC != -1 (:17-22) and takes numpy.median per contig and genome-wide (:24-27).  Here the coverage and GC arrays go to
HBM back to back and ONE call (device_ops.coverage_medians -> tdt_coverage_medians) returns all medians, exact to
the bit; the per-contig ploidy arithmetic and the `<prefix>.ploidies.tab` text (:29-38) stay on the host, written
with the same expressions so the file is byte-identical.
"""
import numpy as np

from . import device_ops

__all__ = ["determine_ploidy", "masked_medians"]


def masked_medians(coverage_data, gc):
    """{contig: float64 bins}, {contig: int8 bins} -> ({contig: median of bins with cov > 0 and gc != -1}, genome-wide
    median) as numpy.float64 (nan where nothing qualifies), in the iteration order of coverage_data."""
    names = list(coverage_data)
    covs, gcs = [], []
    for name in names:
        cov = np.ascontiguousarray(coverage_data[name], dtype=np.float64)
        g = np.asarray(gc[name])
        if len(g) < len(cov):
            # gc[chromosome][i] for i in range(len(coverage)) (:17-18)
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (len(g), len(g)))
        covs.append(cov)
        gcs.append(np.ascontiguousarray(g[:len(cov)], dtype=np.int8))
    bin_off = np.concatenate([[0], np.cumsum([len(c) for c in covs])]).astype(np.int64)
    med, _ = device_ops.coverage_medians(np.concatenate(covs) if covs else np.zeros(0), np.concatenate(gcs) if gcs
                                         else np.zeros(0, dtype=np.int8), bin_off)
    return {name: med[i] for i, name in enumerate(names)}, med[len(names)]


def determine_ploidy(coverage_data, contigs, library, ploidy, prefix, c, reference_fasta, bin_size, bam_header, gc):
    """tiddit_coverage_analysis.pyx:9-41 -> library (avg_coverage_<contig>, avg_coverage, contig_ploidy_<contig>) and
    `<prefix>.ploidies.tab`."""
    f = open("{}.ploidies.tab".format(prefix), "w")
    f.write("Chromosome\tPloidy\tPloidy_rounded\tMean_coverage\n")
    per_contig, genome = masked_medians(coverage_data, gc)
    for chromosome in coverage_data:
        library["avg_coverage_{}".format(chromosome)] = per_contig[chromosome]
        if np.isnan(library["avg_coverage_{}".format(chromosome)]):
            library["avg_coverage_{}".format(chromosome)] = 0

    if not c:
        library["avg_coverage"] = genome
    else:
        library["avg_coverage"] = c

    for chromosome in contigs:
        if chromosome not in coverage_data:
            continue
        avg_coverage_contig = library["avg_coverage_{}".format(chromosome)]
        library["contig_ploidy_{}".format(chromosome)] = int(round(ploidy * avg_coverage_contig / library["avg_coverage"]))
        f.write("{}\t{}\t{}\t{}\n".format(chromosome, avg_coverage_contig / library["avg_coverage"] * ploidy,
                                          library["contig_ploidy_{}".format(chromosome)], avg_coverage_contig))
    f.close()
    return library
