"""tiddit_b200 -- B200 (sm_100a) implementation of TIDDIT's signal-clustering / coverage / GC hot path.

The package mirrors the reference's module-level call surface for that path:

    tiddit_b200.DBSCAN            <- tiddit/DBSCAN.py
    tiddit_b200.tiddit_cluster    <- tiddit/tiddit_cluster.pyx
    tiddit_b200.tiddit_coverage   <- tiddit/tiddit_coverage.pyx
    tiddit_b200.tiddit_gc         <- tiddit/tiddit_gc.pyx
    tiddit_b200.tiddit_coverage_analysis <- tiddit/tiddit_coverage_analysis.pyx   

plus signals.PackedSignals (the records between tiddit_signal and tiddit_cluster as arrays), engine
(host pipeline, CUDA-graph runner, multi-GPU sharding) and bamio / fasta (file access without pysam).

Every compute entry point goes through the C ABI of libtdt_b200.so (include/tdt_b200.h); there is
no CPU fallback: without the library or without a CUDA device the calls raise.
"""
__version__ = "0.1.0"
