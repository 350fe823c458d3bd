"""Drop-in for tiddit/tiddit_gc.pyx (binned_gc :6-33, main :35-42) with the counting on the GPU.

`threads` is accepted for signature compatibility; contigs are processed back to back on the
current CUDA stream (the reference fans them out over joblib processes, tiddit_gc.pyx:36)."""
import numpy as np

from . import _lib, device_ops, fasta

__all__ = ["binned_gc", "main", "gc_bins"]


def gc_bins(seq, bin_size, n_cutoff):
    """int8 GC bins of one in-memory sequence (bytes / str / uint8 array)."""
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    dev, n = device_ops.padded_sequence_device(seq)
    return device_ops.gc_bins_device(dev, n, bin_size, n_cutoff).cpu().numpy()


def binned_gc(fasta_path, contig, bin_size, n_cutoff):
    """tiddit_gc.pyx:6-33 -> [contig, int8 ndarray]."""
    _lib.torch_cuda()
    fa = fasta.open_fasta(fasta_path)
    return [contig, gc_bins(fa.fetch_bytes(contig), bin_size, n_cutoff)]


def main(reference, contigs, threads, bin_size, n_cutoff):
    """tiddit_gc.pyx:35-42 -> {contig: int8 ndarray}."""
    gc_dictionary = {}
    for contig in contigs:
        name, bins = binned_gc(reference, contig, bin_size, n_cutoff)
        gc_dictionary[name] = bins
    return gc_dictionary
