"""Drop-in for tiddit/tiddit_gc.pyx (binned_gc :6-33, main :35-42) with the counting on the GPU.

`threads` is accepted for signature compatibility; contigs are queued back to back on the
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
    """tiddit_gc.pyx:35-42 -> {contig: int8 ndarray}.

    The FASTA is opened once (fasta.open_fasta caches it); every contig's upload and kernel are queued on the current
    stream without waiting for the previous contig's result -- the kernel of contig k runs while the host stages contig
    k + 1 -- and the int8 bins of all contigs come back through pinned buffers after ONE synchronisation.  `threads`
    (the reference's joblib fan-out, tiddit_gc.pyx:36) has nothing to do here and is accepted for compatibility."""
    torch = _lib.torch_cuda()
    fa = fasta.open_fasta(reference)
    pending = []
    for contig in contigs:
        dev, n = device_ops.padded_sequence_device(fa.fetch_bytes(contig))
        bins = device_ops.gc_bins_device(dev, n, bin_size, n_cutoff)
        host = torch.empty(bins.shape, dtype=torch.int8).pin_memory()
        host.copy_(bins, non_blocking=True)
        pending.append((contig, host, bins))   # `bins` stays referenced until the copy has run
    torch.cuda.current_stream().synchronize()
    return {contig: host.numpy().copy() for contig, host, _ in pending}
