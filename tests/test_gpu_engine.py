"""GPU: the host pipeline (chunked H2D / kernels / D2H) and the NCCL-sharded path give the oracle's labels."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunks", [1, 3, 8])
def test_host_pipeline_matches_oracle(chunks, oracle):
    import torch
    from tiddit_b200 import engine, synth
    a, b, off, L = synth.wgs30x_signals(2_000_000)
    want = oracle.cluster_segments(a, b, off, 500, 3)
    pipe = engine.HostPipeline(len(a), n_chunks=chunks)
    a_pin, b_pin = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    out = torch.empty(len(a), dtype=torch.int32).pin_memory()
    for _ in range(2):                       # buffers are reused across calls
        out.fill_(-7)
        pipe.run(a_pin, b_pin, off, 500, 3, L, out)
        assert np.array_equal(out.numpy(), want)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tiddit_b200 import engine, synth
    a, b, off, L = synth.wgs30x_signals(1_000_000)
    got = engine.sharded_labels(a, b, off, 500, 3, L)
    np.save(os.path.join(out_dir, "labels_%d.npy" % rank), got)
    dist.destroy_process_group()


def test_sharded_labels_nccl(tmp_path, oracle):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from tiddit_b200 import synth
    a, b, off, L = synth.wgs30x_signals(1_000_000)
    want = oracle.cluster_segments(a, b, off, 500, 3)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("labels_%d.npy" % r)), want)


def test_host_pipeline_reports_range_errors():
    import torch
    from tiddit_b200 import engine, _lib
    a = np.array([5, 900, 7, 8], dtype=np.int32)
    b = np.array([1, 2, 3, 4], dtype=np.int32)
    pipe = engine.HostPipeline(4, n_chunks=2)
    out = torch.empty(4, dtype=torch.int32).pin_memory()
    with pytest.raises(_lib.TdtError) as e:
        pipe.run(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory(), np.array([0, 2, 4]), 10, 2, 100, out)
    assert e.value.code == _lib.TDT_E_RANGE
    pipe.run(torch.from_numpy(b).pin_memory(), torch.from_numpy(b).pin_memory(), np.array([0, 2, 4]), 10, 2, 100, out)
    assert out.tolist() == [0, 0, 0, 0]


def _nccl_cov_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tiddit_b200 import engine, synth
    contigs = synth.GRCH38[16:22]
    s, e, roff, lens = synth.coverage_reads(3_000_000, contigs=contigs)
    bins, _ = engine.sharded_coverage(s, e, roff, lens, 500)
    np.save(os.path.join(out_dir, "bins_%d.npy" % rank), bins)
    seqs = {"c%d" % i: synth.fasta_sequence(400_000 + 133_337 * i, seed=i) for i in range(5)}
    np.savez(os.path.join(out_dir, "gc_%d.npz" % rank), **engine.sharded_gc(seqs, 50, 0.5))
    dist.destroy_process_group()


def test_sharded_coverage_and_gc_nccl(tmp_path, oracle):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_nccl_cov_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from tiddit_b200 import synth
    contigs = synth.GRCH38[16:22]
    s, e, roff, lens = synth.coverage_reads(3_000_000, contigs=contigs)
    want = []
    for c, ln in enumerate(lens):
        nb = int(np.ceil(ln / 500.0))
        bins = np.zeros(nb)
        oracle.update_coverage_batch(s[roff[c]:roff[c + 1]], e[roff[c]:roff[c + 1]], 500, bins, int(ln - (nb - 1) * 500))
        want.append(bins)
    want = np.concatenate(want)
    seqs = {"c%d" % i: synth.fasta_sequence(400_000 + 133_337 * i, seed=i) for i in range(5)}
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("bins_%d.npy" % r)).view(np.uint64), want.view(np.uint64))
        gc = np.load(tmp_path / ("gc_%d.npz" % r))
        for name, seq in seqs.items():
            assert np.array_equal(gc[name], oracle.gc_bins(seq, 50, 0.5))
