"""CPU, world_size 2 over gloo: the pair-sharding logic of the multi-GPU path (LPT assignment, shard
extraction, rank-major all-gather, reassembly in input order).  The per-rank labeller is injected (the
oracle) because this container has no GPU; on the GPU box the same function runs with the CUDA labeller."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from tiddit_b200 import engine, synth
    a, b, off, L = synth.wgs30x_signals(120_000)
    got = engine.sharded_labels(a, b, off, 500, 3, L, label_fn=lambda pa, pb, so, e, m, mp_: oracle.cluster_segments(pa, pb, so, e, m))
    np.save(os.path.join(out_dir, "labels_%d.npy" % rank), got)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_labels_gloo(world, tmp_path, oracle):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from tiddit_b200 import synth
    a, b, off, L = synth.wgs30x_signals(120_000)
    want = oracle.cluster_segments(a, b, off, 500, 3)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("labels_%d.npy" % r)), want)


def test_shard_plan_balanced():
    from tiddit_b200 import engine, synth
    _, _, off, _ = synth.wgs30x_signals(1_000_000)
    for world in (1, 2, 4, 8):
        plan = engine.ShardPlan(off, world)
        assert sum(plan.counts) == off[-1]
        assert max(plan.counts) <= 1.15 * off[-1] / world + 1          # LPT keeps the shards within 15 %
        where = plan.gather_index()
        assert len(np.unique(where)) == off[-1]
        assert sorted(np.concatenate(plan.pairs).tolist()) == list(range(len(off) - 1))


def test_plan_chunks():
    from tiddit_b200 import engine
    off = np.array([0, 10, 10, 500, 520, 1000, 1000, 1001])
    for k in (1, 2, 3, 7, 20):
        ch = engine.plan_chunks(off, k)
        assert ch[0][0] == 0 and ch[-1][1] == len(off) - 1
        assert all(ch[i][1] == ch[i + 1][0] for i in range(len(ch) - 1))
    assert engine.plan_chunks(np.array([0]), 4) == []


def _cov_gc_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from tiddit_b200 import engine, synth
    contigs = synth.GRCH38[18:22]
    s, e, roff, lens = synth.coverage_reads(300_000, contigs=contigs)

    def acc(ss, ee, my_off, bin_off, ebs, z, n_bins):      # the oracle as the per-rank accumulator (no GPU here)
        bins = np.zeros(n_bins)
        for c in range(len(my_off) - 1):
            oracle.update_coverage_batch(ss[my_off[c]:my_off[c + 1]], ee[my_off[c]:my_off[c + 1]], z,
                                         bins[bin_off[c]:bin_off[c + 1]], int(ebs[c]))
        return bins
    bins, bin_off = engine.sharded_coverage(s, e, roff, lens, 500, accumulate_fn=acc)
    np.save(os.path.join(out_dir, "bins_%d.npy" % rank), bins)
    seqs = {"c%d" % i: synth.fasta_sequence(40_000 + 13_337 * i, seed=i) for i in range(5)}
    gc = engine.sharded_gc(seqs, 50, 0.5, gc_fn=oracle.gc_bins)
    np.savez(os.path.join(out_dir, "gc_%d.npz" % rank), **gc)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_coverage_and_gc_gloo(world, tmp_path, oracle):
    """Read-slice sharding + all-reduce of exact float64 bins, contig sharding + all-gather of GC bins: bit-identical
    to the single-process result on every rank."""
    mp.spawn(_cov_gc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from tiddit_b200 import synth
    contigs = synth.GRCH38[18:22]
    s, e, roff, lens = synth.coverage_reads(300_000, contigs=contigs)
    want = []
    for c, ln in enumerate(lens):
        nb = int(np.ceil(ln / 500.0))
        bins = np.zeros(nb)
        oracle.update_coverage_batch(s[roff[c]:roff[c + 1]], e[roff[c]:roff[c + 1]], 500, bins, int(ln - (nb - 1) * 500))
        want.append(bins)
    want = np.concatenate(want)
    seqs = {"c%d" % i: synth.fasta_sequence(40_000 + 13_337 * i, seed=i) for i in range(5)}
    for r in range(world):
        got = np.load(tmp_path / ("bins_%d.npy" % r))
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        gc = np.load(tmp_path / ("gc_%d.npz" % r))
        assert list(gc) == list(seqs)
        for name, seq in seqs.items():
            assert np.array_equal(gc[name], oracle.gc_bins(seq, 50, 0.5))
