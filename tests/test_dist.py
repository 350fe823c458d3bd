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
