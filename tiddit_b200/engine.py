"""Host-facing engines around the clustering ABI call.

`HostPipeline`   pinned host arrays in, pinned host labels out: the (chrA,chrB) pairs are cut into a few
                 contiguous chunks and the H2D copy of chunk k+1, the kernels of chunk k and the D2H copy of
                 chunk k-1 run on three CUDA streams, so PCIe (both directions) and the SMs overlap.
`sharded_labels` one process per GPU (torch.distributed): pairs are dealt to ranks longest-first (LPT), every
                 rank labels its own pairs, ONE all-gather (NCCL over NVLink on the GPU box, gloo in the CPU
                 tests) returns all labels to every rank in input order.
"""
import numpy as np

from . import _lib, device_ops


def plan_chunks(seg_off, n_chunks):
    """Contiguous pair ranges with roughly equal signal counts -> list of (p0, p1)."""
    seg_off = np.asarray(seg_off, dtype=np.int64)
    P = len(seg_off) - 1
    n = int(seg_off[-1])
    if P <= 0:
        return []
    n_chunks = max(1, min(n_chunks, P))
    cuts = [0]
    for k in range(1, n_chunks):
        p = int(np.searchsorted(seg_off, n * k / n_chunks, side="left"))
        p = min(max(p, cuts[-1] + 1), P - (n_chunks - k))
        cuts.append(p)
    cuts.append(P)
    return [(cuts[i], cuts[i + 1]) for i in range(n_chunks) if cuts[i + 1] > cuts[i]]


class HostPipeline:
    """Reusable buffers + streams for clustering host-resident signal sets of up to n_max signals."""

    def __init__(self, n_max, n_chunks=8):
        torch = _lib.torch_cuda()
        self.torch = torch
        self.n_max = int(n_max)
        self.n_chunks = int(n_chunks)
        self.a_d = torch.empty(self.n_max, dtype=torch.int32, device="cuda")
        self.b_d = torch.empty(self.n_max, dtype=torch.int32, device="cuda")
        self.lab_d = torch.empty(self.n_max, dtype=torch.int32, device="cuda")
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream() for _ in range(3))

    def run(self, posA, posB, seg_off, epsilon, m, max_pos, out):
        """posA / posB / out: pinned CPU int32 tensors; seg_off: numpy int64 (P+1).  Returns `out` (filled when the
        call returns)."""
        torch = self.torch
        n = int(posA.numel())
        if n > self.n_max:
            raise ValueError("HostPipeline sized for %d signals, got %d" % (self.n_max, n))
        device_ops.check_min_pts(m, n)
        seg_off = np.asarray(seg_off, dtype=np.int64)
        chunks = plan_chunks(seg_off, self.n_chunks)
        offs = [torch.from_numpy(seg_off[p0:p1 + 1] - seg_off[p0]).pin_memory() for p0, p1 in chunks]
        cur = torch.cuda.current_stream()
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        ready, done, offs_d = [], [], []
        with torch.cuda.stream(self.s_in):
            for (p0, p1), off in zip(chunks, offs):
                lo, hi = int(seg_off[p0]), int(seg_off[p1])
                self.a_d[lo:hi].copy_(posA[lo:hi], non_blocking=True)
                self.b_d[lo:hi].copy_(posB[lo:hi], non_blocking=True)
                offs_d.append(off.cuda(non_blocking=True))
                ev = torch.cuda.Event()
                ev.record(self.s_in)
                ready.append(ev)
        for k, (p0, p1) in enumerate(chunks):
            lo, hi = int(seg_off[p0]), int(seg_off[p1])
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ready[k])
                if hi > lo:
                    device_ops.cluster_labels_device(self.a_d[lo:hi], self.b_d[lo:hi], offs_d[k], p1 - p0, epsilon, m,
                                                     max_pos, labels_out=self.lab_d[lo:hi])
                ev = torch.cuda.Event()
                ev.record(self.s_run)
                done.append(ev)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done[k])
                out[lo:hi].copy_(self.lab_d[lo:hi], non_blocking=True)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_run)
        self.s_out.synchronize()
        return out


# ---------------------------------------------------------------------------------------------
# multi-GPU: pairs sharded over ranks
# ---------------------------------------------------------------------------------------------
def lpt_assign(sizes, n_ranks):
    """Longest-processing-time greedy: pair -> rank, balancing the signal counts (deterministic)."""
    load = [0] * n_ranks
    owner = np.zeros(len(sizes), dtype=np.int64)
    for p in np.argsort(-np.asarray(sizes), kind="stable"):
        r = int(np.argmin(load))
        owner[p] = r
        load[r] += int(sizes[p])
    return owner


class ShardPlan:
    """Who owns which pair, and where each rank's signals sit in the rank-major gathered buffer."""

    def __init__(self, seg_off, world):
        seg_off = np.asarray(seg_off, dtype=np.int64)
        self.seg_off = seg_off
        self.world = world
        sizes = np.diff(seg_off)
        self.owner = lpt_assign(sizes, world)
        self.pairs = [np.flatnonzero(self.owner == r) for r in range(world)]
        self.counts = [int(sizes[p].sum()) for p in self.pairs]
        self.pad = max(self.counts) if self.counts else 0

    def shard_index(self, rank):
        """Input positions of the signals `rank` owns, pair order kept."""
        mine = self.pairs[rank]
        if len(mine) == 0:
            return np.zeros(0, dtype=np.int64)
        return np.concatenate([np.arange(self.seg_off[p], self.seg_off[p + 1]) for p in mine])

    def shard_seg_off(self, rank):
        sizes = np.diff(self.seg_off)[self.pairs[rank]]
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

    def gather_index(self):
        """Position in the rank-major padded buffer of every input signal (int64, n)."""
        n = int(self.seg_off[-1])
        where = np.empty(n, dtype=np.int64)
        for r in range(self.world):
            idx = self.shard_index(r)
            where[idx] = r * self.pad + np.arange(len(idx))
        return where


def sharded_labels(posA, posB, seg_off, epsilon, m, max_pos=0, group=None, label_fn=None):
    """Labels of ALL signals (numpy int32, input order) on every rank; each rank computes only its own pairs.

    posA / posB / seg_off: the full host arrays (every rank holds them, as every rank parses the same tab files).
    label_fn(posA, posB, seg_off, eps, m, max_pos) -> int32 labels: the per-rank labeller; defaults to the GPU
    path (device_ops.cluster_labels).  The CPU (gloo) tests pass the oracle here to exercise the sharding logic."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    plan = ShardPlan(seg_off, world)
    idx = plan.shard_index(rank)
    off = plan.shard_seg_off(rank)
    fn = label_fn or device_ops.cluster_labels
    mine = fn(np.ascontiguousarray(posA[idx]), np.ascontiguousarray(posB[idx]), off, epsilon, m, max_pos)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    send = torch.full((plan.pad,), -1, dtype=torch.int32, device=dev)
    send[:len(idx)] = torch.from_numpy(np.asarray(mine, dtype=np.int32)).to(dev)
    recv = torch.empty(world * plan.pad, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv.cpu().numpy()[plan.gather_index()]
