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


def plan_chunks(seg_off, n_chunks, taper=1.0):
    """Contiguous pair ranges -> list of (p0, p1).  taper = 1: roughly equal signal counts; taper < 1: the last
    chunk gets about `taper` times the signals of the first (a short last chunk means a short pipeline tail)."""
    seg_off = np.asarray(seg_off, dtype=np.int64)
    P = len(seg_off) - 1
    n = int(seg_off[-1])
    if P <= 0:
        return []
    n_chunks = max(1, min(n_chunks, P))
    w = np.linspace(1.0, taper, n_chunks)
    frac = np.concatenate([[0.0], np.cumsum(w) / w.sum()])
    cuts = [0]
    for k in range(1, n_chunks):
        p = int(np.searchsorted(seg_off, n * frac[k], side="left"))
        p = min(max(p, cuts[-1] + 1), P - (n_chunks - k))
        cuts.append(p)
    cuts.append(P)
    return [(cuts[i], cuts[i + 1]) for i in range(n_chunks) if cuts[i + 1] > cuts[i]]


class GraphRunner:
    """One device-resident clustering call (fixed buffers, fixed shape) captured into a CUDA graph: replay() runs
    the ~40 kernels / memsets of the call without per-launch host latency.  For callers that cluster the same
    buffers repeatedly (benchmarks, streaming windows of equal size)."""

    def __init__(self, posA, posB, seg_off, P, epsilon, m, max_pos, labels_out):
        torch = _lib.torch_cuda()
        self.torch = torch
        self.status = torch.zeros(1, dtype=torch.int32, device=posA.device)
        self.labels = labels_out
        args = (posA, posB, seg_off, P, epsilon, m, max_pos)
        # the graph bakes the workspace pointer in: it must live (and stay put) as long as the graph does
        self.ws = torch.empty(device_ops.cluster_workspace_bytes(posA.numel(), P) + 4096, dtype=torch.uint8,
                              device=posA.device)
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):   # warm-up: kernel attributes, workspace
            device_ops.cluster_labels_device(*args, labels_out=labels_out, status=self.status, ws=self.ws)
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            device_ops.cluster_labels_device(*args, labels_out=labels_out, status=self.status, ws=self.ws)

    def replay(self):
        """Enqueue the captured call on the current stream; labels land in the buffer given at construction."""
        self.graph.replay()
        return self.labels

    def check(self):
        """Synchronise and raise if any replay since construction saw out-of-range coordinates."""
        device_ops.check_async_status(int(self.status.item()))


class HostPipeline:
    """Reusable buffers + streams for clustering host-resident signal sets of up to n_max signals.

    Nothing on the GPU side waits for the host: every chunk's ~40 kernels and memsets go through the asynchronous
    ABI call, captured once per chunk shape into a CUDA graph and replayed (a chunk is launch-bound otherwise);
    inputs travel on one copy stream in chunk order, labels return on another, and the host synchronises once
    at the end.  Chunks taper (the last is the smallest) so little work is left when the last input arrives."""

    def __init__(self, n_max, n_chunks=8, p_max=1 << 16, taper=0.25, use_graphs=True):
        torch = _lib.torch_cuda()
        self.torch = torch
        self.n_max = int(n_max)
        self.n_chunks = int(n_chunks)
        self.a_d = torch.empty(self.n_max, dtype=torch.int32, device="cuda")
        self.b_d = torch.empty(self.n_max, dtype=torch.int32, device="cuda")
        self.lab_d = torch.empty(self.n_max, dtype=torch.int32, device="cuda")
        self.off_pin = torch.empty(p_max + 2 * self.n_chunks + 2, dtype=torch.int64).pin_memory()
        self.off_d = torch.empty_like(self.off_pin, device="cuda")
        self.status_d = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.status_pin = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.taper = float(taper)
        self.use_graphs = bool(use_graphs)
        self._graphs = {}
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream() for _ in range(3))
        self._events = [[torch.cuda.Event() for _ in range(3)] for _ in range(self.n_chunks)]
        # private workspace: the captured chunk graphs hold its address (never the process-wide cached one)
        self.ws = torch.empty(device_ops.cluster_workspace_bytes(self.n_max, min(p_max, self.n_max + 1)) + 4096,
                              dtype=torch.uint8, device="cuda")

    def run(self, posA, posB, seg_off, epsilon, m, max_pos, out):
        """posA / posB / out: pinned CPU int32 tensors; seg_off: numpy int64 (P+1).  Returns `out` (filled when the
        call returns)."""
        torch = self.torch
        n = int(posA.numel())
        if n > self.n_max:
            raise ValueError("HostPipeline sized for %d signals, got %d" % (self.n_max, n))
        device_ops.check_min_pts(m, n)
        seg_off = np.asarray(seg_off, dtype=np.int64)
        chunks = plan_chunks(seg_off, self.n_chunks, self.taper)
        if len(seg_off) + 2 * len(chunks) > self.off_pin.numel():
            raise ValueError("HostPipeline sized for %d pairs" % (self.off_pin.numel() - 2 * self.n_chunks - 2))
        # chunk-relative offsets, back to back in one pinned buffer -> one small copy
        views, pos = [], 0
        off_np = self.off_pin.numpy()
        for p0, p1 in chunks:
            k = p1 - p0 + 1
            off_np[pos:pos + k] = seg_off[p0:p1 + 1] - seg_off[p0]
            views.append((pos, k))
            pos += k
        cur = torch.cuda.current_stream()
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        with torch.cuda.stream(self.s_in):
            self.off_d[:pos].copy_(self.off_pin[:pos], non_blocking=True)
            self.status_d.zero_()
            for k, (p0, p1) in enumerate(chunks):
                lo, hi = int(seg_off[p0]), int(seg_off[p1])
                self.a_d[lo:hi].copy_(posA[lo:hi], non_blocking=True)
                self.b_d[lo:hi].copy_(posB[lo:hi], non_blocking=True)
                self._events[k][0].record(self.s_in)
        for k, (p0, p1) in enumerate(chunks):
            lo, hi = int(seg_off[p0]), int(seg_off[p1])
            ev_in, _, ev_done = self._events[k]
            o0, ok = views[k]
            self.s_run.wait_event(ev_in)
            if hi > lo:
                self._launch_chunk(lo, hi, o0, ok, p1 - p0, epsilon, m, max_pos)
            ev_done.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_done)
                out[lo:hi].copy_(self.lab_d[lo:hi], non_blocking=True)
        with torch.cuda.stream(self.s_out):
            self.status_pin.copy_(self.status_d, non_blocking=True)
        cur.wait_stream(self.s_out)
        self.s_out.synchronize()
        device_ops.check_async_status(int(self.status_pin[0]))
        return out


    def _chunk_call(self, lo, hi, o0, ok, P, epsilon, m, max_pos):
        device_ops.cluster_labels_device(self.a_d[lo:hi], self.b_d[lo:hi], self.off_d[o0:o0 + ok], P, epsilon, m,
                                         max_pos, labels_out=self.lab_d[lo:hi], status=self.status_d, ws=self.ws)

    def _launch_chunk(self, lo, hi, o0, ok, P, epsilon, m, max_pos):
        torch = self.torch
        if not self.use_graphs:
            with torch.cuda.stream(self.s_run):
                self._chunk_call(lo, hi, o0, ok, P, epsilon, m, max_pos)
            return
        key = (lo, hi, o0, ok, P, device_ops.eps_to_int(epsilon), int(m), int(max_pos))
        graph = self._graphs.get(key)
        if graph is None:
            # first time for this chunk shape: run it eagerly (the result of this call), then capture it for later
            # calls; everything the call touches (buffers, workspace, offsets) lives as long as this object
            with torch.cuda.stream(self.s_run):
                self._chunk_call(lo, hi, o0, ok, P, epsilon, m, max_pos)
            self.s_run.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self.s_run):
                self._chunk_call(lo, hi, o0, ok, P, epsilon, m, max_pos)
            if len(self._graphs) > 256:
                self._graphs.clear()
            self._graphs[key] = graph
            return
        with torch.cuda.stream(self.s_run):
            graph.replay()


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


# ---------------------------------------------------------------------------------------------
# multi-GPU: coverage by read slices, GC by contig
# ---------------------------------------------------------------------------------------------
def read_slices(read_off, world, rank):
    """Rank `rank`'s contiguous share of every contig's reads -> list of (lo, hi) per contig.  Every contig is cut
    into `world` nearly equal slices, so the ranks stay balanced whatever the contig sizes are."""
    read_off = np.asarray(read_off, dtype=np.int64)
    out = []
    for c in range(len(read_off) - 1):
        lo, hi = int(read_off[c]), int(read_off[c + 1])
        k = hi - lo
        out.append((lo + k * rank // world, lo + k * (rank + 1) // world))
    return out


def sharded_coverage(start, end, read_off, lengths, bin_size, group=None, accumulate_fn=None):
    """Coverage bins of ALL contigs (numpy float64, contigs back to back) on every rank; each rank accumulates its
    slice of every contig's reads (tiddit_coverage.pyx:48-74 per read) and ONE all-reduce(sum) adds the partial bins.
    Every addend is a float32 quotient or 1.0 and bin totals stay far below 2^53 units of the smallest addend, so
    float64 sums are exact in ANY order: the all-reduce -- ring, tree or in-switch -- returns bins
    bit-identical to the sequential loop.

    start / end: int32 host arrays of all reads, grouped by contig (read_off[C+1]); lengths[C]: contig lengths.
    accumulate_fn(start, end, read_off, bin_off, end_bin_size, bin_size, n_bins) -> float64 bins: the per-rank
    kernel front end; defaults to the GPU path.  The gloo tests inject the oracle."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lengths = np.asarray(lengths, dtype=np.int64)
    nb = np.ceil(lengths / float(bin_size)).astype(np.int64)
    bin_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    ebs = (lengths - (nb - 1) * bin_size).astype(np.int32)
    sl = read_slices(read_off, world, rank)
    s = np.concatenate([start[lo:hi] for lo, hi in sl]) if sl else np.zeros(0, dtype=np.int32)
    e = np.concatenate([end[lo:hi] for lo, hi in sl]) if sl else np.zeros(0, dtype=np.int32)
    my_off = np.concatenate([[0], np.cumsum([hi - lo for lo, hi in sl])]).astype(np.int64)
    on_gpu = dist.get_backend(group) == "nccl"
    if accumulate_fn is None:
        accumulate_fn = _coverage_contigs_gpu
    bins = accumulate_fn(np.ascontiguousarray(s, dtype=np.int32), np.ascontiguousarray(e, dtype=np.int32), my_off, bin_off,
                         ebs, int(bin_size), int(bin_off[-1]))
    if isinstance(bins, np.ndarray):
        bins = torch.from_numpy(bins)
        if on_gpu:
            bins = bins.cuda()
    dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=group)
    return bins.cpu().numpy(), bin_off


def _coverage_contigs_gpu(s, e, read_off, bin_off, ebs, bin_size, n_bins):
    torch = _lib.torch_cuda()
    bins = torch.zeros(n_bins, dtype=torch.float64, device="cuda")
    bad = device_ops.new_first_bad(torch)
    if len(s):
        device_ops.coverage_accumulate_contigs_device(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda(),
                                                      torch.from_numpy(read_off).cuda(), torch.from_numpy(bin_off).cuda(),
                                                      torch.from_numpy(ebs).cuda(), bin_size, bins, bad)
    if int(bad.item()) != device_ops.FIRST_BAD_NONE:
        raise IndexError("Out of bounds on buffer access (axis 0)")
    return bins


def sharded_gc(sequences, bin_size, n_cutoff, group=None, gc_fn=None):
    """GC bins of ALL contigs ({name: int8 ndarray}) on every rank; contigs are dealt to ranks longest-first and the
    int8 bins all-gathered once (tiddit_gc.pyx:35-42 fans contigs out over joblib processes the same way).
    sequences: {name: uint8 array / bytes}; gc_fn(seq, bin_size, n_cutoff) -> int8 bins defaults to the GPU path."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    names = list(sequences)
    lens = [len(sequences[n]) for n in names]
    owner = lpt_assign(lens, world)
    nb = [int(-(-ln // bin_size)) for ln in lens]
    mine = [i for i in range(len(names)) if owner[i] == rank]
    per_rank = [sum(nb[i] for i in range(len(names)) if owner[i] == r) for r in range(world)]
    pad = max(per_rank) if per_rank else 0
    if gc_fn is None:
        from . import tiddit_gc
        gc_fn = tiddit_gc.gc_bins
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    send = torch.zeros(max(pad, 1), dtype=torch.int8, device=dev)
    pos = 0
    for i in mine:
        if nb[i]:
            send[pos:pos + nb[i]] = torch.from_numpy(np.asarray(gc_fn(sequences[names[i]], bin_size, n_cutoff), dtype=np.int8)).to(dev)
        pos += nb[i]
    recv = torch.empty(world * max(pad, 1), dtype=torch.int8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy()
    out, cursor = {}, [r * max(pad, 1) for r in range(world)]
    for i, name in enumerate(names):
        r = int(owner[i])
        out[name] = recv[cursor[r]:cursor[r] + nb[i]].copy()
        cursor[r] += nb[i]
    return out
