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
    ABI call; inputs travel on one copy stream in chunk order, labels return on another, and the host synchronises
    once at the end.  Repeated calls on the same buffers replay ONE CUDA graph of the whole call (see run).  Chunks
    taper (the last is the smallest) so little work is left when the last input arrives."""

    def __init__(self, n_max, n_chunks=8, p_max=1 << 16, taper=0.25, use_graphs=True, small_shard=5_000_000):
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
        self.small_shard = int(small_shard)   # below this many signals two chunks are enough (less fixed cost)
        self._graphs, self._seen = {}, set()
        self.s_in, self.s_run, self.s_out, self.s_main = (torch.cuda.Stream() for _ in range(4))
        self._events = [[torch.cuda.Event() for _ in range(3)] for _ in range(self.n_chunks)]
        # private workspace: the captured chunk graphs hold its address (never the process-wide cached one)
        self.ws = torch.empty(device_ops.cluster_workspace_bytes(self.n_max, min(p_max, self.n_max + 1)) + 4096,
                              dtype=torch.uint8, device="cuda")

    def run(self, posA, posB, seg_off, epsilon, m, max_pos, out):
        """posA / posB / out: pinned CPU int32 tensors; seg_off: numpy int64 (P+1).  Returns `out` (filled when the
        call returns).

        The whole call -- offsets H2D, per chunk: input H2D | ~40 kernels | labels D2H on three streams, status D2H --
        is ONE CUDA graph with memcpy nodes once the same buffers and chunk plan have been seen before (a streaming
        caller re-filling the same pinned buffers): no per-chunk Python, event or launch cost is left on the host,
        which is what bounds small shards (2.5 M signals per rank at N = 8).  The first call with a new shape runs the
        same choreography eagerly."""
        torch = self.torch
        n = int(posA.numel())
        if n > self.n_max:
            raise ValueError("HostPipeline sized for %d signals, got %d" % (self.n_max, n))
        device_ops.check_min_pts(m, n)
        seg_off = np.asarray(seg_off, dtype=np.int64)
        n_chunks = self.n_chunks if n >= self.small_shard else min(self.n_chunks, 2)
        chunks = plan_chunks(seg_off, n_chunks, self.taper)
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
        spans = [(int(seg_off[p0]), int(seg_off[p1]), p1 - p0) for p0, p1 in chunks]
        args = (posA, posB, out, spans, views, pos, epsilon, m, max_pos)
        cur = torch.cuda.current_stream()
        if not self.use_graphs:
            self._enqueue(cur, *args)
        else:
            key = (posA.data_ptr(), posB.data_ptr(), out.data_ptr(), tuple(spans), tuple(views),
                   device_ops.eps_to_int(epsilon), int(m), int(max_pos))
            graph = self._graphs.get(key)
            if graph is None and key in self._seen:
                # second call with this shape: capture the whole choreography (nothing executes during capture)
                graph = torch.cuda.CUDAGraph()
                self.s_main.wait_stream(cur)
                with torch.cuda.graph(graph, stream=self.s_main):
                    self._enqueue(self.s_main, *args)
                if len(self._graphs) > 64:
                    self._graphs.clear()
                self._graphs[key] = graph
            if graph is not None:
                graph.replay()
            else:
                self._seen.add(key)
                self._enqueue(cur, *args)
        cur.synchronize()
        device_ops.check_async_status(int(self.status_pin[0]))
        return out

    def _enqueue(self, main, posA, posB, out, spans, views, n_off, epsilon, m, max_pos):
        """The call's stream choreography, forked from and joined back into `main` (eagerly or under capture)."""
        torch = self.torch
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(main)
        with torch.cuda.stream(self.s_in):
            self.off_d[:n_off].copy_(self.off_pin[:n_off], non_blocking=True)
            self.status_d.zero_()
            for k, (lo, hi, _) in enumerate(spans):
                self.a_d[lo:hi].copy_(posA[lo:hi], non_blocking=True)
                self.b_d[lo:hi].copy_(posB[lo:hi], non_blocking=True)
                self._events[k][0].record(self.s_in)
        for k, (lo, hi, P) in enumerate(spans):
            ev_in, _, ev_done = self._events[k]
            o0, ok = views[k]
            self.s_run.wait_event(ev_in)
            if hi > lo:
                with torch.cuda.stream(self.s_run):
                    self._chunk_call(lo, hi, o0, ok, P, epsilon, m, max_pos)
            ev_done.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_done)
                out[lo:hi].copy_(self.lab_d[lo:hi], non_blocking=True)
        with torch.cuda.stream(self.s_out):
            self.status_pin.copy_(self.status_d, non_blocking=True)
        main.wait_stream(self.s_in)
        main.wait_stream(self.s_run)
        main.wait_stream(self.s_out)

    def _chunk_call(self, lo, hi, o0, ok, P, epsilon, m, max_pos):
        device_ops.cluster_labels_device(self.a_d[lo:hi], self.b_d[lo:hi], self.off_d[o0:o0 + ok], P, epsilon, m,
                                         max_pos, labels_out=self.lab_d[lo:hi], status=self.status_d, ws=self.ws)


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
        self.pad = (max(self.counts) + 31) // 32 * 32 if self.counts else 0     # slot size: whole 128-byte lines

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


class _DevMem:
    """A raw device allocation as a __cuda_array_interface__ object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, n_int32):
        self.__cuda_array_interface__ = {"shape": (int(n_int32),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerExchange:
    """The p2p form of the label exchange: one IPC-exported buffer per rank (tdt_peer_alloc), mapped into every other
    rank (tdt_peer_open), and the push kernel of csrc/tdt_peer.cu (tdt_peer_allgather).  The 64-byte handles travel
    through torch.distributed's object all-gather once, at construction."""

    def __init__(self, pad, world, rank, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        L = _lib.lib()
        self.L, self.torch, self.dist, self.group = L, torch, dist, group
        self.pad, self.world, self.rank = int(pad), int(world), int(rank)
        if self.pad % 4:
            raise ValueError("slot size must be a multiple of 4 int32")
        nbytes = L.tdt_peer_buffer_bytes(self.pad, self.world)
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        rc = L.tdt_peer_alloc(nbytes, ctypes.byref(own), handle)
        self.own = own if rc == 0 else None
        self.opened = []
        # the handle exchange is reached by every rank whatever happened above, so the collectives stay matched
        handles = [None] * world
        dist.all_gather_object(handles, (rank, handle.raw if rc == 0 else None), group=group)
        if rc != 0 or any(raw is None for _, raw in handles):
            raise RuntimeError("tdt_peer_alloc failed on a rank: %s" % L.tdt_last_error().decode("utf-8", "replace"))
        self.ptrs = (ctypes.c_void_p * world)()
        for r, raw in handles:
            if r == rank:
                self.ptrs[r] = own
                continue
            p = ctypes.c_void_p()
            _lib.check(L.tdt_peer_open(raw, ctypes.byref(p)))
            self.ptrs[r] = p
            self.opened.append(p)
        self._mem = _DevMem(own.value, world * self.pad)
        self.gathered = torch.as_tensor(self._mem, device="cuda")
        self.gathered.fill_(-1)
        self.status = torch.zeros(1, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()            # (LabelExchange's all-reduce is the barrier before anybody pushes)

    def run(self):
        _lib.check(self.L.tdt_peer_allgather(self.ptrs, self.pad, self.rank, self.world, _lib.ptr(self.status),
                                             _lib.stream_ptr(self.torch)))
        return self.gathered

    def check(self):
        code = int(self.status.item())
        if code:
            raise _lib.TdtError(_lib.TDT_E_CUDA, "label exchange: a peer did not arrive within 5 s (status %d)" % code)

    def abandon(self):
        """Release without the collective handshakes (a rank failed to set up)."""
        self.torch.cuda.synchronize()
        self.gathered = None
        self._mem = None
        for p in getattr(self, "opened", []):
            self.L.tdt_peer_close(p)
        self.opened = []
        if getattr(self, "own", None) is not None:
            self.L.tdt_peer_free(self.own)
        self.own = None

    def close(self):
        if self.own is None:
            return
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)  # nobody unmaps while a peer may still be writing
        self.gathered = None
        self._mem = None
        for p in self.opened:
            self.L.tdt_peer_close(p)
        self.opened = []
        self.dist.barrier(group=self.group)  # every mapping is gone before the owner frees
        self.L.tdt_peer_free(self.own)
        self.own = None


class LabelExchange:
    """Every rank's finished label shard to every rank (north_star: "a single all-gather of the final label array over
    NVLink").  shard: this rank's int32 slot (pad elements; the clustering call writes its labels straight into it);
    gathered: world * pad int32, rank-major (ShardPlan.gather_index maps it back to input order).

    kind "nccl": dist.all_gather_into_tensor.  kind "p2p" (see PeerExchange): our own push kernel over NVLink peer
    memory -- selected when the peer buffers could be mapped; TDT_LABEL_EXCHANGE=nccl|p2p overrides."""

    def __init__(self, pad, world, rank, group=None):
        import os
        import torch
        self.torch, self.pad, self.world, self.rank, self.group = torch, int(pad), int(world), int(rank), group
        want = os.environ.get("TDT_LABEL_EXCHANGE", "auto")
        self.peer = None
        if want in ("auto", "p2p") and world > 1:
            import torch.distributed as dist
            peer = PeerExchange.__new__(PeerExchange)
            try:
                peer.__init__(self.pad, world, rank, group)
                self.peer = peer
            except Exception as exc:             # no peer access / IPC in this environment: NCCL does the same job
                self.peer_error = "%s: %s" % (type(exc).__name__, str(exc)[:200])
                try:
                    peer.abandon()
                except Exception:
                    pass
            # all ranks take the same path: p2p only if EVERY rank mapped every buffer
            ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                if self.peer is not None:
                    self.peer.abandon()
                    self.peer = None
                if want == "p2p":
                    raise RuntimeError("peer-memory label exchange unavailable: %s" % getattr(self, "peer_error", "a peer failed"))
        if self.peer is not None:
            self.kind = "p2p push kernel over NVLink peer memory (tdt_peer_allgather)"
            self.gathered = self.peer.gathered
            self.shard = self.gathered[rank * self.pad:(rank + 1) * self.pad]
            self.launches = 2
        else:
            self.kind = "nccl all_gather_into_tensor"
            self.shard = torch.full((max(self.pad, 1),), -1, dtype=torch.int32, device="cuda")[:self.pad]
            self.gathered = torch.empty(world * self.pad, dtype=torch.int32, device="cuda")
            self.launches = 1

    def run(self):
        """Enqueue the exchange on the current stream; `gathered` is complete when the stream reaches the end of it."""
        if self.peer is not None:
            self.peer.run()
        else:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.gathered, self.shard, group=self.group)
        return self.gathered

    def close(self):
        if self.peer is not None:
            self.peer.close()
            self.peer = None


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
    if on_gpu:
        # GPU box: the shard goes into this rank's slot of the exchange buffer and is pushed to every peer (p2p kernel
        # when the buffers can be peer-mapped, NCCL all-gather otherwise)
        exch = LabelExchange(plan.pad, world, rank, group)
        exch.shard.fill_(-1)
        exch.shard[:len(idx)] = torch.from_numpy(np.asarray(mine, dtype=np.int32)).cuda()
        out = exch.run().cpu().numpy()[plan.gather_index()]
        if exch.peer is not None:
            exch.peer.check()
        exch.close()
        return out
    dev = torch.device("cpu")
    send = torch.full((plan.pad,), -1, dtype=torch.int32, device=dev)
    send[:len(idx)] = torch.from_numpy(np.asarray(mine, dtype=np.int32)).to(dev)
    recv = torch.empty(world * plan.pad, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv.cpu().numpy()[plan.gather_index()]


# ---------------------------------------------------------------------------------------------
# host arrays -> coverage bins / GC bins, copies overlapped with the kernels
# ---------------------------------------------------------------------------------------------
def coverage_host(start, end, read_off, lengths, bin_size, out=None, chunk_reads=1 << 25):
    """tiddit_coverage.pyx:48-74 for ALL reads of ALL contigs held in (pinned) host memory -> float64 bins of all
    contigs back to back (pinned CPU tensor `out`, allocated when None) and the bin offsets.

    start / end: int32 CPU tensors (pinned for full copy rate) grouped by contig, read_off int64 [C+1], lengths [C].
    The reads travel in chunks of `chunk_reads`: the H2D copy of chunk k+1 overlaps tdt_coverage_accumulate_contigs on
    chunk k (two device buffers, two streams); the bins stay in HBM until the last chunk and come back once."""
    torch = _lib.torch_cuda()
    lengths = np.asarray(lengths, dtype=np.int64)
    read_off = np.asarray(read_off, dtype=np.int64)
    nb = np.ceil(lengths / float(bin_size)).astype(np.int64)
    bin_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    n_bins, n = int(bin_off[-1]), int(read_off[-1])
    ebs_d = torch.from_numpy((lengths - (nb - 1) * bin_size).astype(np.int32)).cuda()
    bin_off_d = torch.from_numpy(bin_off).cuda()
    bins = torch.zeros(n_bins, dtype=torch.float64, device="cuda")
    bad = device_ops.new_first_bad(torch)
    if out is None:
        out = torch.empty(n_bins, dtype=torch.float64).pin_memory()
    k = max(1, min(int(chunk_reads), n))
    bufs = [(torch.empty(k, dtype=torch.int32, device="cuda"), torch.empty(k, dtype=torch.int32, device="cuda"))
            for _ in range(2)]
    # chunk-relative read offsets of every chunk, one small upload
    cuts = list(range(0, n, k)) + [n]
    offs = np.stack([np.clip(read_off, lo, hi) - lo for lo, hi in zip(cuts[:-1], cuts[1:])]) if n else np.zeros((0, len(read_off)), np.int64)
    offs_d = torch.from_numpy(np.ascontiguousarray(offs)).cuda()
    cur = torch.cuda.current_stream()
    s_copy = torch.cuda.Stream()
    s_copy.wait_stream(cur)
    done = [None, None]
    for i, (lo, hi) in enumerate(zip(cuts[:-1], cuts[1:])):
        sb, eb = bufs[i & 1]
        with torch.cuda.stream(s_copy):
            if done[i & 1] is not None:
                s_copy.wait_event(done[i & 1])                 # the kernel that read this buffer two chunks ago
            sb[:hi - lo].copy_(start[lo:hi], non_blocking=True)
            eb[:hi - lo].copy_(end[lo:hi], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(s_copy)
        cur.wait_event(ready)
        device_ops.coverage_accumulate_contigs_device(sb[:hi - lo], eb[:hi - lo], offs_d[i], bin_off_d, ebs_d, int(bin_size),
                                                      bins, bad)
        done[i & 1] = torch.cuda.Event()
        done[i & 1].record(cur)
    out.copy_(bins, non_blocking=True)
    bad_h = int(bad.item())                                    # synchronises
    if bad_h != device_ops.FIRST_BAD_NONE:
        raise IndexError("Out of bounds on buffer access (axis 0)")
    return out, bin_off


def gc_host(seq, bin_size, n_cutoff, out=None, chunk_bytes=32 << 20):
    """tiddit_gc.pyx:6-33 for one contig's bases held in (pinned) host memory (uint8 CPU tensor) -> int8 bins (pinned
    CPU tensor).  Chunks of whole bins (a multiple of 16 bytes, so every chunk starts 16-byte aligned as the kernel's
    bulk copies need) are copied while the previous chunk is being counted; the bins come back once."""
    torch = _lib.torch_cuda()
    n = int(seq.numel())
    z = int(bin_size)
    if z == 0:
        raise ZeroDivisionError("division by zero")
    n_bins = -(-n // z)
    if out is None:
        out = torch.empty(n_bins, dtype=torch.int8).pin_memory()
    bins = torch.zeros(max(n_bins, 1), dtype=torch.int8, device="cuda")
    unit = z * 16
    step = max(unit, int(chunk_bytes) // unit * unit)          # whole bins, 16-byte aligned chunk starts
    bufs = [torch.zeros(step + 32, dtype=torch.uint8, device="cuda") for _ in range(2)]
    cur = torch.cuda.current_stream()
    s_copy = torch.cuda.Stream()
    s_copy.wait_stream(cur)
    done = [None, None]
    for i, lo in enumerate(range(0, n, step)):
        hi = min(n, lo + step)
        buf = bufs[i & 1]
        with torch.cuda.stream(s_copy):
            if done[i & 1] is not None:
                s_copy.wait_event(done[i & 1])
            buf[:hi - lo].copy_(seq[lo:hi], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(s_copy)
        cur.wait_event(ready)
        device_ops.gc_bins_device(buf, hi - lo, z, n_cutoff, out=bins[lo // z:])
        done[i & 1] = torch.cuda.Event()
        done[i & 1].record(cur)
    out.copy_(bins[:n_bins], non_blocking=True)
    cur.synchronize()
    return out


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
