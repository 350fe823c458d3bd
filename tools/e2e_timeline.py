"""Timeline of one HostPipeline.run(): when each chunk's H2D, kernels and D2H finish (CUDA events)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiddit_b200 import engine, synth, device_ops

a, b, off, L = synth.wgs30x_signals(20_000_000)
a_pin, b_pin = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
out = torch.empty(len(a), dtype=torch.int32).pin_memory()
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pipe = engine.HostPipeline(len(a), n_chunks=nch)
for _ in range(3):
    pipe.run(a_pin, b_pin, off, 500, 3, L, out)
torch.cuda.synchronize()
# re-run the pipeline body by hand with timing events
seg_off = off
chunks = engine.plan_chunks(seg_off, nch)
E = lambda: torch.cuda.Event(enable_timing=True)
t0 = E(); evs = []
cur = torch.cuda.current_stream()
t0.record(cur)
for s in (pipe.s_in_a, pipe.s_in_b, pipe.s_run, pipe.s_out):
    s.wait_stream(cur)
offs_d = []
for p0, p1 in chunks:
    offs_d.append(torch.from_numpy(seg_off[p0:p1 + 1] - seg_off[p0]).cuda())
torch.cuda.synchronize()
t0.record(cur)
for s in (pipe.s_in_a, pipe.s_in_b, pipe.s_run, pipe.s_out):
    s.wait_stream(cur)
pipe.status_d.zero_()
rec = []
for k, (p0, p1) in enumerate(chunks):
    lo, hi = int(seg_off[p0]), int(seg_off[p1])
    ea, eb = E(), E()
    with torch.cuda.stream(pipe.s_in_a):
        pipe.a_d[lo:hi].copy_(a_pin[lo:hi], non_blocking=True); ea.record(pipe.s_in_a)
    with torch.cuda.stream(pipe.s_in_b):
        pipe.b_d[lo:hi].copy_(b_pin[lo:hi], non_blocking=True); eb.record(pipe.s_in_b)
    rec.append([ea, eb])
for k, (p0, p1) in enumerate(chunks):
    lo, hi = int(seg_off[p0]), int(seg_off[p1])
    ea, eb = rec[k]
    es, ed, eo = E(), E(), E()
    with torch.cuda.stream(pipe.s_run):
        pipe.s_run.wait_event(ea); pipe.s_run.wait_event(eb)
        es.record(pipe.s_run)
        device_ops.cluster_labels_device(pipe.a_d[lo:hi], pipe.b_d[lo:hi], offs_d[k], p1 - p0, 500, 3, L,
                                         labels_out=pipe.lab_d[lo:hi], status=pipe.status_d)
        ed.record(pipe.s_run)
    with torch.cuda.stream(pipe.s_out):
        pipe.s_out.wait_event(ed)
        out[lo:hi].copy_(pipe.lab_d[lo:hi], non_blocking=True); eo.record(pipe.s_out)
    rec[k] += [es, ed, eo]
torch.cuda.synchronize()
for k, (ea, eb, es, ed, eo) in enumerate(rec):
    print("chunk %d (%d signals): H2D a %.2f b %.2f | kernels %.2f -> %.2f | D2H done %.2f ms" % (
        k, int(seg_off[chunks[k][1]] - seg_off[chunks[k][0]]), t0.elapsed_time(ea), t0.elapsed_time(eb),
        t0.elapsed_time(es), t0.elapsed_time(ed), t0.elapsed_time(eo)))
