"""Where does the host-pipeline time go?  Host enqueue time vs device completion, per chunk count."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiddit_b200 import engine, synth, device_ops, _lib

a, b, off, L = synth.wgs30x_signals(20_000_000)
a_pin, b_pin = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
out = torch.empty(len(a), dtype=torch.int32).pin_memory()
for chunks in (1, 2, 4, 8):
    pipe = engine.HostPipeline(len(a), n_chunks=chunks)
    for _ in range(3):
        pipe.run(a_pin, b_pin, off, 500, 3, L, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); pipe.run(a_pin, b_pin, off, 500, 3, L, out); ts.append(time.perf_counter() - t0)
    print("chunks %d: run() wall %.2f ms (min %.2f)" % (chunks, 1e3 * np.mean(ts), 1e3 * min(ts)))
# host cost of enqueueing one device-resident call (no copies)
a_d, b_d, off_d = a_pin.cuda(), b_pin.cuda(), torch.from_numpy(off).cuda()
lab = torch.empty(len(a), dtype=torch.int32, device="cuda")
st = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(3):
    device_ops.cluster_labels_device(a_d, b_d, off_d, len(off) - 1, 500, 3, L, labels_out=lab, status=st)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    device_ops.cluster_labels_device(a_d, b_d, off_d, len(off) - 1, 500, 3, L, labels_out=lab, status=st)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("async call: host enqueue %.3f ms per call, device %.3f ms per call" % ((t1 - t0) * 100, (t2 - t0) * 100))
# plain copies as the pipeline issues them
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def copies():
    with torch.cuda.stream(s1): a_d.copy_(a_pin, non_blocking=True)
    with torch.cuda.stream(s2): b_d.copy_(b_pin, non_blocking=True)
    s3.wait_stream(s1); s3.wait_stream(s2)
    with torch.cuda.stream(s3): out.copy_(lab, non_blocking=True)
    s3.synchronize()
copies()
t0 = time.perf_counter(); copies(); print("H2D(2 streams) then D2H: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
