"""Host<->device copy rates on this box: one stream vs several, pinned memory."""
import time
import torch

n = 160 << 20
host = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
print("H2D 160MB one stream: %.1f GB/s" % (n / t(lambda: dev.copy_(host, non_blocking=True)) / 1e9))
print("D2H 160MB one stream: %.1f GB/s" % (n / t(lambda: host.copy_(dev, non_blocking=True)) / 1e9))
for k in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    c = n // k
    def multi():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dev[i * c:(i + 1) * c].copy_(host[i * c:(i + 1) * c], non_blocking=True)
    print("H2D 160MB %d streams: %.1f GB/s" % (k, n / t(multi) / 1e9))
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): h2[:n // 2].copy_(dev[:n // 2], non_blocking=True)
print("H2D 160MB + D2H 80MB concurrently: %.2f ms" % (t(both) * 1e3))
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run("lscpu | grep -E 'Model name|Socket|NUMA node\\(s\\)|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
