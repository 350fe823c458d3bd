"""Two clustering calls on the 30X set (for ncu captures of the sort kernels); no timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tiddit_b200 import device_ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
a, b, off, L = synth.wgs30x_signals(n)
A, B, O = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(off).cuda()
for _ in range(2):
    device_ops.cluster_labels_device(A, B, O, len(off) - 1, 500, 3, L)
torch.cuda.synchronize()
print("done")
