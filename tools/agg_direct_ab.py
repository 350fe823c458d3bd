"""A/B of the direct modes / names kernel of tdt_cluster_aggregate (TDT_AGG_DIRECT, csrc/tdt_aggregate.cu) on the
30X-shaped set and on the tumour-shaped set: rows of every setting compared with the all-sorts path (TDT_AGG_DIRECT=0),
whole-call and per-stage times.   python tools/agg_direct_ab.py [n30x] [ntumor] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tiddit_b200 import device_ops, synth, _lib

n30 = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
ntu = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
SETTINGS = os.environ.get("TDT_AB_SETTINGS", "0 32 64 128 256 1024").split()   # "k" or "k:overlap" (TDT_AGG_OVERLAP)
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for name, gen, n, eps, m in (("wgs30x", synth.wgs30x_signals, n30, 500, 3), ("tumor60x", synth.tumor60x_signals, ntu, 1000, 5)):
    if n <= 0:
        continue
    a, b, off, L = gen(n)
    rec = synth.signal_records(a, b, off)
    A, B, O = d(a), d(b), d(off)
    span, nm, flags, same = d(rec["span"]), d(rec["name_id"]), d(rec["flags"]), d(rec["same_chrom"])
    P = len(off) - 1
    labels = device_ops.cluster_labels_device(A, B, O, P, eps, m, L)
    rows = torch.empty((n, 16), dtype=torch.int32, device="cuda")
    mem = torch.empty(n, dtype=torch.int32, device="cuda")
    counts = torch.zeros(4, dtype=torch.int64, device="cuda")
    run = lambda: device_ops.cluster_aggregate_device(labels, A, B, span, nm, flags, O, same, P, 5000, False, 3, L,
                                                      rec["n_names"], rows, mem, counts)
    keep = [c for c in range(16) if c != 3]
    base = None
    res = {}
    for s in SETTINGS:
        os.environ["TDT_AGG_DIRECT"] = s.split(":")[0]
        os.environ.pop("TDT_AGG_OVERLAP", None)
        if ":" in s:
            os.environ["TDT_AGG_OVERLAP"] = s.split(":")[1]
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        C, Mm, err = counts[:3].tolist()
        r = rows[:C][:, keep].clone()
        if base is None:
            base = r
            sizes = rows[:C, 4]
            big = {t: int((sizes > t).sum()) for t in (32, 64, 128, 256, 1024)}
        same_rows = bool(torch.equal(r, base))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        whole, tot = [], {}
        for _ in range(reps):
            flush.add_(1)
            torch.cuda.synchronize()
            ev[0].record(); run(); ev[1].record(); torch.cuda.synchronize()
            whole.append(ev[0].elapsed_time(ev[1]))
            flush.add_(1)
            torch.cuda.synchronize()
            _lib.profile_begin()
            run()
            for k, ms in _lib.profile_end():
                tot[k] = tot.get(k, 0) + ms / reps
        res[s] = {"ms": round(float(np.mean(whole)), 4), "min_ms": round(float(np.min(whole)), 4), "rows_equal_to_sorts": same_rows,
                  "err": err, "stages": {k: round(v, 4) for k, v in tot.items()}}
        print(name, "TDT_AGG_DIRECT[:OVERLAP]=%s" % s, "candidates", C, "members", Mm, res[s], flush=True)
    out[name] = {"n": n, "candidates": C, "members": Mm, "candidates_larger_than": big, "settings": res}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/agg_direct_ab.json", "w"), indent=1)
