#!/bin/bash
# end-of-round check on an N-GPU box: the driver's own sequence -- smoke, gpu tests, default bench at N=1, then the scaling runs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
if [ "${SKIP_TESTS:-0}" != 1 ]; then echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3; fi
echo "== bench N=1 (default flags)"; T0=$SECONDS; timeout 900 python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "bench wall $((SECONDS-T0)) s"; tail -2 gpurun_out/final_n1.err
echo "== reference arm"; T0=$SECONDS; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "reference wall $((SECONDS-T0)) s"; tail -1 gpurun_out/final_ref.err
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    T0=$SECONDS; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > gpurun_out/final_n$n.json 2> gpurun_out/final_n$n.err
    echo "N=$n wall $((SECONDS-T0)) s"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/final_n$n.err | tail -2
  fi
done
python - <<'PY'
import json, glob
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/final_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        continue
    if n == 1: base = d["value"]
    t = d.get("tumor60x", {})
    print("N=%d value %.2f G/s  step %.4f ms  eff %.3f  e2e %.3f ms  verified %s | tumor60x %.2f G/s %.3f ms verified %s | exchange %s" % (
        n, d["value"] / 1e9, d["ms_per_step"], d["value"] / (n * base) if base else 0, d["e2e"]["ms_per_step"], d.get("verified"),
        t.get("value", 0) / 1e9, t.get("ms_per_step", 0), t.get("verified"), (d.get("exchange") or {}).get("ms")))
    if n == 1:
        for leg in ("coverage", "gc", "aggregate", "ploidy_medians", "cluster_main", "bam_coverage"):
            x = d.get(leg, {})
            print("   ", leg, {k: x.get(k) for k in ("ms_per_step", "s_per_call", "s_per_pass", "verified", "speedup_vs_reference")}, (x.get("roofline") or {}).get("frac"), (x.get("e2e") or {}).get("ms_per_step"))
PY
