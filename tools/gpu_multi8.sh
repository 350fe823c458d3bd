#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for ex in ${EXCH:-p2p nccl}; do
  export TDT_LABEL_EXCHANGE=$ex
  n=8
  out=gpurun_out/bench_n${n}_$ex
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 ${BENCH_ARGS} > $out.json 2> $out.err
  echo "N=$n exchange=$ex rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $out.err | tail -5
  python - <<PY
import json
try:
    d = json.loads(open("$out.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ["n_gpus", "value", "ms_per_step", "verified"]}, "e2e_ms", d["e2e"]["ms_per_step"])
    print("  exchange", d.get("exchange"))
    print("  stages", d["roofline"]["stages_ms"])
    t = d.get("tumor60x", {})
    print("  tumor60x", {k: t.get(k) for k in ["value", "ms_per_step", "verified"]}, t.get("e2e", {}).get("ms_per_step"), t.get("exchange"))
    print("  sharded", d.get("sharded"))
except Exception as e:
    print("no result:", e)
PY
done
