#!/bin/bash
# multi-GPU session: NCCL sharding test + bench at N = 1 .. NGPU
mkdir -p gpurun_out
NG=${NGPU:-2}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_engine.py -q --timeout 300 2>&1 | tail -3
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-coverage --no-cpu --no-extra > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    fi
    echo "N=$n rc=$?"; tail -2 gpurun_out/bench_n$n.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['n_gpus','value','ms_per_step','verified']}, d['e2e']['ms_per_step'], d['roofline']['stages_ms'])"
  fi
done
