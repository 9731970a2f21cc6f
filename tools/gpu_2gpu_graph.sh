#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for a in "--amp --graph"; do
  timeout 150 $TR bench_step.py --gpus 2 --steps 10 --warmup 3 $a 2>gpurun_out/g2.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['dtype'], 'graph' if d['cuda_graph'] else 'eager', round(d['value'],2), 'img/s', round(d['ms_per_step'],2), 'ms', d['final'])"
  echo "rc=$?"; tail -3 gpurun_out/g2.err | cut -c1-250
done
