#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err; echo "ref2 rc=$?"
timeout 600 $TR bench_step.py --gpus 2 --variant B --mode train --batch 2 --steps 10 --warmup 3 > gpurun_out/step_2gpu.json 2> gpurun_out/step_2gpu.err; echo "step2 rc=$?"
timeout 600 $TR bench_step.py --gpus 2 --variant B --mode train --batch 2 --steps 10 --warmup 3 --amp >> gpurun_out/step_2gpu.json 2>> gpurun_out/step_2gpu.err; echo "step2amp rc=$?"
tail -n 2 gpurun_out/bench_2gpu.err gpurun_out/step_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json')); print('bench 2gpu value',d['value'],'n',d['n_gpus'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print(open('gpurun_out/bench_2gpu_ref.json').read()[:200])
for l in open('gpurun_out/step_2gpu.json'):
    d=json.loads(l); print(d['metric'],d['n_gpus'],d['dtype'],d['value'],d['ms_per_step'])
PY
