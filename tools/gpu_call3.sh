#!/bin/bash
mkdir -p gpurun_out
for op in ours ref_cuda; do
  timeout 600 python bench_step.py --variant B --mode train --batch 2 --steps 10 --warmup 3 --op $op >> gpurun_out/step.jsonl 2>> gpurun_out/step.err
  timeout 600 python bench_step.py --variant B --mode train --batch 2 --steps 10 --warmup 3 --amp --op $op >> gpurun_out/step.jsonl 2>> gpurun_out/step.err
  timeout 600 python bench_step.py --variant L --mode infer --image 1024 --batch 1 --steps 10 --warmup 3 --op $op >> gpurun_out/step.jsonl 2>> gpurun_out/step.err
  timeout 600 python bench_step.py --variant L --mode train --image 896 --batch 1 --steps 5 --warmup 2 --amp --with-cp --op $op >> gpurun_out/step.jsonl 2>> gpurun_out/step.err
done
tail -3 gpurun_out/step.err
cat gpurun_out/step.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['metric'], d['op'], d['dtype'], round(d['value'],2), 'img/s', round(d['ms_per_step'],2),'ms', d['msda_kernel_launches'])
"
