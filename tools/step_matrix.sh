#!/bin/bash
# Model-level matrix for profiles/rN_step_bench_1gpu.jsonl:  gpurun --timeout 1500 -- 'bash tools/step_matrix.sh'
mkdir -p gpurun_out; : > gpurun_out/step_matrix.jsonl
run() { timeout 400 python bench_step.py --steps 10 --warmup 3 "$@" >> gpurun_out/step_matrix.jsonl 2>> gpurun_out/step_matrix.err || echo "FAILED: $*"; }
for amp in "" "--amp"; do
  run $amp --reference-sequence
  run $amp --op ref_cuda
  run $amp
  run $amp --graph
done
run --tf32 --reference-sequence
run --tf32
run --tf32 --graph
run --variant L --image 896 --batch 1 --amp --with-cp --reference-sequence
run --variant L --image 896 --batch 1 --amp --with-cp
run --variant L --image 896 --batch 1 --amp --with-cp --graph
run --variant L --image 1024 --batch 1 --mode infer --reference-sequence
run --variant L --image 1024 --batch 1 --mode infer
run --variant L --image 1024 --batch 1 --mode infer --graph
run --variant L --image 1024 --batch 1 --mode infer --amp
run --variant L --image 1024 --batch 1 --mode infer --amp --graph
python - <<'PY'
import json
for l in open('gpurun_out/step_matrix.jsonl'):
    d = json.loads(l)
    print(d['metric'], d['dtype'], d['op'], d['adapter'][:9], 'graph' if d['cuda_graph'] else 'eager', 'tf32' if d.get('tf32_gemm') else '', round(d['value'], 2), 'img/s', round(d['ms_per_step'], 2), 'ms')
PY
tail -3 gpurun_out/step_matrix.err
