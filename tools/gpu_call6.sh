#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline > gpurun_out/bench3_bf16.json 2>> gpurun_out/bench3.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench3_ref.json 2>> gpurun_out/bench3.err; echo "ref rc=$?"
timeout 1200 python tools/sweep.py --variants B,S,T,L,L64,HTC --iters 30 --out gpurun_out/sweep_final.json > gpurun_out/sweep_final.log 2>&1; echo "sweep rc=$?"
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v2.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1
python tools/profile_step.py --dtype bf16 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v2_bf16.csv python tools/profile_step.py --dtype bf16 > gpurun_out/ncu1b.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench3.json')); print('value',d['value'],'e2e',d['e2e']['value'],d['clocks'],d['roofline']['frac'],d['roofline']['step_frac'])"
