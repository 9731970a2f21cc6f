#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline --no-other-shapes > gpurun_out/bench5_bf16.json 2>> gpurun_out/bench5.err
timeout 600 python bench.py --steps 20 --warmup 5 --variant S --no-cpu-baseline --no-other-shapes > gpurun_out/bench5_S.json 2>> gpurun_out/bench5.err
timeout 1200 python tools/sweep.py --variants B,S,T,L,L64,HTC --iters 30 --out gpurun_out/sweep_final2.json > gpurun_out/sweep_final2.log 2>&1; echo "sweep rc=$?"
python -c "
import json
for f in ['bench5','bench5_bf16','bench5_S']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f,'value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'frac',round(d['roofline']['frac'],3),'step_frac',round(d['roofline']['step_frac'],3),d['roofline']['kernel'],[round(k['ms']*1e3,1) for k in d['kernels']],d['clocks'])
    if d.get('other_shapes'):
        for o in d['other_shapes']: print('   ',o['variant'],o['dtype'],round(o['us'],1),'us',round(o['gsamples_s'],2),'Gs/s frac',round(o['hbm_frac'],3))
"
