#!/bin/bash
# One GPU-box pass: parity tests, smoke, headline bench (both arms), shape sweep, ncu launch list + full capture.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'
# (gpurun brings back at most 64 MiB of gpurun_out/: the full reports are summarised with tools/ncu_summary.py and removed
#  at the end; tools/gpu_final.sh is the short version of this pass)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2>> gpurun_out/bench.err; echo "bench rc=$?"
timeout 1200 python tools/sweep.py --variants B,S,T,L,L64,HTC --iters 30 --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 14 -c 7 -o gpurun_out/prof python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1   # 7 kernels per step: 2 warm steps skipped
# round 2: the slab-sorted backward against the query-order one at every BASELINE shape (and at 1 / 2 / 4 images), the
# walker / sort kernels under ncu, the north-star configuration's Extractor call under ncu
python tools/bwd_cell_check.py --mode sorted --variants B,S,T,L,L64 --dtypes f32,bf16 --out gpurun_out/bwd_sorted_vs_query_order.jsonl > gpurun_out/bwd_sorted.log 2>&1
for b in 1 2 4; do python tools/bwd_cell_check.py --mode sorted --variants B,S --dtypes f32,bf16 --batch $b --out gpurun_out/bwd_sorted_b$b.jsonl >> gpurun_out/bwd_sorted.log 2>&1; done
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sort -s 4 -c 4 -o gpurun_out/prof_sorted python tools/profile_step.py --warm 1 > gpurun_out/ncu_full.log 2>&1
for v in L L64; do python tools/profile_step.py --variant $v --dtype bf16 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 9 -c 9 -o gpurun_out/prof_${v}_bf16 python tools/profile_step.py --variant $v --dtype bf16 --warm 1 > gpurun_out/ncu_full.log 2>&1; done
ls -la gpurun_out
python tools/profile_adapter_kernels.py > gpurun_out/plain_adapter.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:adapter_ -s 20 -c 10 -o gpurun_out/prof_adapter python tools/profile_adapter_kernels.py > gpurun_out/ncu_adapter.log 2>&1
python tools/profile_block.py --batch 16 --amp 1 --top 30 --out gpurun_out/block_profile_b16_bf16.json > gpurun_out/block_profile_b16_bf16.txt 2>&1
python tools/profile_block.py --batch 2 --amp 1 --top 30 --out gpurun_out/block_profile_b2_bf16.json > gpurun_out/block_profile_b2_bf16.txt 2>&1
python tools/bench_layernorm.py > gpurun_out/layernorm_kernels.jsonl 2> gpurun_out/adapter_bench.err
python tools/bench_dwconv.py --kernels > gpurun_out/dwconv_kernels.jsonl 2>> gpurun_out/adapter_bench.err
python tools/bench_dwconv.py > gpurun_out/dwconv_bench.jsonl 2>> gpurun_out/adapter_bench.err
ls -la gpurun_out
for r in gpurun_out/*.ncu-rep; do python tools/ncu_summary.py $r > ${r%.ncu-rep}.summary.txt 2>&1; rm -f $r; done
ls -la gpurun_out
