#!/bin/bash
# The round's last GPU-box pass (a subset of tools/gpu_round.sh: only what the last kernel changes touch):
# smoke, parity tests, both bench arms, the sweep of every BASELINE shape, the ncu launch list of one bench step and
# full captures of the step's kernels at ViT-Adapter-B fp32 and at the north-star configuration (L 896^2 bf16, 16x32 / 16x64).
#   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh'
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
fi
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2>> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python tools/sweep.py --variants B,S,T,L,L64,HTC --iters 30 --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/profile_step.py --warm 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:msda_ -s 7 -c 7 -o gpurun_out/prof_B_f32 python tools/profile_step.py --warm 1 > gpurun_out/ncu_full.log 2>&1
for v in L L64; do python tools/profile_step.py --variant $v --dtype bf16 --warm 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:msda_ -s 9 -c 9 -o gpurun_out/prof_${v}_bf16 python tools/profile_step.py --variant $v --dtype bf16 --warm 1 > gpurun_out/ncu_full.log 2>&1; done
# gpurun brings back at most 64 MiB: keep the summaries, drop the reports
for r in prof_B_f32 prof_L_bf16 prof_L64_bf16; do python tools/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.summary.txt 2>&1; rm -f gpurun_out/$r.ncu-rep; done
ls -la gpurun_out | tail -30
