#!/bin/bash
# first GPU call: parity tests, bench, sweep, ncu launch list + full capture of the four kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2>> gpurun_out/bench.err
timeout 900 python tools/sweep.py --variants B,S,L,L64,HTC --iters 20 --qc 0,32,64,256 > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 8 -c 4 -o gpurun_out/prof_r1_B_f32 python tools/profile_step.py > gpurun_out/ncu2.log 2>&1
timeout 300 python tools/profile_step.py --ref > gpurun_out/plain_ref.log 2>&1 &&
ncu --set full --clock-control none -k regex:ms_deformable -s 4 -c 4 -o gpurun_out/prof_r1_B_f32_refcuda python tools/profile_step.py --ref > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
