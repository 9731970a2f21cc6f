#!/bin/bash
mkdir -p gpurun_out
python tools/profile_step.py --smem 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 8 -c 4 -o gpurun_out/prof_r1_v2_B_f32 python tools/profile_step.py --smem 1 > gpurun_out/ncu_a.log 2>&1
python tools/profile_step.py --smem 2 --smem-threads 1024 --fwd-only > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_fwd_smem -s 4 -c 2 -o gpurun_out/prof_r1_v2smem_B_f32 python tools/profile_step.py --smem 2 --smem-threads 1024 --fwd-only > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
