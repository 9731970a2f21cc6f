#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
./tools/microbench/mb > gpurun_out/microbench.json 2>&1; echo "mb rc=$?"
timeout 900 python tools/sweep.py --variants B,S,L --iters 20 --minb 3:2,4:3,6:4 --regimes flushed --no-ref --out gpurun_out/sweep2.json > gpurun_out/sweep2.log 2>&1; echo "sweep rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
