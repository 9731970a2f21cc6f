#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep.py --variants B,S,L,L64,HTC --iters 20 --smem 1,0,2:512,2:1024 --regimes flushed --no-ref --out gpurun_out/sweep3.json > gpurun_out/sweep3.log 2>&1; echo "sweep rc=$?"
grep -v "^$" gpurun_out/sweep3.md | cut -d'|' -f2-8 | head -90
