#!/bin/bash
# First GPU pass: parity tests, smoke, short bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" 
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== pytest gpu (no -x, summary only)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_all.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -20 | tee gpurun_out/smoke.log
echo "== bench"
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -30 | tee gpurun_out/bench.log
