#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== barrier microbench"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bar_bench scripts/micro/bar_bench.cu && timeout 120 /tmp/bar_bench | tee gpurun_out/bar_bench.log
echo "== pytest gpu (model tests)"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_v4b.log
echo "== sweep"
timeout 900 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "0,5,0;0,7,0;16,13,0;24,13,0;32,13,0;16,29,0;24,29,0;32,29,0;48,29,0;64,29,0" --trace "0,5,0;16,13,0" 2>&1 | grep -v Warning | tee gpurun_out/mega_sweep_v4b.log | grep -v "^\[mega" | tail -20
grep "^\[mega cta 0\]" gpurun_out/mega_sweep_v4b.log | tail -4
