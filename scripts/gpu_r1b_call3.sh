#!/bin/bash
# v4 persistent decode kernel: parity, knob sweep, per-phase cycle accounting
set -u
mkdir -p gpurun_out
echo "== pytest gpu (model tests)"
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_v4.log
echo "== sweep"
timeout 900 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "0,5,0;0,7,0;0,1,0;0,4,0;8,5,0;16,5,0;32,5,0;16,13,0;32,13,0;16,5,400" --trace "0,5,0;16,5,0" 2>&1 | grep -v Warning | tee gpurun_out/mega_sweep_v4.log | tail -40
