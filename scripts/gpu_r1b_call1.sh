#!/bin/bash
# round 1, session 2, GPU call 1: parity of the rebuilt library, knob sweep of the persistent decode kernel, per-phase timelines
set -u
mkdir -p gpurun_out
echo "== pytest gpu (defaults)"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_default.log
echo "== pytest gpu (model tests with L2 prefetch warp + kv smem prefetch + evict_first)"
PADT_MEGA_PF=8 PADT_MEGA_FLAGS=7 timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_pf.log
echo "== sweep"
timeout 900 python scripts/mega_sweep.py --steps 64 --reps 3 --trace "0,0,0;0,6,0;16,7,0;32,7,0;16,15,0" 2>&1 | grep -v Warning | tee gpurun_out/mega_sweep.log | tail -60
