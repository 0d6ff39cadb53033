#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (model)"
timeout 1500 python -m pytest tests/test_model_gpu.py -q -x --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench"
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
