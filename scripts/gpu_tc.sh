#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tcgen05 gemm tests"
timeout 240 python -m pytest tests/test_gemm_tc_gpu.py -q -x --timeout 120 2>&1 | tail -25 | tee gpurun_out/pytest_tc.log
