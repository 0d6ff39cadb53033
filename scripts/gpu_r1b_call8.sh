#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm tests (persistent)"
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3
echo "== gemm bench one-tile"
PADT_GEMM_TC=2 timeout 300 python scripts/gemm_bench.py --no-cublas 2>&1 | grep '^{' | tee gpurun_out/gemm_onetile.jsonl | tail -1
echo "== gemm bench persistent"
timeout 300 python scripts/gemm_bench.py --no-cublas 2>&1 | grep '^{' | tee gpurun_out/gemm_persistent.jsonl
echo "== model tests"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -3
