#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm tests"
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -m gpu -q -x --timeout 200 2>&1 | tail -3
echo "== gemm bench prev"
timeout 300 python scripts/gemm_bench.py --lib padt_b200/variants/lib_prev.so 2>&1 | grep '^{' | tee gpurun_out/gemm_prev.jsonl
echo "== gemm bench tree"
timeout 300 python scripts/gemm_bench.py --no-cublas 2>&1 | grep '^{' | tee gpurun_out/gemm_tree.jsonl
