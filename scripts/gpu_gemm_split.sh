#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -15
PADT_GEMM_SPLITK=0 timeout 200 python scripts/gemm_bench.py --no-cublas 2>&1 | tail -12 > gpurun_out/gemm_split0.log
timeout 200 python scripts/gemm_bench.py 2>&1 | tail -12 > gpurun_out/gemm_split1.log
cat gpurun_out/gemm_split0.log gpurun_out/gemm_split1.log
