#!/bin/bash
# tcgen05 attention: op-level parity first (under a timeout: a wrong mbarrier protocol hangs), then timing against the mma.sync kernel
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -25 > gpurun_out/attn_tc_tests.log
cat gpurun_out/attn_tc_tests.log
(PADT_ATTN_TC=0 timeout 120 python scripts/attn_bench.py; PADT_ATTN_TC=1 timeout 120 python scripts/attn_bench.py) 2>&1 | tee gpurun_out/attn_tc_bench.log
