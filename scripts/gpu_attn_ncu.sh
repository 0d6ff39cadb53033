#!/bin/bash
set -x
mkdir -p gpurun_out
PADT_ATTN_TC=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -s 3 -c 1 -o gpurun_out/attn_tc_v1 -f python scripts/attn_bench.py > gpurun_out/attn_ncu.log 2>&1
tail -5 gpurun_out/attn_ncu.log
