#!/bin/bash
set -u
mkdir -p gpurun_out
PADT_MEGA_TIMELINE=17 python scripts/decode_only.py --impl 1 --steps 2 --no-graph --reps 1 2>&1 | grep "trace\]" | tail -260 > gpurun_out/mega_trace.log
wc -l gpurun_out/mega_trace.log
