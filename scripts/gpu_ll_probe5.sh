#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ll_probe5.log
timeout 300 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "16,205,0" --trace "16,205,0" --out gpurun_out/ll5.jsonl 2>&1 | grep -v "^==" | tail -8 | tee -a gpurun_out/ll_probe5.log
