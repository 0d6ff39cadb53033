#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/decode_only.py --impl 1 --steps 32 2>&1 | tail -1 | tee gpurun_out/decode_mega.json
python scripts/decode_only.py --impl 0 --steps 32 2>&1 | tail -1 | tee gpurun_out/decode_v1.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decode_mega -s 2 -c 1 -o gpurun_out/mega -f \
   python scripts/decode_only.py --impl 1 --steps 4 --no-graph --reps 1 > gpurun_out/ncu_mega.log 2>&1
tail -3 gpurun_out/ncu_mega.log
ls -la gpurun_out/
