#!/bin/bash
# one ncu --set full capture of the persistent decode kernel (third launch), raw metrics exported to CSV for reading on CPU
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_mega -s 2 -c 1 -o gpurun_out/mega_ncu -f \
   python scripts/decode_only.py --impl 1 --steps 4 --no-graph --reps 1 > gpurun_out/ncu_mega.log 2>&1
tail -2 gpurun_out/ncu_mega.log
ncu -i gpurun_out/mega_ncu.ncu-rep --page raw --csv > gpurun_out/mega_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/
