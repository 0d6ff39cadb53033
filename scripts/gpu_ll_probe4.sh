#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ll_probe4.log
for pf in 16 8 24 32 0; do
  echo "== PADT_MEGA_PF=$pf" | tee -a gpurun_out/ll_probe4.log
  timeout 300 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "$pf,205,0" --out gpurun_out/ll4.jsonl 2>&1 | grep '^{' | tee -a gpurun_out/ll_probe4.log
done
timeout 300 python scripts/mega_sweep.py --steps 8 --reps 1 --configs "16,205,0" --trace "16,205,0" --out gpurun_out/ll_t.jsonl 2>&1 | grep -v "^{" | tail -6 | tee -a gpurun_out/ll_probe4.log
