#!/bin/bash
# LL decode kernel: replica count / poll back-off sweep (one box)
mkdir -p gpurun_out
: > gpurun_out/ll_probe3.log
for cfg in "16 0" "1 0" "4 0" "8 0" "16 100" "16 300"; do
  set -- $cfg
  echo "== PADT_MEGA_LL_REP=$1 PADT_MEGA_LL_SLEEP=$2" | tee -a gpurun_out/ll_probe3.log
  PADT_MEGA_LL_REP=$1 PADT_MEGA_LL_SLEEP=$2 timeout 300 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "16,205,0" --out gpurun_out/ll3.jsonl 2>&1 | grep '^{' | tee -a gpurun_out/ll_probe3.log
done
PADT_MEGA_LL_REP=16 timeout 300 python scripts/mega_sweep.py --steps 8 --reps 1 --configs "16,205,0" --trace "16,205,0" --out gpurun_out/ll_t.jsonl 2>&1 | grep -v "^{" | tail -6 | tee -a gpurun_out/ll_probe3.log
