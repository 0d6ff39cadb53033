#!/bin/bash
# LL decode kernel: what bounds it?  dbg bits: 8 = no attention phase, 64 = no mma in the consumers (results wrong, time meaningful)
mkdir -p gpurun_out
: > gpurun_out/ll_probe2.log
for cfg in "0 16" "72 16" "8 16" "0 0" "0 24" "0 8"; do
  set -- $cfg
  echo "== PADT_MEGA_TIMELINE=$1 PADT_MEGA_PF=$2" | tee -a gpurun_out/ll_probe2.log
  if [ "$1" = "0" ]; then unset PADT_MEGA_TIMELINE; else export PADT_MEGA_TIMELINE=$1; fi
  timeout 300 python scripts/mega_sweep.py --steps 64 --reps 2 --configs "$2,205,0" --out gpurun_out/ll2.jsonl 2>&1 | grep '^{' | tee -a gpurun_out/ll_probe2.log
done
