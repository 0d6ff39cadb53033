#!/bin/bash
# LL decode kernel: timeline + poll back-off sweep (one box)
mkdir -p gpurun_out
: > gpurun_out/ll_probe.log
for s in 0 200 1000; do
  echo "== PADT_MEGA_LL_SLEEP=$s" | tee -a gpurun_out/ll_probe.log
  PADT_MEGA_LL_SLEEP=$s timeout 300 python scripts/mega_sweep.py --steps 64 --reps 2 --configs "16,205,0" --out gpurun_out/ll_$s.jsonl 2>&1 | grep '^{' | tee -a gpurun_out/ll_probe.log
done
PADT_MEGA_LL_SLEEP=0 timeout 300 python scripts/mega_sweep.py --steps 8 --reps 1 --configs "16,205,0" --trace "16,205,0" --out gpurun_out/ll_t.jsonl 2>&1 | grep -v "^{" | tail -12 | tee -a gpurun_out/ll_probe.log
