#!/bin/bash
# same-box A/B of library variants: scripts/gpu_ab.sh "<configs>" lib1 lib2 ...   ("tree" = padt_b200/libpadt_b200.so)
set -u
mkdir -p gpurun_out
CFG="$1"; shift
: > gpurun_out/ab.log
for lib in "$@"; do
  if [ "$lib" = "tree" ]; then L=""; else L="--lib padt_b200/variants/lib_$lib.so"; fi
  timeout 300 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "$CFG" $L --out gpurun_out/ab_$lib.jsonl 2>&1 | grep '^{' | tee -a gpurun_out/ab.log
done
