#!/bin/bash
# same-box A/B: LL kernel (tree; replicas 16 and 1) vs the round-1 barrier kernel, with and without the descriptor-load fix
mkdir -p gpurun_out
: > gpurun_out/ab3.log
run() {  # label, extra env, lib arg
  echo "== $1" | tee -a gpurun_out/ab3.log
  env $2 timeout 300 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "16,205,0" $3 --out gpurun_out/ab3_tmp.jsonl 2>&1 | grep '^{' | tee -a gpurun_out/ab3.log
}
run "LL nrep=16" "PADT_MEGA_LL_REP=16" ""
run "LL nrep=1" "PADT_MEGA_LL_REP=1" ""
run "barrier (round 1)" "X=1" "--lib padt_b200/variants/lib_barrier.so"
run "barrier + descriptor fix" "X=1" "--lib padt_b200/variants/lib_barrierfix.so"
run "LL nrep=2" "PADT_MEGA_LL_REP=2" ""
