#!/bin/bash
# ncu launch list of one short generate (cold-cache, serialised: compare SHARES, not absolutes)
set -u
mkdir -p gpurun_out
export PADT_BENCH_LITE=1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 4 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -3 gpurun_out/launches_bench.log
wc -l gpurun_out/launches.csv
