#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_all.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench.log
PADT_BENCH_LITE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 4 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
wc -l gpurun_out/launches.csv
