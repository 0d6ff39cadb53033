#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_all.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench.log
echo "== ncu full: persistent decode kernel"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_mega -s 2 -c 1 -o gpurun_out/mega_v3 -f \
   python scripts/decode_only.py --impl 1 --steps 4 --no-graph --reps 1 > gpurun_out/ncu_mega.log 2>&1
tail -2 gpurun_out/ncu_mega.log
echo "== ncu launch list (lite bench)"
PADT_BENCH_LITE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 4 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
wc -l gpurun_out/launches.csv
