#!/bin/bash
set -u
mkdir -p gpurun_out
bash scripts/gpu_ab.sh "16,13,0" sw0st0 tree
echo "== pytest gpu"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_all.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
echo "== launch list"
PADT_BENCH_LITE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 4 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -30 | tee gpurun_out/launches_summary.txt
