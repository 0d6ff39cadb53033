#!/bin/bash
# round-end style validation: GPU tests, smoke, bench (N=1), 7B decode timing
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_all.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
echo "== 7B decode"
timeout 300 python scripts/mega_sweep.py --model PaDT_Pro_7B --steps 64 --reps 3 --configs "16,205,0" 2>&1 | grep '^{' | tee gpurun_out/mega_7b.jsonl
echo "== launch list"
PADT_BENCH_LITE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --new-tokens 4 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>/dev/null; head -14 gpurun_out/launches_summary.txt
