#!/bin/bash
# full GPU suite + the default bench line (N = 1)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/full_tests.log
cat gpurun_out/full_tests.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
