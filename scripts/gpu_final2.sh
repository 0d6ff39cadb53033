#!/bin/bash
# round-2 validation pass: full GPU suite, smoke, bench (N = 1, both arms), 7B decode, vl_decode timing, launch list, ncu of the tcgen05 attention
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_all.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench.json
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_reference.json
echo "== 7B decode"
timeout 300 python scripts/mega_sweep.py --model PaDT_Pro_7B --steps 64 --reps 3 --configs "16,205,0" 2>&1 | grep '^{' | tee gpurun_out/mega_7b.jsonl
echo "== vl_decode"
timeout 300 python scripts/vl_decode_bench.py 2>&1 | grep '^{' | tee gpurun_out/vl_decode.jsonl
echo "== attention / gemm micro"
timeout 120 python scripts/attn_bench.py 2>&1 | grep '^{' | tee gpurun_out/attn_bench.jsonl
echo "== launch list"
PADT_BENCH_LITE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --new-tokens 4 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>/dev/null; head -16 gpurun_out/launches_summary.txt
echo "== ncu attention"
PADT_ATTN_TC=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -s 3 -c 2 -o gpurun_out/attn_tc_v2 -f python scripts/attn_bench.py > gpurun_out/attn_ncu.log 2>&1
tail -2 gpurun_out/attn_ncu.log
