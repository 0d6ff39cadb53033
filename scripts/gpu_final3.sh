#!/bin/bash
# end-of-round validation: full GPU suite, smoke, both bench arms, vl_decode, ncu of the GEMM at two shapes
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_all.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench.json | cut -c1-400
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-200
echo "== vl_decode"
timeout 200 python scripts/vl_decode_bench.py 2>&1 | grep '^{' | tee gpurun_out/vl_decode.jsonl
echo "== gemm"
timeout 200 python scripts/gemm_bench.py 2>&1 | grep '^{' > gpurun_out/gemm_final.jsonl; tail -1 gpurun_out/gemm_final.jsonl
echo "== ncu gemm (ViT qkv, LLM o_proj)"
timeout 300 ncu --set full --clock-control none -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/gemm_vit_qkv -f python scripts/gemm_bench.py --no-cublas --iters 1 --only vit.qkv > gpurun_out/gemm_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/gemm_llm_o -f python scripts/gemm_bench.py --no-cublas --iters 1 --only llm.o > gpurun_out/gemm_ncu2.log 2>&1
tail -1 gpurun_out/gemm_ncu1.log gpurun_out/gemm_ncu2.log
