#!/bin/bash
# ncu --set full captures of the three kernel families (one launch each) + 7B decode sanity
set -u
mkdir -p gpurun_out
echo "== ncu: persistent decode kernel"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_mega -s 2 -c 1 -o gpurun_out/ncu_mega_v5 -f \
   python scripts/decode_only.py --impl 1 --steps 4 --no-graph --reps 1 > gpurun_out/ncu_mega.log 2>&1
tail -1 gpurun_out/ncu_mega.log
echo "== ncu: persistent tcgen05 GEMM (llm.gate_up is the 10th shape: warm-ups + iterations precede it; take one launch late in the run)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_persistent -s 200 -c 1 -o gpurun_out/ncu_gemm_tc -f \
   python scripts/gemm_bench.py --no-cublas --iters 20 > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
echo "== ncu: ViT full attention (8th flash_attn launch = block 7) and one window-attention launch"
PADT_BENCH_LITE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:flash_attn -s 6 -c 2 -o gpurun_out/ncu_attn -f \
   python bench.py --steps 1 --warmup 1 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log
for f in ncu_mega_v5 ncu_gemm_tc ncu_attn; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
echo "== 7B decode sanity"
timeout 600 python scripts/mega_sweep.py --model PaDT_Pro_7B --steps 32 --reps 2 --configs "16,77,0;0,5,0" 2>&1 | grep '^{' | tee gpurun_out/mega_7b.jsonl
ls -la gpurun_out | head -30
