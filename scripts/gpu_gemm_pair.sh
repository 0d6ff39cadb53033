#!/bin/bash
# CTA-pair (cta_group::2) GEMM: parity under a timeout (every wait in the kernel traps instead of hanging), then per-shape timing next to the single-CTA kernel
set -x
mkdir -p gpurun_out
PADT_GEMM_PAIR=2 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -15
PADT_GEMM_PAIR=0 timeout 200 python scripts/gemm_bench.py --no-cublas 2>&1 | tail -12 > gpurun_out/gemm_pair0.log
PADT_GEMM_PAIR=2 timeout 200 python scripts/gemm_bench.py --no-cublas 2>&1 | tail -12 > gpurun_out/gemm_pair2.log
PADT_GEMM_PAIR=1 timeout 200 python scripts/gemm_bench.py 2>&1 | tail -12 > gpurun_out/gemm_pair1.log
python - <<'PY'
import json
def load(f):
    return {json.loads(l)["name"]: json.loads(l) for l in open(f) if l.startswith("{") and '"name"' in l}, [json.loads(l) for l in open(f) if '"total_ms_per_request"' in l]
a, ta = load("gpurun_out/gemm_pair0.log"); b, tb = load("gpurun_out/gemm_pair2.log"); c, tc = load("gpurun_out/gemm_pair1.log")
for k in a:
    print(f"{k:16s} single {a[k]['ms']*1e3:7.1f} us  pair {b.get(k,{}).get('ms',0)*1e3:7.1f} us  auto {c.get(k,{}).get('ms',0)*1e3:7.1f} us  cublas {c.get(k,{}).get('cublas_ms',0)*1e3:7.1f} us")
print("totals per request (ms): single", ta and ta[0]["total_ms_per_request"], "pair", tb and tb[0]["total_ms_per_request"], "auto", tc and tc[0]["total_ms_per_request"], "cublas", tc and tc[0]["cublas_total_ms"])
PY
