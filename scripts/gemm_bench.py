#!/usr/bin/env python
"""Time the dense projections of one 1344x1344 request (ViT blocks + LLM prefill shapes) through padt_op_gemm, next to cuBLAS
(torch.matmul, no epilogue) on the same shapes.   python scripts/gemm_bench.py [--lib path.so] [--iters 20]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default="")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--no-cublas", action="store_true")
ap.add_argument("--only", default="", help="comma-separated shape names")
a = ap.parse_args()
from padt_b200 import _capi
if a.lib:
    _capi.LIB_PATH = os.path.abspath(a.lib)
lib = _capi.load()
# (name, M, N, K, Kpad, epi, bias, count per request)
SHAPES = [
    ("vit.qkv", 9216, 3840, 1280, 1280, 0, True, 32), ("vit.proj", 9216, 1280, 1280, 1280, 2, True, 32),
    ("vit.gate_up", 9216, 6840, 1280, 1280, 3, True, 32), ("vit.down", 9216, 1280, 3420, 3424, 2, True, 32),
    ("vit.patch_embed", 9216, 1280, 1176, 1176, 0, False, 1), ("vit.merger0", 2304, 5120, 5120, 5120, 1, True, 1), ("vit.merger1", 2304, 2048, 5120, 5120, 0, True, 1),
    ("llm.qkv", 2350, 2560, 2048, 2048, 0, True, 36), ("llm.o", 2350, 2048, 2048, 2048, 2, False, 36),
    ("llm.gate_up", 2350, 22016, 2048, 2048, 3, False, 36), ("llm.down", 2350, 2048, 11008, 11008, 2, False, 36),
]
st = torch.cuda.current_stream().cuda_stream
tot_ms = tot_cublas = tot_flop = 0.0
rows = []
for name, M, N, K, Kp, epi, has_bias, cnt in SHAPES:
    if a.only and name not in a.only.split(","):
        continue
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.zeros((M, Kp), device="cuda", dtype=torch.bfloat16); A[:, :K] = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    W = torch.zeros((N, Kp), device="cuda", dtype=torch.bfloat16); W[:, :K] = (0.02 * torch.randn((N, K), device="cuda", generator=g)).to(torch.bfloat16)
    bias = torch.randn((N,), device="cuda", generator=g).to(torch.bfloat16) if has_bias else None
    n_out = N // 2 if epi == 3 else N
    ldc = (n_out + 7) // 8 * 8
    C = torch.zeros((M, ldc), device="cuda", dtype=torch.bfloat16)
    resid = torch.randn((M, n_out), device="cuda", generator=g).to(torch.bfloat16) if epi == 2 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def ours():
        _capi.check(lib.padt_op_gemm(epi, A.data_ptr(), Kp, W.data_ptr(), Kp, None if bias is None else bias.data_ptr(),
                                     None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0), C.data_ptr(), ldc, None,
                                     M, N, K, st), "gemm")

    def cublas():
        torch.matmul(A, W.t())

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.iters):
            flush.fill_(1)       # L2 flush between iterations (256 MB > 126 MB L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]
    ms = timeit(ours)
    mc = 0.0 if a.no_cublas else timeit(cublas)
    flop = 2.0 * M * N * K
    rows.append(dict(name=name, M=M, N=N, K=K, epi=epi, ms=round(ms, 4), tflops=round(flop / ms / 1e9, 1), cublas_ms=round(mc, 4),
                     cublas_tflops=round(flop / mc / 1e9, 1) if mc else None, per_request_ms=round(ms * cnt, 2)))
    tot_ms += ms * cnt; tot_cublas += mc * cnt; tot_flop += flop * cnt
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps(dict(lib=os.path.basename(a.lib) if a.lib else "tree", total_ms_per_request=round(tot_ms, 2), cublas_total_ms=round(tot_cublas, 2),
                      tflops=round(tot_flop / tot_ms / 1e9, 1))))
