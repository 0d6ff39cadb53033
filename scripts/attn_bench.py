"""Times the attention op at the two long-sequence shapes of config 2 (ViT full attention: 9216 tokens x 16 heads x d 80, non-causal;
LLM prefill: L tokens x 16 heads (2 kv heads) x d 128, causal) through the C ABI.  Run once with PADT_ATTN_TC=0 and once with 1."""
import math, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from padt_b200 import _capi

lib = _capi.load()
st = torch.cuda.current_stream().cuda_stream
i32 = lambda xs: torch.tensor(xs, dtype=torch.int32, device="cuda")


def run(name, T, H, Hkv, hd, causal, iters=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn((T, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn((T, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn((T, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    o = torch.zeros_like(q)
    qo, ql = i32([0]), i32([T])
    call = lambda: _capi.check(lib.padt_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), H * hd, hd, Hkv * hd, hd, Hkv * hd, hd,
                                                     H * hd, hd, qo.data_ptr(), ql.data_ptr(), qo.data_ptr(), ql.data_ptr(), None, 1, H, H // Hkv, T, hd,
                                                     int(causal), 1.0 / math.sqrt(hd), st), "attention")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 4.0 * T * T * hd * H * (0.5 if causal else 1.0)
    print(json.dumps(dict(shape=name, tc=os.environ.get("PADT_ATTN_TC", "1"), ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1),
                          checksum=float(o.float().abs().mean()))))


run("vit_full_9216x16x80", 9216, 16, 16, 80, False)
run("prefill_2400x16x128_causal", 2400, 16, 2, 128, True)
run("prefill_4096x16x128_causal", 4096, 16, 2, 128, True)
