#!/usr/bin/env python
"""Time vl_decode (PaDT decoder: boxes / scores / masks from VRT hidden states) at the bench workload: 1344x1344 image, N objects x 5 VRTs.
   python scripts/vl_decode_bench.py [--objects 1,4,8]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from padt_b200.config import PRESETS
from padt_b200.modeling import PaDTForConditionalGeneration
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="PaDT_Pro_3B")
ap.add_argument("--objects", default="1,4,8")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
cfg = PRESETS[a.model]()
model = PaDTForConditionalGeneration.from_random(cfg, seed=0, device="cuda:0")
side = 96
P = side * side
g = torch.Generator().manual_seed(1)
pv = torch.randn((P, cfg.vision.patch_dim), generator=g).to(torch.bfloat16).cuda()
grid = torch.tensor([[1, side, side]])
emb, high, pe = model.engine.vit_forward(pv, grid)
protos = model.engine.prototypes(emb)
torch.cuda.synchronize()
D = cfg.text.hidden_size
for n in [int(x) for x in a.objects.split(",")]:
    feats = [[torch.randn((5, D), generator=g).to(torch.bfloat16).cuda() for _ in range(n)]]
    for _ in range(2):
        out = model.vl_decode(feats, protos, high, grid.cuda(), pe)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = model.vl_decode(feats, protos, high, grid.cuda(), pe); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(json.dumps({"objects": n, "ms": round(ms, 3), "ms_per_object": round(ms / n, 3), "ref_tflop_per_object": 0.332,
                      "tflops_vs_reference_flops": round(0.332 * n / ms * 1e3, 1), "mask_shape": list(out["pred_mask"].shape)}), flush=True)
