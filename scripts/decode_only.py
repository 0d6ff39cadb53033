#!/usr/bin/env python
"""Prefill once at the bench workload, then time / profile decode steps only.
   python scripts/decode_only.py --impl 1 --steps 64 [--no-graph] [--model PaDT_Pro_3B] [--batch 1]"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import build_prompt, L_PROMPT, GRID, decode_bytes_per_step, measured_peaks
from padt_b200.config import PRESETS
from padt_b200.modeling import PaDTForConditionalGeneration

ap = argparse.ArgumentParser()
ap.add_argument("--impl", type=int, default=1)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--model", default="PaDT_Pro_3B")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg = PRESETS[a.model]()
model = PaDTForConditionalGeneration.from_random(cfg, seed=0, device="cuda:0")
model.engine.set_option("decode_impl", a.impl)
model.engine.set_option("use_graph", 0 if a.no_graph else 1)
P = GRID[0] * GRID[1] * GRID[2]
g = torch.Generator().manual_seed(1)
pv = torch.randn((P, cfg.vision.patch_dim), generator=g).to(torch.bfloat16).cuda()
ids = build_prompt(cfg, L_PROMPT, P // 4, 100).cuda()
mask = torch.ones_like(ids)
grid = torch.tensor([GRID]).cuda()
T = a.steps + 1
res = []
for rep in range(a.reps):
    model.generate(input_ids=ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid, max_new_tokens=T, return_dict_in_generate=True,
                   eos_token_id=[], _prefill_only=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.engine.decode(a.steps, finish=True)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / a.steps)
ms = sorted(res)[len(res) // 2]
peaks, _ = measured_peaks()
b = decode_bytes_per_step(cfg, P // 4, L_PROMPT + a.steps // 2)
print(json.dumps({"impl": a.impl, "graph": not a.no_graph, "ms_per_step": round(ms, 4), "all": [round(x, 4) for x in res],
                  "GBps": round(b / ms / 1e6, 1), "frac": round(b / ms / 1e6 / peaks["hbm_gbs"], 4)}))
