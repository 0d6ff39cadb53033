#!/usr/bin/env python
"""Sweep the persistent decode kernel's run-time knobs in ONE process (weights are initialised once).
The kernel reads PADT_MEGA_PF / PADT_MEGA_FLAGS / PADT_MEGA_PF_PACE / PADT_MEGA_TIMELINE at launch (= at graph capture, which
happens after every prefill), so setting os.environ between runs is enough.

    python scripts/mega_sweep.py [--steps 64] [--configs "pf,flags,pace;..."] [--trace "pf,flags,pace;..."]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import build_prompt, L_PROMPT, GRID, decode_bytes_per_step, measured_peaks
from padt_b200.config import PRESETS
from padt_b200.modeling import PaDTForConditionalGeneration

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--model", default="PaDT_Pro_3B")
ap.add_argument("--configs", default="")
ap.add_argument("--trace", default="")
ap.add_argument("--out", default="gpurun_out/mega_sweep.jsonl")
ap.add_argument("--lib", default="", help="load this build of the library instead of padt_b200/libpadt_b200.so (A/B runs)")
a = ap.parse_args()

# (prefetch depth in stages, PADT_MEGA_FLAGS bit mask, prefetch pace in cycles); flags: 1 evict_first, 2 no register-resident
# activations, 4 K/V tile prefetch, 8 per-line L2 prefetch, 16 evict_last prefetch, 64 minimal-traffic staging, 128 balanced down_proj
DEFAULT = [(16, 205, 0), (0, 205, 0), (8, 205, 0), (24, 205, 0), (16, 77, 0), (16, 141, 0), (16, 201, 0), (16, 204, 0), (16, 207, 0), (16, 197, 0)]


def parse_cfgs(s):
    return [tuple(int(x) for x in c.split(",")) for c in s.split(";") if c.strip()]


if a.lib:
    from padt_b200 import _capi
    _capi.LIB_PATH = os.path.abspath(a.lib)
cfgs = parse_cfgs(a.configs) if a.configs else DEFAULT
cfg = PRESETS[a.model]()
model = PaDTForConditionalGeneration.from_random(cfg, seed=0, device="cuda:0")
model.engine.set_option("decode_impl", 1)
P = GRID[0] * GRID[1] * GRID[2]
g = torch.Generator().manual_seed(1)
pv = torch.randn((P, cfg.vision.patch_dim), generator=g).to(torch.bfloat16).cuda()
ids = build_prompt(cfg, L_PROMPT, P // 4, 100).cuda()
mask = torch.ones_like(ids)
grid = torch.tensor([GRID]).cuda()
peaks, _ = measured_peaks()
nbytes = decode_bytes_per_step(cfg, P // 4, L_PROMPT + a.steps // 2)


def run(pf, flags, pace, steps, reps, graph=True, timeline=0):
    os.environ["PADT_MEGA_PF"] = str(pf)
    os.environ["PADT_MEGA_FLAGS"] = str(flags)
    os.environ["PADT_MEGA_PF_PACE"] = str(pace)
    if timeline:
        os.environ["PADT_MEGA_TIMELINE"] = str(timeline)
    else:
        os.environ.pop("PADT_MEGA_TIMELINE", None)
    model.engine.set_option("use_graph", 1 if graph else 0)
    res, toks = [], None
    for _ in range(reps):
        out = model.generate(input_ids=ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid, max_new_tokens=steps + 1,
                             return_dict_in_generate=True, eos_token_id=[], _prefill_only=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.engine.decode(steps, finish=True)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / steps)
        toks = model.engine._gen_keep["tokens"][:steps].flatten().tolist()
    return sorted(res)[len(res) // 2], res, toks


os.makedirs(os.path.dirname(a.out), exist_ok=True)
ref_toks = None
with open(a.out, "w") as f:
    for (pf, flags, pace) in cfgs:
        ms, allr, toks = run(pf, flags, pace, a.steps, a.reps)
        same = None
        if toks is not None:
            if ref_toks is None:
                ref_toks = toks
            same = toks == ref_toks
        line = {"lib": os.path.basename(a.lib) if a.lib else "tree", "pf": pf, "flags": flags, "pace": pace, "ms_per_step": round(ms, 4), "all": [round(x, 4) for x in allr],
                "GBps": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peaks["hbm_gbs"], 4), "tokens_same_as_first": same}
        print(json.dumps(line), flush=True)
        f.write(json.dumps(line) + "\n")

for (pf, flags, pace) in parse_cfgs(a.trace):
    print(f"== timeline pf={pf} flags={flags} pace={pace}", flush=True)
    run(pf, flags, pace, 3, 1, graph=False, timeline=17)
    torch.cuda.synchronize()
