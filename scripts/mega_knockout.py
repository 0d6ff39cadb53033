"""Knock-out timing of the persistent decode kernel: ms per token with parts of the kernel switched off through PADT_MEGA_TIMELINE
(4 = no grid barrier, 8 = no attention phase, 64 = no mma, 2 = no epilogue / mma).  Results of a knocked-out run are garbage; only the
time is meaningful.  python scripts/mega_knockout.py [--model PaDT_Pro_3B] [--steps 64]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import build_prompt, L_PROMPT, GRID, decode_bytes_per_step, measured_peaks
from padt_b200.config import PRESETS
from padt_b200.modeling import PaDTForConditionalGeneration

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="PaDT_Pro_3B")
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg = PRESETS[a.model]()
model = PaDTForConditionalGeneration.from_random(cfg, seed=0, device="cuda:0")
model.engine.set_option("decode_impl", 1)
model.engine.set_option("use_graph", 1)
P = GRID[0] * GRID[1] * GRID[2]
g = torch.Generator().manual_seed(1)
pv = torch.randn((P, cfg.vision.patch_dim), generator=g).to(torch.bfloat16).cuda()
ids = build_prompt(cfg, L_PROMPT, P // 4, 100).cuda()
mask = torch.ones_like(ids)
grid = torch.tensor([GRID]).cuda()
peaks, _ = measured_peaks()
nbytes = decode_bytes_per_step(cfg, P // 4, L_PROMPT + a.steps // 2)
NAMES = {0: "full kernel", 4: "no grid barrier", 8: "no attention phase", 64: "no mma", 12: "no barrier, no attention", 68: "no barrier, no mma",
         72: "no attention, no mma", 76: "no barrier, no attention, no mma (= stream + staging + epilogue)", 2: "no epilogue / mma", 14: "no barrier, no attention, no epilogue / mma"}
# knock-outs WITHOUT the grid barrier (bit 4) deadlock: the split-KV counters and the shared down_proj groups rely on the phase order
for dbg in [0, 8, 64, 72, 2, 0]:
    if dbg:
        os.environ["PADT_MEGA_TIMELINE"] = str(dbg)
    else:
        os.environ.pop("PADT_MEGA_TIMELINE", None)
    res = []
    for _ in range(a.reps):
        model.generate(input_ids=ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid, max_new_tokens=a.steps + 1, return_dict_in_generate=True,
                       eos_token_id=[], _prefill_only=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.engine.decode(a.steps, finish=True)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / a.steps)
    ms = sorted(res)[len(res) // 2]
    print(json.dumps({"knockout": dbg, "what": NAMES.get(dbg, ""), "ms_per_token": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peaks["hbm_gbs"], 4)}), flush=True)
