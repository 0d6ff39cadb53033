"""Golden-vector generator: runs the UNMODIFIED reference classes (``/root/reference/src/PaDT/models``) under
``oracle/ref_shim.py`` on seeded tiny configs, checks ``oracle/padt_oracle.py`` against them bit-for-bit and
writes the fixtures ``tests/golden/*.pt``.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference):   python oracle/make_golden.py

What is pinned by the reference itself (its own Python, verbatim): custom_visual_forward (padt.py:48-106),
forward_main prefill + decode branches (padt.py:147-340) incl. prototype projection, extended embedding/logits and
mask, vl_decode (padt.py:342-412), PaDTDecoder (padt_decoder.py), processor id shifts and the parser.
What the shim has to supply because the pinned third-party versions are absent: FA2 attention
numerics (restated in oracle.fa2_attention, used by both sides), flash-attn's apply_rotary_emb, and the
transformers-4.50.0 ``get_rope_index`` (restated; KAT in tests/test_rope_index.py).  The greedy loop restates
padt.py:671-757 because the copied HF-4.50 ``generate`` cannot run on transformers 5.5.0.
"""
import os
import sys

# same oneDNN kernel family as the tests (tests/conftest.py): bit-exactness of bf16 GEMMs depends on the ISA path
os.environ.setdefault("ONEDNN_MAX_CPU_ISA", "AVX512_CORE")

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
import padt_oracle as O  # noqa: E402
from padt_b200.config import padt_small, padt_tiny  # noqa: E402
from padt_b200.weights import random_weights  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


from ref_runner import build_reference, ref_greedy  # noqa: E402,F401  (shared with bench.py and the GPU floor test)


def make_inputs(cfg, grids, prompt_lens, seed, pre_len=3, text_hi=None):
    """Chat-template-shaped rows: text, <vision_start>, M x <image_pad>, <vision_end>, text; left padded."""
    g = torch.Generator().manual_seed(seed)
    V = cfg.text.vocab_size
    text_hi = V - 16 if text_hi is None else text_hi
    rows = []
    for (t, h, w), pl in zip(grids, prompt_lens):
        M = t * h * w // 4
        pre = torch.randint(0, text_hi, (pre_len,), generator=g).tolist()
        post = torch.randint(0, text_hi, (pl,), generator=g).tolist()
        rows.append(pre + [cfg.vision_start_token_id] + [cfg.image_token_id] * M + [cfg.vision_end_token_id] + post)
    L = max(len(r) for r in rows)
    ids = torch.full((len(rows), L), cfg.pad_token_id, dtype=torch.long)
    mask = torch.zeros((len(rows), L), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, L - len(r):] = torch.tensor(r)
        mask[i, L - len(r):] = 1
    P = sum(t * h * w for t, h, w in grids)
    pixel_values = torch.randn((P, cfg.vision.patch_dim), generator=g).to(torch.bfloat16)
    grid = torch.tensor(grids, dtype=torch.long)
    return ids, mask, pixel_values, grid


MISMATCHES = None      # list -> collect instead of raising (the full-size case: a mismatch after minutes of CPU work is reported, not fatal)


def eq(a, b, what):
    if MISMATCHES is not None:
        try:
            _eq(a, b, what)
        except AssertionError as e:
            MISMATCHES.append(str(e))
            print("[golden] MISMATCH", e)
        return
    _eq(a, b, what)


def _eq(a, b, what):
    if isinstance(a, (tuple, list)):
        assert len(a) == len(b), what
        for i, (x, y) in enumerate(zip(a, b)):
            _eq(x, y, f"{what}[{i}]")
        return
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    same = torch.equal(a, b) if a.dtype != torch.float32 and a.dtype != torch.bfloat16 else \
        bool(((a == b) | (torch.isinf(a) & torch.isinf(b))).all())
    if not same:
        d = (a.float() - b.float()).abs()
        d = d[torch.isfinite(d)]
        raise AssertionError(f"oracle != reference for {what}: max abs diff {float(d.max())}")


def case(name, cfg, wseed, grids, prompt_lens, steps, iseed, objects, pre_len=3, text_hi=None, store_inputs=True):
    import time
    t0 = time.time()
    w = random_weights(cfg, seed=wseed)
    print(f"[golden] {name}: weights {time.time() - t0:.0f}s", flush=True)
    model, proc = build_reference(cfg, w)
    print(f"[golden] {name}: reference model built {time.time() - t0:.0f}s", flush=True)
    ids, mask, pv, grid = make_inputs(cfg, grids, prompt_lens, iseed, pre_len=pre_len, text_hi=text_hi)
    V = cfg.text.vocab_size
    text_hi = V - 16 if text_hi is None else text_hi
    with torch.no_grad():
        # ---- ViT alone (V1-V6)
        r_emb, r_high, r_pe = model.visual(pv, grid_thw=grid)
        o_emb, o_high, o_pe = O.vit_forward(w, cfg, pv, grid)
        eq(o_emb, r_emb, "image_embeds"); eq(o_high, r_high, "high_res"); eq(o_pe[0], r_pe[0], "cos"); eq(o_pe[1], r_pe[1], "sin")
        # ---- free-running greedy (E1-E4, R1, T1, T2, S1)
        rg = ref_greedy(model, cfg, ids, mask, pv, grid, steps)
        og = O.greedy_generate(w, cfg, ids, mask, pv, grid, steps, eos_token_ids=[])
        eq(og["sequences"], rg["sequences"], "sequences")
        valid = mask.bool()
        eq(og["hidden_steps"][0][valid], rg["hidden_steps"][0][valid], "prefill hidden (non-pad rows)")
        for s in range(1, len(rg["hidden_steps"])):
            eq(og["hidden_steps"][s], rg["hidden_steps"][s], f"hidden step {s}")
        for s in range(steps):
            eq(og["logits_steps"][s], rg["logits_steps"][s], f"logits step {s}")
        eq(og["past_image_embeds"], rg["past_image_embeds"], "prototypes")
        eq(og["past_logit_mask"], rg["past_logit_mask"], "logit mask")
        # ---- teacher-forced run that feeds VRT ids back in (exercises ext[input_ids] for ids >= V)
        off = O.patch_offsets(cfg, grid)
        g = torch.Generator().manual_seed(iseed + 7)
        forced = torch.randint(0, text_hi, (len(grids), steps), generator=g)
        for b in range(len(grids)):
            Mb = off[b + 1] - off[b]
            pos = torch.randperm(steps, generator=g)[: steps // 2]
            forced[b, pos] = V + off[b] + torch.randint(0, Mb, (len(pos),), generator=g)
        rf = ref_greedy(model, cfg, ids, mask, pv, grid, steps, forced=forced)
        of = O.greedy_generate(w, cfg, ids, mask, pv, grid, steps, eos_token_ids=[], forced_tokens=forced)
        eq(of["sequences"], rf["sequences"], "forced sequences")
        eq(torch.stack(of["argmax_steps"]), torch.stack(rf["argmax_steps"]), "forced argmax")
        for s in range(1, len(rf["hidden_steps"])):
            eq(of["hidden_steps"][s], rf["hidden_steps"][s], f"forced hidden step {s}")
        # ---- vl_decode (D1-D4): VRT features = final-norm hidden of the step that predicted each VRT
        feats = []
        for b, objs in enumerate(objects):
            sample = []
            for (s0, n) in objs:
                sample.append(torch.cat([rf["hidden_steps"][s][b] for s in range(s0, s0 + n)], dim=0))
            feats.append(sample)
        rd = model.vl_decode(feats, rf["past_image_embeds"], rf["past_high_res_image_embeds"], grid, rf["past_visual_pe"])
        od = O.vl_decode(w, cfg, feats, of["past_image_embeds"], of["past_high_res_image_embeds"], grid, of["past_visual_pe"])
        for k in ("pred_boxes", "pred_score", "pred_mask"):
            eq(od[k], rd[k], k)
        eq(od["pred_mask_valid_hw"][0], rd["pred_mask_valid_hw"][0], "Hs"); eq(od["pred_mask_valid_hw"][1], rd["pred_mask_valid_hw"][1], "Ws")
        assert od["sample_idx"] == rd["sample_idx"]
    top = [torch.topk(l, 8, dim=-1) for l in rg["logits_steps"]]
    fx = dict(
        name=name, cfg_name=cfg.name, tie=cfg.text.tie_word_embeddings, vocab=V, wseed=wseed,
        input_ids=ids, attention_mask=mask, pixel_values=pv, grid=grid, steps=steps,
        image_embeds=r_emb, high_res=r_high, cos=r_pe[0], sin=r_pe[1],
        prototypes=rg["past_image_embeds"], sequences=rg["sequences"],
        last_hidden=torch.stack([h[:, -1, :] for h in rg["hidden_steps"]]),
        top8_values=torch.stack([t.values for t in top]), top8_indices=torch.stack([t.indices for t in top]),
        forced=forced, forced_argmax=torch.stack(rf["argmax_steps"]),
        forced_last_hidden=torch.stack([h[:, -1, :] for h in rf["hidden_steps"]]),
        forced_top2=torch.stack([torch.topk(l, 2, dim=-1).values for l in rf["logits_steps"]]),
        forced_top8_values=torch.stack([torch.topk(l, 8, dim=-1).values for l in rf["logits_steps"]]),
        forced_top8_indices=torch.stack([torch.topk(l, 8, dim=-1).indices for l in rf["logits_steps"]]),
        oracle_mismatches=list(MISMATCHES) if MISMATCHES is not None else [],
        objects=objects, pred_boxes=rd["pred_boxes"], pred_score=rd["pred_score"], pred_mask=rd["pred_mask"],
        Hs=rd["pred_mask_valid_hw"][0], Ws=rd["pred_mask_valid_hw"][1], sample_idx=rd["sample_idx"],
    )
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".pt")
    torch.save(fx, path)
    status = "oracle == reference (bit-exact)" if not fx["oracle_mismatches"] else f"{len(fx['oracle_mismatches'])} ORACLE MISMATCHES"
    print(f"[golden] {name}: {status}; wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); "
          f"sequence tail {rg['sequences'][0, -steps:].tolist()}; boxes {rd['pred_boxes'].float().min():.3f}..{rd['pred_boxes'].float().max():.3f} "
          f"scores {rd['pred_score'].float().flatten().tolist()} mask std {rd['pred_mask'].float().std():.3f} |max| {rd['pred_mask'].float().abs().max():.3f} "
          f"({time.time() - t0:.0f}s)", flush=True)


def vit_case(name, cfg, wseed, grid_hw, iseed, row_stride=16):
    """ViT only (V1-V6) at a large grid on a shallow ViT of full width: the 1344x1344 window plan (144 windows) and one
    full-attention block over 9216 tokens, through the UNMODIFIED custom_visual_forward.  The fixture keeps every
    `row_stride`-th high-res row, all merged embeddings and the fp32 rotary tables' checksum (the inputs are re-generated
    from the seed by the test; their SHA-256 is stored)."""
    import hashlib
    import time
    t0 = time.time()
    w = random_weights(cfg, seed=wseed)
    model, _ = build_reference(cfg, w)
    g = torch.Generator().manual_seed(iseed)
    P = grid_hw[0] * grid_hw[1]
    pv = torch.randn((P, cfg.vision.patch_dim), generator=g).to(torch.bfloat16)
    grid = torch.tensor([[1, grid_hw[0], grid_hw[1]]], dtype=torch.long)
    with torch.no_grad():
        r_emb, r_high, r_pe = model.visual(pv, grid_thw=grid)
        o_emb, o_high, o_pe = O.vit_forward(w, cfg, pv, grid)
    eq(o_emb, r_emb, "image_embeds"); eq(o_high, r_high, "high_res"); eq(o_pe[0], r_pe[0], "cos"); eq(o_pe[1], r_pe[1], "sin")
    rows = torch.arange(0, P, row_stride)
    fx = dict(name=name, wseed=wseed, iseed=iseed, grid=grid, pixel_sha256=hashlib.sha256(pv.view(torch.int16).numpy().tobytes()).hexdigest(),
              rows=rows, high_res_rows=r_high[rows].clone(), image_embeds=r_emb.clone(),
              cos_rows=r_pe[0][rows].clone(), sin_rows=r_pe[1][rows].clone(),
              high_res_absmax=float(r_high.float().abs().max()), high_res_rowsum=r_high.float().sum(dim=1),
              oracle_mismatches=list(MISMATCHES) if MISMATCHES is not None else [])
    path = os.path.join(GOLD, name + ".pt")
    torch.save(fx, path)
    print(f"[golden] {name}: ViT {grid_hw} oracle == reference: {not fx['oracle_mismatches']}; wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB) ({time.time() - t0:.0f}s)", flush=True)


def dec_case(name, cfg, wseed, grid_hw, objects, iseed):
    """PaDT decoder only (D1-D4) at full image size through the UNMODIFIED vl_decode / PaDTDecoder: 9216 high-res + 2304 low-res
    memory tokens, `objects` = VRT count per object.  Inputs come from the seed; the fixture holds the outputs and the input hash."""
    import hashlib
    import time
    t0 = time.time()
    w = random_weights(cfg, seed=wseed)
    model, _ = build_reference(cfg, w)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import decoder_inputs          # seeded inputs shared with tests/test_model_gpu.py
    feats, low, high, grid, pe = decoder_inputs(cfg, grid_hw, objects, iseed)
    with torch.no_grad():
        rd = model.vl_decode(feats, low, high, grid, pe)
        od = O.vl_decode(w, cfg, feats, low, high, grid, pe)
    for k in ("pred_boxes", "pred_score", "pred_mask"):
        eq(od[k], rd[k], k)
    sha = hashlib.sha256(b"".join(t.contiguous().view(torch.int16).numpy().tobytes() for t in [low, high] + feats[0])).hexdigest()
    fx = dict(name=name, cfg_name=cfg.name, wseed=wseed, iseed=iseed, grid_hw=list(grid_hw), objects=list(objects), input_sha256=sha,
              pred_boxes=rd["pred_boxes"], pred_score=rd["pred_score"], pred_mask=rd["pred_mask"], Hs=rd["pred_mask_valid_hw"][0],
              Ws=rd["pred_mask_valid_hw"][1], sample_idx=rd["sample_idx"], oracle_mismatches=list(MISMATCHES) if MISMATCHES is not None else [])
    path = os.path.join(GOLD, name + ".pt")
    torch.save(fx, path)
    print(f"[golden] {name}: decoder {grid_hw} x {objects} oracle == reference: {not fx['oracle_mismatches']}; wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB); boxes {rd['pred_boxes'].float().min():.3f}..{rd['pred_boxes'].float().max():.3f} scores "
          f"{rd['pred_score'].float().flatten().tolist()} mask |max| {rd['pred_mask'].float().abs().max():.3f} ({time.time() - t0:.0f}s)", flush=True)


def parser_cases(path_name="parser_cases", n_cases=200, seed=11):
    """P1/P2 cross-run: random piece streams through the UNMODIFIED reference parser / processor wrapper
    (padt_processor.py:4-151) and through padt_b200.processor; must agree; the reference's outputs are committed as JSON."""
    import json
    import random
    import re
    from padt_b200 import processor as ours
    _, _, proc = ref_shim.install()

    class Tok:
        eos_token = "<|im_end|>"

        def __init__(self, n):
            self.vocab = {f"tok{i}": i for i in range(n)}

        def get_vocab(self):
            return self.vocab

        def add_tokens(self, toks):
            for t in toks:
                sname = t if isinstance(t, str) else t.content
                if sname not in self.vocab:
                    self.vocab[sname] = len(self.vocab)

    class Proc:
        def __init__(self, V, strings):
            self.tokenizer = Tok(V); self.V = V; self.strings = strings

        def batch_decode(self, ids):
            return [self.strings[int(i)] if int(i) < self.V else f"<|VRT_{int(i) - self.V}|>" for i in ids]

        def __call__(self, text=None, return_tensors=None, **kw):
            return {"input_ids": torch.tensor([[self.V + int(k) for k in re.findall(r"<\|VRT_(\d+)\|>", text)]])}

    rnd = random.Random(seed)
    alphabet = ["<", "answer", ">", "</", ' "', '"', " cat", " dog", " on", " the", " left", ".", " ", "<|im_end|>", "think", '"a', 'b"', "<think", "\n"]
    V = len(alphabet)
    strings = {i: a for i, a in enumerate(alphabet)}
    cases = []
    for ci in range(n_cases):
        B = rnd.choice([1, 2, 3])
        T = rnd.randint(4, 40)
        grids = [(1, 2 * rnd.randint(2, 8), 2 * rnd.randint(2, 8)) for _ in range(B)]
        grid = torch.tensor(grids)
        Ms = [g[1] * g[2] // 4 for g in grids]
        ids = []
        for b in range(B):
            row = []
            while len(row) < T:
                r = rnd.random()
                if r < 0.12:     # a well-formed object mention
                    row += [4] + [rnd.choice([6, 7, 8, 9, 10]) for _ in range(rnd.randint(1, 3))] + [5, 12]
                    row += [V + rnd.randrange(Ms[b]) for _ in range(rnd.randint(1, 5))]
                elif r < 0.18:
                    row += [0, 1, 2]
                elif r < 0.22:
                    row += [3, 1, 2]
                elif r < 0.30:
                    row.append(V + rnd.randrange(Ms[b]))
                else:
                    row.append(rnd.randrange(V))
            ids.append(row[:T])
        ids = torch.tensor(ids)
        think = torch.tensor([rnd.random() < 0.5 for _ in range(B)])
        hidden = [(torch.arange(B, dtype=torch.float32).view(B, 1, 1) * 1000 + float(s),) for s in range(T)]
        protos = torch.arange(sum(Ms), dtype=torch.float32)[:, None]
        outs = []
        for mod, cls in ((proc, proc.VisonTextProcessingClass), (ours, ours.VisonTextProcessingClass)):
            pr = cls(Proc(V, strings)); pr.model_embed_token_size = V
            comp, feats, labels, vrts, vf = mod.parseVRTintoCompletion(pr, ids.clone(), hidden, think, protos, grid)
            outs.append(dict(completions=comp, labels=labels, vrts=vrts,
                             feat_steps=[[f.flatten().tolist() for f in fs] for fs in feats],
                             proto_rows=[[f.flatten().tolist() for f in fs] for fs in vf]))
            # id shifts (in place): local -> global -> local
            x = ids.clone()
            gl = pr.assign_to_global_vrt_id(x, grid).clone()
            back = pr.assign_to_local_vrt_id(x, grid)
            assert torch.equal(back, ids)
            outs[-1]["global_ids"] = gl.tolist()
        assert outs[0] == outs[1], f"parser case {ci}: reference {outs[0]} != padt_b200 {outs[1]}"
        cases.append(dict(V=V, ids=ids.tolist(), grids=[list(g) for g in grids], thinking=think.tolist(), expect=outs[0]))
    path = os.path.join(GOLD, path_name + ".json")
    with open(path, "w") as f:
        json.dump(dict(alphabet=alphabet, cases=cases), f)
    n_obj = sum(len(l) for c in cases for l in c["expect"]["labels"])
    print(f"[golden] {path_name}: {n_cases} random piece streams, reference parser == padt_b200 parser ({n_obj} objects parsed); wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    only = sys.argv[1:]                     # optional: names of the cases to (re)generate
    # B=1, tied head, ragged windows (grid 8x12 -> merged 4x6: windows of 16 and 8 merged units)
    if not only or "tiny_b1" in only:
        case("tiny_b1", padt_tiny(tie=True), 0, [(1, 8, 12)], [9], 12, 1, [[(2, 3), (7, 2)]])
    # B=2, left padded, two different grids, untied head
    if not only or "tiny_b2" in only:
        case("tiny_b2", padt_tiny(tie=False), 1, [(1, 8, 12), (1, 4, 8)], [5, 11], 10, 2, [[(1, 2)], [(3, 1), (5, 3)]])
    # mid-size: GQA group 4, 4 LLM layers, 6 ViT blocks (two full-attention), untied head, two rows with different grids incl. a
    # window-divisible side (16 patches = 2 windows of 8 -> the empty pad window of get_window_index), left padded
    if not only or "small_b2" in only:
        case("small_b2", padt_small(tie=False), 2, [(1, 12, 16), (1, 8, 8)], [7, 4], 8, 3, [[(1, 3)], [(2, 2), (5, 2)]])
    if not only or "parser_cases" in only:
        parser_cases()
    # ViT at the bench geometry: 1344x1344 -> grid 96x96 (9216 patches, 144 windows of 64, 2304 merged tokens), full ViT width
    # (1280 / 16 heads / 3420), 2 blocks (window, full attention).  ~1 min, ~25 GB of CPU memory.
    if "vit1344" in only:
        from padt_b200.config import padt_vit1344
        cfg = padt_vit1344()
        vit_case("vit1344", cfg, 4, (96, 96), 6)
    # PaDT decoder at 7B dims on a 1344x1344 image (BASELINE config 4 geometry), two objects (5 and 3 VRTs).  A few minutes of CPU.
    if "dec7b_1344" in only:
        from padt_b200.config import padt_dec7b_1344
        dec_case("dec7b_1344", padt_dec7b_1344(), 5, (96, 96), [5, 3], 9)
    # BASELINE.json configs[0]: PaDT_Pro_3B dims, one 336x336 image (grid 24x24, 144 image tokens), chat-template-shaped prompt
    # L = 14 + 1 + 144 + 1 + 30 = 190, 32-token greedy.  ~10 min and ~25 GB of CPU memory; opt-in.
    if "full3b_c1" in only:
        global MISMATCHES
        MISMATCHES = []
        from padt_b200.config import padt_pro_3b
        case("full3b_c1", padt_pro_3b(), 0, [(1, 24, 24)], [30], 32, 1, [[(3, 5), (14, 4), (25, 1)]], pre_len=14, text_hi=151643)


if __name__ == "__main__":
    main()
