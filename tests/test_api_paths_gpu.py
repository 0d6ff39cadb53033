"""§8f rows 1 and 4 + the checkpoint entry path, on the GPU through the drop-in API:
  * teacher-forced ``forward(is_main=True)`` (padt.py:147-340): all-position extended logits + final-norm states against the oracle;
  * the processor path of ``generate`` (padt.py:717,740-745): identical ids to the device-side greedy path when the processors change
    nothing, sampling respects the per-row VRT mask, repetition penalty steers away from repeated ids;
  * ``from_pretrained`` on a synthetic safetensors checkpoint in BOTH key layouts (flat 4.50 / nested 5.x) + generation_config.json."""
import json
import os

import pytest
import torch

from helpers import assert_close_bf16, bf16_ulp, load_golden, report

pytestmark = pytest.mark.gpu


def _model(fx):
    from padt_b200.config import padt_small, padt_tiny
    from padt_b200.modeling import PaDTForConditionalGeneration
    from padt_b200.weights import random_weights
    cfg = padt_small(tie=fx["tie"]) if fx["cfg_name"] == "PaDT_small" else padt_tiny(tie=fx["tie"], vocab=fx["vocab"])
    w_cpu = random_weights(cfg, seed=fx["wseed"])
    return PaDTForConditionalGeneration(cfg, {k: v.cuda() for k, v in w_cpu.items()}), cfg, w_cpu


def _kw(fx, **more):
    kw = dict(input_ids=fx["input_ids"].cuda(), attention_mask=fx["attention_mask"].cuda(), pixel_values=fx["pixel_values"].cuda(),
              image_grid_thw=fx["grid"].cuda())
    kw.update(more)
    return kw


@pytest.mark.parametrize("name", ["tiny_b2", "small_b2"])
def test_forward_main_against_oracle(name):
    import padt_oracle as O
    fx = load_golden(name)
    model, cfg, w_cpu = _model(fx)
    out = model(**_kw(fx), output_hidden_states=True, return_dict=True)          # forward(is_main=True) is the default dispatch (padt.py:141-145)
    st = O.prefill(w_cpu, cfg, fx["input_ids"], fx["attention_mask"], fx["pixel_values"], fx["grid"])
    ref_logits = O.extended_logits(w_cpu, cfg, st["hidden"], st["prototypes"], st["logit_mask"])      # [B,L,V+sum M], every position (padt.py:292-301)
    B, L = fx["input_ids"].shape
    V = cfg.text.vocab_size
    assert out.logits.shape == ref_logits.shape == (B, L, V + st["prototypes"].shape[0])
    assert out.hidden_states[-1].shape == st["hidden"].shape
    real = fx["attention_mask"].bool()
    ours, ref = out.logits.float().cpu(), ref_logits.float()
    assert torch.equal(torch.isinf(ours), torch.isinf(ref)), "-inf pattern = the per-row VRT mask, at every position"
    fin = torch.isfinite(ref) & real[:, :, None]
    r_h = assert_close_bf16(out.hidden_states[-1].cpu()[real], st["hidden"][real], "final-norm states of the prompt")
    d = (ours[fin] - ref[fin]).abs()
    scale = float(ref[fin].abs().max())
    # arg-max of every real position agrees up to near-ties of the reference's logits
    am_o, am_r = ours.argmax(-1)[real], ref.argmax(-1)[real]
    top2 = torch.topk(ref, 2, dim=-1).values[real]
    flips = 0
    for i in (am_o != am_r).nonzero().flatten().tolist():
        assert float(top2[i, 0] - top2[i, 1]) <= 2 * bf16_ulp(float(top2[i, 0])), f"position {i}: arg-max differs on a gap of {float(top2[i, 0] - top2[i, 1])}"
        flips += 1
    report(f"{name} forward_main all-position logits", positions=int(real.sum()), max_abs=float(d.max()), rel_to_max=float(d.max()) / scale, argmax_equal=int((am_o == am_r).sum()),
           near_tie_flips=flips, hidden=r_h)
    assert float(d.max()) <= 0.03 * scale
    assert torch.equal(out.past_logit_mask.cpu(), st["logit_mask"]) and out.past_image_embeds.shape == st["prototypes"].shape
    assert out.past_high_res_image_embeds.shape == st["high_res"].shape and out.rope_deltas.cpu().tolist() == st["rope_deltas"].tolist()
    with pytest.raises(NotImplementedError):
        model(**_kw(fx), labels=fx["input_ids"].cuda())


@pytest.mark.parametrize("impl", [1, 0], ids=["persistent_kernel", "per_layer_kernels"])
def test_processor_path_equals_device_greedy(impl):
    """A custom processor that changes nothing must reproduce the device-side greedy ids exactly: same logits, first-index arg-max on the
    host, id fed back through the sampler's teacher-forcing slot."""
    fx = load_golden("tiny_b2")
    model, cfg, _ = _model(fx)
    model.engine.set_option("decode_impl", impl)
    T = 24
    base = model.generate(**_kw(fx), max_new_tokens=T, return_dict_in_generate=True, eos_token_id=[])
    ident = model.generate(**_kw(fx), max_new_tokens=T, return_dict_in_generate=True, eos_token_id=[], logits_processor=[lambda ids, s: s])
    assert torch.equal(ident.sequences, base.sequences)
    # (do_sample with top_k=1 is NOT equivalent to greedy here: TopKLogitsWarper keeps every entry equal to the k-th value, bf16-rounded
    # logits tie exactly, and multinomial then draws among the tied ids while arg-max takes the first -- same behaviour as the reference.)
    # EOS bookkeeping on the processor path: row 0 stops at its 3rd token and is padded afterwards (padt.py:748-757)
    L = fx["input_ids"].shape[1]
    eos0 = int(base.sequences[0, L + 2])
    a = model.generate(**_kw(fx), max_new_tokens=T, return_dict_in_generate=True, eos_token_id=[eos0], pad_token_id=cfg.pad_token_id)
    b = model.generate(**_kw(fx), max_new_tokens=T, return_dict_in_generate=True, eos_token_id=[eos0], pad_token_id=cfg.pad_token_id,
                       logits_processor=[lambda ids, s: s])
    assert torch.equal(a.sequences, b.sequences)


def test_sampling_respects_mask_and_repetition_penalty():
    fx = load_golden("tiny_b2")
    model, cfg, _ = _model(fx)
    V = cfg.text.vocab_size
    L = fx["input_ids"].shape[1]
    T = 200
    torch.manual_seed(0)
    out = model.generate(**_kw(fx), max_new_tokens=T, return_dict_in_generate=True, eos_token_id=[], do_sample=True, temperature=50.0)
    gen = out.sequences[:, L:].cpu()
    allow = out.past_logit_mask.cpu()
    for b in range(gen.shape[0]):
        assert bool(allow[b, gen[b]].all()), "a sampled id must be allowed by the row's logit mask"
    n_vrt = int((gen >= V).sum())
    assert int(torch.unique(gen).numel()) > 50 and n_vrt > 0, "hot sampling spreads over the extended vocabulary, VRT rows included"
    torch.manual_seed(0)
    again = model.generate(**_kw(fx), max_new_tokens=T, return_dict_in_generate=True, eos_token_id=[], do_sample=True, temperature=50.0)
    assert torch.equal(again.sequences, out.sequences), "same seed, same draws"
    # greedy + a strong repetition penalty: the fixed point of the plain greedy stream (tiny_b2 repeats a 4-cycle) must break up
    plain = model.generate(**_kw(fx), max_new_tokens=40, return_dict_in_generate=True, eos_token_id=[])
    pen = model.generate(**_kw(fx), max_new_tokens=40, return_dict_in_generate=True, eos_token_id=[], repetition_penalty=5.0)
    u_plain, u_pen = int(torch.unique(plain.sequences[0, L:]).numel()), int(torch.unique(pen.sequences[0, L:]).numel())
    report("sampling / repetition penalty", sampled_unique=int(torch.unique(gen).numel()), sampled_vrt=n_vrt, greedy_unique=u_plain, penalised_unique=u_pen)
    assert u_pen > u_plain


@pytest.mark.parametrize("layout", ["flat_4.50", "nested_5.x"])
def test_from_pretrained_synthetic_checkpoint(tmp_path, layout):
    st = pytest.importorskip("safetensors.torch")
    from padt_b200.modeling import PaDTForConditionalGeneration
    fx = load_golden("tiny_b1")
    model, cfg, w_cpu = _model(fx)
    t, v, d = cfg.text, cfg.vision, cfg.decoder
    text = {"hidden_size": t.hidden_size, "intermediate_size": t.intermediate_size, "num_hidden_layers": t.num_layers, "num_attention_heads": t.num_heads,
            "num_key_value_heads": t.num_kv_heads, "vocab_size": t.vocab_size, "rms_norm_eps": t.rms_norm_eps}
    hc = {"_name_or_path": "synthetic", "tie_word_embeddings": t.tie_word_embeddings, "eos_token_id": cfg.eos_token_ids[0], "pad_token_id": cfg.pad_token_id,
          "image_token_id": cfg.image_token_id, "video_token_id": cfg.video_token_id, "vision_start_token_id": cfg.vision_start_token_id,
          "vision_end_token_id": cfg.vision_end_token_id, "use_visual_prototype_projection": True,
          "vision_config": {"depth": v.depth, "hidden_size": v.hidden_size, "intermediate_size": v.intermediate_size, "num_heads": v.num_heads, "in_chans": 3,
                            "patch_size": v.patch_size, "temporal_patch_size": 2, "spatial_merge_size": 2, "window_size": v.window_size,
                            "fullatt_block_indexes": list(v.fullatt_block_indexes), "out_hidden_size": v.out_hidden_size},
          "vl_decoder": {"hidden_size": d.hidden_size, "intermediate_size": d.intermediate_size, "num_heads": d.num_heads, "use_mask_loss": True}}
    if layout == "flat_4.50":
        hc.update(text)
        hc.update({"rope_theta": t.rope_theta, "rope_scaling": {"type": "mrope", "mrope_section": list(t.mrope_section)}})
        sd = dict(w_cpu)
    else:
        hc["text_config"] = dict(text, rope_parameters={"rope_type": "default", "rope_theta": t.rope_theta, "mrope_section": list(t.mrope_section)})
        sd = {("model." + k if k.startswith("visual.") else "model.language_model." + k[len("model."):] if k.startswith("model.") else k): x
              for k, x in w_cpu.items()}
    keys = sorted(sd)
    half = len(keys) // 2                                           # two shards, like a real checkpoint
    st.save_file({k: sd[k].contiguous() for k in keys[:half]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    st.save_file({k: sd[k].contiguous() for k in keys[half:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    (tmp_path / "config.json").write_text(json.dumps(hc))
    eos_pair = [cfg.eos_token_ids[0], 3]
    (tmp_path / "generation_config.json").write_text(json.dumps({"eos_token_id": eos_pair, "pad_token_id": cfg.pad_token_id, "repetition_penalty": 1.0}))
    loaded = PaDTForConditionalGeneration.from_pretrained(str(tmp_path), torch_dtype=torch.bfloat16, attn_implementation="flash_attention_2", device_map={"": 0})
    assert loaded.model.embed_tokens.weight.shape[0] == t.vocab_size                       # eval/test_demo.py:31
    assert loaded.generation_config.eos_token_id == eos_pair and loaded.padt_config.eos_token_ids == tuple(eos_pair)
    a = model.generate(**_kw(fx), max_new_tokens=fx["steps"], return_dict_in_generate=True, eos_token_id=[], output_hidden_states=True)
    b = loaded.generate(**_kw(fx), max_new_tokens=fx["steps"], return_dict_in_generate=True, eos_token_id=[], output_hidden_states=True)
    assert torch.equal(a.sequences, b.sequences) and torch.equal(a.hidden_states[-1][-1], b.hidden_states[-1][-1])
    assert torch.equal(b.sequences.cpu(), fx["sequences"]), "the loaded checkpoint reproduces the reference fixture's greedy ids"
    with pytest.raises(KeyError):
        PaDTForConditionalGeneration.from_state_dict(cfg, {k: x for k, x in sd.items() if "vis_norm" not in k})


def test_vit_plan_cache_is_bounded_and_eviction_is_transparent():
    """The per-grid-signature plan cache holds 16 entries (an eval run over images of many sizes makes nearly every batch a new key): after
    more distinct grids than that, the first grid is rebuilt from scratch and must give bit-identical outputs."""
    fx = load_golden("tiny_b1")
    model, cfg, _ = _model(fx)
    g = torch.Generator().manual_seed(3)

    def run(h, w):
        gg = torch.Generator().manual_seed(h * 100 + w)
        pv = torch.randn((h * w, cfg.vision.patch_dim), generator=gg).to(torch.bfloat16).cuda()
        return [t.clone() for t in (lambda e, hr, pe: (e, hr, pe[0], pe[1]))(*model.engine.vit_forward(pv, torch.tensor([[1, h, w]])))]

    first = run(8, 12)
    for k in range(20):                                   # 20 more signatures than the cache holds
        run(4 + 2 * (k % 5), 4 + 2 * (k // 5) + 2 * (k % 3))
    again = run(8, 12)
    torch.cuda.synchronize()
    for a, b in zip(first, again):
        assert torch.equal(a, b)
