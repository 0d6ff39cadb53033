"""Runs the UNMODIFIED reference classes (``PaDT.models.padt`` / ``padt_decoder``) under ``oracle/ref_shim.py``.

TEST / BENCH INFRASTRUCTURE shared by ``oracle/make_golden.py`` (CPU, bit-exact pinning of the oracle), ``bench.py`` (``--impl reference``
on the host cores; ``hf_gpu_baseline`` on the same B200) and ``tests/test_reference_floor_gpu.py``.  Nothing in ``padt_b200/`` imports it.

``build_reference(cfg, weights, device, attn)``:
  * attn="fa2_numerics" (CPU): HF's eager attention is replaced by the oracle's restatement of FA2 numerics (flash-attn is CUDA-only);
  * attn="flash_attention_2" (GPU): the real flash-attn 2 kernels for the ViT, the LLM and the PaDT decoder, as in the reference's setup;
  * attn="sdpa": torch SDPA for ViT/LLM (``eval/test_demo.py`` on CPU), shim varlen attention for the decoder.
``fast_init=True`` skips HF's random initialisation (177 s at 3B) -- every parameter is overwritten by ``weights`` anyway.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for _p in (ROOT, HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import ref_shim  # noqa: E402
import padt_oracle as O  # noqa: E402


def _eager_fa2(module, query, key, value, attention_mask, scaling, dropout=0.0, **kwargs):
    """Replacement for HF `eager_attention_forward` (HF:182-204) with FA2 numerics."""
    B = query.shape[0]
    outs = []
    for b in range(B):
        mask = None
        if attention_mask is not None:
            mask = attention_mask[b, 0] > -1.0
        outs.append(O.fa2_attention(query[b], key[b], value[b], scaling, mask))
    out = torch.stack(outs, 0).transpose(1, 2).contiguous()
    return out, None


def build_reference(cfg, weights, device="cpu", attn="fa2_numerics", fast_init=False, need=("visual", "model", "lm_head", "vis_", "vl_decoder")):
    """`need`: key prefixes of `weights` that must be present (a decode-only CPU sample passes only the LLM weights)."""
    padt, dec, proc = ref_shim.install(use_flash_attn=(attn == "flash_attention_2"))
    import transformers.models.qwen2_5_vl.modeling_qwen2_5_vl as hf
    from transformers import Qwen2_5_VLConfig
    if attn == "fa2_numerics":
        hf.eager_attention_forward = _eager_fa2
    hf_attn = {"fa2_numerics": "eager", "flash_attention_2": "flash_attention_2", "sdpa": "sdpa"}[attn]
    t, v, d = cfg.text, cfg.vision, cfg.decoder
    hc = Qwen2_5_VLConfig(
        text_config=dict(vocab_size=t.vocab_size, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
                         num_hidden_layers=t.num_layers, num_attention_heads=t.num_heads,
                         num_key_value_heads=t.num_kv_heads, rms_norm_eps=t.rms_norm_eps,
                         rope_parameters=dict(rope_type="default", rope_theta=t.rope_theta,
                                              mrope_section=list(t.mrope_section)),
                         tie_word_embeddings=t.tie_word_embeddings, use_sliding_window=False,
                         max_position_embeddings=4096, pad_token_id=None),
        vision_config=dict(depth=v.depth, hidden_size=v.hidden_size, intermediate_size=v.intermediate_size,
                           num_heads=v.num_heads, out_hidden_size=v.out_hidden_size,
                           fullatt_block_indexes=list(v.fullatt_block_indexes), window_size=v.window_size,
                           patch_size=v.patch_size, spatial_merge_size=v.spatial_merge_size,
                           temporal_patch_size=v.temporal_patch_size),
        image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id,
        vision_start_token_id=cfg.vision_start_token_id, vision_end_token_id=cfg.vision_end_token_id,
        tie_word_embeddings=t.tie_word_embeddings,
    )
    hc.hidden_size = t.hidden_size
    hc.vocab_size = t.vocab_size
    hc.vl_decoder = dict(name="PaDTDecoder", attn_implementation="flash_attention_2", hidden_size=d.hidden_size,
                         intermediate_size=d.intermediate_size, llm_hidden_state=t.hidden_size,
                         num_heads=d.num_heads, spatial_merge_size=v.spatial_merge_size, use_mask_loss=True)
    hc.use_visual_prototype_projection = cfg.use_visual_prototype_projection
    hc._attn_implementation = hf_attn
    hc.text_config._attn_implementation = hf_attn
    hc.vision_config._attn_implementation = hf_attn
    # like from_pretrained(torch_dtype=bf16): instantiate under a bf16 default dtype so parameters are bf16 while
    # the explicitly-fp32 rotary inv_freq buffers stay fp32 (a later .to(bf16) would round them)
    torch.set_default_dtype(torch.bfloat16)
    try:
        if fast_init:
            from transformers.initialization import no_init_weights
            with no_init_weights():
                model = padt.PaDTForConditionalGeneration(hc).eval()
        else:
            model = padt.PaDTForConditionalGeneration(hc).eval()
    finally:
        torch.set_default_dtype(torch.float32)
    model.vl_decoder.dtype = torch.bfloat16
    assert model.model.visual.rotary_pos_emb.inv_freq.dtype == torch.float32
    assert model.model.language_model.rotary_emb.inv_freq.dtype == torch.float32
    # 4.50 attribute layout expected by padt.py
    model.__dict__["visual"] = model.model.visual
    model.model.embed_tokens = model.model.language_model.embed_tokens
    model.rope_deltas = None
    def _rope_index(input_ids=None, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None, attention_mask=None):
        # restated transformers-4.50 function (host integer loop, like the original): evaluate on the CPU, return on the caller's device
        dev = input_ids.device
        pos, delta = O.get_rope_index(cfg, input_ids.cpu(), image_grid_thw.cpu(), None if attention_mask is None else attention_mask.cpu())
        return pos.to(dev), delta.to(dev)
    model.get_rope_index = _rope_index
    sd = {}
    for k, val in weights.items():
        if k.startswith("visual."):
            nk = "model." + k
        elif k.startswith("model."):
            nk = "model.language_model." + k[len("model."):]
        else:
            nk = k
        sd[nk] = val
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "inv_freq" not in m and m != "model.embed_tokens.weight"  # alias of language_model.embed_tokens
               and not (t.tie_word_embeddings and m == "lm_head.weight")]
    def _needed(k):
        k = k[len("model."):] if k.startswith("model.visual") else ("model." + k[len("model.language_model."):] if k.startswith("model.language_model.") else k)
        return any(k.startswith(n) for n in need)
    missing = [m for m in missing if _needed(m)]
    assert not missing and not unexpected, (missing, unexpected)
    if t.tie_word_embeddings:
        model.lm_head.weight = model.model.language_model.embed_tokens.weight
    if str(device) != "cpu":
        model = model.to(device)
        # .to() keeps dtypes; the fp32 rotary buffers must stay fp32 on the device too
        assert model.model.visual.rotary_pos_emb.inv_freq.dtype == torch.float32
    return model, proc


def ref_greedy(model, cfg, input_ids, attention_mask, pixel_values, grid, steps, forced=None, timing=None, keep=True):
    """padt.py:671-757 restated around the reference's own forward_main (eos disabled -> fixed length).
    `timing` (dict, CUDA only): filled with CUDA-event times `prefill_ms` (ViT + prototypes + prompt forward + first arg-max) and
    `decode_ms` (the remaining steps).  keep=False drops the per-step hidden states / logits (long benchmark runs)."""
    from transformers import DynamicCache
    B, L = input_ids.shape
    cache = DynamicCache(config=model.config)
    model.rope_deltas = None
    dev = input_ids.device
    out = model.forward_main(input_ids=input_ids, attention_mask=attention_mask, pixel_values=pixel_values,
                             image_grid_thw=grid, past_key_values=cache, use_cache=True, output_hidden_states=True,
                             return_dict=True, cache_position=torch.arange(L, device=dev))
    past = dict(past_image_embeds=out.past_image_embeds, past_logit_mask=out.past_logit_mask,
                past_high_res_image_embeds=out.past_high_res_image_embeds, past_visual_pe=out.past_visual_pe)
    seq, mask = input_ids.clone(), attention_mask.clone()
    hid, logit_steps, arg_steps = [out.hidden_states[-1]], [], []
    ev = None
    if timing is not None and dev.type == "cuda":
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0] = timing["start_event"]
    for s in range(steps):
        scores = out.logits[:, -1, :].clone().float()
        if keep:
            logit_steps.append(scores)
        nxt = torch.argmax(scores, dim=-1)
        arg_steps.append(nxt.clone())
        if ev is not None and s == 0:
            ev[1].record()
        if forced is not None:
            nxt = forced[:, s].clone()
        seq = torch.cat([seq, nxt[:, None]], dim=-1)
        if s == steps - 1:
            break
        mask = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype, device=dev)], dim=-1)
        out = model.forward_main(input_ids=nxt[:, None], attention_mask=mask, past_key_values=cache, use_cache=True,
                                 output_hidden_states=True, return_dict=True,
                                 cache_position=torch.tensor([L + s], device=dev), **past)
        if keep:
            hid.append(out.hidden_states[-1])
    if ev is not None:
        ev[2].record()
        torch.cuda.synchronize()
        timing["prefill_ms"] = ev[0].elapsed_time(ev[1])
        timing["decode_ms"] = ev[1].elapsed_time(ev[2])
    return dict(sequences=seq, hidden_steps=hid, logits_steps=logit_steps, argmax_steps=arg_steps, **past)


def synthetic_decode_sample(model, cfg, ctx, n_protos, max_steps, seconds):
    """Bounded CPU sample of the reference's OWN decode step: forward_main's past_image_embeds branch (padt.py:221-229, 268-301) + arg-max
    (padt.py:713,745) at context `ctx`, on a DynamicCache pre-filled with random keys / values (the CPU prefill of a 2350-token prompt takes
    minutes).  Returns (steps run, seconds)."""
    import time
    from transformers import DynamicCache
    t = cfg.text
    B = 1
    g = torch.Generator().manual_seed(0)
    cache = DynamicCache(config=model.config)
    for li in range(t.num_layers):
        k = torch.randn((B, t.num_kv_heads, ctx, t.head_dim), generator=g).to(torch.bfloat16)
        v = torch.randn((B, t.num_kv_heads, ctx, t.head_dim), generator=g).to(torch.bfloat16)
        cache.update(k, v, li)
    protos = (0.02 * torch.randn((n_protos, t.hidden_size), generator=g)).to(torch.bfloat16)
    logit_mask = torch.ones((B, t.vocab_size + n_protos), dtype=torch.bool)
    # the decode branch only passes these two through (padt.py:223), but it calls .detach() on the first
    hr = torch.zeros((4, cfg.vision.hidden_size), dtype=torch.bfloat16)
    pe = (torch.zeros((4, cfg.vision.head_dim)), torch.zeros((4, cfg.vision.head_dim)))
    past = dict(past_image_embeds=protos, past_logit_mask=logit_mask, past_high_res_image_embeds=hr, past_visual_pe=pe)
    model.rope_deltas = torch.zeros((B, 1), dtype=torch.long)
    tok = torch.tensor([123])
    mask = torch.ones((B, ctx), dtype=torch.long)
    n = 0
    t0 = time.time()
    with torch.no_grad():
        while True:
            mask = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype)], dim=-1)
            out = model.forward_main(input_ids=tok[:, None], attention_mask=mask, past_key_values=cache, use_cache=True, output_hidden_states=True,
                                     return_dict=True, cache_position=torch.tensor([ctx + n]), **past)
            tok = torch.argmax(out.logits[:, -1, :].clone().float(), dim=-1)
            n += 1
            if time.time() - t0 > seconds or n >= max_steps:
                break
    return n, time.time() - t0


