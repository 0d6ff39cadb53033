"""Weight inventory of the hot path (HF-4.50 key layout the reference checkpoints use) + synthetic init.

Key layout follows SURVEY.md §5 / Appendix A: ``visual.*``, ``model.layers.*``, ``model.embed_tokens.*``,
``lm_head.*``, ``vis_norm.*``, ``vis_proj.{0,1}.*``, ``vl_decoder.*`` (reference modules:
``src/PaDT/models/padt.py:115-132``, ``src/PaDT/models/padt_decoder.py:131-185``; HF Qwen2.5-VL for the rest).

There is no checkpoint offline, so benchmarks/tests use ``random_weights``: N(0, 0.02) matrices, small random
biases, norm weights near 1, and ``vis_norm.weight ~ 0.02`` so prototype rows sit on the scale of text rows
(the reference zero-inits it, padt.py:134-139, which makes every VRT logit exactly 0).
The PaDT decoder (``vl_decoder.*``) gets variance-preserving matrices N(0, 1/fan_in) and N(0, 0.1) biases: with
N(0, 0.02) its heads collapse to sigmoid(~0) = 0.5 boxes and ~1e-2 mask logits, which no parity test can tell from
a constant; with this init boxes span ~0.05..0.95, scores are O(1) and mask logits O(1) (std ~1.3, |max| ~6).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import PadtConfig


def weight_shapes(cfg: PadtConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    v, t, d = cfg.vision, cfg.text, cfg.decoder
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    # ---- ViT (HF:345-378)
    s["visual.patch_embed.proj.weight"] = (v.hidden_size, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size)
    for i in range(v.depth):
        p = f"visual.blocks.{i}."
        s[p + "norm1.weight"] = (v.hidden_size,)
        s[p + "norm2.weight"] = (v.hidden_size,)
        s[p + "attn.qkv.weight"] = (3 * v.hidden_size, v.hidden_size)
        s[p + "attn.qkv.bias"] = (3 * v.hidden_size,)
        s[p + "attn.proj.weight"] = (v.hidden_size, v.hidden_size)
        s[p + "attn.proj.bias"] = (v.hidden_size,)
        s[p + "mlp.gate_proj.weight"] = (v.intermediate_size, v.hidden_size)
        s[p + "mlp.gate_proj.bias"] = (v.intermediate_size,)
        s[p + "mlp.up_proj.weight"] = (v.intermediate_size, v.hidden_size)
        s[p + "mlp.up_proj.bias"] = (v.intermediate_size,)
        s[p + "mlp.down_proj.weight"] = (v.hidden_size, v.intermediate_size)
        s[p + "mlp.down_proj.bias"] = (v.hidden_size,)
    mh = v.hidden_size * v.spatial_merge_size ** 2
    s["visual.merger.ln_q.weight"] = (v.hidden_size,)
    s["visual.merger.mlp.0.weight"] = (mh, mh)
    s["visual.merger.mlp.0.bias"] = (mh,)
    s["visual.merger.mlp.2.weight"] = (v.out_hidden_size, mh)
    s["visual.merger.mlp.2.bias"] = (v.out_hidden_size,)
    # ---- LLM (HF:831-855)
    D, I, hd = t.hidden_size, t.intermediate_size, t.head_dim
    s["model.embed_tokens.weight"] = (t.vocab_size, D)
    for i in range(t.num_layers):
        p = f"model.layers.{i}."
        s[p + "input_layernorm.weight"] = (D,)
        s[p + "self_attn.q_proj.weight"] = (t.num_heads * hd, D)
        s[p + "self_attn.q_proj.bias"] = (t.num_heads * hd,)
        s[p + "self_attn.k_proj.weight"] = (t.num_kv_heads * hd, D)
        s[p + "self_attn.k_proj.bias"] = (t.num_kv_heads * hd,)
        s[p + "self_attn.v_proj.weight"] = (t.num_kv_heads * hd, D)
        s[p + "self_attn.v_proj.bias"] = (t.num_kv_heads * hd,)
        s[p + "self_attn.o_proj.weight"] = (D, t.num_heads * hd)
        s[p + "post_attention_layernorm.weight"] = (D,)
        s[p + "mlp.gate_proj.weight"] = (I, D)
        s[p + "mlp.up_proj.weight"] = (I, D)
        s[p + "mlp.down_proj.weight"] = (D, I)
    s["model.norm.weight"] = (D,)
    if not t.tie_word_embeddings:
        s["lm_head.weight"] = (t.vocab_size, D)
    # ---- VRT prototype projection (padt.py:117-125)
    if cfg.use_visual_prototype_projection:
        s["vis_norm.weight"] = (D,)
        s["vis_norm.bias"] = (D,)
        s["vis_proj.0.weight"] = (cfg.lora_r, D)
        s["vis_proj.1.weight"] = (D, cfg.lora_r)
    # ---- PaDT decoder (padt_decoder.py:131-185)
    H, DI = d.hidden_size, d.intermediate_size
    q = "vl_decoder."
    s[q + "vp_embedding.weight"] = (1, H)
    s[q + "bbox_score_mask_tokens.weight"] = (3, H)
    s[q + "input_projection.0.weight"] = (D,)
    s[q + "input_projection.1.weight"] = (H, D)
    s[q + "input_projection.1.bias"] = (H,)
    s[q + "input_projection.3.weight"] = (H, H)
    s[q + "input_projection.3.bias"] = (H,)
    for blk in ("low_res_transformer", "high_res_transformer1", "high_res_transformer2"):
        p = q + blk + "."
        for n in range(1, 7):
            s[p + f"norm{n}.weight"] = (H,)
        for att in ("self_attn", "cross_attn_query_to_image", "cross_attn_image_to_query"):
            for lin in ("q_proj", "k_proj", "v_proj", "proj"):
                s[p + f"{att}.{lin}.weight"] = (H, H)
                s[p + f"{att}.{lin}.bias"] = (H,)
        s[p + "mlp.0.weight"] = (DI, H)
        s[p + "mlp.0.bias"] = (DI,)
        s[p + "mlp.2.weight"] = (H, DI)
        s[p + "mlp.2.bias"] = (H,)
    s[q + "high_res_norm.weight"] = (H,)
    for head, out in (("bbox_prediction", 4), ("mask_output_mlp", H // 16)):
        s[q + head + ".0.weight"] = (H, H)
        s[q + head + ".0.bias"] = (H,)
        s[q + head + ".2.weight"] = (H, H)
        s[q + head + ".2.bias"] = (H,)
        s[q + head + ".4.weight"] = (out, H)
        s[q + head + ".4.bias"] = (out,)
    s[q + "score_prediction.weight"] = (1, H)
    s[q + "score_prediction.bias"] = (1,)
    s[q + "mask_output_upscaling1.0.weight"] = (H // 4 * 4, H)
    s[q + "mask_output_upscaling1.0.bias"] = (H // 4 * 4,)
    s[q + "mask_output_upscaling1.1.weight"] = (H // 4 * 4,)
    s[q + "mask_output_upscaling2.0.weight"] = (H // 16 * 4, H // 4)
    s[q + "mask_output_upscaling2.0.bias"] = (H // 16 * 4,)
    return s


def _is_norm(name: str) -> bool:
    leaf = name.rsplit(".", 2)
    return (
        "norm" in name.split(".")[-2]
        or name.endswith("ln_q.weight")
        or name.endswith("input_projection.0.weight")
        or name.endswith("mask_output_upscaling1.1.weight")
    ) and not name.startswith("vis_norm")


def num_params(cfg: PadtConfig) -> int:
    n = 0
    for shp in weight_shapes(cfg).values():
        k = 1
        for x in shp:
            k *= x
        n += k
    return n


def random_weights(cfg: PadtConfig, seed: int = 0, device="cpu", dtype=torch.bfloat16,
                   std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded synthetic init.  Generated per tensor with its own sub-seed so that a given tensor is identical
    no matter which other tensors exist (tied/untied variants share everything else)."""
    cfg.validate()
    dev = torch.device(device)
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shp) in enumerate(weight_shapes(cfg).items()):
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1000003 + idx)
        r = torch.randn(shp, generator=g, device=dev, dtype=torch.float32)
        if name == "vis_norm.weight":
            w = 0.02 * (1.0 + 0.05 * r)
        elif name == "vis_norm.bias":
            w = torch.zeros_like(r)
        elif _is_norm(name):
            w = 1.0 + 0.05 * r
        elif name.endswith(".bias"):
            w = (0.1 if name.startswith("vl_decoder.") else std) * r
        elif name in ("vl_decoder.vp_embedding.weight", "vl_decoder.bbox_score_mask_tokens.weight"):
            w = r  # nn.Embedding default init N(0,1)
        elif name.startswith("vl_decoder.") and len(shp) == 2:
            w = r / float(shp[1]) ** 0.5
        else:
            w = std * r
        out[name] = w.to(dtype)
    return out
