"""CPU ORACLE for the PaDT generate/decode hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import
this module.  The product (``padt_b200/``) never does; it fails loudly when its CUDA library is missing.

What it is: a plain-PyTorch functional restatement (weights = dict of tensors in the HF-4.50 key layout, no
nn.Module, no HF import) of the reference algorithm, each function citing the reference lines it follows:

  * ``/root/reference/src/PaDT/models/padt.py``            (custom_visual_forward, forward_main, vl_decode, _sample_vision_token)
  * ``/root/reference/src/PaDT/models/padt_decoder.py``    (PaDTDecoder, PaDTDecoderBlock, PaDTDecoderFlashAttention2)
  * ``/root/reference/src/PaDT/models/padt_processor.py``  (id shift, parser)
  * third-party arithmetic the reference calls but does not vendor (pins: ``transformers==4.50.0`` setup.py:19,
    ``flash-attn==2.7.4.post1`` setup.sh:6): Qwen2.5-VL ViT / LLM blocks, M-RoPE, ``get_rope_index`` (4.50.0
    semantics, NOT the 5.5.0 function installed here), FA2 attention numerics (fp32 scores/softmax, P rounded
    to bf16 before P@V, fp32 accumulate).  "HF:n" below = line n of the installed
    transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py (block math identical to 4.50).

Pinning: ``oracle/make_golden.py`` runs the UNMODIFIED reference classes (via ``oracle/ref_shim.py``) on seeded
tiny configs in the build container and stores inputs/outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` requires this module to reproduce them bit-for-bit on CPU.  The reference ships
no numeric golden vectors of its own; its only KATs (parser text, id shift) are in
``tests/test_processor.py``.

All rounding points follow the bf16 HF path (SURVEY.md §8a notes): every Linear output is rounded to the
activation dtype, RMSNorm normalises in fp32 -> casts -> multiplies by the weight, LLM RoPE is evaluated in the
activation dtype, ViT / decoder RoPE in fp32.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
W = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
def rmsnorm(x: Tensor, weight: Tensor, eps: float = 1e-6) -> Tensor:
    """HF:66-71 (Qwen2_5_VLRMSNorm == Qwen2RMSNorm used by padt_decoder.py:71-74)."""
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return weight * h.to(dt)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return F.linear(x, w, b)


def rotate_half(x: Tensor) -> Tensor:
    """HF:149-153."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def fa2_attention(q: Tensor, k: Tensor, v: Tensor, scale: float, mask: Optional[Tensor] = None) -> Tensor:
    """FlashAttention-2 numerics restated (flash-attn 2.7.4 `flash_attn_varlen_func` / HF FA2 path; called at
    padt_decoder.py:55 and by every HF attention layer): scores and softmax statistics in fp32, probabilities
    rounded to the input dtype before P@V, fp32 accumulation, normalisation by the fp32 row sum, output rounded
    to the input dtype.  q [H,Sq,D]; k,v [Hkv,Sk,D] (GQA: H % Hkv == 0); mask bool [Sq,Sk] (True = attend) or None.
    Rows with no valid key return 0 (FA2 never computes padded rows)."""
    H, Hkv = q.shape[0], k.shape[0]
    if H != Hkv:
        rep = H // Hkv
        k = k[:, None].expand(Hkv, rep, *k.shape[1:]).reshape(H, *k.shape[1:])
        v = v[:, None].expand(Hkv, rep, *v.shape[1:]).reshape(H, *v.shape[1:])
    s = torch.matmul(q.float(), k.float().transpose(1, 2)) * scale
    if mask is not None:
        s = s.masked_fill(~mask[None], float("-inf"))
    m = s.max(dim=-1, keepdim=True).values
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    p = p.to(q.dtype).float()
    o = torch.matmul(p, v.float())
    o = torch.where(l > 0, o / l, torch.zeros_like(o))
    return o.to(q.dtype)


def apply_rotary_emb_fa(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """flash_attn.layers.rotary.apply_rotary_emb, non-interleaved, as called at padt_decoder.py:43,50 with fp32
    x [1,S,H,Dh] and cos/sin [S,R]: rotates the first 2R dims (NeoX halves)."""
    ro = cos.shape[-1] * 2
    x1, x2 = x[..., : ro // 2], x[..., ro // 2: ro]
    c = cos[None, :, None, :]
    s = sin[None, :, None, :]
    o1 = x1 * c - x2 * s
    o2 = x1 * s + x2 * c
    return torch.cat([o1, o2, x[..., ro:]], dim=-1)


# --------------------------------------------------------------------------------------------------
# ViT  (padt.py:48-106 custom_visual_forward; HF:91-146, 207-322, 382-453)
# --------------------------------------------------------------------------------------------------
def vit_rot_pos_emb(grid_thw: Sequence[Sequence[int]], head_dim: int, merge: int) -> Tensor:
    """HF:382-409 + HF:117-130.  Returns fp32 [P, head_dim/2] in merge-block order."""
    pos_ids = []
    for t, h, w in grid_thw:
        hpos = torch.arange(h).unsqueeze(1).expand(-1, w)
        hpos = hpos.reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        wpos = torch.arange(w).unsqueeze(0).expand(h, -1)
        wpos = wpos.reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        pos_ids.append(torch.stack([hpos, wpos], dim=-1).repeat(t, 1))
    pos_ids = torch.cat(pos_ids, dim=0)
    max_grid = max(max(h, w) for _, h, w in grid_thw)
    dim = head_dim // 2
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
    seq = torch.arange(max_grid, dtype=inv_freq.dtype)
    freqs = torch.outer(seq, inv_freq)
    return freqs[pos_ids].flatten(1)


def vit_window_index(grid_thw: Sequence[Sequence[int]], window_size: int, merge: int, patch: int):
    """HF:411-453.  Returns (window_index [M] int64, cu_window_seqlens list incl. duplicates)."""
    window_index = []
    cu = [0]
    wid = 0
    vws = window_size // merge // patch
    for gt, gh, gw in grid_thw:
        lh, lw = gh // merge, gw // merge
        index = torch.arange(gt * lh * lw).reshape(gt, lh, lw)
        pad_h = vws - lh % vws
        pad_w = vws - lw % vws
        nwh = (lh + pad_h) // vws
        nww = (lw + pad_w) // vws
        ip = F.pad(index, (0, pad_w, 0, pad_h), "constant", -100)
        ip = ip.reshape(gt, nwh, vws, nww, vws).permute(0, 1, 3, 2, 4).reshape(gt, nwh * nww, vws, vws)
        seqlens = (ip != -100).sum([2, 3]).reshape(-1)
        ip = ip.reshape(-1)
        window_index.append(ip[ip != -100] + wid)
        tmp = seqlens.cumsum(0) * (merge * merge) + cu[-1]
        cu.extend(tmp.tolist())
        wid += gt * lh * lw
    return torch.cat(window_index, dim=0), cu


def vit_forward(w: W, cfg, pixel_values: Tensor, grid_thw: Tensor):
    """padt.py:48-106.  Returns (image_embeds [M,D] raster-merged order, high_res [P,Hv] WINDOW order,
    (cos, sin) fp32 [P, head_dim] WINDOW order)."""
    v = cfg.vision
    dt = w["visual.patch_embed.proj.weight"].dtype
    grid = [[int(a) for a in g] for g in grid_thw.tolist()]
    unit = v.spatial_merge_size ** 2
    # patch embed HF:108-114
    x = pixel_values.to(dt).view(-1, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size)
    ks = (v.temporal_patch_size, v.patch_size, v.patch_size)
    x = F.conv3d(x, w["visual.patch_embed.proj.weight"], None, stride=ks).view(-1, v.hidden_size)
    rot = vit_rot_pos_emb(grid, v.head_dim, v.spatial_merge_size)
    window_index, cu_win = vit_window_index(grid, v.window_size, v.spatial_merge_size, v.patch_size)
    cu_win = torch.unique_consecutive(torch.tensor(cu_win, dtype=torch.int32)).tolist()
    S = x.shape[0]
    x = x.reshape(S // unit, unit, -1)[window_index].reshape(S, -1)
    rot = rot.reshape(S // unit, unit, -1)[window_index].reshape(S, -1)
    emb = torch.cat((rot, rot), dim=-1)
    cos, sin = emb.cos(), emb.sin()
    cu_full = [0]
    for t, h, ww in grid:
        for _ in range(t):
            cu_full.append(cu_full[-1] + h * ww)
    scale = v.head_dim ** -0.5
    for li in range(v.depth):
        p = f"visual.blocks.{li}."
        cu = cu_full if li in v.fullatt_block_indexes else cu_win
        # attention HF:221-287 (+ fp32 rope HF:156-167)
        n1 = rmsnorm(x, w[p + "norm1.weight"], 1e-6)
        qkv = linear(n1, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"])
        q, k, vv = qkv.reshape(S, 3, v.num_heads, -1).permute(1, 0, 2, 3).unbind(0)
        c, s_ = cos.unsqueeze(-2).float(), sin.unsqueeze(-2).float()
        qf, kf = q.float(), k.float()
        q = ((qf * c) + (rotate_half(qf) * s_)).to(dt)
        k = ((kf * c) + (rotate_half(kf) * s_)).to(dt)
        out = torch.empty_like(q)
        for a, b in zip(cu[:-1], cu[1:]):
            out[a:b] = fa2_attention(q[a:b].transpose(0, 1), k[a:b].transpose(0, 1), vv[a:b].transpose(0, 1),
                                     scale).transpose(0, 1)
        x = x + linear(out.reshape(S, -1), w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
        # MLP HF:87-88
        n2 = rmsnorm(x, w[p + "norm2.weight"], 1e-6)
        g = F.silu(linear(n2, w[p + "mlp.gate_proj.weight"], w[p + "mlp.gate_proj.bias"]))
        u = linear(n2, w[p + "mlp.up_proj.weight"], w[p + "mlp.up_proj.bias"])
        x = x + linear(g * u, w[p + "mlp.down_proj.weight"], w[p + "mlp.down_proj.bias"])
    high_res = x
    # merger HF:144-146
    m = rmsnorm(x, w["visual.merger.ln_q.weight"], 1e-6).view(-1, v.hidden_size * unit)
    m = linear(m, w["visual.merger.mlp.0.weight"], w["visual.merger.mlp.0.bias"])
    m = F.gelu(m)
    m = linear(m, w["visual.merger.mlp.2.weight"], w["visual.merger.mlp.2.bias"])
    m = m[torch.argsort(window_index)]
    return m, high_res, (cos, sin)


def prototypes(w: W, cfg, image_embeds: Tensor) -> Tensor:
    """padt.py:187-191."""
    if not cfg.use_visual_prototype_projection:
        return image_embeds.clone()
    p = F.layer_norm(image_embeds, (image_embeds.shape[-1],), w["vis_norm.weight"], w["vis_norm.bias"], 1e-5)
    return p + linear(linear(p, w["vis_proj.0.weight"]), w["vis_proj.1.weight"])


# --------------------------------------------------------------------------------------------------
# M-RoPE position ids: transformers==4.50.0 get_rope_index semantics (called at padt.py:263)
# --------------------------------------------------------------------------------------------------
def get_rope_index(cfg, input_ids: Tensor, image_grid_thw: Tensor, attention_mask: Optional[Tensor] = None):
    """Images only (second_per_grid_t = 0 -> temporal index = start).  Padded slots hold 1; text after an
    image resumes at max+1; delta_i = max+1 - padded_len.  KAT: SURVEY.md Appendix A."""
    merge = cfg.vision.spatial_merge_size
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    B, L = input_ids.shape
    position_ids = torch.ones(3, B, L, dtype=input_ids.dtype)
    deltas = []
    image_index = 0
    for i in range(B):
        ids = input_ids[i][attention_mask[i] == 1]
        toks = ids.tolist()
        vs = [j for j, tkn in enumerate(toks) if tkn == cfg.vision_start_token_id]
        image_nums = sum(1 for j in vs if j + 1 < len(toks) and toks[j + 1] == cfg.image_token_id)
        pos_list: List[Tensor] = []
        st = 0
        for _ in range(image_nums):
            ed = toks.index(cfg.image_token_id, st)
            t, h, ww = [int(a) for a in image_grid_thw[image_index].tolist()]
            image_index += 1
            gt, gh, gw = t, h // merge, ww // merge
            text_len = ed - st
            st_idx = int(pos_list[-1].max()) + 1 if pos_list else 0
            pos_list.append(torch.arange(text_len).view(1, -1).expand(3, -1) + st_idx)
            t_index = (torch.arange(gt).view(-1, 1).expand(-1, gh * gw) * 0).long().flatten()
            h_index = torch.arange(gh).view(1, -1, 1).expand(gt, -1, gw).flatten()
            w_index = torch.arange(gw).view(1, 1, -1).expand(gt, gh, -1).flatten()
            pos_list.append(torch.stack([t_index, h_index, w_index]) + text_len + st_idx)
            st = ed + gt * gh * gw
        if st < len(toks):
            st_idx = int(pos_list[-1].max()) + 1 if pos_list else 0
            pos_list.append(torch.arange(len(toks) - st).view(1, -1).expand(3, -1) + st_idx)
        llm_pos = torch.cat(pos_list, dim=1).reshape(3, -1)
        position_ids[:, i, attention_mask[i] == 1] = llm_pos.to(position_ids.dtype)
        deltas.append(int(llm_pos.max()) + 1 - L)
    return position_ids, torch.tensor(deltas, dtype=input_ids.dtype).unsqueeze(1)


def mrope_cos_sin(cfg, position_ids: Tensor, dtype) -> Tuple[Tensor, Tensor]:
    """HF:595-608 + section selection of HF:659-665.  position_ids [3,B,L] -> cos, sin [B,L,head_dim] in `dtype`."""
    t = cfg.text
    hd = t.head_dim
    inv_freq = 1.0 / (t.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).to(torch.float) / hd))
    inv = inv_freq[None, None, :, None].float().expand(3, position_ids.shape[1], -1, 1)
    pos = position_ids[:, :, None, :].float()
    freqs = (inv @ pos).transpose(2, 3)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(dtype), emb.sin().to(dtype)
    sec = list(t.mrope_section) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos.split(sec, dim=-1))], dim=-1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin.split(sec, dim=-1))], dim=-1)
    return cos, sin


# --------------------------------------------------------------------------------------------------
# LLM (HF:672-943) with a plain KV cache
# --------------------------------------------------------------------------------------------------
class KVCache:
    def __init__(self, n_layers: int):
        self.k: List[Optional[Tensor]] = [None] * n_layers
        self.v: List[Optional[Tensor]] = [None] * n_layers

    def update(self, li: int, k: Tensor, v: Tensor):
        if self.k[li] is None:
            self.k[li], self.v[li] = k, v
        else:
            self.k[li] = torch.cat([self.k[li], k], dim=2)
            self.v[li] = torch.cat([self.v[li], v], dim=2)
        return self.k[li], self.v[li]

    def seq_len(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[2]


def llm_forward(w: W, cfg, inputs_embeds: Tensor, position_ids: Tensor, attention_mask: Tensor,
                cache: KVCache, all_hidden: bool = False) -> Tensor:
    """self.model(...) at padt.py:279-290.  inputs_embeds [B,q,D]; attention_mask [B, past+q] (1 = real token).
    Returns the final-norm hidden states [B,q,D]."""
    t = cfg.text
    B, q_len, D = inputs_embeds.shape
    dt = inputs_embeds.dtype
    hd, nh, nkv = t.head_dim, t.num_heads, t.num_kv_heads
    cos, sin = mrope_cos_sin(cfg, position_ids, dt)
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    past = cache.seq_len()
    total = past + q_len
    qpos = torch.arange(past, total)
    causal = qpos[:, None] >= torch.arange(total)[None, :]
    scale = hd ** -0.5
    h = inputs_embeds
    for li in range(t.num_layers):
        p = f"model.layers.{li}."
        n = rmsnorm(h, w[p + "input_layernorm.weight"], t.rms_norm_eps)
        q = linear(n, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"])
        k = linear(n, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"])
        v = linear(n, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
        q = q.view(B, q_len, nh, hd).transpose(1, 2)
        k = k.view(B, q_len, nkv, hd).transpose(1, 2)
        v = v.view(B, q_len, nkv, hd).transpose(1, 2)
        q = (q * cos) + (rotate_half(q) * sin)          # HF:667-668, evaluated in the activation dtype
        k = (k * cos) + (rotate_half(k) * sin)
        kk, vv = cache.update(li, k, v)
        outs = []
        for b in range(B):
            mask = causal & (attention_mask[b, :total] == 1)[None, :]
            outs.append(fa2_attention(q[b], kk[b], vv[b], scale, mask))
        a = torch.stack(outs, 0).transpose(1, 2).reshape(B, q_len, nh * hd)
        h = h + linear(a, w[p + "self_attn.o_proj.weight"])
        n2 = rmsnorm(h, w[p + "post_attention_layernorm.weight"], t.rms_norm_eps)
        g = F.silu(linear(n2, w[p + "mlp.gate_proj.weight"]))
        u = linear(n2, w[p + "mlp.up_proj.weight"])
        h = h + linear(g * u, w[p + "mlp.down_proj.weight"])
    return rmsnorm(h, w["model.norm.weight"], t.rms_norm_eps)


def patch_offsets(cfg, image_grid_thw: Tensor) -> List[int]:
    """padt.py:199 / padt_processor.py:39: exclusive cumsum of merged-patch counts."""
    unit = cfg.vision.spatial_merge_size ** 2
    off = [0]
    for t, h, ww in image_grid_thw.tolist():
        off.append(off[-1] + int(t) * int(h) * int(ww) // unit)
    return off


def head_weight(w: W, cfg) -> Tensor:
    return w["model.embed_tokens.weight"] if cfg.text.tie_word_embeddings else w["lm_head.weight"]


def extended_logits(w: W, cfg, hidden: Tensor, protos: Tensor, logit_mask: Tensor) -> Tensor:
    """padt.py:292-301."""
    ext = torch.cat([head_weight(w, cfg), protos], dim=0)
    logits = hidden @ ext.T
    logits.masked_fill_((~logit_mask[:, None, :]).expand(-1, logits.shape[1], -1), -float("inf"))
    return logits


def build_logit_mask(cfg, image_grid_thw: Tensor, n_ext: int) -> Tensor:
    """padt.py:196-201."""
    V = cfg.text.vocab_size
    off = patch_offsets(cfg, image_grid_thw)
    mask = torch.zeros((image_grid_thw.shape[0], n_ext), dtype=torch.bool)
    mask[:, :V] = True
    for i in range(len(off) - 1):
        mask[i, V + off[i]: V + off[i + 1]] = True
    return mask


def prefill(w: W, cfg, input_ids: Tensor, attention_mask: Tensor, pixel_values: Tensor, image_grid_thw: Tensor):
    """forward_main, pixel_values branch (padt.py:183-219, 256-301).  Returns dict of state + last-position logits."""
    image_embeds, high_res, pe = vit_forward(w, cfg, pixel_values, image_grid_thw)
    protos = prototypes(w, cfg, image_embeds)
    ext = torch.cat([w["model.embed_tokens.weight"], protos], dim=0)
    logit_mask = build_logit_mask(cfg, image_grid_thw, ext.shape[0])
    assert int(input_ids.max()) < ext.shape[0]
    x = ext[input_ids]
    n_tok = int((input_ids == cfg.image_token_id).sum())
    if n_tok != image_embeds.shape[0]:
        raise ValueError(f"Image features and image tokens do not match: tokens: {n_tok}, features {image_embeds.shape[0]}")
    img_mask = (input_ids == cfg.image_token_id).unsqueeze(-1).expand_as(x)
    x = x.masked_scatter(img_mask, image_embeds.to(x.dtype))
    position_ids, rope_deltas = get_rope_index(cfg, input_ids, image_grid_thw, attention_mask)
    cache = KVCache(cfg.text.num_layers)
    hidden = llm_forward(w, cfg, x, position_ids, attention_mask, cache)
    # the reference projects EVERY prompt position (padt.py:292-301, no logits_to_keep); restating that, rather than only the last
    # row, also keeps the CPU GEMM shape -- and therefore oneDNN's summation order -- identical to the reference's
    logits_last = extended_logits(w, cfg, hidden, protos, logit_mask)[:, -1, :]
    return dict(image_embeds=image_embeds, high_res=high_res, visual_pe=pe, prototypes=protos, logit_mask=logit_mask,
                position_ids=position_ids, rope_deltas=rope_deltas, cache=cache, hidden=hidden, logits_last=logits_last)


def decode_step(w: W, cfg, st: dict, tokens: Tensor, attention_mask: Tensor):
    """forward_main, past_image_embeds branch (padt.py:221-229, 268-277, 279-301) for one new token per row.
    tokens [B] global ids; attention_mask [B, past+1]."""
    protos = st["prototypes"]
    ext = torch.cat([w["model.embed_tokens.weight"], protos], dim=0)
    x = ext[tokens[:, None]]
    cache: KVCache = st["cache"]
    B = tokens.shape[0]
    cache_pos0 = cache.seq_len()
    delta = cache_pos0 + st["rope_deltas"]                      # [B,1]
    position_ids = (torch.arange(1).view(1, -1).expand(B, -1) + delta).unsqueeze(0).expand(3, -1, -1)
    hidden = llm_forward(w, cfg, x, position_ids, attention_mask, cache)
    logits = extended_logits(w, cfg, hidden, protos, st["logit_mask"])[:, -1, :]
    return hidden, logits


def greedy_generate(w: W, cfg, input_ids: Tensor, attention_mask: Tensor, pixel_values: Tensor,
                    image_grid_thw: Tensor, max_new_tokens: int, eos_token_ids: Optional[Sequence[int]] = None,
                    pad_token_id: Optional[int] = None, forced_tokens: Optional[Tensor] = None):
    """Greedy loop of padt.py:671-757 (do_sample=False, no logits processors).  `forced_tokens` [B,T] teacher-forces
    the fed-back token (the argmax of every step is still recorded)."""
    eos = list(cfg.eos_token_ids if eos_token_ids is None else eos_token_ids)
    pad = cfg.pad_token_id if pad_token_id is None else pad_token_id
    B = input_ids.shape[0]
    st = prefill(w, cfg, input_ids, attention_mask, pixel_values, image_grid_thw)
    seq = input_ids.clone()
    mask = attention_mask.clone()
    unfinished = torch.ones(B, dtype=torch.long)
    hidden_steps = [st["hidden"]]
    logits_steps = []
    argmax_steps = []
    logits = st["logits_last"]
    for step in range(max_new_tokens):
        scores = logits.clone().float()                                   # padt.py:713
        logits_steps.append(scores)
        nxt = torch.argmax(scores, dim=-1)                                # padt.py:745
        argmax_steps.append(nxt.clone())
        if forced_tokens is not None:
            nxt = forced_tokens[:, step].clone()
        if len(eos) > 0:
            nxt = nxt * unfinished + pad * (1 - unfinished)               # padt.py:749
        seq = torch.cat([seq, nxt[:, None]], dim=-1)
        if len(eos) > 0:
            is_eos = torch.isin(nxt, torch.tensor(eos, dtype=nxt.dtype))
            unfinished = unfinished & ~is_eos
        done = int(unfinished.max()) == 0 or step == max_new_tokens - 1
        if done:
            break
        mask = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype)], dim=-1)
        hidden, logits = decode_step(w, cfg, st, nxt, mask)
        hidden_steps.append(hidden)
    return dict(sequences=seq, hidden_steps=hidden_steps, logits_steps=logits_steps, argmax_steps=argmax_steps,
                past_image_embeds=st["prototypes"], past_logit_mask=st["logit_mask"],
                past_high_res_image_embeds=st["high_res"], past_visual_pe=st["visual_pe"], state=st)


# --------------------------------------------------------------------------------------------------
# PaDT decoder (padt_decoder.py) and vl_decode (padt.py:342-412)
# --------------------------------------------------------------------------------------------------
def _dec_attention(w: W, p: str, nh: int, query, key, cu_q, cu_k, q_pos, k_pos, rot: Tuple[bool, bool]):
    """PaDTDecoderFlashAttention2.forward, padt_decoder.py:20-60."""
    q = linear(query if rot[0] else query + q_pos, w[p + "q_proj.weight"], w[p + "q_proj.bias"])
    k = linear(key if rot[1] else key + k_pos, w[p + "k_proj.weight"], w[p + "k_proj.bias"])
    v = linear(key, w[p + "v_proj.weight"], w[p + "v_proj.bias"])
    q = q.reshape(query.shape[0], nh, -1)
    k = k.reshape(key.shape[0], nh, -1)
    v = v.reshape(key.shape[0], nh, -1)
    if rot[0]:
        cos, sin = q_pos
        cos = cos.chunk(2, dim=-1)[0].contiguous()
        sin = sin.chunk(2, dim=-1)[0].contiguous()
        q = apply_rotary_emb_fa(q.unsqueeze(0).float(), cos.float(), sin.float()).type_as(q).squeeze(0)
    if rot[1]:
        cos, sin = k_pos
        cos = cos.chunk(2, dim=-1)[0].contiguous()
        sin = sin.chunk(2, dim=-1)[0].contiguous()
        k = apply_rotary_emb_fa(k.unsqueeze(0).float(), cos.float(), sin.float()).type_as(k).squeeze(0)
    scale = 1.0 / math.sqrt(q.shape[-1])
    out = torch.empty_like(q)
    cq, ck = [int(a) for a in cu_q], [int(a) for a in cu_k]
    for i in range(len(cq) - 1):
        out[cq[i]:cq[i + 1]] = fa2_attention(q[cq[i]:cq[i + 1]].transpose(0, 1), k[ck[i]:ck[i + 1]].transpose(0, 1),
                                             v[ck[i]:ck[i + 1]].transpose(0, 1), scale).transpose(0, 1)
    return linear(out.reshape(query.shape[0], -1), w[p + "proj.weight"], w[p + "proj.bias"])


def _dec_block(w: W, p: str, nh: int, query, memory, cu_q, cu_m, q_pos, m_pos):
    """PaDTDecoderBlock.forward, padt_decoder.py:95-128 (update_memory=True for all three blocks, :151-153)."""
    qn = rmsnorm(query, w[p + "norm1.weight"])
    query = query + _dec_attention(w, p + "self_attn.", nh, qn, qn, cu_q, cu_q, q_pos, q_pos, (False, False))
    qn = rmsnorm(query, w[p + "norm2.weight"])
    mn = rmsnorm(memory, w[p + "norm3.weight"])
    query = query + _dec_attention(w, p + "cross_attn_query_to_image.", nh, qn, mn, cu_q, cu_m, q_pos, m_pos, (False, True))
    n4 = rmsnorm(query, w[p + "norm4.weight"])
    query = query + linear(F.gelu(linear(n4, w[p + "mlp.0.weight"], w[p + "mlp.0.bias"])), w[p + "mlp.2.weight"], w[p + "mlp.2.bias"])
    qn = rmsnorm(query, w[p + "norm5.weight"])
    mn = rmsnorm(memory, w[p + "norm6.weight"])
    memory = memory + _dec_attention(w, p + "cross_attn_image_to_query.", nh, mn, qn, cu_m, cu_q, m_pos, q_pos, (True, False))
    return query, memory


def _input_projection(w: W, x: Tensor) -> Tensor:
    q = "vl_decoder.input_projection."
    x = rmsnorm(x, w[q + "0.weight"])
    x = F.gelu(linear(x, w[q + "1.weight"], w[q + "1.bias"]))
    return linear(x, w[q + "3.weight"], w[q + "3.bias"])


def _mlp3(w: W, p: str, x: Tensor) -> Tensor:
    x = F.gelu(linear(x, w[p + "0.weight"], w[p + "0.bias"]))
    x = F.gelu(linear(x, w[p + "2.weight"], w[p + "2.bias"]))
    return linear(x, w[p + "4.weight"], w[p + "4.bias"])


def padt_decoder_forward(w: W, cfg, object_vp_feat: List[Tensor], cu_low: Tensor, cu_high: Tensor,
                         visual_pe: Tuple[Tensor, Tensor], cu_patch: Tensor, obj_thws: Tensor):
    """PaDTDecoder.forward, padt_decoder.py:187-276."""
    d = cfg.decoder
    dt = cu_low.dtype
    H = d.hidden_size
    ms = cfg.vision.spatial_merge_size
    q_ = "vl_decoder."
    vp_num = [int(i.shape[0]) for i in object_vp_feat]
    max_vp = max(vp_num)
    n_obj = len(object_vp_feat)
    add = 3
    qmask = torch.zeros((n_obj, max_vp + add), dtype=torch.bool)
    all_vp = _input_projection(w, torch.cat(object_vp_feat))
    cu_query = []
    acc = 0
    for oi, on in enumerate(vp_num):
        cu_query.append(w[q_ + "bbox_score_mask_tokens.weight"])
        cu_query.append(all_vp[acc: acc + on] + w[q_ + "vp_embedding.weight"])
        qmask[oi, : add + on] = 1
        acc += on
    cu_query = torch.cat(cu_query, dim=0)
    cu_q = F.pad(qmask.sum(dim=-1).cumsum(dim=-1), (1, 0), "constant", 0).to(torch.int32)
    cu_lowp = (cu_patch // ms ** 2).to(torch.int32)
    low = _input_projection(w, cu_low)
    ped = visual_pe[0].shape[-1]
    low_pe = (visual_pe[0].reshape(-1, ms ** 2, ped)[:, 0, :], visual_pe[1].reshape(-1, ms ** 2, ped)[:, 0, :])
    out = cu_query
    out, low = _dec_block(w, q_ + "low_res_transformer.", d.num_heads, out, low, cu_q, cu_lowp, cu_query, low_pe)
    high = rmsnorm(low.unsqueeze(1).repeat_interleave(ms ** 2, dim=1).flatten(0, 1) + cu_high, w[q_ + "high_res_norm.weight"])
    out, high = _dec_block(w, q_ + "high_res_transformer1.", d.num_heads, out, high, cu_q, cu_patch, cu_query, visual_pe)
    out, high = _dec_block(w, q_ + "high_res_transformer2.", d.num_heads, out, high, cu_q, cu_patch, cu_query, visual_pe)
    oib = torch.zeros((n_obj, max_vp + add, H), dtype=dt)
    oib[qmask] = out
    tok = oib[:, :3]
    bbox = torch.sigmoid(_mlp3(w, q_ + "bbox_prediction.", tok[:, 0]))
    score = linear(tok[:, 1], w[q_ + "score_prediction.weight"], w[q_ + "score_prediction.bias"])
    if not d.use_mask_loss:
        return bbox, score, None, ()
    mask_out = _mlp3(w, q_ + "mask_output_mlp.", tok[:, 2])
    N, Dm = high.shape
    me = linear(high, w[q_ + "mask_output_upscaling1.0.weight"], w[q_ + "mask_output_upscaling1.0.bias"])
    me = F.gelu(rmsnorm(me, w[q_ + "mask_output_upscaling1.1.weight"]))
    me = me.reshape(N, 2, 2, Dm // 4).permute(1, 2, 0, 3)
    me = F.gelu(linear(me, w[q_ + "mask_output_upscaling2.0.weight"], w[q_ + "mask_output_upscaling2.0.bias"]))
    me = me.reshape(2, 2, N, 2, 2, Dm // 16).permute(0, 3, 1, 4, 2, 5).flatten(0, 1).flatten(1, 2)
    mepp = me.permute(2, 0, 1, 3).contiguous()                      # N,4,4,D/16
    pn = (cu_patch[1:] - cu_patch[:-1]).to(torch.int64)
    sum_pn = int(cu_patch[-1])
    obj_ids = torch.repeat_interleave(torch.arange(n_obj), pn)
    exp = mask_out.index_select(0, obj_ids)
    offsets = cu_patch[:-1].to(torch.int64)
    pos = torch.arange(sum_pn, dtype=torch.int64) - offsets[obj_ids]
    Hs = torch.tensor([int(thw[1]) for thw in obj_thws], dtype=torch.int64)
    Ws = torch.tensor([int(thw[2]) for thw in obj_thws], dtype=torch.int64)
    Wpp = Ws[obj_ids]
    row, col = pos // Wpp, pos % Wpp
    mlpp = (mepp * exp[:, None, None, :]).sum(dim=-1)
    Hm, Wm = int(Hs.max()), int(Ws.max())
    padded = torch.zeros((n_obj, 4, 4, Hm, Wm), dtype=mlpp.dtype)
    padded[obj_ids, :, :, row, col] = mlpp
    masks = padded.permute(0, 3, 1, 4, 2).contiguous().reshape(n_obj, Hm * 4, Wm * 4)
    return bbox, score, masks, (Hs, Ws)


def vl_decode(w: W, cfg, object_vp_feats: List[List[Tensor]], low_res: Tensor, high_res: Tensor,
              image_grid_thws: Tensor, visual_pes: Tuple[Tensor, Tensor]):
    """padt.py:342-412 (non-empty branch; the empty branch returns zero-size tensors, :405-412)."""
    flat = sum(object_vp_feats, [])
    dt = low_res.dtype
    if len(flat) == 0:
        return dict(pred_boxes=torch.zeros((0, 4), dtype=dt), pred_score=torch.zeros((0, 1), dtype=dt),
                    pred_mask=torch.zeros((0, 8, 8), dtype=dt), pred_mask_valid_hw=(), sample_idx=[])
    off = 0
    sidx, lows, highs, pc, ps, cu, thws = [], [], [], [], [], [], []
    for si, (feats, thw) in enumerate(zip(object_vp_feats, image_grid_thws)):
        pn = int(thw[0]) * int(thw[1]) * int(thw[2])
        n = len(feats)
        lo = low_res[off // 4: (off + pn) // 4]
        hi = high_res[off: off + pn]
        sidx.extend([si] * n)
        lows.append(lo.unsqueeze(0).repeat_interleave(n, dim=0).flatten(0, 1))
        highs.append(hi.unsqueeze(0).repeat_interleave(n, dim=0).flatten(0, 1))
        pc.append(visual_pes[0][off: off + pn].unsqueeze(0).repeat_interleave(n, dim=0).flatten(0, 1))
        ps.append(visual_pes[1][off: off + pn].unsqueeze(0).repeat_interleave(n, dim=0).flatten(0, 1))
        cu.extend([pn] * n)
        off += pn
        thws.extend([thw] * n)
    cu_patch = F.pad(torch.Tensor(cu).cumsum(dim=0), (1, 0), "constant", 0).to(torch.int32)
    bbox, score, masks, hw = padt_decoder_forward(
        w, cfg, flat, torch.cat(lows, 0), torch.cat(highs, 0), (torch.cat(pc, 0), torch.cat(ps, 0)), cu_patch,
        torch.stack(list(thws), dim=0))
    return dict(pred_boxes=bbox, pred_score=score, pred_mask=masks, pred_mask_valid_hw=hw, sample_idx=sidx)


# --------------------------------------------------------------------------------------------------
# processor arithmetic (padt_processor.py:36-50)
# --------------------------------------------------------------------------------------------------
def assign_to_global_vrt_id(V: int, merge: int, input_ids: Tensor, image_grid_thw: Tensor) -> Tensor:
    m = input_ids >= V
    if m.sum() > 0:
        off = F.pad((image_grid_thw.cumprod(-1)[:, -1] // merge ** 2).cumsum(dim=-1), (1, 0), "constant", 0)
        input_ids[m] += off[:-1, None].expand(-1, input_ids.shape[1])[m]
    return input_ids


def assign_to_local_vrt_id(V: int, merge: int, input_ids: Tensor, image_grid_thw: Tensor) -> Tensor:
    m = input_ids >= V
    if m.sum() > 0:
        off = F.pad((image_grid_thw.cumprod(-1)[:, -1] // merge ** 2).cumsum(dim=-1), (1, 0), "constant", 0)
        input_ids[m] -= off[:-1, None].expand(-1, input_ids.shape[1])[m]
    return input_ids
