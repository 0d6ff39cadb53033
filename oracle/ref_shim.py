"""Compat shim that lets the UNMODIFIED reference hot-path code (`/root/reference/src/PaDT/models/*.py`)
run on the stack installed in this image (transformers 5.5.0, no deepspeed, flash-attn optional).

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` in the build container (the only place
`/root/reference` exists) to pin `oracle/padt_oracle.py` against the real reference and to write the
fixtures under `tests/golden/`.  Nothing in the product (`padt_b200/`) imports this file.

Recipe = SURVEY.md §8c:
  1. stub `deepspeed` (imported at padt.py:4 / padt_decoder.py:3, unused on the hot path);
  2. inject the three names padt_decoder.py:8 imports from the HF-4.50 module
     (`Qwen2RMSNorm`, `apply_rotary_emb`, `flash_attn_varlen_func`) into the 5.5.0 module;
  3. build the config with top-level hidden_size / vocab_size / vl_decoder (padt.py:117-130);
  4. after construction alias `visual`, `model.embed_tokens`, `rope_deltas`, `get_rope_index`
     (4.50 layout that padt.py:185,193,263 expects).
"""
import sys
import types
import math
import torch
import torch.nn.functional as F

import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_src():
    """/root/reference/src in the build container; the staged byte-for-byte copy under oracle/_ref/ elsewhere (oracle/stage_ref.py)."""
    if os.path.isdir("/root/reference/src/PaDT/models"):
        return "/root/reference/src"
    staged = os.path.join(_HERE, "_ref")
    if os.path.exists(os.path.join(staged, "PaDT", "models", "padt.py")):
        return staged
    raise FileNotFoundError("the reference model files are neither under /root/reference nor staged under oracle/_ref "
                            "(run `python oracle/stage_ref.py` in the build container)")


REF_SRC = None   # resolved by install()


def _apply_rotary_emb(x, cos, sin):
    # flash_attn.layers.rotary.apply_rotary_emb (non-interleaved): rotate the first 2*cos.shape[-1] dims.
    # x: (B, S, H, Dh) ; cos/sin: (S, R)
    ro = cos.shape[-1] * 2
    x1, x2 = x[..., : ro // 2], x[..., ro // 2: ro]
    c = cos[None, :, None, :]
    s = sin[None, :, None, :]
    o1 = x1 * c - x2 * s
    o2 = x1 * s + x2 * c
    return torch.cat([o1, o2, x[..., ro:]], dim=-1)


def _flash_attn_varlen_func(q, k, v, cu_q, cu_k, max_q, max_k, causal=False, softmax_scale=None, **kw):
    # per-segment non-causal attention with the FA2 rounding points (fp32 softmax, P rounded to the
    # input dtype before P@V, fp32 accumulate, output in the input dtype).
    assert not causal
    H, Dh = q.shape[1], q.shape[2]
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(Dh)
    out = torch.empty_like(q)
    cq = cu_q.tolist()
    ck = cu_k.tolist()
    for i in range(len(cq) - 1):
        qs = q[cq[i]:cq[i + 1]].transpose(0, 1).float()   # H, Sq, D
        ks = k[ck[i]:ck[i + 1]].transpose(0, 1).float()
        vs = v[ck[i]:ck[i + 1]].transpose(0, 1).float()
        s = torch.matmul(qs, ks.transpose(1, 2)) * scale
        m = s.max(dim=-1, keepdim=True).values
        p = torch.exp(s - m)
        l = p.sum(dim=-1, keepdim=True)
        p = p.to(q.dtype).float()
        o = torch.matmul(p, vs) / l
        out[cq[i]:cq[i + 1]] = o.transpose(0, 1).to(q.dtype)
    return out


def install(use_flash_attn: bool = False):
    """use_flash_attn=True (GPU baseline): the three names padt_decoder.py:8 imports are the REAL flash-attn functions
    (flash_attn.flash_attn_varlen_func, flash_attn.layers.rotary.apply_rotary_emb), as in the reference's own environment."""
    global REF_SRC
    REF_SRC = _ref_src()
    if "deepspeed" not in sys.modules:
        sys.modules["deepspeed"] = types.ModuleType("deepspeed")
    import transformers.models.qwen2_5_vl.modeling_qwen2_5_vl as hf
    hf.Qwen2RMSNorm = hf.Qwen2_5_VLRMSNorm
    if use_flash_attn:
        from flash_attn import flash_attn_varlen_func
        from flash_attn.layers.rotary import apply_rotary_emb
        hf.apply_rotary_emb = apply_rotary_emb
        hf.flash_attn_varlen_func = flash_attn_varlen_func
    else:
        hf.apply_rotary_emb = _apply_rotary_emb
        hf.flash_attn_varlen_func = _flash_attn_varlen_func
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import importlib
    # import the model files directly: `import PaDT` would pull in the trainer (deepspeed/accelerate/trl)
    pkg = types.ModuleType("PaDT")
    pkg.__path__ = [REF_SRC + "/PaDT"]
    sys.modules.setdefault("PaDT", pkg)
    mpkg = types.ModuleType("PaDT.models")
    mpkg.__path__ = [REF_SRC + "/PaDT/models"]
    sys.modules.setdefault("PaDT.models", mpkg)
    padt = importlib.import_module("PaDT.models.padt")
    dec = importlib.import_module("PaDT.models.padt_decoder")
    proc = importlib.import_module("PaDT.models.padt_processor")
    # padt_decoder.py:8 binds the three names at import time: rebind them in case the module was imported under the other mode
    dec.apply_rotary_emb = hf.apply_rotary_emb
    dec.flash_attn_varlen_func = hf.flash_attn_varlen_func
    return padt, dec, proc
