"""Thin Python host layer over the C ABI (``include/padt_b200.h``): owns the handle, passes borrowed device
pointers of torch tensors, precomputes the two small rotary tables on the host.  No compute happens here."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _capi
from .config import PadtConfig
from .weights import weight_shapes


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class PadtEngine:
    MAX_POS = 8192          # initial rows of the LLM rotary table; grown on demand (the reference has no length limit)
    MAX_GRID = 1024         # rows of the ViT rotary table (patches per image side)

    def __init__(self, cfg: PadtConfig, weights: Dict[str, torch.Tensor]):
        cfg.validate()
        self.cfg = cfg
        self.lib = _capi.load()
        if not torch.cuda.is_available():
            raise _capi.PadtError("padt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.dims = _capi.dims_from_config(cfg)
        names = list(weight_shapes(cfg).keys())
        n = self.lib.padt_weight_count(C.byref(self.dims))
        if n != len(names):
            raise _capi.PadtError(f"weight inventory mismatch: library {n} vs python {len(names)}")
        ptrs = (C.c_void_p * n)()
        self._keep: List[torch.Tensor] = []
        self._keep_by_name: Dict[str, torch.Tensor] = {}
        for i, name in enumerate(names):
            cname = self.lib.padt_weight_name(C.byref(self.dims), i).decode()
            if cname != name:
                raise _capi.PadtError(f"weight order mismatch at {i}: library '{cname}' vs python '{name}'")
            w = weights[name]
            shape = tuple(weight_shapes(cfg)[name])
            if tuple(w.shape) != shape or w.dtype != torch.bfloat16 or not w.is_cuda or not w.is_contiguous():
                raise _capi.PadtError(f"weight {name}: expected contiguous cuda bf16 {shape}, got {tuple(w.shape)} {w.dtype} {w.device}")
            ptrs[i] = w.data_ptr()
            self._keep.append(w)
            self._keep_by_name[name] = w
        self.device = self._keep[0].device
        self.max_pos = self.MAX_POS
        self.llm_cos, self.llm_sin = self._llm_rope_table(self.max_pos)
        # ViT rotary tables (HF:117-130 with dim = head_dim/2)
        vd = cfg.vision.head_dim // 2
        vinv = 1.0 / (10000.0 ** (torch.arange(0, vd, 2, dtype=torch.float) / vd))
        vfreqs = torch.outer(torch.arange(self.MAX_GRID, dtype=vinv.dtype), vinv)
        self.vit_cos = vfreqs.cos().to(self.device).contiguous()
        self.vit_sin = vfreqs.sin().to(self.device).contiguous()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.padt_create(C.byref(self.dims), ptrs, n, self.llm_cos.data_ptr(), self.llm_sin.data_ptr(), self.MAX_POS,
                                             C.byref(handle)), "padt_create")
        self.handle = handle
        self._gen_keep = None

    def _llm_rope_table(self, n_pos: int):
        """LLM rotary tables, evaluated like HF:595-608 (fp32 on the host, cast to bf16)."""
        t = self.cfg.text
        hd = t.head_dim
        inv_freq = 1.0 / (t.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).to(torch.float) / hd))
        freqs = torch.outer(torch.arange(n_pos, dtype=torch.float32), inv_freq.float())
        return (freqs.cos().to(torch.bfloat16).to(self.device).contiguous(), freqs.sin().to(torch.bfloat16).to(self.device).contiguous())

    def _ensure_positions(self, n_pos: int):
        if n_pos <= self.max_pos:
            return
        new = (n_pos + self.MAX_POS - 1) // self.MAX_POS * self.MAX_POS
        cos, sin = self._llm_rope_table(new)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.padt_set_rope_table(self.handle, cos.data_ptr(), sin.data_ptr(), new), "padt_set_rope_table")
        self.llm_cos, self.llm_sin, self.max_pos = cos, sin, new

    def close(self):
        if getattr(self, "handle", None):
            self.lib.padt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        _capi.check(self.lib.padt_set_option(self.handle, name.encode(), int(value)), "padt_set_option")

    # ------------------------------------------------------------------ V1-V6
    def vit_forward(self, pixel_values: torch.Tensor, grid_thw: torch.Tensor):
        cfg = self.cfg
        grid = [int(x) for x in grid_thw.reshape(-1).tolist()]
        n_img = len(grid) // 3
        P = sum(grid[3 * i] * grid[3 * i + 1] * grid[3 * i + 2] for i in range(n_img))
        pv = pixel_values.to(device=self.device, dtype=torch.bfloat16).contiguous()
        if pv.shape != (P, cfg.vision.patch_dim):
            raise ValueError(f"pixel_values shape {tuple(pv.shape)} does not match grid_thw ({P} patches x {cfg.vision.patch_dim})")
        M = P // (cfg.vision.spatial_merge_size ** 2)
        image_embeds = torch.empty((M, cfg.text.hidden_size), dtype=torch.bfloat16, device=self.device)
        high_res = torch.empty((P, cfg.vision.hidden_size), dtype=torch.bfloat16, device=self.device)
        cos = torch.empty((P, cfg.vision.head_dim), dtype=torch.float32, device=self.device)
        sin = torch.empty_like(cos)
        g = _capi.i64_array(grid)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.padt_vit_forward(self.handle, pv.data_ptr(), g, n_img, self.vit_cos.data_ptr(), self.vit_sin.data_ptr(),
                                                  self.MAX_GRID, image_embeds.data_ptr(), high_res.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                                  _stream()), "padt_vit_forward")
        return image_embeds, high_res, (cos, sin)

    # ------------------------------------------------------------------ E1
    def prototypes(self, image_embeds: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(image_embeds)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.padt_prototypes(self.handle, image_embeds.data_ptr(), image_embeds.shape[0], out.data_ptr(), _stream()),
                        "padt_prototypes")
        return out

    # ------------------------------------------------------------------ prefill / decode
    def prefill(self, input_ids: torch.Tensor, image_slot: torch.Tensor, pos3: torch.Tensor, kv_start: Sequence[int],
                rope_delta: Sequence[int], vrt_lo: Sequence[int], vrt_hi: Sequence[int], image_embeds: torch.Tensor,
                prototypes: torch.Tensor, max_steps: int, eos_ids: Sequence[int], pad_id: int, want_hidden: bool = True,
                forced: Optional[torch.Tensor] = None, want_logits: bool = False):
        B, L = input_ids.shape
        D = self.cfg.text.hidden_size
        dev = self.device
        # positions never exceed the padded length + generated tokens (rope_delta <= 0, padt.py:266-276)
        self._ensure_positions(L + max_steps + 1)
        bufs = dict(
            tokens=torch.zeros((max_steps, B), dtype=torch.int64, device=dev),
            argmax=torch.zeros((max_steps, B), dtype=torch.int64, device=dev),
            hidden_steps=torch.zeros((max_steps, B, D), dtype=torch.bfloat16, device=dev) if want_hidden else None,
            last_hidden=torch.zeros((B, D), dtype=torch.bfloat16, device=dev),
            prefill_hidden=torch.zeros((B, L, D), dtype=torch.bfloat16, device=dev) if want_hidden else None,
            forced=None if forced is None else forced.to(device=dev, dtype=torch.int64).contiguous(),
            # masked bf16 logits of the latest forward (sampling / logits processors); columns no row may emit are never written: -inf
            logits=torch.full((B, (self.cfg.text.vocab_size + prototypes.shape[0] + 7) // 8 * 8), float("-inf"), dtype=torch.bfloat16, device=dev)
            if want_logits else None,
            ids=input_ids.to(device=dev, dtype=torch.int64).contiguous(),
            slot=image_slot.to(device=dev, dtype=torch.int32).contiguous(),
            pos3=pos3.to(device=dev, dtype=torch.int32).contiguous(),
            image_embeds=image_embeds, prototypes=prototypes,
        )
        if bufs["forced"] is not None and tuple(bufs["forced"].shape) != (max_steps, B):
            raise ValueError("forced tokens must be [max_steps, B]")
        gb = _capi.PadtGenBuffers()
        gb.tokens_out = bufs["tokens"].data_ptr()
        gb.argmax_out = bufs["argmax"].data_ptr()
        gb.hidden_steps = _ptr(bufs["hidden_steps"])
        gb.last_hidden = bufs["last_hidden"].data_ptr()
        gb.prefill_hidden = _ptr(bufs["prefill_hidden"])
        gb.forced = _ptr(bufs["forced"])
        gb.max_steps = max_steps
        gb.logits_out = _ptr(bufs["logits"])
        gb.logits_ld = 0 if bufs["logits"] is None else bufs["logits"].shape[1]
        with torch.cuda.device(dev):
            _capi.check(self.lib.padt_prefill(
                self.handle, bufs["ids"].data_ptr(), bufs["slot"].data_ptr(), bufs["pos3"].data_ptr(),
                _capi.i32_array(kv_start), _capi.i32_array(rope_delta), _capi.i32_array(vrt_lo), _capi.i32_array(vrt_hi),
                image_embeds.data_ptr(), prototypes.data_ptr(), prototypes.shape[0], B, L,
                _capi.i64_array(list(eos_ids)), len(eos_ids), int(pad_id), C.byref(gb), _stream()), "padt_prefill")
        self._gen_keep = bufs   # keep the borrowed buffers alive for the decode calls
        return bufs

    def decode(self, n_steps: int, finish: bool = False):
        with torch.cuda.device(self.device):
            _capi.check(self.lib.padt_decode(self.handle, int(n_steps), int(bool(finish)), _stream()), "padt_decode")

    def status(self) -> Tuple[int, int]:
        a, b = C.c_int32(), C.c_int32()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.padt_decode_status(self.handle, C.byref(a), C.byref(b), _stream()), "padt_decode_status")
        return a.value, b.value

    # ------------------------------------------------------------------ T2 for every position (teacher-forced forward, padt.py:292-301)
    def extended_logits(self, hidden: torch.Tensor, prototypes: torch.Tensor, vrt_lo: Sequence[int], vrt_hi: Sequence[int]) -> torch.Tensor:
        """hidden [B,L,D] final-norm states -> logits [B,L,V+sum M] bf16 = hidden @ cat(head, prototypes).T with -inf outside each row's own
        VRT range.  Two GEMMs of the library write the two column ranges of ONE buffer (the concatenated table is never materialised)."""
        B, L, D = hidden.shape
        V, n_p = self.cfg.text.vocab_size, prototypes.shape[0]
        head = self._keep_by_name["model.embed_tokens.weight" if self.cfg.text.tie_word_embeddings else "lm_head.weight"]
        ld = (V + n_p + 7) // 8 * 8
        out = torch.empty((B * L, ld), dtype=torch.bfloat16, device=self.device)
        h2 = hidden.reshape(B * L, D).contiguous()
        st = _stream()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.padt_op_gemm(0, h2.data_ptr(), D, head.data_ptr(), D, None, None, 0, out.data_ptr(), ld, None, B * L, V, D, st), "logits (head rows)")
            if n_p > 0:
                _capi.check(self.lib.padt_op_gemm(0, h2.data_ptr(), D, prototypes.data_ptr(), D, None, None, 0, out.data_ptr() + 2 * V, ld, None,
                                                  B * L, n_p, D, st), "logits (prototype rows)")
                _capi.check(self.lib.padt_op_logit_mask(out.data_ptr(), ld, L, B, V, V + n_p, _capi.i32_array(vrt_lo), _capi.i32_array(vrt_hi), st),
                            "logit mask")
        return out.view(B, L, ld)[:, :, : V + n_p]

    # ------------------------------------------------------------------ D1-D4
    def vl_decode(self, vrt_feats: torch.Tensor, obj_vrt_count: Sequence[int], obj_sample: Sequence[int], low: torch.Tensor,
                  high: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, grid_thw: torch.Tensor, want_mask: bool = True):
        cfg = self.cfg
        n_obj = len(obj_vrt_count)
        grid = [int(x) for x in grid_thw.reshape(-1).tolist()]
        n_img = len(grid) // 3
        hmax = max(grid[3 * s + 1] for s in obj_sample)
        wmax = max(grid[3 * s + 2] for s in obj_sample)
        dev = self.device
        boxes = torch.empty((n_obj, 4), dtype=torch.bfloat16, device=dev)
        scores = torch.empty((n_obj, 1), dtype=torch.bfloat16, device=dev)
        masks = torch.empty((n_obj, 4 * hmax, 4 * wmax), dtype=torch.bfloat16, device=dev) if want_mask else None
        vf = vrt_feats.to(device=dev, dtype=torch.bfloat16).contiguous()
        low = low.contiguous(); high = high.contiguous(); cos = cos.contiguous(); sin = sin.contiguous()
        assert vf.shape[0] == sum(obj_vrt_count) and vf.shape[1] == cfg.text.hidden_size
        with torch.cuda.device(dev):
            _capi.check(self.lib.padt_vl_decode(
                self.handle, vf.data_ptr(), _capi.i32_array(obj_vrt_count), _capi.i32_array(obj_sample), n_obj, low.data_ptr(),
                high.data_ptr(), cos.data_ptr(), sin.data_ptr(), _capi.i64_array(grid), n_img, boxes.data_ptr(), scores.data_ptr(),
                _ptr(masks), hmax, wmax, _stream()), "padt_vl_decode")
        return boxes, scores, masks, hmax, wmax
