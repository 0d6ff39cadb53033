"""``PaDTForConditionalGeneration`` -- drop-in model class for the reference's generate/decode path.

Mirrors the public surface callers of ``src/PaDT/models/padt.py`` use (``eval/test_demo.py:20-113``,
``eval/evaluation_scripts/utils.py:221-266``): ``from_pretrained``, ``.generate(**processor_batch, ...)``,
``.vl_decode(...)``, ``forward(..., is_main=False)``, ``.config``, ``.model.embed_tokens.weight``, ``.device``, ``.dtype``.
Weight loading and request orchestration stay in Python/PyTorch; every FLOP runs in ``libpadt_b200.so``.
"""
from __future__ import annotations

import glob
import json
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _capi
from .config import PRESETS, DecoderConfig, PadtConfig, TextConfig, VisionConfig
from .engine import PadtEngine
from .rope_index import get_rope_index
from .weights import random_weights, weight_shapes


class CustomGenerateDecoderOnlyOutput(dict):
    """Same fields as the reference's dataclass (padt.py:40-45, 786-798); supports ``out['sequences']`` and ``out.sequences``."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


def _hf_style_config(cfg: PadtConfig) -> SimpleNamespace:
    t, v, d = cfg.text, cfg.vision, cfg.decoder
    return SimpleNamespace(
        hidden_size=t.hidden_size, vocab_size=t.vocab_size, image_token_id=cfg.image_token_id,
        video_token_id=cfg.video_token_id, vision_start_token_id=cfg.vision_start_token_id,
        vision_end_token_id=cfg.vision_end_token_id, tie_word_embeddings=t.tie_word_embeddings,
        eos_token_id=list(cfg.eos_token_ids), pad_token_id=cfg.pad_token_id,
        use_visual_prototype_projection=cfg.use_visual_prototype_projection,
        vision_config=SimpleNamespace(spatial_merge_size=v.spatial_merge_size, hidden_size=v.hidden_size, num_heads=v.num_heads,
                                      patch_size=v.patch_size, window_size=v.window_size, depth=v.depth),
        vl_decoder=dict(name="PaDTDecoder", attn_implementation="padt_b200", hidden_size=d.hidden_size,
                        intermediate_size=d.intermediate_size, llm_hidden_state=t.hidden_size, num_heads=d.num_heads,
                        spatial_merge_size=v.spatial_merge_size, use_mask_loss=d.use_mask_loss),
        padt=cfg,
    )


class PaDTForConditionalGeneration:
    def __init__(self, cfg: PadtConfig, weights: Dict[str, torch.Tensor], device=None):
        self.padt_config = cfg
        self.config = _hf_style_config(cfg)
        self.dtype = torch.bfloat16
        self.weights = weights
        self.engine = PadtEngine(cfg, weights)
        self.device = self.engine.device
        # attribute paths callers touch: model.model.embed_tokens.weight.shape[0] (eval/test_demo.py:31)
        self.model = SimpleNamespace(embed_tokens=SimpleNamespace(weight=weights["model.embed_tokens.weight"]))
        self.generation_config = SimpleNamespace(eos_token_id=list(cfg.eos_token_ids), pad_token_id=cfg.pad_token_id)
        self.rope_deltas = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_random(cls, cfg_or_name="PaDT_Pro_3B", seed: int = 0, device="cuda:0"):
        cfg = PRESETS[cfg_or_name]() if isinstance(cfg_or_name, str) else cfg_or_name
        return cls(cfg, random_weights(cfg, seed=seed, device=device))

    @classmethod
    def from_state_dict(cls, cfg: PadtConfig, state_dict: Dict[str, torch.Tensor], device="cuda:0"):
        w = {}
        for name, shape in weight_shapes(cfg).items():
            src = name
            if src not in state_dict:   # transformers >= 4.52 nests the towers under model.{visual,language_model}
                if name.startswith("visual."):
                    src = "model." + name
                elif name.startswith("model."):
                    src = "model.language_model." + name[len("model."):]
            if src not in state_dict:
                raise KeyError(f"checkpoint is missing '{name}'")
            t = state_dict[src].to(device=device, dtype=torch.bfloat16).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: checkpoint shape {tuple(t.shape)} != {tuple(shape)}")
            w[name] = t
        return cls(cfg, w)

    @classmethod
    def from_pretrained(cls, path: str, torch_dtype=torch.bfloat16, attn_implementation: Optional[str] = None, device_map=None,
                        config=None, **kwargs):
        """HF-layout checkpoint directory (config.json + *.safetensors, HF-4.50 key names); same call shape as
        eval/test_demo.py:20-25.  Weight IO is host plumbing; needs the `safetensors` package."""
        if torch_dtype not in (torch.bfloat16, "bfloat16", None):
            raise ValueError("padt_b200 computes in bf16 (the dtype the reference evaluates in, eval/test_demo.py:22)")
        device = "cuda:0"
        if isinstance(device_map, dict) and "" in device_map:
            dm = device_map[""]
            device = f"cuda:{dm}" if isinstance(dm, int) else str(dm)
        with open(os.path.join(path, "config.json")) as f:
            hc = json.load(f)
        cfg = config_from_hf_dict(hc)
        try:
            from safetensors.torch import load_file
        except Exception as e:  # pragma: no cover
            raise _capi.PadtError("from_pretrained needs the `safetensors` package") from e
        sd = {}
        for fn in sorted(glob.glob(os.path.join(path, "*.safetensors"))):
            sd.update(load_file(fn))
        return cls.from_state_dict(cfg, sd, device=device)

    # ------------------------------------------------------------------ dispatch (padt.py:141-145)
    def forward(self, *args, is_main=True, **kwargs):
        if is_main:
            raise NotImplementedError("teacher-forced forward_main is the training path (SURVEY.md §8f row 1); use generate()")
        return self.vl_decode(*args, **kwargs)

    __call__ = forward

    def eval(self):
        return self

    # ------------------------------------------------------------------ generate (padt.py:414-616 -> 618-800)
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor = None, attention_mask: Optional[torch.Tensor] = None, pixel_values: torch.Tensor = None,
                 image_grid_thw: torch.Tensor = None, use_cache: bool = True, max_new_tokens: int = 20, do_sample: bool = False,
                 output_hidden_states: bool = False, return_dict_in_generate: bool = False, synced_gpus: bool = False,
                 eos_token_id=None, pad_token_id=None, forced_tokens: Optional[torch.Tensor] = None, prototype_group=None,
                 global_image_grid_thw: Optional[torch.Tensor] = None, _prefill_only: bool = False, **unused):
        """``prototype_group`` (a torch.distributed group) switches on the sample-sharded mode: this rank's
        rows are generated exactly as in a single-process run of the LOCAL batch -- ``sequences``, ``past_image_embeds``,
        ``past_logit_mask`` are rank-local, so the reference flow ``assign_to_local_vrt_id(seq, local_grid)`` ->
        ``parseVRTintoCompletion`` -> ``vl_decode(feats, out.past_image_embeds, ..., local_grid, ...)`` is correct on every
        rank -- while ONE all-gather of the prototype rows runs on a side stream.  ``out.global_view()`` waits for it and
        returns the single-process view of the whole batch: (global table, row offset of this rank, sequences with
        VRT ids shifted by that offset).  ``global_image_grid_thw`` [B_global,3] (rank-strided samples) lets the row counts be
        derived from the grids, without a count collective."""
        cfg = self.padt_config
        if do_sample:
            raise NotImplementedError("only greedy decoding (do_sample=False), the mode every reference script uses (utils.py:233-236)")
        if pixel_values is None or image_grid_thw is None:
            raise ValueError("PaDT generate needs pixel_values and image_grid_thw (text-only prompts crash in the reference too, padt.py:230-232,292)")
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, L]")
        B, L = input_ids.shape
        if image_grid_thw.shape[0] != B:
            raise ValueError("one image per sample (padt.py:196-201, padt_sft_trainer.py:341)")
        dev = self.device
        ids_cpu = input_ids.detach().to("cpu", torch.int64)
        mask_cpu = torch.ones_like(ids_cpu) if attention_mask is None else attention_mask.detach().to("cpu", torch.int64)
        grid_cpu = image_grid_thw.detach().to("cpu", torch.int64)
        # ---- ViT + prototypes
        image_embeds, high_res, visual_pe = self.engine.vit_forward(pixel_values, grid_cpu)
        protos = self.engine.prototypes(image_embeds)
        merge2 = cfg.vision.spatial_merge_size ** 2
        counts = (grid_cpu[:, 0] * grid_cpu[:, 1] * grid_cpu[:, 2] // merge2).tolist()
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + int(c))
        exchange = None
        if prototype_group is not None:
            # sample-sharded inference: ONE all-gather of past_image_embeds, on a side stream -- decode reads only
            # the local rows, so the collective overlaps the local prefill + decode instead of sitting between ViT and prefill
            import torch.distributed as dist
            from .distributed import PrototypeExchange, rank_row_counts
            rc = None
            if global_image_grid_thw is not None:
                rc = rank_row_counts(global_image_grid_thw.detach().to("cpu", torch.int64), dist.get_world_size(prototype_group),
                                     cfg.vision.spatial_merge_size)
            exchange = PrototypeExchange(protos, prototype_group, rc)
        n_table = protos.shape[0]
        V = cfg.text.vocab_size
        if int(ids_cpu.max()) >= V + n_table:
            raise AssertionError("input_ids.max() must be below the extended vocabulary (padt.py:203)")
        is_img = ids_cpu == cfg.image_token_id
        n_tok = int(is_img.sum())
        if n_tok != image_embeds.shape[0]:
            raise ValueError(f"Image features and image tokens do not match: tokens: {n_tok}, features {image_embeds.shape[0]}")
        slot = (torch.cumsum(is_img.reshape(-1).to(torch.int64), 0) - 1).to(torch.int32)
        slot = torch.where(is_img.reshape(-1), slot, torch.full_like(slot, -1)).reshape(B, L)
        # left padding only (the reference pads left, eval/test_demo.py:77): mask rows must be 0..0 1..1
        n_real = mask_cpu.sum(dim=1)
        kv_start = (L - n_real).tolist()
        for b in range(B):
            if int(mask_cpu[b, int(kv_start[b]):].min()) != 1:
                raise ValueError("attention_mask must be left-padded")
        pos3, deltas = get_rope_index(ids_cpu, grid_cpu, mask_cpu, cfg.image_token_id, cfg.vision_start_token_id, cfg.vision.spatial_merge_size)
        self.rope_deltas = deltas.to(dev)
        eos = list(cfg.eos_token_ids) if eos_token_id is None else ([eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id))
        pad = cfg.pad_token_id if pad_token_id is None else pad_token_id
        T = int(max_new_tokens)
        forced = None
        if forced_tokens is not None:
            forced = forced_tokens.t().contiguous()            # [B,T] -> [T,B]
        bufs = self.engine.prefill(input_ids, slot, pos3.reshape(3, B * L), kv_start, deltas.reshape(-1).tolist(), offs[:-1], offs[1:],
                                   image_embeds, protos, T, eos, pad, want_hidden=output_hidden_states, forced=forced)
        if _prefill_only:
            return None
        # ---- decode: chunks without host sync; poll the finished flag between chunks only when EOS can stop us
        remaining = T - 1
        stopped_early = False
        chunk = 32 if len(eos) > 0 else max(remaining, 1)
        while remaining > 0:
            n = min(chunk, remaining)
            self.engine.decode(n, finish=False)
            remaining -= n
            if len(eos) > 0 and remaining > 0:
                _, unfinished = self.engine.status()
                if unfinished == 0:
                    stopped_early = True
                    break
        if not stopped_early:
            self.engine.decode(0, finish=True)
        n_sampled, _ = self.engine.status()
        toks = bufs["tokens"][:n_sampled]                       # [t,B]
        t_gen = n_sampled
        if len(eos) > 0:
            # generation stops at the first step after which every row has emitted EOS (padt.py:756-757)
            is_eos = torch.isin(toks, torch.tensor(eos, device=toks.device))
            done = torch.cummax(is_eos.to(torch.int8), dim=0).values.all(dim=1)
            idx = torch.nonzero(done)
            if idx.numel() > 0:
                t_gen = int(idx[0]) + 1
        sequences = torch.cat([input_ids.to(dev), toks[:t_gen].t()], dim=1)
        hidden_states = None
        if output_hidden_states:
            hs = bufs["hidden_steps"]
            hidden_states = tuple([(bufs["prefill_hidden"],)] + [(hs[k][:, None, :],) for k in range(1, t_gen)])
        logit_mask = torch.zeros((B, V + n_table), dtype=torch.bool, device=dev)   # padt.py:196-201
        logit_mask[:, :V] = True
        for b in range(B):
            logit_mask[b, V + offs[b]: V + offs[b + 1]] = True
        out = CustomGenerateDecoderOnlyOutput(
            sequences=sequences, scores=None, logits=None, attentions=None, hidden_states=hidden_states, past_key_values=None,
            past_image_embeds=protos, past_logit_mask=logit_mask, past_high_res_image_embeds=high_res, past_visual_pe=visual_pe,
            argmax_tokens=bufs["argmax"][:t_gen].t(),
        )
        if exchange is not None:
            def global_view(_ex=exchange, _seq=sequences, _V=V):
                table, base = _ex.wait()
                return table, base, torch.where(_seq >= _V, _seq + base, _seq)
            out["global_view"] = global_view
        return out if return_dict_in_generate else sequences

    # ------------------------------------------------------------------ vl_decode (padt.py:342-412)
    @torch.no_grad()
    def vl_decode(self, object_vp_feats, low_res_image_embeds, high_res_image_embeds, image_grid_thws, visual_pes):
        dev, dt = self.device, self.dtype
        flat = [f for sample in object_vp_feats for f in sample]
        if len(flat) == 0:   # padt.py:405-412
            return {"pred_boxes": torch.zeros((0, 4), device=dev, dtype=dt), "pred_score": torch.zeros((0, 1), device=dev, dtype=dt),
                    "pred_mask": torch.zeros((0, 8, 8), device=dev, dtype=dt), "pred_mask_valid_hw": (), "sample_idx": []}
        grid_cpu = image_grid_thws.detach().to("cpu", torch.int64)
        counts, samples = [], []
        for si, sample in enumerate(object_vp_feats):
            for f in sample:
                counts.append(int(f.shape[0]))
                samples.append(si)
        feats = torch.cat([f.to(dev) for f in flat], dim=0)
        want_mask = self.padt_config.decoder.use_mask_loss
        boxes, scores, masks, hmax, wmax = self.engine.vl_decode(feats, counts, samples, low_res_image_embeds, high_res_image_embeds,
                                                                 visual_pes[0], visual_pes[1], grid_cpu, want_mask=want_mask)
        Hs = torch.tensor([int(grid_cpu[s, 1]) for s in samples], device=dev, dtype=torch.int64)
        Ws = torch.tensor([int(grid_cpu[s, 2]) for s in samples], device=dev, dtype=torch.int64)
        return {"pred_boxes": boxes, "pred_score": scores, "pred_mask": masks, "pred_mask_valid_hw": (Hs, Ws) if want_mask else (),
                "sample_idx": samples}


def config_from_hf_dict(hc: dict) -> PadtConfig:
    """config.json of a PaDT checkpoint (Qwen2.5-VL layout, either flat 4.50 style or nested text_config) -> PadtConfig."""
    tc = hc.get("text_config", hc)
    vc = hc["vision_config"]
    dec = hc.get("vl_decoder", {})
    rope = tc.get("rope_scaling") or tc.get("rope_parameters") or {}
    vision = VisionConfig(depth=vc["depth"], hidden_size=vc["hidden_size"], intermediate_size=vc["intermediate_size"],
                          num_heads=vc["num_heads"], in_channels=vc.get("in_channels", vc.get("in_chans", 3)), patch_size=vc["patch_size"],
                          temporal_patch_size=vc.get("temporal_patch_size", 2), spatial_merge_size=vc.get("spatial_merge_size", 2),
                          window_size=vc.get("window_size", 112), fullatt_block_indexes=tuple(vc.get("fullatt_block_indexes", (7, 15, 23, 31))),
                          out_hidden_size=vc.get("out_hidden_size", tc["hidden_size"]))
    text = TextConfig(hidden_size=tc["hidden_size"], intermediate_size=tc["intermediate_size"], num_layers=tc["num_hidden_layers"],
                      num_heads=tc["num_attention_heads"], num_kv_heads=tc["num_key_value_heads"], vocab_size=tc["vocab_size"],
                      rms_norm_eps=tc.get("rms_norm_eps", 1e-6), rope_theta=tc.get("rope_theta", rope.get("rope_theta", 1e6)),
                      mrope_section=tuple(rope.get("mrope_section", (16, 24, 24))),
                      tie_word_embeddings=bool(hc.get("tie_word_embeddings", tc.get("tie_word_embeddings", False))))
    decoder = DecoderConfig(hidden_size=dec.get("hidden_size", 1280), intermediate_size=dec.get("intermediate_size", 3420),
                            num_heads=dec.get("num_heads", 16), use_mask_loss=dec.get("use_mask_loss", True))
    eos = hc.get("eos_token_id", 151645)
    eos = tuple(eos) if isinstance(eos, (list, tuple)) else (eos,)
    return PadtConfig(name=hc.get("_name_or_path", "PaDT"), vision=vision, text=text, decoder=decoder,
                      image_token_id=hc.get("image_token_id", 151655), video_token_id=hc.get("video_token_id", 151656),
                      vision_start_token_id=hc.get("vision_start_token_id", 151652), vision_end_token_id=hc.get("vision_end_token_id", 151653),
                      eos_token_ids=eos, pad_token_id=hc.get("pad_token_id", 151643) or 151643,
                      use_visual_prototype_projection=hc.get("use_visual_prototype_projection", True))
