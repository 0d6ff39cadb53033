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
from .sampling import LogitsPipeline
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
        # generation defaults the reference takes from generation_config.json (padt.py:436-447): overridden by from_pretrained
        self.generation_config = SimpleNamespace(eos_token_id=list(cfg.eos_token_ids), pad_token_id=cfg.pad_token_id, do_sample=False,
                                                 repetition_penalty=1.0, temperature=1.0, top_k=None, top_p=1.0)
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
        gen = load_generation_config(path, cfg)
        if gen.get("eos_token_id") is not None or gen.get("pad_token_id") is not None:
            import dataclasses
            cfg = dataclasses.replace(cfg, eos_token_ids=tuple(gen["eos_token_id"]), pad_token_id=gen["pad_token_id"])
        model = cls.from_state_dict(cfg, sd, device=device)
        for k, v in gen.items():
            setattr(model.generation_config, k, v)
        return model

    # ------------------------------------------------------------------ dispatch (padt.py:141-145)
    def forward(self, *args, is_main=True, **kwargs):
        if is_main:
            return self.forward_main(*args, **kwargs)
        return self.vl_decode(*args, **kwargs)

    __call__ = forward

    def eval(self):
        return self

    # ------------------------------------------------------------------ shared front end of generate / forward_main (padt.py:183-219, 256-277)
    def _encode(self, input_ids, attention_mask, pixel_values, image_grid_thw):
        """ViT + prototypes + the host-side integer tables of one batch: image slots of the <|image_pad|> tokens (masked_scatter order,
        padt.py:219), left-pad counts, M-RoPE ids (4.50 get_rope_index), per-row VRT ranges (padt.py:196-201)."""
        cfg = self.padt_config
        if pixel_values is None or image_grid_thw is None:
            raise ValueError("PaDT needs pixel_values and image_grid_thw (text-only prompts crash in the reference too, padt.py:230-232,292)")
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, L]")
        B, L = input_ids.shape
        if image_grid_thw.shape[0] != B:
            raise ValueError("one image per sample (padt.py:196-201, padt_sft_trainer.py:341)")
        ids_cpu = input_ids.detach().to("cpu", torch.int64)
        mask_cpu = torch.ones_like(ids_cpu) if attention_mask is None else attention_mask.detach().to("cpu", torch.int64)
        grid_cpu = image_grid_thw.detach().to("cpu", torch.int64)
        image_embeds, high_res, visual_pe = self.engine.vit_forward(pixel_values, grid_cpu)
        protos = self.engine.prototypes(image_embeds)
        merge2 = cfg.vision.spatial_merge_size ** 2
        counts = (grid_cpu[:, 0] * grid_cpu[:, 1] * grid_cpu[:, 2] // merge2).tolist()
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + int(c))
        V = cfg.text.vocab_size
        if int(ids_cpu.max()) >= V + protos.shape[0]:
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
        self.rope_deltas = deltas.to(self.device)
        return SimpleNamespace(B=B, L=L, image_embeds=image_embeds, high_res=high_res, visual_pe=visual_pe, protos=protos, offs=offs, slot=slot,
                               kv_start=kv_start, pos3=pos3, deltas=deltas, grid_cpu=grid_cpu)

    def _logit_mask(self, B, offs, n_table):
        V = self.padt_config.text.vocab_size
        logit_mask = torch.zeros((B, V + n_table), dtype=torch.bool, device=self.device)   # padt.py:196-201
        logit_mask[:, :V] = True
        for b in range(B):
            logit_mask[b, V + offs[b]: V + offs[b + 1]] = True
        return logit_mask

    # ------------------------------------------------------------------ teacher-forced forward (padt.py:147-340), SURVEY.md §8f row 1
    @torch.no_grad()
    def forward_main(self, input_ids: torch.Tensor = None, attention_mask: Optional[torch.Tensor] = None, pixel_values: torch.Tensor = None,
                     image_grid_thw: torch.Tensor = None, labels=None, output_hidden_states: bool = True, return_dict: bool = True,
                     past_key_values=None, **unused):
        """All-position extended logits [B, L, V + sum M] (bf16, -inf outside each row's own VRT range: padt.py:292-301) and the final-norm
        hidden states of the prompt -- what ``PaDTSFTTrainer.compute_loss`` consumes (padt_sft_trainer.py:469-487) -- plus the four
        ``past_*`` attributes (padt.py:322-340).  Forward only (inference engine: no autograd graph); the decode branch of forward_main
        (``past_image_embeds`` given) is what ``generate`` runs on the device and is not exposed here.  ``hidden_states`` holds the final-norm
        states as its only entry, so ``out.hidden_states[-1]`` is what the reference returns there."""
        if past_key_values is not None or unused.get("past_image_embeds") is not None:
            raise NotImplementedError("the cached decode branch of forward_main runs inside generate(); call generate() for incremental decoding")
        if labels is not None:
            raise NotImplementedError("the reference's labels branch views the logits as [-1, vocab_size] and fails with VRT columns (padt.py:304-316)")
        e = self._encode(input_ids, attention_mask, pixel_values, image_grid_thw)
        cfg = self.padt_config
        bufs = self.engine.prefill(input_ids, e.slot, e.pos3.reshape(3, e.B * e.L), e.kv_start, e.deltas.reshape(-1).tolist(), e.offs[:-1], e.offs[1:],
                                   e.image_embeds, e.protos, 1, [], cfg.pad_token_id, want_hidden=True)
        hidden = bufs["prefill_hidden"]
        logits = self.engine.extended_logits(hidden, e.protos, e.offs[:-1], e.offs[1:])
        out = CustomGenerateDecoderOnlyOutput(
            loss=None, logits=logits, past_key_values=None, hidden_states=(hidden,) if output_hidden_states else None, attentions=None,
            rope_deltas=self.rope_deltas, past_image_embeds=e.protos, past_logit_mask=self._logit_mask(e.B, e.offs, e.protos.shape[0]),
            past_high_res_image_embeds=e.high_res, past_visual_pe=e.visual_pe)
        return out if return_dict else (logits,)

    # ------------------------------------------------------------------ generate (padt.py:414-616 -> 618-800)
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor = None, attention_mask: Optional[torch.Tensor] = None, pixel_values: torch.Tensor = None,
                 image_grid_thw: torch.Tensor = None, use_cache: bool = True, max_new_tokens: int = 20, do_sample: Optional[bool] = None,
                 output_hidden_states: bool = False, return_dict_in_generate: bool = False, synced_gpus: bool = False,
                 eos_token_id=None, pad_token_id=None, forced_tokens: Optional[torch.Tensor] = None, prototype_group=None,
                 global_image_grid_thw: Optional[torch.Tensor] = None, temperature: Optional[float] = None, top_k: Optional[int] = None,
                 top_p: Optional[float] = None, repetition_penalty: Optional[float] = None, logits_processor=None,
                 prototype_exchange: str = "lazy", _prefill_only: bool = False, **unused):
        """``prototype_group`` (a torch.distributed group) switches on the sample-sharded mode: this rank's
        rows are generated exactly as in a single-process run of the LOCAL batch -- ``sequences``, ``past_image_embeds``,
        ``past_logit_mask`` are rank-local, so the reference flow ``assign_to_local_vrt_id(seq, local_grid)`` ->
        ``parseVRTintoCompletion`` -> ``vl_decode(feats, out.past_image_embeds, ..., local_grid, ...)`` is correct on every
        rank.  ``out.global_view()`` (a collective call: every rank makes it) all-gathers the prototype rows -- or, with
        ``prototype_exchange="eager"``, waits for the all-gather that was issued on a side stream during generate -- and
        returns the single-process view of the whole batch: (global table, row offset of this rank, sequences with
        VRT ids shifted by that offset).  ``global_image_grid_thw`` [B_global,3] (rank-strided samples) lets the row counts be
        derived from the grids, without a count collective.

        Token selection (padt.py:717,740-745).  Plain greedy (no sampling, no processors: what every reference script uses,
        utils.py:233-236) stays on the device: masked arg-max inside the decode kernel, no host round trip per token.  ``do_sample`` /
        ``temperature`` / ``top_k`` / ``top_p`` / ``repetition_penalty`` / ``logits_processor`` (defaults from ``generation_config``,
        as in the reference) switch to the processor path: the kernel also writes the masked bf16 logit rows, the processors and
        ``torch.multinomial`` / ``argmax`` run on them (padt_b200/sampling.py), and the chosen ids are fed back through the
        teacher-forcing slot of the sampler -- still without a host synchronisation per token."""
        cfg = self.padt_config
        gc = self.generation_config
        pipe = LogitsPipeline(do_sample=gc.do_sample if do_sample is None else do_sample,
                              repetition_penalty=gc.repetition_penalty if repetition_penalty is None else repetition_penalty,
                              temperature=gc.temperature if temperature is None else temperature, top_k=gc.top_k if top_k is None else top_k,
                              top_p=gc.top_p if top_p is None else top_p, custom=logits_processor)
        e = self._encode(input_ids, attention_mask, pixel_values, image_grid_thw)
        B, L = e.B, e.L
        dev = self.device
        image_embeds, high_res, visual_pe, protos, offs = e.image_embeds, e.high_res, e.visual_pe, e.protos, e.offs
        exchange = None
        if prototype_group is not None:
            # sample-sharded inference: every output is rank-local and decode reads only the local prototype rows, so
            # generate itself holds NO collective; the ONE all-gather of past_image_embeds that rebuilds the single-process view runs when
            # out.global_view() asks for it (a collective call).  prototype_exchange="eager": issue it now on a side stream instead, where
            # it overlaps the local prefill + decode (couples the ranks once per request).
            import torch.distributed as dist
            from .distributed import PrototypeExchange, rank_row_counts
            rc = None
            if global_image_grid_thw is not None:
                rc = rank_row_counts(global_image_grid_thw.detach().to("cpu", torch.int64), dist.get_world_size(prototype_group),
                                     cfg.vision.spatial_merge_size)
            eager = prototype_exchange == "eager" or bool(os.environ.get("PADT_EAGER_ALLGATHER")) or bool(os.environ.get("PADT_SYNC_ALLGATHER"))
            exchange = PrototypeExchange(protos, prototype_group, rc, lazy=not eager)
            if os.environ.get("PADT_SYNC_ALLGATHER"):      # A/B switch: order the collective before the prefill, as round 1 did
                exchange.wait()
        n_table = protos.shape[0]
        V = cfg.text.vocab_size
        eos = list(gc.eos_token_id) if eos_token_id is None else ([eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id))
        pad = gc.pad_token_id if pad_token_id is None else pad_token_id
        T = int(max_new_tokens)
        if not pipe.is_identity and forced_tokens is not None:
            raise ValueError("forced_tokens (teacher forcing) cannot be combined with sampling / logits processors")
        forced = None
        if forced_tokens is not None:
            forced = forced_tokens.t().contiguous()            # [B,T] -> [T,B]
        elif not pipe.is_identity:
            forced = torch.zeros((T, B), dtype=torch.int64, device=dev)     # filled one row per step by the processor path
        bufs = self.engine.prefill(input_ids, e.slot, e.pos3.reshape(3, B * L), e.kv_start, e.deltas.reshape(-1).tolist(), offs[:-1], offs[1:],
                                   image_embeds, protos, T, eos, pad, want_hidden=output_hidden_states, forced=forced,
                                   want_logits=not pipe.is_identity)
        if _prefill_only:
            return None
        if pipe.is_identity:
            stopped_early = self._decode_greedy(T, eos)
        else:
            stopped_early = self._decode_with_processors(pipe, bufs, input_ids.to(dev), T, eos, pad, V + n_table)
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
        out = CustomGenerateDecoderOnlyOutput(
            sequences=sequences, scores=None, logits=None, attentions=None, hidden_states=hidden_states, past_key_values=None,
            past_image_embeds=protos, past_logit_mask=self._logit_mask(B, offs, n_table), past_high_res_image_embeds=high_res,
            past_visual_pe=visual_pe, argmax_tokens=bufs["argmax"][:t_gen].t(),
        )
        if exchange is not None:
            def global_view(_ex=exchange, _seq=sequences, _V=V):
                table, base = _ex.wait()
                return table, base, torch.where(_seq >= _V, _seq + base, _seq)
            out["global_view"] = global_view
        return out if return_dict_in_generate else sequences

    def _decode_greedy(self, T: int, eos) -> bool:
        """Device-side greedy: chunks without host sync; the finished flag is polled between chunks only when EOS can stop us."""
        remaining = T - 1
        chunk = 32 if len(eos) > 0 else max(remaining, 1)
        while remaining > 0:
            n = min(chunk, remaining)
            self.engine.decode(n, finish=False)
            remaining -= n
            if len(eos) > 0 and remaining > 0:
                _, unfinished = self.engine.status()
                if unfinished == 0:
                    return True
        return False

    def _decode_with_processors(self, pipe: LogitsPipeline, bufs, prompt_ids: torch.Tensor, T: int, eos, pad, n_ext: int) -> bool:
        """padt.py:713-752 around the device step: scores = logits[:, -1].float() -> processors -> multinomial / argmax -> the chosen id goes
        into the sampler's teacher-forcing slot `forced[t]`, which the next decode launch reads when it samples token t (finished rows are
        turned into pad ids by the device-side bookkeeping exactly as in the greedy path).  One decode launch per token, no host sync."""
        logits, forced = bufs["logits"], bufs["forced"]
        seq = prompt_ids
        eos_t = torch.tensor(eos, device=seq.device) if len(eos) > 0 else None
        unfinished = torch.ones((seq.shape[0],), dtype=torch.bool, device=seq.device)
        for t in range(T):
            scores = logits[:, :n_ext].float()                                    # padt.py:713 (.clone().float())
            nxt = pipe.select(pipe.process(seq, scores))                          # padt.py:717,740-745
            forced[t].copy_(nxt)
            if eos_t is not None:                                                 # mirror of padt.py:748-757 for the processors' input_ids
                nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad))
                unfinished = unfinished & ~torch.isin(nxt, eos_t)
            seq = torch.cat([seq, nxt[:, None]], dim=1)
            if t == T - 1:
                break
            self.engine.decode(1, finish=False)                                   # samples token t (= forced[t]), runs the layers, new logits
            if eos_t is not None and (t % 32) == 31 and not bool(unfinished.any()):
                return True
        return False

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


def load_generation_config(path: str, cfg: PadtConfig) -> dict:
    """generation_config.json of a checkpoint directory -> the generate() defaults the reference's HF ``generate`` takes from it
    (padt.py:436-447): EOS ids (Qwen2.5-VL ships [151645, 151643]: <|im_end|> AND <|endoftext|>), pad id, sampling switches, repetition
    penalty.  Missing file / keys fall back to the PadtConfig defaults, never to config.json's single eos id."""
    out = {"eos_token_id": None, "pad_token_id": None}
    fn = os.path.join(path, "generation_config.json")
    if not os.path.exists(fn):
        return out
    with open(fn) as f:
        g = json.load(f)
    eos = g.get("eos_token_id")
    if eos is not None:
        out["eos_token_id"] = [int(x) for x in (eos if isinstance(eos, (list, tuple)) else [eos])]
        pad = g.get("pad_token_id")
        out["pad_token_id"] = int(pad) if pad is not None else cfg.pad_token_id
    elif g.get("pad_token_id") is not None:
        out["eos_token_id"], out["pad_token_id"] = list(cfg.eos_token_ids), int(g["pad_token_id"])
    for k in ("do_sample", "repetition_penalty", "temperature", "top_k", "top_p"):
        if g.get(k) is not None:
            out[k] = g[k]
    return out


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
    # config.json of Qwen2.5-VL carries a single eos id (<|im_end|>); the reference's generate() takes the PAIR from generation_config.json
    # (load_generation_config).  A single id here is therefore not authoritative: keep the (151645, 151643) default unless a list is given.
    eos = hc.get("eos_token_id")
    eos = tuple(int(x) for x in eos) if isinstance(eos, (list, tuple)) else PadtConfig().eos_token_ids
    return PadtConfig(name=hc.get("_name_or_path", "PaDT"), vision=vision, text=text, decoder=decoder,
                      image_token_id=hc.get("image_token_id", 151655), video_token_id=hc.get("video_token_id", 151656),
                      vision_start_token_id=hc.get("vision_start_token_id", 151652), vision_end_token_id=hc.get("vision_end_token_id", 151653),
                      eos_token_ids=eos, pad_token_id=hc.get("pad_token_id", 151643) or 151643,
                      use_visual_prototype_projection=hc.get("use_visual_prototype_projection", True))
