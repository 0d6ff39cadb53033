"""Host-side API surface: ``VisonTextProcessingClass`` (sic) and ``parseVRTintoCompletion``.

Drop-in for ``src/PaDT/models/padt_processor.py`` of the reference (same names, arguments, return values, in-place
mutation and error behaviour).  Pure host integer/string work; nothing here touches the GPU kernels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def _merged_patch_offsets(image_grid_thw: torch.Tensor, merge: int) -> torch.Tensor:
    """Exclusive prefix sums of per-sample merged-patch counts, with the total appended (padt_processor.py:39,47,70)."""
    per_sample = image_grid_thw.cumprod(-1)[:, -1] // (merge ** 2)
    return torch.nn.functional.pad(per_sample.cumsum(dim=-1), (1, 0), "constant", 0)


def _added_token(text: str, special: bool):
    try:  # the reference builds transformers AddedToken objects (padt_processor.py:20,27)
        from transformers.tokenization_utils import AddedToken
        return AddedToken(text, lstrip=False, rstrip=False, special=special, normalized=False)
    except Exception:  # fake tokenizers in the tests accept plain strings
        return text


class VisonTextProcessingClass(object):
    """Wraps an HF processor and grows its tokenizer with ``<|VRT_i|>`` tokens (padt_processor.py:4-57)."""

    def __init__(self, processing_class, spatial_merge_size: int = 2):
        self.processing_class = processing_class
        self.spatial_merge_size = spatial_merge_size
        self.model_embed_token_size = len(processing_class.tokenizer.get_vocab())

    def __getattr__(self, name: str):
        # passthrough (apply_chat_template, batch_decode, tokenizer, ...), padt_processor.py:10-13
        inner = self.__dict__.get("processing_class")
        if inner is not None and hasattr(inner, name):
            return getattr(inner, name)
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def prepare(self, model_embed_token_size: int) -> bool:
        """Pad the tokenizer up to the embedding-table row count (padt_processor.py:15-21)."""
        self.model_embed_token_size = model_embed_token_size
        missing = model_embed_token_size - len(self.tokenizer.get_vocab())
        assert "<|empty_token_0|>" in self.tokenizer.vocab or missing > 0
        if missing > 0:
            self.tokenizer.add_tokens([_added_token("<|empty_token_%d|>" % i, special=True) for i in range(missing)])
        return True

    def set_image_grid_thw(self, image_grid_thw: torch.Tensor) -> bool:
        """Make sure ``<|VRT_0|> .. <|VRT_{max merged patches - 1}|>`` exist (padt_processor.py:23-28)."""
        need = image_grid_thw.cumprod(-1).max(dim=0)[0][-1] // (self.spatial_merge_size) ** 2
        tok = self.processing_class.tokenizer
        have = len(tok.get_vocab()) - self.model_embed_token_size
        if have < need:
            tok.add_tokens([_added_token("<|VRT_%d|>" % i, special=False) for i in range(have, int(need))])
        return True

    def __call__(self, *args, **kwargs):
        out = self.processing_class(*args, **kwargs)
        if "image_grid_thw" in out:
            self.set_image_grid_thw(out["image_grid_thw"])
        return out

    def _shift(self, input_ids: torch.Tensor, image_grid_thw: torch.Tensor, sign: int) -> torch.Tensor:
        is_vrt = input_ids >= self.model_embed_token_size
        if is_vrt.sum() > 0:
            starts = _merged_patch_offsets(image_grid_thw, self.spatial_merge_size)[:-1, None].expand(-1, input_ids.shape[1])
            starts = starts.to(input_ids.device)
            if sign > 0:
                input_ids[is_vrt] += starts[is_vrt]
            else:
                input_ids[is_vrt] -= starts[is_vrt]
        return input_ids

    def assign_to_global_vrt_id(self, input_ids: torch.Tensor, image_grid_thw: torch.Tensor) -> torch.Tensor:
        """local -> batch-global VRT ids, IN PLACE (padt_processor.py:36-42)."""
        return self._shift(input_ids, image_grid_thw, +1)

    def assign_to_local_vrt_id(self, input_ids: torch.Tensor, image_grid_thw: torch.Tensor) -> torch.Tensor:
        """batch-global -> local VRT ids, IN PLACE (padt_processor.py:44-50)."""
        return self._shift(input_ids, image_grid_thw, -1)

    def pid2vrt(self, patch_ids) -> str:
        ids = [patch_ids] if type(patch_ids) == int else [int(i) for i in patch_ids]
        return "".join("<|VRT_%d|>" % i for i in ids)


def _scan_sample(processor, pieces: Sequence[str], hidden_states, batch_idx: int, skip_thinking: bool,
                 image_prototype, image_grid_thw, vision_patch_nums):
    """Token-string state machine of padt_processor.py:85-141 for one sample.

    `pieces[i]` raises IndexError past the end exactly like the reference's list lookups do; the caller turns any
    exception into empty results for the sample (padt_processor.py:146-150)."""
    feats: List[torch.Tensor] = []
    labels: List[str] = []
    vrts: List[str] = []
    vrt_feats: List[torch.Tensor] = []
    eos = processor.tokenizer.eos_token
    i = 0
    in_answer = False
    in_name = False
    label = ""
    n = len(pieces)
    while i < n:
        cur = pieces[i]
        if eos in cur:
            break
        # "<answer>" arrives as three pieces: '<', 'answer', '>'
        if in_answer is False and "<" in cur and "</" not in cur and "answer" in pieces[i + 1] and ">" in pieces[i + 2]:
            in_answer = True
            i += 3
            continue
        if in_answer is True or skip_thinking:
            if "</" in cur and "answer" in pieces[i + 1] and ">" in pieces[i + 2]:
                in_answer = False
                break
            if '"' in cur and in_name is False:
                in_name = True
                label = cur.split('"')[1]
                i += 1
                continue
            if '"' in cur and in_name is True:
                in_name = False
                label += cur.split('"')[0]
                label = label.strip()
                i += 1
                continue
            if "<|VRT_" in cur:
                in_name = False
                states = []
                text = ""
                while "<|VRT_" in pieces[i]:
                    states.append(hidden_states[i][-1][batch_idx])   # state that predicted this VRT (padt_processor.py:125)
                    text += pieces[i]
                    i += 1
                feats.append(torch.cat(states, dim=0))
                labels.append(label)
                vrts.append(text)
                if image_prototype is not None and image_grid_thw is not None:
                    ids = processor(text=text, return_tensors="pt")["input_ids"].to(image_grid_thw.device)[0, ...]
                    ids = ids + vision_patch_nums[batch_idx] - processor.model_embed_token_size
                    vrt_feats.append(image_prototype[ids])
                continue
            if in_name:
                label += cur
        i += 1
    return feats, labels, vrts, vrt_feats


def parseVRTintoCompletion(processor, completion_ids, hidden_states, need_thinking_mask=None, image_prototype=None,
                           image_grid_thw=None):
    """Completion ids -> (completions, per-object VRT hidden states, labels, VRT strings, prototype rows).

    Same contract as padt_processor.py:60-151: `completion_ids` are LOCAL ids [B,T]; `hidden_states[step][-1][b]` is the
    final-norm state of generation step `step`; a sample whose scan raises yields four empty lists."""
    completions, all_feats, all_labels, all_vrts, all_vrt_feats = [], [], [], [], []
    vision_patch_nums = None
    if image_grid_thw is not None:
        vision_patch_nums = torch.nn.functional.pad((image_grid_thw.cumprod(-1)[:, -1] // 4).cumsum(-1), (1, 0), "constant", 0)
    if need_thinking_mask is None:
        need_thinking_mask = torch.ones(len(completion_ids)).to(torch.bool)
    for b, completion in enumerate(completion_ids):
        pieces = processor.batch_decode(completion)
        completions.append("".join(pieces))
        skip_thinking = not need_thinking_mask[b].item()
        try:
            f, l, v, vf = _scan_sample(processor, pieces, hidden_states, b, skip_thinking, image_prototype, image_grid_thw,
                                       vision_patch_nums)
        except Exception:
            f, l, v, vf = [], [], [], []
        all_feats.append(f)
        all_labels.append(l)
        all_vrts.append(v)
        all_vrt_feats.append(vf)
    return completions, all_feats, all_labels, all_vrts, all_vrt_feats
