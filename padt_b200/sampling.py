"""Logits processors and token selection of the reference's sampling loop (host/PyTorch orchestration, device tensors).

The reference's ``_sample_vision_token`` (padt.py:618-800) is the transformers-4.50 ``_sample``: per step ``next_token_scores =
logits_processor(input_ids, next_token_logits)`` (padt.py:717), then ``softmax`` + ``torch.multinomial`` when ``do_sample`` else ``argmax``
(padt.py:740-745).  The processor list comes from the generation config (padt.py:570-580): repetition penalty, then -- for sampling -- the
temperature / top-k / top-p warpers, in that order (transformers ``generation/utils.py::_get_logits_processor``; third-party, not vendored by
the reference: the published algorithms are restated here and cross-checked against the installed transformers classes in
tests/test_sampling.py).  ``scores`` are fp32 rows of the EXTENDED vocabulary (text rows + the VRT rows of the batch), ``input_ids`` the
sequence so far with batch-global VRT ids, so a VRT that was already emitted is penalised like any other repeated token.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch


def repetition_penalty_(input_ids: torch.Tensor, scores: torch.Tensor, penalty: float) -> torch.Tensor:
    """RepetitionPenaltyLogitsProcessor: score of every id already in the sequence is divided by `penalty` if positive, multiplied if negative."""
    picked = torch.gather(scores, 1, input_ids)
    picked = torch.where(picked < 0, picked * penalty, picked / penalty)
    return scores.scatter(1, input_ids, picked)


def temperature_(scores: torch.Tensor, temperature: float) -> torch.Tensor:
    return scores / temperature


def top_k_(scores: torch.Tensor, top_k: int, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1) -> torch.Tensor:
    k = min(max(int(top_k), min_tokens_to_keep), scores.size(-1))
    remove = scores < torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(remove, filter_value)


def top_p_(scores: torch.Tensor, top_p: float, filter_value: float = -float("inf"), min_tokens_to_keep: int = 1) -> torch.Tensor:
    sorted_logits, sorted_indices = torch.sort(scores, descending=False)
    cumulative = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    sorted_remove = cumulative <= (1 - top_p)
    sorted_remove[..., -min_tokens_to_keep:] = 0
    remove = sorted_remove.scatter(1, sorted_indices, sorted_remove)
    return scores.masked_fill(remove, filter_value)


class LogitsPipeline:
    """The processor list of one generate() call.  `custom` entries are HF-style callables ``f(input_ids, scores) -> scores``."""

    def __init__(self, do_sample: bool = False, repetition_penalty: Optional[float] = None, temperature: Optional[float] = None,
                 top_k: Optional[int] = None, top_p: Optional[float] = None, custom: Optional[Sequence[Callable]] = None):
        self.do_sample = bool(do_sample)
        self.repetition_penalty = None if repetition_penalty in (None, 1.0) else float(repetition_penalty)
        # the warpers only exist in sampling mode (transformers: they are added `if generation_config.do_sample`)
        self.temperature = None if (not do_sample or temperature in (None, 1.0)) else float(temperature)
        self.top_k = None if (not do_sample or not top_k) else int(top_k)
        self.top_p = None if (not do_sample or top_p is None or top_p >= 1.0) else float(top_p)
        self.custom: List[Callable] = list(custom or [])
        if self.repetition_penalty is not None and self.repetition_penalty <= 0:
            raise ValueError("`repetition_penalty` has to be a strictly positive float")
        if self.temperature is not None and self.temperature <= 0:
            raise ValueError("`temperature` has to be a strictly positive float")
        if self.top_p is not None and not (0 <= self.top_p <= 1.0):
            raise ValueError("`top_p` has to be a float > 0 and < 1")

    @property
    def is_identity(self) -> bool:
        """True when token selection is a plain arg-max of the raw logits: the device-side fast path applies."""
        return not self.do_sample and self.repetition_penalty is None and not self.custom

    def process(self, input_ids: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
        if self.repetition_penalty is not None:
            scores = repetition_penalty_(input_ids, scores, self.repetition_penalty)
        for f in self.custom:
            scores = f(input_ids, scores)
        if self.temperature is not None:
            scores = temperature_(scores, self.temperature)
        if self.top_k is not None:
            scores = top_k_(scores, self.top_k)
        if self.top_p is not None:
            scores = top_p_(scores, self.top_p)
        return scores

    def select(self, scores: torch.Tensor) -> torch.Tensor:
        """padt.py:740-745."""
        if self.do_sample:
            probs = torch.nn.functional.softmax(scores, dim=-1)
            return torch.multinomial(probs, num_samples=1).squeeze(1)
        return torch.argmax(scores, dim=-1)
