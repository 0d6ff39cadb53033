"""§8f-4 host logic on CPU: the restated logits processors against the installed transformers classes (the third-party code the reference's
sampling loop calls, padt.py:570-580,717), the pipeline order, and multinomial selection under a fixed seed."""
import pytest
import torch

from padt_b200.sampling import LogitsPipeline, repetition_penalty_, temperature_, top_k_, top_p_


def _hf():
    lp = pytest.importorskip("transformers.generation.logits_process")
    return lp


def _case(seed, B=3, V=97, T=11):
    g = torch.Generator().manual_seed(seed)
    scores = torch.randn((B, V), generator=g) * 3
    scores[0, 5:9] = float("-inf")                                  # masked VRT columns stay -inf through every processor
    ids = torch.randint(0, V, (B, T), generator=g)
    return ids, scores


@pytest.mark.parametrize("seed", range(5))
def test_processors_match_transformers(seed):
    lp = _hf()
    ids, scores = _case(seed)
    assert torch.equal(repetition_penalty_(ids, scores.clone(), 1.3), lp.RepetitionPenaltyLogitsProcessor(1.3)(ids, scores.clone()))
    assert torch.equal(temperature_(scores.clone(), 0.7), lp.TemperatureLogitsWarper(0.7)(ids, scores.clone()))
    for k in (1, 5, 1000):
        assert torch.equal(top_k_(scores.clone(), k), lp.TopKLogitsWarper(k)(ids, scores.clone()))
    for p in (0.1, 0.5, 0.95):
        assert torch.equal(top_p_(scores.clone(), p), lp.TopPLogitsWarper(p)(ids, scores.clone()))


def test_pipeline_order_and_selection():
    lp = _hf()
    ids, scores = _case(7)
    pipe = LogitsPipeline(do_sample=True, repetition_penalty=1.05, temperature=0.8, top_k=20, top_p=0.9)
    ref = lp.LogitsProcessorList([lp.RepetitionPenaltyLogitsProcessor(1.05), lp.TemperatureLogitsWarper(0.8), lp.TopKLogitsWarper(20),
                                  lp.TopPLogitsWarper(0.9)])
    out = pipe.process(ids, scores.clone())
    assert torch.equal(out, ref(ids, scores.clone()))
    torch.manual_seed(3)
    a = pipe.select(out)
    torch.manual_seed(3)
    b = torch.multinomial(torch.nn.functional.softmax(out, dim=-1), num_samples=1).squeeze(1)
    assert torch.equal(a, b)
    assert bool(torch.isfinite(out.gather(1, a[:, None])).all()), "a filtered column can never be drawn"


def test_greedy_modes():
    ids, scores = _case(1)
    assert LogitsPipeline().is_identity
    assert not LogitsPipeline(repetition_penalty=1.05).is_identity
    # warpers are sampling-only: greedy + temperature is still an arg-max of the raw scores
    p = LogitsPipeline(do_sample=False, temperature=0.5, top_k=3, top_p=0.5)
    assert p.is_identity and torch.equal(p.select(p.process(ids, scores)), scores.argmax(-1))
    # custom HF-style processor
    bump = lambda i, s: s.index_fill(1, torch.tensor([2]), 1e9)
    p = LogitsPipeline(custom=[bump])
    assert not p.is_identity and p.select(p.process(ids, scores)).tolist() == [2, 2, 2]
    with pytest.raises(ValueError):
        LogitsPipeline(do_sample=True, temperature=0.0)
