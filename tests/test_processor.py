"""P1/P2: processor wrapper + completion parser against the reference's own KATs (eval/outputs/demo/completion.txt, the id-shift
example of SURVEY.md §8c) and against the oracle's arithmetic.  Host-only."""
import re

import torch

import padt_oracle as O
from padt_b200 import VisonTextProcessingClass, parseVRTintoCompletion


class FakeTokenizer:
    eos_token = "<|im_end|>"

    def __init__(self, n):
        self.vocab = {f"tok{i}": i for i in range(n)}

    def get_vocab(self):
        return self.vocab

    def add_tokens(self, toks):
        for t in toks:
            s = t if isinstance(t, str) else t.content
            if s not in self.vocab:
                self.vocab[s] = len(self.vocab)


class FakeProcessor:
    """Only what the parser/processor wrapper touch."""

    def __init__(self, V, strings):
        self.tokenizer = FakeTokenizer(V)
        self.V = V
        self.strings = strings     # id -> piece for ids < V

    def batch_decode(self, ids):
        return [self.strings.get(int(i), f"<|VRT_{int(i) - self.V}|>") if int(i) < self.V else f"<|VRT_{int(i) - self.V}|>" for i in ids]

    def __call__(self, text=None, return_tensors=None, **kw):
        ids = [self.V + int(k) for k in re.findall(r"<\|VRT_(\d+)\|>", text)]
        return {"input_ids": torch.tensor([ids])}

    def apply_chat_template(self, *a, **k):
        return "templated"


DEMO = 'The "the car is on the left side of the horse" refers to <|VRT_122|><|VRT_107|><|VRT_138|><|VRT_256|><|VRT_135|> in this image.<|im_end|>'


def _tokenise_demo(V):
    """Split the golden completion of eval/outputs/demo/completion.txt into tokenizer-like pieces."""
    pieces = ['The', ' "', 'the', ' car', ' is', ' on', ' the', ' left', ' side', ' of', ' the', ' horse', '"', ' refers', ' to', ' ',
              '<|VRT_122|>', '<|VRT_107|>', '<|VRT_138|>', '<|VRT_256|>', '<|VRT_135|>', ' in', ' this', ' image', '.', '<|im_end|>']
    assert "".join(pieces) == DEMO
    strings, ids = {}, []
    nxt = 0
    for p in pieces:
        m = re.fullmatch(r"<\|VRT_(\d+)\|>", p)
        if m:
            ids.append(V + int(m.group(1)))
        else:
            strings[nxt] = p
            ids.append(nxt)
            nxt += 1
    return strings, torch.tensor([ids])


def test_parser_demo_kat():
    V, D = 1000, 8
    strings, ids = _tokenise_demo(V)
    proc = VisonTextProcessingClass(FakeProcessor(V, strings))
    proc.model_embed_token_size = V
    T = ids.shape[1]
    hidden = [(torch.full((1, 1, D), float(s)),) for s in range(T)]
    grid = torch.tensor([[1, 30, 46]])
    protos = torch.arange(345, dtype=torch.float32)[:, None].expand(-1, D)
    comp, feats, labels, vrts, vf = parseVRTintoCompletion(proc, ids, hidden, torch.Tensor([False]), protos, grid)
    assert comp == [DEMO]
    assert labels == [["the car is on the left side of the horse"]]
    assert vrts == [["<|VRT_122|><|VRT_107|><|VRT_138|><|VRT_256|><|VRT_135|>"]]
    assert feats[0][0].shape == (5, D) and feats[0][0][:, 0].tolist() == [16.0, 17.0, 18.0, 19.0, 20.0]   # steps 16-20 predicted the VRTs
    assert vf[0][0][:, 0].tolist() == [122.0, 107.0, 138.0, 256.0, 135.0]


def test_parser_answer_tags_and_exception_swallowing():
    V, D = 100, 4
    strings = {0: "<", 1: "answer", 2: ">", 3: ' "', 4: "cat", 5: '"', 6: "</", 7: "think", 8: "x"}
    proc = VisonTextProcessingClass(FakeProcessor(V, strings))
    proc.model_embed_token_size = V
    # thinking mode: objects only inside <answer> ... </answer>
    ids = torch.tensor([[8, V + 1, 0, 1, 2, 3, 4, 5, V + 2, V + 3, 6, 1, 2, 8]])
    hidden = [(torch.full((1, 1, D), float(s)),) for s in range(ids.shape[1])]
    comp, feats, labels, vrts, vf = parseVRTintoCompletion(proc, ids, hidden, torch.tensor([True]))
    assert labels == [["cat"]] and vrts == [["<|VRT_2|><|VRT_3|>"]] and feats[0][0][:, 0].tolist() == [8.0, 9.0] and vf == [[]]
    # a VRT run that reaches the end of the completion indexes past the list in the reference -> the sample is emptied (padt_processor.py:146-150)
    ids2 = torch.tensor([[3, 4, 5, V + 2, V + 3]])
    hidden2 = [(torch.zeros((1, 1, D)),) for _ in range(5)]
    comp, feats, labels, vrts, vf = parseVRTintoCompletion(proc, ids2, hidden2, torch.tensor([False]))
    assert feats == [[]] and labels == [[]] and vrts == [[]] and vf == [[]] and comp[0].endswith("<|VRT_3|>")


def test_id_shift_kat_and_inplace():
    """V=1000, grids (30,46),(24,24): local [[5,1003],[1001,7]] <-> global [[5,1003],[1346,7]]."""
    proc = VisonTextProcessingClass(FakeProcessor(1000, {}))
    grid = torch.tensor([[1, 30, 46], [1, 24, 24]])
    ids = torch.tensor([[5, 1003], [1001, 7]])
    out = proc.assign_to_global_vrt_id(ids, grid)
    assert out is ids and ids.tolist() == [[5, 1003], [1346, 7]]
    ref = O.assign_to_global_vrt_id(1000, 2, torch.tensor([[5, 1003], [1001, 7]]), grid)
    assert torch.equal(ref, ids)
    back = proc.assign_to_local_vrt_id(ids, grid)
    assert back is ids and ids.tolist() == [[5, 1003], [1001, 7]]


def test_processor_wrapper_growth_and_passthrough():
    inner = FakeProcessor(50, {})
    proc = VisonTextProcessingClass(inner, 2)
    assert proc.model_embed_token_size == 50
    proc.prepare(64)                               # pads the tokenizer to the embedding rows with <|empty_token_i|>
    assert len(inner.tokenizer.get_vocab()) == 64 and "<|empty_token_13|>" in inner.tokenizer.vocab
    proc.set_image_grid_thw(torch.tensor([[1, 8, 12], [1, 4, 4]]))
    assert "<|VRT_23|>" in inner.tokenizer.vocab and "<|VRT_24|>" not in inner.tokenizer.vocab
    assert proc.pid2vrt([3, 4]) == "<|VRT_3|><|VRT_4|>" and proc.pid2vrt(7) == "<|VRT_7|>"
    assert proc.apply_chat_template() == "templated"            # __getattr__ passthrough
    try:
        proc.nope
        assert False
    except AttributeError:
        pass


def test_parser_against_reference_cross_run():
    """tests/golden/parser_cases.json: 200 random piece streams that oracle/make_golden.py ran through the UNMODIFIED reference
    parser and processor wrapper (padt_processor.py:4-151); this module must reproduce every output (completions, labels, VRT strings,
    which hidden-state steps were gathered, which prototype rows were looked up, the global id shift)."""
    import json
    import os
    from helpers import ROOT
    with open(os.path.join(ROOT, "tests", "golden", "parser_cases.json")) as f:
        data = json.load(f)
    strings = {i: a for i, a in enumerate(data["alphabet"])}
    n_obj = 0
    for ci, c in enumerate(data["cases"]):
        V = c["V"]
        ids = torch.tensor(c["ids"])
        grid = torch.tensor(c["grids"])
        B, T = ids.shape
        Ms = [g[1] * g[2] // 4 for g in c["grids"]]
        proc = VisonTextProcessingClass(FakeProcessor(V, strings))
        proc.model_embed_token_size = V
        hidden = [(torch.arange(B, dtype=torch.float32).view(B, 1, 1) * 1000 + float(s),) for s in range(T)]
        protos = torch.arange(sum(Ms), dtype=torch.float32)[:, None]
        comp, feats, labels, vrts, vf = parseVRTintoCompletion(proc, ids.clone(), hidden, torch.tensor(c["thinking"]), protos, grid)
        e = c["expect"]
        assert comp == e["completions"] and labels == e["labels"] and vrts == e["vrts"], ci
        assert [[f.flatten().tolist() for f in fs] for fs in feats] == e["feat_steps"], ci
        assert [[f.flatten().tolist() for f in fs] for fs in vf] == e["proto_rows"], ci
        x = ids.clone()
        assert proc.assign_to_global_vrt_id(x, grid).tolist() == e["global_ids"], ci
        assert torch.equal(proc.assign_to_local_vrt_id(x, grid), ids), ci
        n_obj += sum(len(l) for l in labels)
    assert len(data["cases"]) >= 200 and n_obj > 100
