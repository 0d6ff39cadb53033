"""Batched inference harness (eval/evaluation_scripts/utils.py:176-266): the oracle's box / mask / RLE arithmetic against
hand-derived known answers and torch (CPU), the CUDA mask post-processing against the reference line executed on the same GPU and the
JSONL records of a whole batch against the reference loop restated with torch ops (GPU)."""
import ctypes as C
import json
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import eval_oracle as EO  # noqa: E402


# ------------------------------------------------------------------------------------------------ CPU
def test_rle_known_answers_and_round_trip():
    # hand-derived from maskApi.c rleToString: one 5-bit group per char (+48), 0x20 = more groups, deltas to the count two back from i = 3
    assert EO.rle_to_string([9]) == "9"                     # 3x3 zeros
    assert EO.rle_to_string([0, 9]) == "09"                 # 3x3 ones: empty zero run first
    assert EO.rle_to_string([32]) == "P1"                   # 32 = 0b1_00000: low group 0 | 0x20 -> 'P', then 1 -> '1'
    assert EO.rle_to_string([5, 1, 5, 1, 2]) == "5150M"     # i = 3: 1 - 1 = 0 -> '0'; i = 4: 2 - 5 = -3 -> 0b11101, sign group closes -> 'M'
    assert EO.rle_counts(np.zeros((3, 3), np.uint8)) == [9] and EO.rle_counts(np.ones((3, 3), np.uint8)) == [0, 9]
    m = np.array([[0, 1], [1, 1], [0, 0]], np.uint8)        # column-major flattening: 0 1 0 | 1 1 0
    assert EO.rle_counts(m) == [1, 1, 1, 2, 1]
    rng = np.random.default_rng(0)
    for t in range(200):
        h, w = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        m = (rng.random((h, w)) < rng.choice([0.0, 0.03, 0.5, 0.97, 1.0])).astype(np.uint8)
        c = EO.rle_counts(m)
        s = EO.rle_to_string(c)
        assert sum(c) == h * w and EO.rle_from_string(s) == c and np.array_equal(EO.rle_decode(c, h, w), m)
        assert all(48 <= ord(ch) < 48 + 64 for ch in s)


def test_rle_string_c_abi_equals_oracle():
    from padt_b200 import _capi
    lib = _capi.load()
    rng = np.random.default_rng(1)
    for t in range(200):
        h, w = int(rng.integers(1, 50)), int(rng.integers(1, 50))
        m = (rng.random((h, w)) < rng.choice([0.0, 0.02, 0.5, 1.0])).astype(np.uint8)
        if t % 3 == 0:
            m[:] = 0
            m[h // 4: h // 2 + 1, w // 3: w // 2 + 1] = 1
        want = EO.rle_to_string(EO.rle_counts(m))
        flat = m.flatten(order="F")
        starts = np.nonzero(flat != np.concatenate([[0], flat[:-1]]))[0].astype(np.int32)
        buf = C.create_string_buffer(6 * (len(starts) + 1) + 8)
        n = lib.padt_rle_string(starts.ctypes.data_as(C.POINTER(C.c_int32)), len(starts), h * w, buf, len(buf))
        assert n == len(want) and buf.value.decode() == want
    small = C.create_string_buffer(2)
    starts = np.array([1, 2, 3, 4], np.int32)
    assert lib.padt_rle_string(starts.ctypes.data_as(C.POINTER(C.c_int32)), 4, 10, small, 2) == -1     # loud, no overflow


def test_eval_box_equals_reference_expression():
    from padt_b200.eval_harness import eval_box
    torch.manual_seed(0)
    boxes = torch.rand(500, 4).to(torch.bfloat16)
    boxes[:20, 2:] = torch.tensor([1.5, 1.25], dtype=torch.bfloat16)          # w, h > 1 clip
    for i, box in enumerate(boxes):
        w, h = 640 - i, 333 + i
        ref = (max(box[0].item() - box[2].item() / 2, 0), max(box[1].item() - box[3].item() / 2, 0), min(box[2].item(), 1), min(box[3].item(), 1))
        ref = (round(ref[0] * w), round(ref[1] * h), round(ref[2] * w), round(ref[3] * h))                 # utils.py:257-258 verbatim
        assert eval_box(box.tolist(), w, h) == ref == EO.eval_box([v.item() for v in box], w, h)


def test_oracle_mask_against_torch_cpu_float_path():
    """The reference line runs ATen's CUDA kernel (pinned on the GPU below).  On the CPU the closest executable statement is the same
    interpolate on float32 inputs, rounded to bf16: identical masks up to fp32 contraction noise at the threshold."""
    torch.manual_seed(0)
    tot = bad = 0
    for (H, W, h, w, sig) in [(6, 8, 100, 133, 3.0), (24, 24, 480, 640, 1.0), (30, 46, 420, 644, 0.05), (10, 12, 640, 480, 0.02)]:
        m = (torch.randn(4 * H + 8, 4 * W + 4) * sig).to(torch.bfloat16)
        up = torch.nn.functional.interpolate(m.float()[None, None, :4 * H, :4 * W], size=(h, w), mode="bilinear")[0, 0].to(torch.bfloat16)
        ref = (up.sigmoid() > 0.5).numpy().astype(np.uint8)
        got = EO.mask_to_binary(m.float().numpy(), 4 * H, 4 * W, h, w)
        tot += ref.size
        bad += int((ref != got).sum())
    assert bad <= max(1, tot // 100000), (bad, tot)


def test_bf16_sigmoid_threshold_identity():
    """bf16(sigmoid(x)) > 0.5  <=>  x > 2^-7 for every bf16 x (the CUDA kernel thresholds the logit instead of evaluating the sigmoid)."""
    bits = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16)
    x = bits.view(torch.bfloat16)
    x = x[torch.isfinite(x.float())]
    assert torch.equal(x.sigmoid() > 0.5, x.float() > 2.0 ** -7)


def test_shard_indices_and_completion_records():
    from padt_b200.eval_harness import completion_records, shard_indices
    for n, bs, ws in [(10, 1, 4), (500, 8, 8), (7, 2, 2), (1, 4, 8)]:
        seen = []
        for r in range(ws):
            all_number = math.ceil(n / (ws * bs)) * ws * bs                              # utils.py:181-182
            assert list(shard_indices(n, bs, r, ws)) == list(range(r * bs, all_number, ws * bs))
            seen += [i for s in shard_indices(n, bs, r, ws) for i in range(s, min(s + bs, n))]
        assert sorted(seen) == list(range(n))
    recs = completion_records(["a cat<|im_end|>", "x<|endoftext|><|im_end|>"], [7, "id9"])
    assert recs == [{"image_id": 7, "completion": "a cat"}, {"image_id": "id9", "completion": "x"}]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("H,W,h,w,sig", [(6, 8, 100, 133, 3.0), (24, 24, 480, 640, 1.0), (30, 46, 420, 644, 0.05), (96, 96, 1344, 1344, 2.0),
                                         (96, 96, 375, 500, 0.02), (4, 6, 37, 53, 0.01), (5, 5, 20, 20, 1.0)])
def test_gpu_mask_postprocess_against_reference_line(H, W, h, w, sig):
    from helpers import report
    from padt_b200.eval_harness import MaskRleEncoder
    torch.manual_seed(H * 100 + W)
    n_obj = 3
    pm = (torch.randn(n_obj, 4 * H + 8, 4 * W + 4, device="cuda") * sig).to(torch.bfloat16)
    pm[1, : 2 * H] += 1.0                                                                     # a large connected region as well as noise
    enc = MaskRleEncoder("cuda")
    rles, dense = enc.encode(pm, [(H, W)] * n_obj, [(h, w)] * n_obj, return_dense=True)
    bad_ref = bad_oracle = 0
    for i in range(n_obj):
        # utils.py:260 verbatim on the same device
        ref = torch.nn.functional.interpolate(pm[i][None, None, :H * 4, :W * 4], size=(h, w), mode="bilinear")[0, 0].sigmoid() > 0.5
        ref = ref.cpu().numpy().astype(np.uint8)
        got = dense[i].cpu().numpy()
        bad_ref += int((got != ref).sum())
        bad_oracle += int((got != EO.mask_to_binary(pm[i].float().cpu().numpy(), 4 * H, 4 * W, h, w)).sum())
        # RLE of OUR mask through the device compaction + C string == the oracle's encoder on the same mask, and it decodes back
        assert rles[i] == {"size": [h, w], "counts": EO.rle_to_string(EO.rle_counts(got))}
        assert np.array_equal(EO.rle_decode(EO.rle_from_string(rles[i]["counts"]), h, w), got)
    report("mask post-processing", shape=[H, W, h, w], pixels=n_obj * h * w, mismatch_vs_torch_cuda=bad_ref, mismatch_vs_oracle=bad_oracle)
    assert bad_ref == 0, "thresholded masks must equal the reference line's on the same GPU"
    assert bad_oracle <= max(1, n_obj * h * w // 100000)


@pytest.mark.gpu
def test_gpu_infer_batch_records_equal_reference_loop():
    """generate -> parse -> vl_decode -> JSON records for a 2-image batch; the records are compared with the reference's per-object loop
    (utils.py:253-265) written out with torch ops on the same outputs, RLE by the oracle."""
    from helpers import load_golden
    from test_processor import FakeProcessor
    from padt_b200 import VisonTextProcessingClass
    from padt_b200.config import padt_tiny
    from padt_b200.eval_harness import infer_batch
    from padt_b200.modeling import PaDTForConditionalGeneration
    from padt_b200.weights import random_weights
    fx = load_golden("tiny_b2")
    cfg = padt_tiny(tie=fx["tie"], vocab=fx["vocab"])
    model = PaDTForConditionalGeneration(cfg, {k: v.cuda() for k, v in random_weights(cfg, seed=fx["wseed"]).items()})
    V = cfg.text.vocab_size
    strings = {10: "The", 11: ' "', 12: "cat", 13: '"', 14: " refers", 15: " to", 16: " ", 17: " and", 18: "dog", cfg.eos_token_ids[0]: "<|im_end|>", cfg.pad_token_id: "<|endoftext|>"}
    proc = VisonTextProcessingClass(FakeProcessor(V, strings))
    proc.model_embed_token_size = V
    grid = fx["grid"]
    m = (grid[:, 1] * grid[:, 2] // 4).tolist()
    # row 0: one object (3 VRTs); row 1: two objects; ids are LOCAL here and shifted to global below exactly as the harness does
    rows = [[10, 11, 12, 13, 14, 15, 16, V + 1, V + 3, V + 2, cfg.eos_token_ids[0]] + [cfg.pad_token_id] * 7,
            [10, 11, 12, 13, 14, 15, 16, V + 0, V + 2, 17, 11, 18, 13, 16, V + 4, V + 1, V + 3, cfg.eos_token_ids[0]]]
    forced = proc.assign_to_global_vrt_id(torch.tensor(rows, dtype=torch.int64), grid)
    assert max(m) > 4
    inputs = dict(input_ids=fx["input_ids"].clone().cuda(), attention_mask=fx["attention_mask"].cuda(), pixel_values=fx["pixel_values"].cuda(),
                  image_grid_thw=grid.cuda())
    sizes_wh = [(133, 100), (60, 47)]
    ids = [101, "img_b"]
    comps, recs = infer_batch(model, proc, inputs, sizes_wh, ids, max_new_tokens=len(rows[0]), forced_tokens=forced.cuda(), eos_token_id=[])
    assert [c["image_id"] for c in comps] == ids and comps[0]["completion"].startswith('The "cat" refers to <|VRT_1|>')
    assert len(recs) == 3 and [r["image_id"] for r in recs] == [101, "img_b", "img_b"] and [r["category"] for r in recs] == ["cat", "cat", "dog"]
    # the reference loop on the same model outputs
    out = model.generate(**{**inputs, "input_ids": proc.assign_to_global_vrt_id(inputs["input_ids"].clone(), inputs["image_grid_thw"])}, use_cache=True,
                         max_new_tokens=len(rows[0]), do_sample=False, output_hidden_states=True, return_dict_in_generate=True, forced_tokens=forced.cuda(),
                         eos_token_id=[])
    from padt_b200 import parseVRTintoCompletion
    L = inputs["input_ids"].shape[1]
    comp_ids = proc.assign_to_local_vrt_id(out["sequences"], inputs["image_grid_thw"])[:, L:]
    completions, feats, labels, _, _ = parseVRTintoCompletion(proc, comp_ids, out["hidden_states"], torch.Tensor([False, False]))
    dec = model.vl_decode(feats, out.past_image_embeds, out.past_high_res_image_embeds, inputs["image_grid_thw"], out.past_visual_pe)
    want = []
    for box, score, label, mask, mask_hw, sample_idx in zip(dec["pred_boxes"], dec["pred_score"].sigmoid(), sum(labels, []), dec["pred_mask"],
                                                            torch.stack([dec["pred_mask_valid_hw"][0], dec["pred_mask_valid_hw"][1]], dim=-1), dec["sample_idx"]):
        eval_box = (max(box[0].item() - box[2].item() / 2, 0), max(box[1].item() - box[3].item() / 2, 0), min(box[2].item(), 1), min(box[3].item(), 1))
        w, h = sizes_wh[sample_idx]
        eval_box = (round(eval_box[0] * w), round(eval_box[1] * h), round(eval_box[2] * w), round(eval_box[3] * h))
        mk = torch.nn.functional.interpolate(mask[None, None, :mask_hw[0] * 4, :mask_hw[1] * 4], size=(h, w), mode="bilinear")[0, 0].sigmoid() > 0.5
        mk = mk.cpu().numpy().astype(np.uint8)
        rle = {"size": [h, w], "counts": EO.rle_to_string(EO.rle_counts(mk))}
        want.append(json.dumps({"image_id": ids[sample_idx], "score": score.item(), "category": label, "bbox": eval_box, "mask": rle}))
    assert [json.dumps(r) for r in recs] == want


# ------------------------------------------------------------------------------------------------ CPU: the dataset loop with a stub model
class _StubOut(dict):
    def __getattr__(self, k):
        return self[k]


class _StubModel:
    """Stands in for PaDTForConditionalGeneration where no GPU is needed: generate returns the prompt + two text tokens, vl_decode the
    reference's empty result (padt.py:405-412)."""
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def generate(self, input_ids=None, max_new_tokens=None, **kw):
        self.calls.append(input_ids.clone())
        seq = torch.cat([input_ids, torch.tensor([[3, 4]] * input_ids.shape[0])], dim=1)
        hidden = tuple([(torch.zeros(input_ids.shape[0], input_ids.shape[1], 8),)] + [(torch.zeros(input_ids.shape[0], 1, 8),)])
        return _StubOut(sequences=seq, hidden_states=hidden, past_image_embeds=None, past_high_res_image_embeds=None, past_visual_pe=None)

    def vl_decode(self, feats, *a):
        return {"pred_boxes": torch.zeros(0, 4), "pred_score": torch.zeros(0, 1), "pred_mask": torch.zeros(0, 8, 8), "pred_mask_valid_hw": (), "sample_idx": []}


class _StubDataset:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        idx = list(range(self.n))[sl]
        return {"id": [100 + i for i in idx], "n": len(idx)}


def test_infer_dataset_files_and_rank_striding(tmp_path):
    """utils.py:176-266 outer loop: rank r takes batches r, r + world, ...; one completion line per sample in dataset order within the rank;
    both files are created (and truncated) even when a rank has nothing to write; the result file stays empty without objects."""
    from test_processor import FakeProcessor
    from padt_b200 import VisonTextProcessingClass
    from padt_b200.eval_harness import infer_dataset
    V = 50
    strings = {i: f"w{i} " for i in range(V)}
    strings[4] = "<|im_end|>"
    proc = VisonTextProcessingClass(FakeProcessor(V, strings))
    proc.model_embed_token_size = V
    ds = _StubDataset(7)

    def build_inputs(batch):
        n = batch["n"]
        return ({"input_ids": torch.tensor([[1, 2]] * n), "attention_mask": torch.ones(n, 2, dtype=torch.long),
                 "pixel_values": torch.zeros(4 * n, 8), "image_grid_thw": torch.tensor([[1, 2, 2]] * n)}, [(64, 48)] * n)

    seen = []
    for rank in range(2):
        model = _StubModel()
        f_res, f_comp = infer_dataset(model, ds, proc, build_inputs, str(tmp_path), batch_size=2, datasetname="toy", suffix="s", rank=rank, world_size=2)
        assert os.path.basename(f_res) == f"toy_{rank}_pred_results_s.json" and os.path.basename(f_comp) == f"toy_{rank}_pred_comp_s.json"
        lines = [json.loads(l) for l in open(f_comp)]
        seen += [l["image_id"] for l in lines]
        assert all(l["completion"] == "w3 " for l in lines)                 # <|im_end|> stripped (utils.py:251)
        assert open(f_res).read() == ""
        assert len(model.calls) == (2 if rank == 0 else 2)
    assert sorted(seen) == [100 + i for i in range(7)]                       # every sample exactly once across the ranks


def test_default_build_inputs_follows_reference_batch_construction(tmp_path):
    """utils.py:199-229: chat template, image loading, LANCZOS up-scaling of images with a side below 28 px (aspect ratio kept), processor
    call arguments, tensors moved to the device, PIL sizes (w, h) returned for the box / mask scaling."""
    Image = pytest.importorskip("PIL.Image")
    from padt_b200.eval_harness import default_build_inputs
    paths = []
    for i, (w, h) in enumerate([(64, 48), (20, 50), (100, 10)]):
        p = str(tmp_path / f"im{i}.png")
        Image.fromarray(np.full((h, w, 3), 40 * i, np.uint8)).save(p)
        paths.append(p)
    calls = {}

    class Proc:
        def apply_chat_template(self, prompts, tokenize, add_generation_prompt):
            calls["template"] = (prompts, tokenize, add_generation_prompt)
            return ["T:" + str(p) for p in prompts]

        def __call__(self, **kw):
            calls["proc"] = kw
            n = len(kw["images"])
            return {"input_ids": torch.ones(n, 3, dtype=torch.long), "attention_mask": torch.ones(n, 3, dtype=torch.long),
                    "pixel_values": torch.zeros(4 * n, 8), "image_grid_thw": torch.tensor([[1, 2, 2]] * n)}

    batch = {"prompt": ["p0", "p1", "p2"], "image_path": [[paths[0]], [paths[1]], [paths[2]]], "id": [1, 2, 3]}
    inputs, sizes = default_build_inputs(Proc(), torch.device("cpu"))(batch)
    assert calls["template"] == (["p0", "p1", "p2"], False, True)
    kw = calls["proc"]
    assert kw["text"] == ["T:p0", "T:p1", "T:p2"] and kw["return_tensors"] == "pt" and kw["padding"] is True and kw["padding_side"] == "left"
    assert kw["add_special_tokens"] is False
    assert sizes == [(64, 48), (28, int(50 * (28 / 20))), (int(100 * (28 / 10)), 28)] == [im.size for im in kw["images"]]
    assert set(inputs) == {"input_ids", "attention_mask", "pixel_values", "image_grid_thw"} and inputs["input_ids"].device.type == "cpu"
