"""Batched inference harness: the reference's ``infer_dataset`` loop
(/root/reference/eval/evaluation_scripts/utils.py:176-266) on this backend.

Kept: the per-batch flow (processor call -> ``assign_to_global_vrt_id`` -> ``generate`` -> ``assign_to_local_vrt_id`` ->
``parseVRTintoCompletion`` -> ``vl_decode``), the rank striding over the dataset (utils.py:181-182), and both JSONL outputs byte for byte:
``*_pred_comp_*.json`` lines ``{"image_id", "completion"}`` (utils.py:249-251) and ``*_pred_results_*.json`` lines ``{"image_id", "score",
"category", "bbox": [x, y, w, h] px, "mask": {"size": [h, w], "counts": "<COCO RLE>"}}`` (utils.py:255-265), so ``eval_coco.py`` /
``eval_refcoco.py`` read them unchanged.
Dropped: the accelerate / DeepSpeed ZeRO-3 wrapper and the per-batch ``GatheredParameters`` (utils.py:162-169, 196) -- weights are
resident; rank and world size come from ``torch.distributed`` (or the arguments).
Moved to the GPU: the mask post-processing (``F.interpolate(..., 'bilinear').sigmoid() > 0.5`` + RLE, utils.py:260-264):
``padt_mask_to_runs`` leaves only the run starts to copy back, ``padt_rle_string`` builds pycocotools' string.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _capi
from .processor import parseVRTintoCompletion


def eval_box(box, img_w: int, img_h: int):
    """utils.py:257-258.  ``box`` = one row of ``pred_boxes`` (cx, cy, w, h in 0..1); Python float arithmetic and ``round`` as there."""
    cx, cy, bw, bh = (float(v) for v in box)
    b = (max(cx - bw / 2, 0), max(cy - bh / 2, 0), min(bw, 1), min(bh, 1))
    return (round(b[0] * img_w), round(b[1] * img_h), round(b[2] * img_w), round(b[3] * img_h))


class MaskRleEncoder:
    """``pred_mask`` rows -> {"size": [h, w], "counts": str} (utils.py:260-264), work on the GPU, one sync per batch of objects."""

    MAX_START_BYTES = 256 << 20   # run-start scratch: objects are processed in groups that fit

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self._ws = None
        self._dense = None
        self._starts = None

    def _ensure(self, n_pix: int, ws_bytes: int, n_obj: int):
        if self._ws is None or self._ws.numel() < ws_bytes:
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        if self._dense is None or self._dense.numel() < n_pix:
            self._dense = torch.empty(n_pix, dtype=torch.uint8, device=self.device)
        if self._starts is None or self._starts.shape[0] < n_obj or self._starts.shape[1] < n_pix:
            self._starts = torch.empty((n_obj, n_pix), dtype=torch.int32, device=self.device)

    def encode(self, pred_mask: torch.Tensor, valid_hw: Sequence[Sequence[int]], image_hw: Sequence[Sequence[int]], return_dense: bool = False):
        """pred_mask [N, 4Hmax, 4Wmax] bf16 (device); valid_hw[i] = (H_i, W_i) patch grid of object i; image_hw[i] = (h, w) of its image."""
        lib = _capi.load()
        n_obj = int(pred_mask.shape[0])
        if n_obj == 0:
            return ([], []) if return_dense else []
        if pred_mask.dtype != torch.bfloat16 or not pred_mask.is_cuda:
            raise TypeError("pred_mask must be a bf16 CUDA tensor (vl_decode's output)")
        pred_mask = pred_mask.contiguous()
        max_pix = max(int(h) * int(w) for h, w in image_hw)
        ws_bytes = max(int(lib.padt_mask_runs_workspace_bytes(int(h), int(w))) for h, w in image_hw)
        group = max(1, min(n_obj, self.MAX_START_BYTES // (4 * max_pix)))     # objects per host sync
        self._ensure(max_pix, ws_bytes, group)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        out, dense_out = [], []
        for g0 in range(0, n_obj, group):
            g1 = min(n_obj, g0 + group)
            n_runs = torch.zeros(g1 - g0, dtype=torch.int32, device=self.device)
            for i in range(g0, g1):
                (gh, gw), (h, w) = valid_hw[i], image_hw[i]
                dense = self._dense if not return_dense else torch.empty(int(h) * int(w), dtype=torch.uint8, device=self.device)
                _capi.check(lib.padt_mask_to_runs(pred_mask[i].data_ptr(), pred_mask.stride(1), int(gh) * 4, int(gw) * 4, int(h), int(w),
                                                  dense.data_ptr(), self._starts[i - g0].data_ptr(), n_runs[i - g0:].data_ptr(), self._ws.data_ptr(),
                                                  self._ws.numel(), stream), "padt_mask_to_runs")
                if return_dense:
                    dense_out.append(dense.view(int(w), int(h)).t())      # [h, w] view of the column-major mask
            counts_host = n_runs.cpu()                                  # one sync per group
            for i in range(g0, g1):
                h, w = int(image_hw[i][0]), int(image_hw[i][1])
                k = int(counts_host[i - g0])
                starts = np.ascontiguousarray(self._starts[i - g0, :k].cpu().numpy(), dtype=np.int32)
                buf = C.create_string_buffer(6 * (k + 1) + 8)
                n = lib.padt_rle_string(starts.ctypes.data_as(C.POINTER(C.c_int32)), k, h * w, buf, len(buf))
                if n < 0:
                    raise _capi.PadtError(lib.padt_last_error().decode("utf-8", "replace"))
                out.append({"size": [h, w], "counts": buf.value.decode("ascii")})
        return (out, dense_out) if return_dense else out


def result_records(decoded: dict, labels: List[List[str]], image_sizes_wh: Sequence[Sequence[int]], image_ids: Sequence, encoder: MaskRleEncoder):
    """utils.py:253-265: one record per decoded object, in ``vl_decode`` order."""
    n = int(decoded["pred_boxes"].shape[0])
    if n == 0:
        return []
    flat_labels = sum(labels, [])
    boxes = decoded["pred_boxes"].float().cpu()
    scores = decoded["pred_score"].sigmoid().cpu()                    # bf16 sigmoid, then .item(), as the reference
    hs, ws = decoded["pred_mask_valid_hw"]
    valid_hw = list(zip(hs.tolist(), ws.tolist()))
    sample_idx = list(decoded["sample_idx"])
    image_hw = [(int(image_sizes_wh[s][1]), int(image_sizes_wh[s][0])) for s in sample_idx]
    rles = encoder.encode(decoded["pred_mask"], valid_hw, image_hw)
    recs = []
    for i in range(n):
        s = sample_idx[i]
        w, h = int(image_sizes_wh[s][0]), int(image_sizes_wh[s][1])
        recs.append({"image_id": image_ids[s], "score": scores[i].item(), "category": flat_labels[i], "bbox": eval_box(boxes[i].tolist(), w, h),
                     "mask": rles[i]})
    return recs


def completion_records(completions: Sequence[str], image_ids: Sequence):
    """utils.py:249-251."""
    return [{"image_id": image_ids[i], "completion": c.replace("<|endoftext|>", "").replace("<|im_end|>", "")} for i, c in enumerate(completions)]


@torch.no_grad()
def infer_batch(model, processor, prompt_inputs: dict, image_sizes_wh: Sequence[Sequence[int]], image_ids: Sequence, max_new_tokens: int = 1024,
                encoder: Optional[MaskRleEncoder] = None, **generate_kwargs):
    """utils.py:229-247 + 249-265 for one batch.  ``prompt_inputs``: the processor's output (input_ids, attention_mask, pixel_values,
    image_grid_thw) already on the model's device.  Returns (completion records, result records)."""
    encoder = encoder or MaskRleEncoder(model.device)
    prompt_ids = prompt_inputs["input_ids"]
    batch_size = prompt_ids.size(0)
    prompt_inputs = dict(prompt_inputs)
    prompt_inputs["input_ids"] = processor.assign_to_global_vrt_id(prompt_inputs["input_ids"], prompt_inputs["image_grid_thw"])
    out = model.generate(**prompt_inputs, use_cache=True, max_new_tokens=max_new_tokens, do_sample=False, output_hidden_states=True,
                         return_dict_in_generate=True, synced_gpus=False, **generate_kwargs)
    prompt_length = prompt_ids.size(1)
    prompt_completion_ids = processor.assign_to_local_vrt_id(out["sequences"], prompt_inputs["image_grid_thw"])
    completion_ids = prompt_completion_ids[:, prompt_length:]
    completions, feats, labels, vrts, vrts_feats = parseVRTintoCompletion(processor, completion_ids, out["hidden_states"],
                                                                          torch.Tensor([False] * batch_size))
    decoded = model.vl_decode(feats, out.past_image_embeds, out.past_high_res_image_embeds, prompt_inputs["image_grid_thw"], out.past_visual_pe)
    return completion_records(completions, image_ids), result_records(decoded, labels, image_sizes_wh, image_ids, encoder)


def default_build_inputs(processor, device, open_image: Optional[Callable] = None) -> Callable:
    """The per-batch input construction of the reference loop (utils.py:199-229) as a ``build_inputs`` callback for ``infer_dataset``:
    chat template over ``batch['prompt']``, images opened from ``batch['image_path']`` (a list of paths per sample), images with a side
    below 28 px up-scaled with LANCZOS keeping the aspect ratio (utils.py:204-217), processor call with left padding, tensors to
    ``device``.  Returns (prompt_inputs, [(w, h) per image]).
    The reference resizes inside a bare ``try / except`` with ``new_w`` / ``new_h`` only defined for small images, so a large image is
    left alone through a NameError -- unless an earlier small image left stale values behind; here a large image is always left alone."""
    def build(batch):
        import PIL.Image
        opener = open_image or PIL.Image.open
        prompts_text = processor.apply_chat_template(batch["prompt"], tokenize=False, add_generation_prompt=True)
        images = []
        for sample in batch["image_path"]:
            for path in sample:
                img = opener(path)
                w, h = img.size
                if w < 28 or h < 28:
                    if w < h:
                        new_w, new_h = 28, int(h * (28 / w))
                    else:
                        new_h, new_w = 28, int(w * (28 / h))
                    img = img.resize((new_w, new_h), PIL.Image.Resampling.LANCZOS)
                images.append(img)
        inputs = processor(text=prompts_text, images=images, return_tensors="pt", padding=True, padding_side="left", add_special_tokens=False)
        inputs = {k: (v.to(device) if hasattr(v, "to") else v) for k, v in dict(inputs).items()}
        return inputs, [im.size for im in images]
    return build


def shard_indices(n_items: int, batch_size: int, rank: int, world_size: int):
    """utils.py:181-182: the first index of every batch this rank takes (the total is padded up to a multiple of world_size * batch)."""
    all_number = math.ceil(n_items / (world_size * batch_size)) * world_size * batch_size
    return range(rank * batch_size, all_number, world_size * batch_size)


def infer_dataset(model, dataset, processor, build_inputs: Callable, output_dir: str, batch_size: int = 1, datasetname: str = "coco", suffix: str = "",
                  rank: Optional[int] = None, world_size: Optional[int] = None, max_new_tokens: int = 1024):
    """utils.py:176-266.  ``dataset[idx: idx + batch]`` must give a dict with 'id' (list) as in the reference; ``build_inputs(batch)`` returns
    (prompt_inputs on the device, [(w, h) per image]) -- the chat template / image loading / processor call of utils.py:199-229;
    ``default_build_inputs(processor, device)`` is that code for a real checkpoint's processor.  Writes the two JSONL files of this rank."""
    if rank is None or world_size is None:
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    os.makedirs(output_dir, exist_ok=True)
    f_res = os.path.join(output_dir, f"{datasetname}_{rank}_pred_results_{suffix}.json")
    f_comp = os.path.join(output_dir, f"{datasetname}_{rank}_pred_comp_{suffix}.json")
    open(f_res, "w+").close()
    open(f_comp, "w+").close()
    encoder = MaskRleEncoder(model.device)
    for idx in shard_indices(len(dataset), batch_size, rank, world_size):
        if idx >= len(dataset):
            continue
        batch = dataset[idx: idx + batch_size]
        prompt_inputs, sizes_wh = build_inputs(batch)
        comps, recs = infer_batch(model, processor, prompt_inputs, sizes_wh, batch["id"], max_new_tokens, encoder)
        with open(f_comp, "a+") as f:
            for r in comps:
                f.write(json.dumps(r) + "\n")
        with open(f_res, "a+") as f:
            for r in recs:
                f.write(json.dumps(r) + "\n")
    return f_res, f_comp
