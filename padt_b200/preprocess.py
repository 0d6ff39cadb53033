"""Image pre-processing on the GPU.

Drop-in for the image half of the reference's processor call (eval/test_demo.py:87-93, eval/evaluation_scripts/utils.py:221-229:
``processing_class(text=..., images=...)`` -> transformers 4.50 ``Qwen2VLImageProcessor``): same ``pixel_values`` / ``image_grid_thw``
keys, same values (Pillow's bicubic resampler bit for bit, then the processor's float32 rescale / normalise), computed by
``padt_preprocess_image`` in one fused pass per image.  The tokenizer half of the processor is untouched.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _capi

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280):
    """image_processing_qwen2_vl.py:62-88 through the C ABI (``padt_smart_resize``); raises ValueError like the reference."""
    out = (C.c_int32 * 2)()
    lib = _capi.load()
    if lib.padt_smart_resize(int(height), int(width), int(factor), int(min_pixels), int(max_pixels), out) != 0:
        raise ValueError(lib.padt_last_error().decode("utf-8", "replace"))
    return int(out[0]), int(out[1])


def _to_hwc_uint8(img) -> torch.Tensor:
    """PIL image / numpy [H,W,3] / torch [H,W,3] uint8 -> contiguous torch uint8 [H,W,3] (RGB, as ``do_convert_rgb`` leaves it)."""
    if hasattr(img, "convert") and hasattr(img, "size"):          # PIL.Image without importing PIL here
        img = np.array(img.convert("RGB"))                        # a writable copy (torch.from_numpy warns on PIL's read-only view)
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(np.ascontiguousarray(img))
    if not isinstance(img, torch.Tensor) or img.dtype != torch.uint8 or img.dim() != 3 or img.shape[-1] != 3:
        raise TypeError("images must be PIL images or uint8 [H, W, 3] arrays / tensors")
    return img.contiguous()


class GpuImageProcessor:
    """``processor.image_processor`` stand-in: ``__call__(images) -> {"pixel_values", "image_grid_thw"}`` with the work on ``device``.

    ``dtype=torch.bfloat16`` (default) emits what the ViT consumes; ``torch.float32`` emits exactly what the reference processor returns.
    """

    def __init__(self, device="cuda", min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280, patch_size: int = 14,
                 temporal_patch_size: int = 2, merge_size: int = 2, image_mean: Sequence[float] = OPENAI_CLIP_MEAN,
                 image_std: Sequence[float] = OPENAI_CLIP_STD, rescale_factor: float = 1 / 255, dtype: torch.dtype = torch.bfloat16):
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("dtype must be torch.bfloat16 or torch.float32")
        self.device = torch.device(device)
        self.min_pixels, self.max_pixels = int(min_pixels), int(max_pixels)
        self.patch_size, self.temporal_patch_size, self.merge_size = patch_size, temporal_patch_size, merge_size
        self.image_mean, self.image_std, self.rescale_factor = tuple(image_mean), tuple(image_std), float(rescale_factor)
        self.dtype = dtype
        self._ws = None

    def output_size(self, height: int, width: int):
        return smart_resize(height, width, self.patch_size * self.merge_size, self.min_pixels, self.max_pixels)

    def __call__(self, images: Iterable, return_tensors: str = "pt"):
        lib = _capi.load()
        imgs = [_to_hwc_uint8(im) for im in (images if isinstance(images, (list, tuple)) else [images])]
        sizes = [self.output_size(im.shape[0], im.shape[1]) for im in imgs]
        grids = [(1, oh // self.patch_size, ow // self.patch_size) for oh, ow in sizes]
        row_len = 3 * self.temporal_patch_size * self.patch_size * self.patch_size
        n_rows = [g[1] * g[2] for g in grids]
        pv = torch.empty((sum(n_rows), row_len), dtype=self.dtype, device=self.device)
        mean = (C.c_float * 3)(*self.image_mean)
        std = (C.c_float * 3)(*self.image_std)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        off = 0
        with torch.cuda.device(self.device):
            for im, (oh, ow), n in zip(imgs, sizes, n_rows):
                h, w = int(im.shape[0]), int(im.shape[1])
                dev = im.to(self.device, non_blocking=True) if im.device != self.device else im
                need = int(lib.padt_preprocess_workspace_bytes(h, w, oh, ow))
                if self._ws is None or self._ws.numel() < need:
                    self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                _capi.check(lib.padt_preprocess_image(dev.data_ptr(), h, w, oh, ow, mean, std, self.rescale_factor, self.patch_size,
                                                      self.temporal_patch_size, self.merge_size, pv[off:].data_ptr(), int(self.dtype == torch.float32),
                                                      self._ws.data_ptr(), self._ws.numel(), stream), "padt_preprocess_image")
                dev.record_stream(torch.cuda.current_stream(self.device))
                off += n
        return {"pixel_values": pv, "image_grid_thw": torch.tensor(grids, dtype=torch.int64)}
