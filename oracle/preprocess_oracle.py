"""TEST INFRASTRUCTURE ONLY (checker for SURVEY.md §8f-3; nothing under padt_b200/ may import this).

CPU restatement, in numpy integer / IEEE arithmetic, of the image pre-processing the reference runs on the host for every request:
``processing_class(text=..., images=...)`` in eval/test_demo.py:87-93 and eval/evaluation_scripts/utils.py:221-229 is transformers'
Qwen2-VL image processor (transformers==4.50.0 pinned by the reference's setup.py:19; third-party, not vendored in /root/reference):

  smart_resize  ->  PIL ``Image.resize(..., BICUBIC)`` on uint8 RGB  ->  * 1/255 (float64, cast to float32)  ->  (x - mean) / std in float32
  ->  patchify to [grid_h * grid_w, 3 * 2 * 14 * 14] with rows in (h/2, w/2, 2, 2) order and the single frame repeated twice in time.

Published algorithms restated here:
  * smart_resize: transformers/models/qwen2_vl/image_processing_qwen2_vl.py ``smart_resize`` (4.50.0: lines 62-88 as cited by SURVEY.md §8f-3);
  * PIL bicubic: Pillow src/libImaging/Resample.c — ``precompute_coeffs`` (double), ``normalize_coeffs_8bpc`` (22-bit fixed point),
    ``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc`` (int32 accumulate from 1 << 21, arithmetic shift, clip to 0..255), horizontal
    pass first, then the vertical pass over the uint8 intermediate;
  * rescale / normalize: transformers/image_transforms.py ``rescale`` (float64 product, float32 cast) and ``normalize`` (float32 sub, div).

Parity is PINNED: tests/test_preprocess.py checks every function here against Pillow and against transformers' own PIL-backend
processor (``Qwen2VLImageProcessorPil`` in transformers 5.x = the 4.50 slow processor's arithmetic) on random and structured images.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280):
    if max(height, width) / min(height, width) > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}")
    h_bar = round(height / factor) * factor          # Python round: half to even
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size): returns (ksize, bounds [out,2], coeffs int32 [out,ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """8bpc resample along axis 0 of a [n, ...] uint8 array."""
    _, bounds, kk = precompute_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * int(kk[xx, x])
        # int32 wrap-around never happens: |sum| <= 255 * sum|k| << 2^31; arithmetic shift, then clip8
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_bicubic_resize(img_hwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Image.resize((out_w, out_h), BICUBIC) of a uint8 [H, W, 3] image: horizontal pass, then vertical pass (ImagingResample)."""
    assert img_hwc.dtype == np.uint8 and img_hwc.ndim == 3
    h, w, _ = img_hwc.shape
    if (h, w) == (out_h, out_w):
        return img_hwc.copy()
    tmp = img_hwc
    if w != out_w:
        tmp = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(tmp.transpose(1, 0, 2)), out_w).transpose(1, 0, 2))
    if h != out_h:
        tmp = _resample_axis0(tmp, out_h)
    return tmp


def rescale_normalize(img_hwc_u8: np.ndarray, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD, rescale_factor: float = 1 / 255) -> np.ndarray:
    x = (img_hwc_u8.astype(np.float64) * rescale_factor).astype(np.float32)
    return (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)


def patchify(img_hwc_f32: np.ndarray, patch: int = 14, temporal: int = 2, merge: int = 2):
    """[H, W, 3] float32 -> ([grid_h * grid_w, 3 * temporal * patch * patch], (1, grid_h, grid_w)); the frame is repeated `temporal` times."""
    h, w, c = img_hwc_f32.shape
    gh, gw = h // patch, w // patch
    chw = np.ascontiguousarray(img_hwc_f32.transpose(2, 0, 1))
    frames = np.stack([chw] * temporal, axis=0)[None]                      # [1, T, C, H, W]
    p = frames.reshape(1, 1, temporal, c, gh // merge, merge, patch, gw // merge, merge, patch)
    p = p.transpose(0, 1, 4, 7, 5, 8, 3, 2, 6, 9)
    return p.reshape(gh * gw, c * temporal * patch * patch), (1, gh, gw)


def preprocess(img_hwc_u8: np.ndarray, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD):
    h, w, _ = img_hwc_u8.shape
    oh, ow = smart_resize(h, w, 28, min_pixels, max_pixels)
    resized = pil_bicubic_resize(img_hwc_u8, oh, ow)
    return patchify(rescale_normalize(resized, mean, std))
