"""TEST INFRASTRUCTURE ONLY (checker for SURVEY.md §8f-2; nothing under padt_b200/ may import this).

CPU restatement of the result post-processing of the reference's batched inference harness,
/root/reference/eval/evaluation_scripts/utils.py:249-266 (the per-object loop after ``vl_decode``):

  * utils.py:257 box: (cx, cy, w, h) in 0..1 -> clipped (x, y, w, h), scaled by the PIL image size, Python ``round`` (half to even);
  * utils.py:260 mask: ``F.interpolate(mask[None, None, :4H, :4W], size=(h, w), mode='bilinear')[0, 0].sigmoid() > 0.5`` on the bf16
    mask logits (restated below in numpy float32 following ATen's upsample_bilinear2d: scale = in / out in float, source index
    max(scale * (dst + 0.5) - 0.5, 0), two lerps, result rounded to bf16; bf16 sigmoid > 0.5 is exactly "bf16 value > 2^-7":
    sigmoid(2^-7) = 0.50195311 rounds to 0.5, the next bf16 above rounds to 0.50390625);
  * utils.py:262-264 ``pycocotools.mask.encode(np.asfortranarray(mask))`` + ``counts.decode()``: COCO compressed RLE.  pycocotools is a
    third-party dependency that is NOT in /root/reference nor installed here (version unpinned by the reference); the published algorithm
    (cocoapi common/maskApi.c ``rleEncode`` / ``rleToString`` / ``rleFrString``) is restated: column-major runs starting with the zero
    run; every count (difference to the count two back from the fourth on) in 5-bit groups, low group first, bit 0x20 = more groups
    follow, + 48 to land in ASCII.
  * utils.py:265 one JSON line per object: {"image_id", "score", "category", "bbox", "mask": {"size": [h, w], "counts": str}}.

Pinning: ``mask_to_binary`` is checked on the GPU box against the reference line itself, ``F.interpolate(...).sigmoid() > 0.5`` on bf16 CUDA
tensors (tests/test_eval_harness.py::test_gpu_mask_postprocess_against_reference_line: 0 mismatching pixels of 7.8 M, both for the CUDA
path and for this restatement), and on the CPU against the same interpolate on float32 inputs (torch's CPU bf16 kernel rounds its
intermediates to bf16 and is NOT what the reference runs).  ``eval_box`` is checked against the reference expression verbatim.  The RLE
codec has no reference implementation to run here: it is pinned by hand-derived known answers and the encode -> decode round trip only
("parity unpinned" against pycocotools' binary output).
"""
import json

import numpy as np


# ---------------------------------------------------------------------------------------------------------------- utils.py:257-258
def eval_box(box4, img_w: int, img_h: int):
    """box4: four Python floats (cx, cy, w, h) as ``tensor.item()`` returns them."""
    cx, cy, bw, bh = box4
    b = (max(cx - bw / 2, 0), max(cy - bh / 2, 0), min(bw, 1), min(bh, 1))
    return (round(b[0] * img_w), round(b[1] * img_h), round(b[2] * img_w), round(b[3] * img_h))


# ---------------------------------------------------------------------------------------------------------------- utils.py:260
def _bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bf16, returned as float32."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def _fma(a, b, c):
    """float32 fused multiply-add (one rounding): the product of two float32 is exact in float64."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def bilinear_bf16(logits_bf16_as_f32: np.ndarray, out_h: int, out_w: int, fma: bool = True) -> np.ndarray:
    """upsample_bilinear2d(align_corners=False) of a [in_h, in_w] bf16 map (given as float32 values), fp32 math, bf16 result
    (ATen/native/cuda/UpSampleBilinear2d.cu: the CUDA kernel the reference line runs).  ``fma=True`` applies the contraction nvcc
    makes of that expression (a * b + c * d -> fma(a, b, c * d)); ``fma=False`` rounds every product."""
    src = logits_bf16_as_f32.astype(np.float32)
    in_h, in_w = src.shape
    one, half = np.float32(1.0), np.float32(0.5)

    def axis(in_n, out_n):
        scale = np.float32(np.float32(in_n) / np.float32(out_n))
        d = (np.arange(out_n, dtype=np.float32) + half).astype(np.float32)
        r = _fma(np.full_like(d, scale), d, np.full_like(d, -half)) if fma else (scale * d - half).astype(np.float32)
        r = np.maximum(r, np.float32(0.0)).astype(np.float32)
        i0 = r.astype(np.int32)
        ip = (i0 < in_n - 1).astype(np.int32)
        l1 = (r - i0.astype(np.float32)).astype(np.float32)
        return i0, ip, l1, (one - l1).astype(np.float32)

    h1, h1p, hl1, hl0 = axis(in_h, out_h)
    w1, w1p, wl1, wl0 = axis(in_w, out_w)
    a, b = src[h1][:, w1], src[h1][:, w1 + w1p]
    c, d = src[h1 + h1p][:, w1], src[h1 + h1p][:, w1 + w1p]
    wl0b, wl1b = np.broadcast_to(wl0[None, :], a.shape), np.broadcast_to(wl1[None, :], a.shape)
    hl0b, hl1b = np.broadcast_to(hl0[:, None], a.shape), np.broadcast_to(hl1[:, None], a.shape)
    if fma:
        top = _fma(wl0b, a, (wl1b * b).astype(np.float32))
        bot = _fma(wl0b, c, (wl1b * d).astype(np.float32))
        val = _fma(hl0b, top, (hl1b * bot).astype(np.float32))
    else:
        top = ((wl0b * a).astype(np.float32) + (wl1b * b).astype(np.float32)).astype(np.float32)
        bot = ((wl0b * c).astype(np.float32) + (wl1b * d).astype(np.float32)).astype(np.float32)
        val = ((hl0b * top).astype(np.float32) + (hl1b * bot).astype(np.float32)).astype(np.float32)
    return _bf16_round(val)


def mask_to_binary(mask_logits_bf16_as_f32: np.ndarray, valid_h4: int, valid_w4: int, img_h: int, img_w: int) -> np.ndarray:
    up = bilinear_bf16(mask_logits_bf16_as_f32[:valid_h4, :valid_w4], img_h, img_w)
    return (up > np.float32(2.0 ** -7)).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------- utils.py:262-264
def rle_counts(mask_hw: np.ndarray):
    """maskApi.c rleEncode: runs of the column-major flattening, first run counts zeros (may be 0 long)."""
    flat = np.asarray(mask_hw, dtype=np.uint8).flatten(order="F")
    counts, prev, c = [], 0, 0
    for v in flat.tolist():
        if v != prev:
            counts.append(c)
            c, prev = 0, v
        c += 1
    counts.append(c)
    return counts


def rle_to_string(counts) -> str:
    """maskApi.c rleToString."""
    out = []
    for i, cnt in enumerate(counts):
        x = int(cnt)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c = x & 0x1F
            x >>= 5                      # arithmetic shift (Python ints floor-shift like C's long)
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return "".join(out)


def rle_from_string(s: str):
    """maskApi.c rleFrString (for the round-trip test)."""
    counts, p = [], 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def rle_decode(counts, h: int, w: int) -> np.ndarray:
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in counts:
        flat[pos:pos + c] = v
        pos += c
        v ^= 1
    return flat.reshape((h, w), order="F")


def result_line(image_id, score: float, category, box4, mask_logits_bf16_as_f32, valid_hw, img_w: int, img_h: int) -> str:
    m = mask_to_binary(mask_logits_bf16_as_f32, int(valid_hw[0]) * 4, int(valid_hw[1]) * 4, img_h, img_w)
    rle = {"size": [int(img_h), int(img_w)], "counts": rle_to_string(rle_counts(m))}
    return json.dumps({"image_id": image_id, "score": score, "category": category, "bbox": eval_box(box4, img_w, img_h), "mask": rle})
