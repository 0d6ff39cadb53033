"""Writes tests/golden/preprocess_small.npz: input images and the pixel_values / grids transformers' own PIL-backend Qwen2-VL image
processor (the arithmetic of the reference's pinned transformers 4.50 slow processor) returns for them.  Run here (needs Pillow +
transformers); the fixture travels to the GPU box.   python oracle/make_preprocess_golden.py"""
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cases():
    rng = np.random.default_rng(20251017)
    out = {}
    for name, (h, w) in {"up_37x53": (37, 53), "down_300x500": (300, 500), "demo_420x644": (420, 644), "tall_500x90": (500, 90)}.items():
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if name.startswith("down"):   # smooth content as well as noise
            yy, xx = np.mgrid[0:h, 0:w]
            img[..., 0] = (xx * 255 // w).astype(np.uint8)
            img[..., 1] = (yy * 255 // h).astype(np.uint8)
        out[name] = img
    return out


def main():
    try:
        from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil as Proc
        proc = Proc(min_pixels=56 * 56, max_pixels=28 * 28 * 256)
    except ImportError:   # transformers 4.x: the slow processor is the default class
        from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor as Proc
        proc = Proc(min_pixels=56 * 56, max_pixels=28 * 28 * 256)
    blob = {}
    for name, img in cases().items():
        r = proc(images=[Image.fromarray(img)], return_tensors="np")
        blob[name + ".image"] = img
        blob[name + ".pixel_values"] = r["pixel_values"].astype(np.float32)
        blob[name + ".grid"] = np.asarray(r["image_grid_thw"], dtype=np.int64)
        print(name, img.shape, "->", r["image_grid_thw"].tolist(), r["pixel_values"].shape)
    path = os.path.join(ROOT, "tests", "golden", "preprocess_small.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
