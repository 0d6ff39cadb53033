"""Image pre-processing: GPU path (pinned host uint8 image -> H2D -> padt_preprocess_image -> bf16 patch rows) next to the reference's
host path (transformers' PIL-backend Qwen2-VL processor = the 4.50 slow processor) on the same images.  One JSON line per case."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from padt_b200.preprocess import GpuImageProcessor

CASES = [("1344x1344 -> 1344x1344 (config 2 image, max_pixels = 1344^2)", 1344, 1344, 1344 * 1344),
         ("3000x4000 photo -> default max_pixels (28*28*1280)", 3000, 4000, 28 * 28 * 1280),
         ("480x640 (COCO-size) -> default", 480, 640, 28 * 28 * 1280)]
try:
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil as Proc
except ImportError:
    Proc = None

for name, h, w, mp in CASES:
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).pin_memory()
    proc = GpuImageProcessor(max_pixels=mp)
    for _ in range(3):
        out = proc([img])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        out = proc([img])
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / iters
    cpu_ms = None
    if Proc is not None:
        pil = Image.fromarray(img.numpy())
        ref = Proc(max_pixels=mp)
        ref(images=[pil], return_tensors="pt")
        t = time.perf_counter()
        for _ in range(3):
            r = ref(images=[pil], return_tensors="pt")
        cpu_ms = (time.perf_counter() - t) / 3 * 1e3
        same = torch.equal(out["pixel_values"].cpu(), r["pixel_values"].to(torch.bfloat16))
    print(json.dumps(dict(case=name, grid=out["image_grid_thw"].tolist(), gpu_ms_incl_h2d=round(gpu_ms, 3), reference_cpu_ms=None if cpu_ms is None else round(cpu_ms, 1),
                          speedup=None if cpu_ms is None else round(cpu_ms / gpu_ms, 1), bit_equal_to_reference=None if cpu_ms is None else bool(same),
                          in_bytes=h * w * 3, out_bytes=out["pixel_values"].numel() * 2)))
