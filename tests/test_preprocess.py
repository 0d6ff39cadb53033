"""Image pre-processing: the oracle against Pillow / transformers' own processor and the committed golden (CPU),
the CUDA path against the oracle bit for bit (GPU, through the C ABI)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import preprocess_oracle as PO  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "preprocess_small.npz")
MAXPIX_GOLDEN = 28 * 28 * 256


def _golden_cases():
    z = np.load(GOLDEN)
    names = sorted({k.split(".")[0] for k in z.files})
    return [(n, z[n + ".image"], z[n + ".pixel_values"], z[n + ".grid"]) for n in names]


# ------------------------------------------------------------------------------------------------ CPU: the oracle is pinned
def test_oracle_matches_reference_processor_golden():
    for name, img, pv, grid in _golden_cases():
        got, g = PO.preprocess(img, max_pixels=MAXPIX_GOLDEN)
        assert list(g) == grid[0].tolist(), name
        assert got.dtype == np.float32 and np.array_equal(got, pv), name


def test_oracle_resize_matches_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for (h, w, oh, ow) in [(37, 53, 56, 84), (480, 640, 476, 644), (700, 900, 336, 420), (100, 100, 280, 280), (64, 1500, 28, 644), (57, 57, 57, 57)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(PO.pil_bicubic_resize(img, oh, ow), ref), (h, w, oh, ow)


def test_oracle_matches_live_transformers_processor():
    pytest.importorskip("PIL.Image")
    from PIL import Image
    try:
        from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil as Proc
    except ImportError:
        pytest.skip("transformers without the PIL-backend Qwen2-VL processor")
    proc = Proc()
    rng = np.random.default_rng(1)
    for (h, w) in [(420, 644), (29, 300), (901, 333)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        r = proc(images=[Image.fromarray(img)], return_tensors="np")
        got, g = PO.preprocess(img)
        assert r["image_grid_thw"].tolist() == [list(g)]
        assert np.array_equal(got, r["pixel_values"])


def test_smart_resize_c_abi_equals_python_statement():
    from padt_b200.preprocess import smart_resize
    rnd = random.Random(3)
    checked = 0
    for _ in range(5000):
        h, w = rnd.randint(1, 6000), rnd.randint(1, 6000)
        mp, Mp = rnd.choice([56 * 56, 3136, 28 * 28 * 4]), rnd.choice([12845056, 1003520, 1344 * 1344])
        try:
            want = PO.smart_resize(h, w, 28, mp, Mp)
        except ValueError:
            with pytest.raises(ValueError):
                smart_resize(h, w, 28, mp, Mp)
            continue
        assert smart_resize(h, w, 28, mp, Mp) == want
        checked += 1
    assert checked > 4000
    # half-to-even rounding of round(h / 28): 42 / 28 = 1.5 -> 2, 70 / 28 = 2.5 -> 2
    assert smart_resize(42, 280, 28, 1, 10 ** 9) == (56, 280) and smart_resize(70, 280, 28, 1, 10 ** 9) == (56, 280)


# ------------------------------------------------------------------------------------------------ GPU: CUDA path == oracle, bit for bit
@pytest.mark.gpu
def test_gpu_preprocess_golden_bit_exact():
    from padt_b200.preprocess import GpuImageProcessor
    f32 = GpuImageProcessor(max_pixels=MAXPIX_GOLDEN, dtype=torch.float32)
    b16 = GpuImageProcessor(max_pixels=MAXPIX_GOLDEN, dtype=torch.bfloat16)
    cases = _golden_cases()
    out = f32([c[1] for c in cases])
    assert out["image_grid_thw"].tolist() == [c[3][0].tolist() for c in cases]
    want = np.concatenate([c[2] for c in cases], axis=0)
    got = out["pixel_values"].cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)        # float32: what the reference processor returns
    got16 = b16([c[1] for c in cases])["pixel_values"].cpu()
    assert torch.equal(got16, torch.from_numpy(want).to(torch.bfloat16))  # bf16: its cast at the patch embed (HF:108-114)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,max_pixels", [(1344, 1344, 1344 * 1344), (1344, 1344, 28 * 28 * 1280), (3000, 4000, 28 * 28 * 1280), (200, 3000, 28 * 28 * 1280),
                                            (31, 47, 28 * 28 * 1280), (640, 480, 28 * 28 * 1280)])
def test_gpu_preprocess_random_sizes_bit_exact(h, w, max_pixels):
    from padt_b200.preprocess import GpuImageProcessor
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[: h // 2, :, 0] = (np.arange(w) * 255 // max(w - 1, 1)).astype(np.uint8)[None, :]   # a smooth ramp next to the noise
    want, grid = PO.preprocess(img, max_pixels=max_pixels)
    out = GpuImageProcessor(max_pixels=max_pixels, dtype=torch.float32)(torch.from_numpy(img))
    assert out["image_grid_thw"].tolist() == [list(grid)]
    assert np.array_equal(out["pixel_values"].cpu().numpy(), want)


@pytest.mark.gpu
def test_gpu_preprocess_feeds_generate():
    """The bf16 rows go straight into generate(): same tokens as feeding the oracle's pixel_values."""
    from padt_b200 import PaDTForConditionalGeneration
    from padt_b200.config import PRESETS
    from padt_b200.weights import random_weights
    from padt_b200.preprocess import GpuImageProcessor
    cfg = PRESETS["PaDT_tiny"]()
    w = {k: v.cuda() for k, v in random_weights(cfg, seed=0).items()}
    model = PaDTForConditionalGeneration(cfg, w)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (100, 150, 3), dtype=np.uint8)
    proc = GpuImageProcessor(max_pixels=28 * 28 * 16, patch_size=cfg.vision.patch_size, temporal_patch_size=cfg.vision.temporal_patch_size,
                             merge_size=cfg.vision.spatial_merge_size)
    out = proc([img])
    want, grid = PO.preprocess(img, max_pixels=28 * 28 * 16)
    assert torch.equal(out["pixel_values"].cpu(), torch.from_numpy(want).to(torch.bfloat16))
    n_img = grid[1] * grid[2] // 4
    ids = torch.tensor([[5, 6, cfg.vision_start_token_id] + [cfg.image_token_id] * n_img + [cfg.vision_end_token_id, 7, 8]], dtype=torch.int64)
    kw = dict(input_ids=ids.cuda(), attention_mask=torch.ones_like(ids).cuda(), image_grid_thw=out["image_grid_thw"].cuda(), max_new_tokens=6,
              eos_token_id=[])
    a = model.generate(pixel_values=out["pixel_values"], **kw)
    b = model.generate(pixel_values=torch.from_numpy(want).to(torch.bfloat16).cuda(), **kw)
    assert torch.equal(a.sequences if hasattr(a, "sequences") else a, b.sequences if hasattr(b, "sequences") else b)


def test_processor_host_side_input_handling():
    """Host logic of GpuImageProcessor that needs no GPU: accepted image containers, RGB conversion, output sizes, argument checks."""
    from padt_b200.preprocess import GpuImageProcessor, _to_hwc_uint8
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2)
    arr = rng.integers(0, 256, (30, 40, 3), dtype=np.uint8)
    for img in (arr, torch.from_numpy(arr), Image.fromarray(arr)):
        t = _to_hwc_uint8(img)
        assert t.dtype == torch.uint8 and t.shape == (30, 40, 3) and torch.equal(t, torch.from_numpy(arr))
    grey = Image.fromarray(arr[..., 0])                                   # do_convert_rgb: a grey image becomes three equal channels
    assert torch.equal(_to_hwc_uint8(grey), torch.from_numpy(np.stack([arr[..., 0]] * 3, axis=-1)))
    for bad in (arr.astype(np.float32), arr[..., :2], torch.zeros(3, 30, 40, dtype=torch.uint8)):
        with pytest.raises(TypeError):
            _to_hwc_uint8(bad)
    p = GpuImageProcessor(device="cpu", max_pixels=28 * 28 * 1280)        # construction and size arithmetic never touch the device
    assert p.output_size(1344, 1344) == PO.smart_resize(1344, 1344, 28, 56 * 56, 28 * 28 * 1280) == (980, 980)
    assert GpuImageProcessor(device="cpu", max_pixels=1344 * 1344).output_size(1344, 1344) == (1344, 1344)
    with pytest.raises(ValueError):
        GpuImageProcessor(device="cpu", dtype=torch.float16)
    with pytest.raises(ValueError):
        p.output_size(10, 4000)                                           # aspect ratio > 200, like the reference
