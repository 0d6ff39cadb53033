"""Model configuration for the PaDT generate/decode hot path.

The reference keeps these numbers in the HF checkpoint's ``config.json`` (Qwen2.5-VL-3B/7B) plus the
``vl_decoder`` dict injected at ``src/PaDT/trainer/padt_sft_trainer.py:149-162`` and fixed up at
``src/PaDT/models/padt.py:127-130``.  SURVEY.md §8 lists the 3B / 7B constants used for the presets.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Tuple


@dataclass(frozen=True)
class VisionConfig:
    depth: int = 32
    hidden_size: int = 1280
    intermediate_size: int = 3420
    num_heads: int = 16
    in_channels: int = 3
    patch_size: int = 14
    temporal_patch_size: int = 2
    spatial_merge_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: Tuple[int, ...] = (7, 15, 23, 31)
    out_hidden_size: int = 2048

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def patch_dim(self) -> int:
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size


@dataclass(frozen=True)
class TextConfig:
    hidden_size: int = 2048
    intermediate_size: int = 11008
    num_layers: int = 36
    num_heads: int = 16
    num_kv_heads: int = 2
    vocab_size: int = 151936
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    mrope_section: Tuple[int, int, int] = (16, 24, 24)
    tie_word_embeddings: bool = True

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads


@dataclass(frozen=True)
class DecoderConfig:
    """``vl_decoder`` dict of the reference (padt_sft_trainer.py:151-160)."""
    hidden_size: int = 1280
    intermediate_size: int = 3420
    num_heads: int = 16
    use_mask_loss: bool = True

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads


@dataclass(frozen=True)
class PadtConfig:
    name: str = "PaDT_Pro_3B"
    vision: VisionConfig = field(default_factory=VisionConfig)
    text: TextConfig = field(default_factory=TextConfig)
    decoder: DecoderConfig = field(default_factory=DecoderConfig)
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_end_token_id: int = 151653
    eos_token_ids: Tuple[int, ...] = (151645, 151643)
    pad_token_id: int = 151643
    use_visual_prototype_projection: bool = True
    lora_r: int = 64  # padt.py:120

    def to_dict(self):
        return asdict(self)

    def validate(self) -> None:
        v, t, d = self.vision, self.text, self.decoder
        assert v.hidden_size % v.num_heads == 0 and t.hidden_size % t.num_heads == 0
        assert t.num_heads % t.num_kv_heads == 0
        assert sum(t.mrope_section) * 2 == t.head_dim, "mrope sections must cover head_dim/2"
        assert v.out_hidden_size == t.hidden_size
        assert d.hidden_size == v.hidden_size, "padt_decoder.py:220 adds low-res (decoder dim) to high-res (ViT dim)"
        assert d.head_dim == v.head_dim, "visual_pe (ViT head_dim) is applied to decoder heads (padt_decoder.py:39-51)"
        assert d.hidden_size % 16 == 0


def padt_pro_3b() -> PadtConfig:
    return PadtConfig()


def padt_pro_7b() -> PadtConfig:
    return PadtConfig(
        name="PaDT_Pro_7B",
        vision=VisionConfig(out_hidden_size=3584),
        text=TextConfig(hidden_size=3584, intermediate_size=18944, num_layers=28, num_heads=28, num_kv_heads=4,
                        vocab_size=152064, tie_word_embeddings=False),
        decoder=DecoderConfig(),
    )


def padt_tiny(tie: bool = True, vocab: int = 512) -> PadtConfig:
    """Structure-preserving miniature (head dims 80 / 128 kept so the same CUDA instantiations run)."""
    return PadtConfig(
        name="PaDT_tiny",
        vision=VisionConfig(depth=4, hidden_size=160, intermediate_size=216, num_heads=2,
                            fullatt_block_indexes=(1, 3), out_hidden_size=256),
        text=TextConfig(hidden_size=256, intermediate_size=640, num_layers=3, num_heads=2, num_kv_heads=1,
                        vocab_size=vocab, tie_word_embeddings=tie),
        decoder=DecoderConfig(hidden_size=160, intermediate_size=216, num_heads=2),
        image_token_id=vocab - 12, video_token_id=vocab - 11, vision_start_token_id=vocab - 14,
        vision_end_token_id=vocab - 13, eos_token_ids=(vocab - 10,), pad_token_id=vocab - 9,
    )


def padt_small(tie: bool = False) -> PadtConfig:
    """Mid-size config for GPU parity tests (GQA group 4, several layers, untied head)."""
    vocab = 4096
    return PadtConfig(
        name="PaDT_small",
        vision=VisionConfig(depth=6, hidden_size=320, intermediate_size=856, num_heads=4,
                            fullatt_block_indexes=(2, 5), out_hidden_size=1024),
        text=TextConfig(hidden_size=1024, intermediate_size=2752, num_layers=4, num_heads=8, num_kv_heads=2,
                        vocab_size=vocab, tie_word_embeddings=tie),
        decoder=DecoderConfig(hidden_size=320, intermediate_size=856, num_heads=4),
        image_token_id=vocab - 12, video_token_id=vocab - 11, vision_start_token_id=vocab - 14,
        vision_end_token_id=vocab - 13, eos_token_ids=(vocab - 10,), pad_token_id=vocab - 9,
    )


def padt_vit1344() -> PadtConfig:
    """Full-width ViT (1280 / 16 heads / 3420) cut to 2 blocks (window, full attention) over a miniature LLM: the config of the
    1344x1344 ViT parity fixture (tests/golden/vit1344.pt, oracle/make_golden.py)."""
    vocab = 512
    return PadtConfig(
        name="PaDT_vit1344",
        vision=VisionConfig(depth=2, fullatt_block_indexes=(1,), out_hidden_size=256),
        text=TextConfig(hidden_size=256, intermediate_size=640, num_layers=1, num_heads=2, num_kv_heads=1, vocab_size=vocab),
        decoder=DecoderConfig(),
        image_token_id=vocab - 12, video_token_id=vocab - 11, vision_start_token_id=vocab - 14,
        vision_end_token_id=vocab - 13, eos_token_ids=(vocab - 10,), pad_token_id=vocab - 9,
    )


def padt_dec7b_1344() -> PadtConfig:
    """PaDT decoder at PaDT_Pro_7B dimensions (llm_hidden_state 3584, decoder 1280 / 3420 / 16 heads) over a miniature ViT / LLM: the
    config of the 1344x1344 decoder parity fixture (BASELINE config 4 geometry; tests/golden/dec7b_1344.pt, oracle/make_golden.py)."""
    vocab = 512
    return PadtConfig(
        name="PaDT_dec7b_1344",
        vision=VisionConfig(depth=1, fullatt_block_indexes=(0,), out_hidden_size=3584),
        text=TextConfig(hidden_size=3584, intermediate_size=1024, num_layers=1, num_heads=28, num_kv_heads=4, vocab_size=vocab,
                        tie_word_embeddings=True),
        decoder=DecoderConfig(),
        image_token_id=vocab - 12, video_token_id=vocab - 11, vision_start_token_id=vocab - 14,
        vision_end_token_id=vocab - 13, eos_token_ids=(vocab - 10,), pad_token_id=vocab - 9,
    )


PRESETS = {
    "PaDT_Pro_3B": padt_pro_3b,
    "PaDT_Pro_7B": padt_pro_7b,
    "PaDT_dec7b_1344": padt_dec7b_1344,
    "PaDT_tiny": padt_tiny,
    "PaDT_small": padt_small,
    "PaDT_vit1344": padt_vit1344,
}
