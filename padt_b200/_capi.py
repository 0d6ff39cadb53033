"""ctypes binding of ``libpadt_b200.so`` (C ABI declared in ``include/padt_b200.h``).

The product path has no CPU fallback: if the CUDA library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpadt_b200.so")

_lib: Optional[C.CDLL] = None


class PadtError(RuntimeError):
    pass


class PadtDims(C.Structure):
    _fields_ = [
        ("vit_depth", C.c_int32), ("vit_hidden", C.c_int32), ("vit_inter", C.c_int32), ("vit_heads", C.c_int32),
        ("vit_in_ch", C.c_int32), ("vit_patch", C.c_int32), ("vit_tpatch", C.c_int32), ("vit_merge", C.c_int32),
        ("vit_window", C.c_int32), ("vit_out_hidden", C.c_int32), ("vit_n_fullatt", C.c_int32),
        ("vit_fullatt", C.c_int32 * 8),
        ("D", C.c_int32), ("I", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("vocab", C.c_int32), ("rms_eps", C.c_float),
        ("mrope0", C.c_int32), ("mrope1", C.c_int32), ("mrope2", C.c_int32), ("tie_word_embeddings", C.c_int32),
        ("dec_hidden", C.c_int32), ("dec_inter", C.c_int32), ("dec_heads", C.c_int32),
        ("use_proto_proj", C.c_int32), ("lora_r", C.c_int32),
    ]


class PadtGenBuffers(C.Structure):
    _fields_ = [
        ("tokens_out", C.c_void_p), ("argmax_out", C.c_void_p), ("hidden_steps", C.c_void_p), ("last_hidden", C.c_void_p),
        ("prefill_hidden", C.c_void_p), ("forced", C.c_void_p), ("max_steps", C.c_int32),
        ("logits_out", C.c_void_p), ("logits_ld", C.c_int32),
    ]


#: every symbol include/padt_b200.h declares (tests/test_capi_symbols.py checks the library exports all of them)
EXPORTED_SYMBOLS = (
    "padt_last_error", "padt_version", "padt_launch_count", "padt_weight_count", "padt_weight_name", "padt_create", "padt_destroy", "padt_set_rope_table",
    "padt_vit_plan_host", "padt_vit_forward", "padt_prototypes", "padt_prefill", "padt_decode", "padt_decode_status", "padt_set_option",
    "padt_vl_decode", "padt_op_gemm", "padt_op_rmsnorm", "padt_op_attention", "padt_op_attention_split", "padt_op_rope_neox", "padt_op_logit_mask",
    "padt_debug_mega_plan", "padt_debug_mega_partition", "padt_smart_resize", "padt_preprocess_workspace_bytes", "padt_preprocess_image",
    "padt_mask_runs_workspace_bytes", "padt_mask_to_runs", "padt_rle_string",
)


def dims_from_config(cfg) -> PadtDims:
    v, t, d = cfg.vision, cfg.text, cfg.decoder
    dm = PadtDims()
    dm.vit_depth, dm.vit_hidden, dm.vit_inter, dm.vit_heads = v.depth, v.hidden_size, v.intermediate_size, v.num_heads
    dm.vit_in_ch, dm.vit_patch, dm.vit_tpatch, dm.vit_merge = v.in_channels, v.patch_size, v.temporal_patch_size, v.spatial_merge_size
    dm.vit_window, dm.vit_out_hidden = v.window_size, v.out_hidden_size
    fa = list(v.fullatt_block_indexes)
    assert len(fa) <= 8
    dm.vit_n_fullatt = len(fa)
    for i, x in enumerate(fa):
        dm.vit_fullatt[i] = x
    dm.D, dm.I, dm.n_layers, dm.n_heads, dm.n_kv_heads, dm.vocab = (t.hidden_size, t.intermediate_size, t.num_layers,
                                                                      t.num_heads, t.num_kv_heads, t.vocab_size)
    dm.rms_eps = t.rms_norm_eps
    dm.mrope0, dm.mrope1, dm.mrope2 = t.mrope_section
    dm.tie_word_embeddings = int(t.tie_word_embeddings)
    dm.dec_hidden, dm.dec_inter, dm.dec_heads = d.hidden_size, d.intermediate_size, d.num_heads
    dm.use_proto_proj, dm.lora_r = int(cfg.use_visual_prototype_projection), cfg.lora_r
    return dm


def load() -> C.CDLL:
    """Load the CUDA library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PadtError(f"{LIB_PATH} is missing: build it with `python -m padt_b200.build` "
                        f"(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.padt_last_error.restype = C.c_char_p
    lib.padt_version.restype = C.c_char_p
    lib.padt_launch_count.restype = C.c_ulonglong
    lib.padt_weight_name.restype = C.c_char_p
    lib.padt_weight_name.argtypes = [C.POINTER(PadtDims), C.c_int]
    lib.padt_weight_count.argtypes = [C.POINTER(PadtDims)]
    lib.padt_create.argtypes = [C.POINTER(PadtDims), C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                C.POINTER(C.c_void_p)]
    lib.padt_destroy.argtypes = [C.c_void_p]
    lib.padt_set_rope_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.padt_vit_forward.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.padt_vit_plan_host.argtypes = [C.POINTER(PadtDims), C.POINTER(C.c_int64), C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int32)] * 2
    lib.padt_prototypes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.padt_prefill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_int64), C.c_int, C.c_int64, C.POINTER(PadtGenBuffers), C.c_void_p]
    lib.padt_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.padt_decode_status.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]
    lib.padt_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.padt_vl_decode.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_void_p]
    lib.padt_op_gemm.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.padt_op_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.padt_op_attention.argtypes = [C.c_void_p] * 4 + [C.c_int64] * 8 + [C.c_void_p] * 5 + [C.c_int] * 6 + [C.c_float, C.c_void_p]
    lib.padt_op_logit_mask.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]
    lib.padt_op_attention_split.argtypes = [C.c_void_p] * 4 + [C.c_int64] * 8 + [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                                                                                                       C.c_void_p]
    lib.padt_op_rope_neox.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.padt_smart_resize.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int32)]
    lib.padt_preprocess_workspace_bytes.argtypes = [C.c_int] * 4
    lib.padt_preprocess_workspace_bytes.restype = C.c_size_t
    lib.padt_preprocess_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_double,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.padt_mask_runs_workspace_bytes.argtypes = [C.c_int, C.c_int]
    lib.padt_mask_runs_workspace_bytes.restype = C.c_size_t
    lib.padt_mask_to_runs.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]
    lib.padt_rle_string.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_char_p, C.c_int]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("padt_last_error", "padt_version", "padt_weight_name", "padt_launch_count", "padt_preprocess_workspace_bytes",
                        "padt_mask_runs_workspace_bytes"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().padt_last_error().decode("utf-8", "replace")
        raise PadtError(f"{what} failed (rc={rc}): {msg}")


def i32_array(vals):
    arr = (C.c_int32 * len(vals))(*[int(v) for v in vals])
    return arr


def i64_array(vals):
    arr = (C.c_int64 * max(1, len(vals)))(*[int(v) for v in vals])
    return arr
