"""Checkpoint entry path, host side (eval/test_demo.py:20-31): config.json in the flat transformers-4.50 layout and in the nested 5.x layout
map to the same PadtConfig; generation_config.json supplies the EOS pair / pad id / sampling defaults the reference's generate() uses."""
import json

from padt_b200.config import PadtConfig, padt_pro_3b, padt_pro_7b
from padt_b200.modeling import config_from_hf_dict, load_generation_config


def _flat(cfg: PadtConfig) -> dict:
    t, v, d = cfg.text, cfg.vision, cfg.decoder
    return {
        "_name_or_path": cfg.name, "hidden_size": t.hidden_size, "intermediate_size": t.intermediate_size, "num_hidden_layers": t.num_layers,
        "num_attention_heads": t.num_heads, "num_key_value_heads": t.num_kv_heads, "vocab_size": t.vocab_size, "rms_norm_eps": t.rms_norm_eps,
        "rope_theta": t.rope_theta, "rope_scaling": {"type": "mrope", "mrope_section": list(t.mrope_section)},
        "tie_word_embeddings": t.tie_word_embeddings, "eos_token_id": 151645, "pad_token_id": 151643, "image_token_id": 151655,
        "video_token_id": 151656, "vision_start_token_id": 151652, "vision_end_token_id": 151653,
        "vision_config": {"depth": v.depth, "hidden_size": v.hidden_size, "intermediate_size": v.intermediate_size, "num_heads": v.num_heads,
                          "in_chans": 3, "patch_size": 14, "temporal_patch_size": 2, "spatial_merge_size": 2, "window_size": 112,
                          "fullatt_block_indexes": [7, 15, 23, 31], "out_hidden_size": v.out_hidden_size},
        "vl_decoder": {"name": "PaDTDecoder", "hidden_size": d.hidden_size, "intermediate_size": d.intermediate_size, "num_heads": d.num_heads,
                       "llm_hidden_state": t.hidden_size, "spatial_merge_size": 2, "use_mask_loss": True},
        "use_visual_prototype_projection": True,
    }


def _nested(cfg: PadtConfig) -> dict:
    f = _flat(cfg)
    text_keys = ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads", "vocab_size", "rms_norm_eps")
    tc = {k: f.pop(k) for k in text_keys}
    tc["rope_parameters"] = {"rope_type": "default", "rope_theta": f.pop("rope_theta"), "mrope_section": f.pop("rope_scaling")["mrope_section"]}
    f["text_config"] = tc
    f["vision_config"]["in_channels"] = f["vision_config"].pop("in_chans")
    return f


def test_config_json_flat_and_nested_layouts():
    for preset in (padt_pro_3b(), padt_pro_7b()):
        a, b = config_from_hf_dict(_flat(preset)), config_from_hf_dict(_nested(preset))
        for c in (a, b):
            assert c.text == preset.text and c.vision == preset.vision and c.decoder == preset.decoder
            # config.json's single eos id is not authoritative: the reference stops on the generation_config pair
            assert c.eos_token_ids == (151645, 151643) and c.pad_token_id == 151643
        assert a == b


def test_generation_config_merge(tmp_path):
    cfg = padt_pro_3b()
    assert load_generation_config(str(tmp_path), cfg) == {"eos_token_id": None, "pad_token_id": None}      # no file: PadtConfig defaults stay
    (tmp_path / "generation_config.json").write_text(json.dumps(
        {"bos_token_id": 151643, "pad_token_id": 151643, "do_sample": True, "eos_token_id": [151645, 151643], "repetition_penalty": 1.05,
         "temperature": 0.1, "top_p": 0.001, "top_k": 1}))
    g = load_generation_config(str(tmp_path), cfg)
    assert g["eos_token_id"] == [151645, 151643] and g["pad_token_id"] == 151643
    assert g["do_sample"] is True and g["repetition_penalty"] == 1.05 and g["temperature"] == 0.1 and g["top_k"] == 1 and g["top_p"] == 0.001
    (tmp_path / "generation_config.json").write_text(json.dumps({"eos_token_id": 7}))
    g = load_generation_config(str(tmp_path), cfg)
    assert g["eos_token_id"] == [7] and g["pad_token_id"] == cfg.pad_token_id
