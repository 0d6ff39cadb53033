"""CPU: the C-ABI library loads, exports every symbol include/padt_b200.h declares, agrees with the Python weight inventory,
and its host-only ViT index tables match the oracle / HF-derived known answers.  No GPU compute is called."""
import ctypes as C
import os
import re

import pytest
import torch

import padt_oracle as O
from padt_b200 import _capi
from padt_b200.config import padt_pro_3b, padt_pro_7b, padt_small, padt_tiny
from padt_b200.weights import num_params, weight_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from padt_b200 import build
    build.build(force=False, verbose=False)
    return _capi.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "padt_b200.h")).read()
    declared = set(re.findall(r"\b(padt_[a-z0-9_]+)\s*\(", header))
    declared.discard("padt_stream_t")
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


@pytest.mark.parametrize("cfg_fn", [padt_tiny, padt_small, padt_pro_3b, padt_pro_7b])
def test_weight_inventory_matches(lib, cfg_fn):
    cfg = cfg_fn()
    cfg.validate()
    d = _capi.dims_from_config(cfg)
    names = list(weight_shapes(cfg))
    assert lib.padt_weight_count(C.byref(d)) == len(names)
    for i, n in enumerate(names):
        assert lib.padt_weight_name(C.byref(d), i).decode() == n


def test_param_counts():
    # the PaDT decoder has 98,013,353 parameters at llm_hidden_state = 2048
    cfg = padt_pro_3b()
    dec = sum(int(torch.tensor(s).prod()) for n, s in weight_shapes(cfg).items() if n.startswith("vl_decoder."))
    assert dec == 98013353
    assert 3.7e9 < num_params(cfg) < 4.0e9


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_capi.PadtError):
        _capi.load()


def _plan(lib, cfg, grids):
    d = _capi.dims_from_config(cfg)
    g = _capi.i64_array([x for t in grids for x in t])
    P = sum(t * h * w for t, h, w in grids)
    M = P // 4
    erm, hp, wp = (C.c_int32 * P)(), (C.c_int32 * P)(), (C.c_int32 * P)()
    mrm, wl = (C.c_int32 * M)(), (C.c_int32 * M)()
    nw, npch = C.c_int32(), C.c_int32()
    _capi.check(lib.padt_vit_plan_host(C.byref(d), g, len(grids), erm, hp, wp, mrm, wl, C.byref(nw), C.byref(npch)), "plan")
    assert npch.value == P
    return list(erm), list(hp), list(wp), list(mrm), list(wl)[: nw.value]


@pytest.mark.parametrize("grids,first8,n_win,lens", [
    ([(1, 24, 24)], [0, 1, 2, 3, 12, 13, 14, 15], 9, {64}),
    ([(1, 96, 96)], [0, 1, 2, 3, 48, 49, 50, 51], 144, {64}),
    ([(1, 30, 46)], [0, 1, 2, 3, 23, 24, 25, 26], 24, {36, 48, 64}),
])
def test_window_index_kats(lib, grids, first8, n_win, lens):
    """SURVEY.md §8a row V3 known answers (probed on the HF implementation)."""
    erm, hp, wp, mrm, wl = _plan(lib, padt_pro_3b(), grids)
    assert mrm[:8] == first8 and len(wl) == n_win and set(wl) == lens


def test_vit_plan_matches_oracle(lib):
    cfg = padt_tiny()
    for grids in ([(1, 8, 12)], [(1, 8, 12), (1, 4, 8)], [(1, 10, 6), (2, 4, 4), (1, 16, 16)]):
        erm, hp, wp, mrm, wl = _plan(lib, cfg, grids)
        wi, cu = O.vit_window_index(grids, cfg.vision.window_size, 2, cfg.vision.patch_size)
        assert mrm == wi.tolist()
        cu = torch.unique_consecutive(torch.tensor(cu)).tolist()
        assert wl == [b - a for a, b in zip(cu[:-1], cu[1:])]
        # embed_row_map is the scatter form of hidden[window_index] at merge-unit granularity (padt.py:70-72)
        P = sum(t * h * w for t, h, w in grids)
        x = torch.arange(P)
        gathered = x.reshape(P // 4, 4)[wi].reshape(-1)
        scattered = torch.empty(P, dtype=torch.long)
        scattered[torch.tensor(erm)] = x
        assert torch.equal(gathered, scattered)
        # rotary ids in window order reproduce the oracle's table (HF:382-409)
        rot = O.vit_rot_pos_emb(grids, 80, 2).reshape(P // 4, 4, -1)[wi].reshape(P, -1)
        inv = 1.0 / (10000.0 ** (torch.arange(0, 40, 2, dtype=torch.float) / 40))
        mine = torch.cat([torch.outer(torch.tensor(hp, dtype=torch.float), inv), torch.outer(torch.tensor(wp, dtype=torch.float), inv)], dim=1)
        assert torch.equal(rot, mine)


def test_bench_workload_constants_match_the_design():
    """bench.py's algorithmic bytes per decode step (the roofline numerator) and prompt builder: DESIGN.md §4 / SURVEY.md §8d."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from padt_b200.config import PRESETS
    c3, c7 = PRESETS["PaDT_Pro_3B"](), PRESETS["PaDT_Pro_7B"]()
    b3 = bench.decode_bytes_per_step(c3, 2304, 2350 + 128)
    assert abs(b3 - 6.27e9) < 0.01e9, b3                      # 3B, B=1, 1344 px: 6.27 GB per token
    layer_params = (b3 - 2 * 151936 * 2048 - 2 * 2304 * 2048 - (2478 + 1) * 36864) / 2
    assert abs(layer_params - 2.7748e9) < 1e6                 # 2.7748 B layer parameters
    b7 = bench.decode_bytes_per_step(c7, 2304, 2350 + 128)
    assert abs(b7 - 14.30e9) < 0.02e9, b7                     # 7B: 14.30 GB per token
    ids = bench.build_prompt(c3, bench.L_PROMPT, 2304, seed=100)
    assert ids.shape == (1, 2350) and int((ids == c3.image_token_id).sum()) == 2304
    assert int(ids[0, 14]) == c3.vision_start_token_id and int(ids[0, 14 + 1 + 2304]) == c3.vision_end_token_id
