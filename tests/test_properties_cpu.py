"""CPU property tests of the host-side integer work on random grids / ids (hypothesis): ViT window plan through the C ABI,
VRT id shifts, M-RoPE position ids.  Size-independent invariants the reference's algorithms guarantee."""
import ctypes as C

import torch
from hypothesis import given, settings, strategies as st

import padt_oracle as O
from padt_b200 import _capi
from padt_b200.config import padt_pro_3b, padt_tiny
from padt_b200.processor import VisonTextProcessingClass
from padt_b200.rope_index import get_rope_index

even = st.integers(min_value=1, max_value=40).map(lambda x: 2 * x)
grid_st = st.tuples(st.just(1), even, even)


def _plan(lib, cfg, grids):
    d = _capi.dims_from_config(cfg)
    g = _capi.i64_array([x for t in grids for x in t])
    P = sum(t * h * w for t, h, w in grids)
    M = P // 4
    erm, hp, wp = (C.c_int32 * P)(), (C.c_int32 * P)(), (C.c_int32 * P)()
    mrm, wl = (C.c_int32 * M)(), (C.c_int32 * M)()
    nw, npch = C.c_int32(), C.c_int32()
    _capi.check(lib.padt_vit_plan_host(C.byref(d), g, len(grids), erm, hp, wp, mrm, wl, C.byref(nw), C.byref(npch)), "plan")
    return list(erm), list(hp), list(wp), list(mrm), list(wl)[: nw.value]


@settings(max_examples=40, deadline=None)
@given(st.lists(grid_st, min_size=1, max_size=3))
def test_window_plan_invariants(grids):
    lib = _capi.load()
    cfg = padt_pro_3b()
    erm, hp, wp, mrm, wl = _plan(lib, cfg, grids)
    P = sum(t * h * w for t, h, w in grids)
    M = P // 4
    assert sorted(mrm) == list(range(M)), "window_index is a permutation of the merge units"
    assert sorted(erm) == list(range(P)), "the patch scatter map is a permutation"
    assert sum(wl) == P and all(0 < x <= 64 for x in wl), "windows tile the patches; at most 4x4 merge units = 64 patches each"
    # every window stays inside one image, and (h, w) ids are inside that image's grid
    off = 0
    for (t, h, w) in grids:
        n = t * h * w
        assert max(hp[off:off + n]) == h - 1 and max(wp[off:off + n]) == w - 1 and min(hp[off:off + n]) == 0 and min(wp[off:off + n]) == 0
        off += n
    # same answer as the oracle's restatement of get_window_index (HF:411-453)
    wi, _ = O.vit_window_index(list(grids), cfg.vision.window_size, 2, cfg.vision.patch_size)
    assert mrm == wi.tolist()


@settings(max_examples=60, deadline=None)
@given(st.lists(grid_st, min_size=1, max_size=4), st.integers(min_value=0, max_value=2**31 - 1))
def test_vrt_id_shift_round_trip(grids, seed):
    V = 1000

    class Tok:
        def __init__(self):
            self.vocab = {}

        def get_vocab(self):
            return self.vocab

        def add_tokens(self, toks):
            for t in toks:
                self.vocab.setdefault(t, len(self.vocab))

    class Proc:
        def __init__(self):
            self.tokenizer = Tok()
            self.tokenizer.vocab = {f"t{i}": i for i in range(V)}

    proc = VisonTextProcessingClass(Proc())
    g = torch.Generator().manual_seed(seed)
    grid = torch.tensor(grids)
    counts = [t * h * w // 4 for t, h, w in grids]
    B, T = len(grids), 9
    ids = torch.randint(0, V, (B, T), generator=g)
    for b in range(B):       # sprinkle valid LOCAL VRT ids of the row's own image
        for _ in range(3):
            ids[b, int(torch.randint(0, T, (1,), generator=g))] = V + int(torch.randint(0, counts[b], (1,), generator=g))
    local = ids.clone()
    glob = proc.assign_to_global_vrt_id(ids, grid)
    assert glob is ids                                             # in place, like the reference (padt_processor.py:36-42)
    offs = [sum(counts[:b]) for b in range(B)]
    for b in range(B):
        is_vrt = local[b] >= V
        assert torch.equal(glob[b][~is_vrt], local[b][~is_vrt])
        assert torch.equal(glob[b][is_vrt], local[b][is_vrt] + offs[b])
        assert bool(((glob[b][is_vrt] - V) < offs[b] + counts[b]).all())
    assert torch.equal(O.assign_to_global_vrt_id(V, 2, local.clone(), grid), glob)
    back = proc.assign_to_local_vrt_id(glob, grid)
    assert torch.equal(back, local)


@settings(max_examples=40, deadline=None)
@given(grid_st, st.integers(min_value=0, max_value=12), st.integers(min_value=0, max_value=12), st.integers(min_value=0, max_value=9))
def test_rope_index_invariants(grid, n_pre, n_post, n_pad):
    cfg = padt_tiny()
    t, h, w = grid
    Mi = h * w // 4
    row = list(range(1, n_pre + 1)) + [cfg.vision_start_token_id] + [cfg.image_token_id] * Mi + [cfg.vision_end_token_id] + list(range(1, n_post + 1))
    ids = torch.tensor([[cfg.pad_token_id] * n_pad + row])
    mask = torch.tensor([[0] * n_pad + [1] * len(row)])
    pos, delta = get_rope_index(ids, torch.tensor([list(grid)]), mask, cfg.image_token_id, cfg.vision_start_token_id, 2)
    p = pos[:, 0, n_pad:]
    assert bool((pos[:, 0, :n_pad] == 1).all())
    st_ = n_pre + 1                                   # first image token index inside the row
    # text before the image: all planes equal, running index
    assert p[:, :st_].tolist() == [list(range(st_))] * 3
    img = p[:, st_:st_ + Mi]
    assert bool((img[0] == st_).all())                # temporal plane stays at the start value (second_per_grid_t = 0)
    assert img[1].tolist() == [st_ + r for r in range(h // 2) for _ in range(w // 2)]
    assert img[2].tolist() == [st_ + c for _ in range(h // 2) for c in range(w // 2)]
    nxt = st_ + max(h // 2, w // 2)                   # text resumes at max + 1 over all planes
    tail = p[:, st_ + Mi:]
    assert tail.tolist() == [list(range(nxt, nxt + tail.shape[1]))] * 3
    assert int(delta[0, 0]) == int(p.max()) + 1 - ids.shape[1]
    p2, d2 = O.get_rope_index(cfg, ids, torch.tensor([list(grid)]), mask)
    assert torch.equal(pos, p2) and torch.equal(delta, d2)
