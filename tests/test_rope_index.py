"""R1: M-RoPE position ids (transformers 4.50.0 semantics).  KAT from SURVEY.md Appendix A + product-vs-oracle on random rows."""
import torch

import padt_oracle as O
from padt_b200.config import padt_tiny
from padt_b200.rope_index import get_rope_index

IMG, VS, VE = 500, 498, 499


def _cfg():
    return padt_tiny()


def test_appendix_a_kat():
    cfg = _cfg()
    row = [7, 8, cfg.vision_start_token_id] + [cfg.image_token_id] * 8 + [cfg.vision_end_token_id, 9, 512 + 3]
    ids = torch.tensor([row])
    grid = torch.tensor([[1, 4, 8]])
    for fn in (lambda i, g, m: get_rope_index(i, g, m, cfg.image_token_id, cfg.vision_start_token_id, 2), lambda i, g, m: O.get_rope_index(cfg, i, g, m)):
        pos, delta = fn(ids, grid, torch.ones_like(ids))
        assert pos[0, 0].tolist() == [0, 1, 2, 3, 3, 3, 3, 3, 3, 3, 3, 7, 8, 9]
        assert pos[1, 0].tolist() == [0, 1, 2, 3, 3, 3, 3, 4, 4, 4, 4, 7, 8, 9]
        assert pos[2, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 3, 4, 5, 6, 7, 8, 9]
        assert int(delta[0, 0]) == 10 - 14
        # left padded to 32: padded slots hold 1 on every plane, delta = 10 - 32, first decode position = 32 - 22 = 10
        pad = 32 - len(row)
        ids_p = torch.tensor([[cfg.pad_token_id] * pad + row])
        mask = torch.tensor([[0] * pad + [1] * len(row)])
        pos, delta = fn(ids_p, grid, mask)
        assert bool((pos[:, 0, :pad] == 1).all())
        assert pos[2, 0, pad:].tolist() == [0, 1, 2, 3, 4, 5, 6, 3, 4, 5, 6, 7, 8, 9]
        assert int(delta[0, 0]) == -22 and 32 + int(delta[0, 0]) == 10


def test_product_matches_oracle_random():
    cfg = _cfg()
    g = torch.Generator().manual_seed(0)
    for trial in range(20):
        B = int(torch.randint(1, 4, (1,), generator=g))
        rows, grids = [], []
        for b in range(B):
            h = 2 * int(torch.randint(1, 7, (1,), generator=g))
            w = 2 * int(torch.randint(1, 7, (1,), generator=g))
            grids.append([1, h, w])
            pre = torch.randint(0, 400, (int(torch.randint(0, 6, (1,), generator=g)),), generator=g).tolist()
            post = torch.randint(0, 400, (int(torch.randint(0, 9, (1,), generator=g)),), generator=g).tolist()
            rows.append(pre + [cfg.vision_start_token_id] + [cfg.image_token_id] * (h * w // 4) + [cfg.vision_end_token_id] + post)
        L = max(len(r) for r in rows)
        ids = torch.full((B, L), cfg.pad_token_id, dtype=torch.long)
        mask = torch.zeros((B, L), dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, L - len(r):] = torch.tensor(r)
            mask[i, L - len(r):] = 1
        grid = torch.tensor(grids)
        p1, d1 = get_rope_index(ids, grid, mask, cfg.image_token_id, cfg.vision_start_token_id, 2)
        p2, d2 = O.get_rope_index(cfg, ids, grid, mask)
        assert torch.equal(p1, p2) and torch.equal(d1, d2)
