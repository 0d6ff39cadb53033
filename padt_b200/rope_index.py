"""M-RoPE position ids for image+text prompts (host integer work, once per request).

The reference calls ``self.get_rope_index`` of **transformers 4.50.0** at ``src/PaDT/models/padt.py:263`` and turns
the returned deltas into decode positions at ``padt.py:268-277``.  Semantics:

* text tokens: t = h = w = running index;
* the k-th ``<|image_pad|>`` run of a row (grid t x h x w, merged by ``spatial_merge_size``): all three planes start
  at ``st = previous max + 1``; t stays at ``st`` (images have second_per_grid_t = 0), h = st + merged row,
  w = st + merged column;
* the next text token resumes at ``max + 1`` over all planes;
* padded slots (attention_mask == 0) hold 1;
* ``rope_delta = max + 1 - padded_len`` so that ``cache_position + delta`` is the next position.

(The 5.5.0 function installed in this image scales the temporal index by tokens_per_second and uses another delta
convention; it must not be used.)
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def get_rope_index(input_ids: torch.Tensor, image_grid_thw: Optional[torch.Tensor], attention_mask: Optional[torch.Tensor],
                   image_token_id: int, vision_start_token_id: int, spatial_merge_size: int = 2
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (position_ids int64 [3,B,L], rope_deltas int64 [B,1]); inputs are CPU tensors."""
    ids = input_ids.detach().cpu().numpy()
    B, L = ids.shape
    mask = np.ones_like(ids) if attention_mask is None else attention_mask.detach().cpu().numpy()
    grids = None if image_grid_thw is None else image_grid_thw.detach().cpu().numpy().astype(np.int64)
    pos = np.ones((3, B, L), dtype=np.int64)
    deltas = np.zeros((B, 1), dtype=np.int64)
    next_image = 0
    for b in range(B):
        cols = np.nonzero(mask[b] == 1)[0]
        row = ids[b, cols]
        n = row.shape[0]
        out = np.zeros((3, n), dtype=np.int64)
        # an image is announced by <|vision_start|> immediately followed by <|image_pad|>
        starts = np.nonzero(row[:-1] == vision_start_token_id)[0] if n > 1 else np.zeros((0,), dtype=np.int64)
        n_images = int(np.sum(row[starts + 1] == image_token_id)) if starts.size else 0
        cursor = 0        # next unprocessed token
        nxt = 0           # next free position value
        for _ in range(n_images):
            rest = np.nonzero(row[cursor:] == image_token_id)[0]
            first = cursor + int(rest[0])
            t, h, w = (int(x) for x in grids[next_image])
            next_image += 1
            gh, gw = h // spatial_merge_size, w // spatial_merge_size
            k = first - cursor
            out[:, cursor:first] = nxt + np.arange(k)[None, :]
            base = nxt + k
            count = t * gh * gw
            tt, yy, xx = np.meshgrid(np.zeros(t, dtype=np.int64), np.arange(gh), np.arange(gw), indexing="ij")
            out[0, first:first + count] = base + tt.reshape(-1)
            out[1, first:first + count] = base + yy.reshape(-1)
            out[2, first:first + count] = base + xx.reshape(-1)
            cursor = first + count
            nxt = int(out[:, first:cursor].max()) + 1 if count > 0 else base
        if cursor < n:
            out[:, cursor:] = nxt + np.arange(n - cursor)[None, :]
        pos[:, b, cols] = out
        deltas[b, 0] = (int(out.max()) + 1 if n > 0 else 0) - L
    return torch.from_numpy(pos), torch.from_numpy(deltas)
