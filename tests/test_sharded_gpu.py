"""Sample-sharded path on 2 GPUs (NCCL): every rank's generate + vl_decode must equal a single-GPU run of the same sample, and the
all-gathered table / global id view must equal the single-process batch view of the reference (padt_processor.py:36-42,
padt.py:350-382, utils.py:181-182).  Skips on a 1-GPU box; the driver's multi-GPU tiers and `gpurun --gpus 2` run it."""
import os
import socket

import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from padt_b200 import VisonTextProcessingClass
    from padt_b200.config import padt_tiny
    from padt_b200.distributed import shard_samples
    from padt_b200.modeling import PaDTForConditionalGeneration
    from padt_b200.weights import random_weights
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        fx = load_golden("tiny_b2")                       # two samples with different grids, left padded
        cfg = padt_tiny(tie=fx["tie"], vocab=fx["vocab"])
        V = cfg.text.vocab_size
        w = {k: v.to(dev) for k, v in random_weights(cfg, seed=fx["wseed"]).items()}
        model = PaDTForConditionalGeneration(cfg, w)
        mine = shard_samples(2, rank, world)
        assert mine == [rank]
        b = rank
        grid_all = fx["grid"]
        P_off = [0, int(grid_all[0].prod())]
        P_b = int(grid_all[b].prod())
        M_off = [0, int(grid_all[0].prod()) // 4]
        pv = fx["pixel_values"][P_off[b]: P_off[b] + P_b].to(dev)
        real = fx["attention_mask"][b].bool()
        ids = fx["input_ids"][b][real][None].clone()      # un-padded local row, local VRT id space
        forced_g = fx["forced"][b][None].clone()          # fixture ids are batch-global: V + M_off[b] + k
        forced = torch.where(forced_g >= V, forced_g - M_off[b], forced_g)
        kw = dict(input_ids=ids.to(dev), attention_mask=torch.ones_like(ids).to(dev), pixel_values=pv, image_grid_thw=grid_all[b][None].to(dev),
                  max_new_tokens=fx["steps"], output_hidden_states=True, return_dict_in_generate=True, eos_token_id=[], forced_tokens=forced.to(dev))
        single = model.generate(**kw)
        shard = model.generate(**kw, prototype_group=dist.group.WORLD, global_image_grid_thw=grid_all)
        # rank-local outputs == single-GPU run of the same sample, bit for bit
        assert torch.equal(shard.sequences, single.sequences) and torch.equal(shard.argmax_tokens, single.argmax_tokens)
        assert torch.equal(shard.past_image_embeds, single.past_image_embeds) and torch.equal(shard.past_logit_mask, single.past_logit_mask)
        # the documented follow-up flow with the LOCAL grid works on every rank
        proc_grid = grid_all[b][None]
        objs = fx["objects"][b]
        feats = [[torch.cat([shard.hidden_states[s][-1][0] for s in range(s0, s0 + n)], dim=0) for (s0, n) in objs]]
        d_sh = model.vl_decode(feats, shard.past_image_embeds, shard.past_high_res_image_embeds, proc_grid.to(dev), shard.past_visual_pe)
        feats1 = [[torch.cat([single.hidden_states[s][-1][0] for s in range(s0, s0 + n)], dim=0) for (s0, n) in objs]]
        d_1 = model.vl_decode(feats1, single.past_image_embeds, single.past_high_res_image_embeds, proc_grid.to(dev), single.past_visual_pe)
        for k in ("pred_boxes", "pred_score", "pred_mask"):
            assert torch.equal(d_sh[k], d_1[k]), k
        # global view: all-gathered table in rank order + ids shifted by this rank's row offset == the single-process batch view
        table, base, seq_g = shard.global_view()
        torch.cuda.synchronize()
        assert base == M_off[b] and table.shape[0] == int(grid_all.prod(dim=1).sum()) // 4
        assert torch.equal(table[base: base + single.past_image_embeds.shape[0]], single.past_image_embeds)
        L = ids.shape[1]
        assert torch.equal(seq_g[0, L:].cpu(), forced_g[0]), "global view must reproduce the batch-global ids of the fixture"
        # the eager exchange (all-gather on a side stream during generate) gives the same view
        eager = model.generate(**kw, prototype_group=dist.group.WORLD, global_image_grid_thw=grid_all, prototype_exchange="eager")
        table_e, base_e, seq_e = eager.global_view()
        torch.cuda.synchronize()
        assert base_e == base and torch.equal(table_e, table) and torch.equal(seq_e, seq_g) and torch.equal(eager.sequences, single.sequences)
        q.put((rank, "ok", table.float().sum().item()))
    except Exception as e:     # surface the failure in the parent
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_generate_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
    for rank, status, _ in res:
        assert status == "ok", f"rank {rank}: {status}"
    assert res[0][2] == res[1][2], "both ranks must hold the same global table"
