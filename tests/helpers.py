"""Shared helpers for the parity tests."""
import math
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden(name):
    return torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)


def bf16_ulp(x: float) -> float:
    """Spacing of bf16 numbers around |x| (8 significant bits)."""
    ax = abs(float(x))
    if ax == 0.0 or not math.isfinite(ax):
        return 2.0 ** -133
    return 2.0 ** (math.floor(math.log2(ax)) - 7)


def close_report(a: torch.Tensor, b: torch.Tensor):
    a32, b32 = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    diff = (a32 - b32).abs()
    scale = b32.abs().max().clamp_min(1e-12)
    cos = torch.nn.functional.cosine_similarity(a32, b32, dim=0).item() if a32.numel() > 1 else 1.0
    return dict(max_abs=float(diff.max()), rel_to_max=float(diff.max() / scale), mean_abs=float(diff.mean()), cos=cos,
                frac_exact=float((a32 == b32).float().mean()))


def assert_close_bf16(a, b, what, rel_to_max=0.03, cos=0.999):
    """bf16 pipelines agree to accumulation-order noise: every intermediate is rounded to 8 bits, so after a stack of
    layers elements differ by a few ulps of the *largest* magnitudes; we bound the worst element relative to max|ref|
    and require near-perfect direction agreement."""
    r = close_report(a, b)
    assert r["rel_to_max"] <= rel_to_max and r["cos"] >= cos, f"{what}: {r}"
    return r


def report(what: str, **kv):
    """Append one JSON line to the parity report (gpurun_out/parity_report.jsonl on the GPU box; copied to profiles/)."""
    import json
    path = os.environ.get("PADT_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as f:
            f.write(json.dumps(dict(what=what, **kv)) + "\n")
    except OSError:
        pass
    print("[parity]", what, kv)


def assert_box_like(a, b, what, ulps: float = 2.0, abs_tol: float = 1e-3, check: bool = True):
    """North-star float bar for boxes / scores / mask logits: |cand - ref| <= 1e-3, or <= `ulps` bf16 ulps of the reference
    value where bf16 itself is coarser than 1e-3 (the reference emits bf16: spacing at 0.5 is 3.9e-3, at 4.0 3.1e-2).
    Returns the report dict (max abs diff, how many elements exceed 1e-3, how many needed the ulp clause)."""
    a32, b32 = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    d = (a32 - b32).abs()
    ulp = torch.exp2(torch.floor(torch.log2(b32.abs().clamp_min(2.0 ** -126))) - 7)
    tol = torch.maximum(torch.full_like(d, abs_tol), ulps * ulp)
    bad = d > tol
    rep = dict(n=int(d.numel()), max_abs=float(d.max()), n_gt_1e3=int((d > abs_tol).sum()), n_bad=int(bad.sum()),
               ref_min=float(b32.min()), ref_max=float(b32.max()), worst_ulps=float((d / ulp).max()))
    report(what, **rep)
    if check:
        assert not bool(bad.any()), f"{what}: {rep}"
    return rep


def mask_iou(a, b):
    """IoU of the sigmoid > 0.5 (logit > 0) masks."""
    ma, mb = a.detach().float().cpu() > 0, b.detach().float().cpu() > 0
    inter, union = (ma & mb).sum().item(), (ma | mb).sum().item()
    return 1.0 if union == 0 else inter / union


def check_tokens_margin(cand: torch.Tensor, ref: torch.Tensor, top2: torch.Tensor, what: str, ulps: float = 1.0):
    """Greedy parity with a near-tie escape: a candidate arg-max may differ from the oracle's only where the oracle's own
    top-1/top-2 logit gap is <= `ulps` bf16 ulps (an exact or 1-ulp tie that accumulation order can flip); everywhere else the
    ids must be identical.  cand/ref [T,B] ints; top2 [T,B,2] oracle logits.  Returns (n_equal, n_tie_flips)."""
    cand, ref = cand.cpu(), ref.cpu()
    T, B = ref.shape
    equal = flips = 0
    for t in range(T):
        for b in range(B):
            if int(cand[t, b]) == int(ref[t, b]):
                equal += 1
                continue
            v1, v2 = float(top2[t, b, 0]), float(top2[t, b, 1])
            gap = v1 - v2
            assert gap <= ulps * bf16_ulp(v1), (f"{what}: step {t} row {b}: candidate {int(cand[t, b])} != oracle {int(ref[t, b])} "
                                                f"with oracle top-2 gap {gap} (> {ulps} bf16 ulps = {ulps * bf16_ulp(v1)})")
            flips += 1
    report(what + " tokens", steps=T * B, exact=equal, near_tie_flips=flips, escape_ulps=ulps)
    return equal, flips


def decoder_inputs(cfg, grid_hw, objects, iseed):
    """Seeded inputs of a decoder-only case (shared by the generator and the test: only the outputs are stored)."""
    g = torch.Generator().manual_seed(iseed)
    P, D, Hd = grid_hw[0] * grid_hw[1], cfg.text.hidden_size, cfg.decoder.hidden_size
    low = torch.randn((P // 4, D), generator=g).to(torch.bfloat16)
    high = torch.randn((P, Hd), generator=g).to(torch.bfloat16)
    theta = torch.rand((P, Hd // cfg.decoder.num_heads // 2), generator=g) * 6.0
    emb = torch.cat([theta, theta], dim=-1)
    feats = [[(2.0 * torch.randn((n, D), generator=g)).to(torch.bfloat16) for n in objects]]
    grid = torch.tensor([[1, grid_hw[0], grid_hw[1]]], dtype=torch.long)
    return feats, low, high, grid, (emb.cos().float(), emb.sin().float())
