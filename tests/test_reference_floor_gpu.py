"""The noise floor of the parity comparison, measured with the REFERENCE ITSELF: the unmodified reference classes (oracle/_ref staged copy,
compat shim) run on the GPU with the real flash-attn kernels and cuBLAS, against the fixture the same code produced on the CPU
(oneDNN + the restated FA2 numerics).  Both sides are "the reference"; what differs is only the fp32 summation order of the GEMMs and
attention kernels -- exactly what differs between the reference and padt_b200.  The numbers (hidden-state error, logit error on the
top-8 ids, how many greedy ids flip) are what a correct bf16 implementation on another backend can be expected to show, so they are the
yardstick for tests/test_model_gpu.py::test_full_size_3b_config1_against_reference_golden.  Report + sanity asserts; skips when the
reference files are not staged (oracle/stage_ref.py) or flash-attn is missing."""
import os

import pytest
import torch

from helpers import ROOT, bf16_ulp, close_report, load_golden, report

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_reference_gpu_vs_reference_cpu_full3b_c1():
    if not os.path.exists(os.path.join(ROOT, "tests", "golden", "full3b_c1.pt")):
        pytest.skip("tests/golden/full3b_c1.pt not generated")
    try:
        import flash_attn  # noqa: F401
        import ref_runner as R
        import ref_shim
        ref_shim._ref_src()
    except Exception as e:
        pytest.skip(f"reference not runnable here: {e}")
    from padt_b200.config import padt_pro_3b
    from padt_b200.weights import random_weights
    fx = load_golden("full3b_c1")
    cfg = padt_pro_3b()
    w = random_weights(cfg, seed=fx["wseed"])                      # the fixture's weights (CPU generator)
    model, _ = R.build_reference(cfg, w, device="cuda", attn="flash_attention_2", fast_init=True)
    del w
    steps = fx["steps"]
    dev = torch.device("cuda")
    with torch.no_grad():
        out = R.ref_greedy(model, cfg, fx["input_ids"].to(dev), fx["attention_mask"].to(dev), fx["pixel_values"].to(dev), fx["grid"].to(dev), steps,
                           forced=fx["forced"].to(dev))
    torch.cuda.synchronize()
    rep = dict(prototypes=close_report(out["past_image_embeds"], fx["prototypes"]),
               high_res=close_report(out["past_high_res_image_embeds"], fx["high_res"]))
    exact = flips = 0
    worst_rel = worst_err = worst_gap = 0.0
    for t in range(steps):
        h = out["hidden_steps"][t][:, -1, :].float().cpu()
        r = close_report(h, fx["forced_last_hidden"][t])
        worst_rel = max(worst_rel, r["rel_to_max"])
        idx = fx["forced_top8_indices"][t, 0]
        logits = out["logits_steps"][t][0].cpu()
        err = float((logits[idx] - fx["forced_top8_values"][t, 0]).abs().max())
        worst_err = max(worst_err, err)
        pick, want = int(out["argmax_steps"][t][0]), int(fx["forced_argmax"][t, 0])
        if pick == want:
            exact += 1
        else:
            flips += 1
            pos = (idx == pick).nonzero()
            gap = float(fx["forced_top8_values"][t, 0, 0] - fx["forced_top8_values"][t, 0, int(pos[0])]) if pos.numel() else float("inf")
            worst_gap = max(worst_gap, gap)
    report("FLOOR full3b_c1: reference on GPU (FA2 + cuBLAS) vs reference on CPU, teacher-forced", steps=steps, exact=exact, flips=flips,
           worst_hidden_rel_to_max=worst_rel, worst_logit_err_top8=worst_err, worst_flip_gap=worst_gap, **rep)
    # sanity only: the two backends of the SAME code agree to bf16 noise
    assert worst_rel <= 0.15 and rep["high_res"]["cos"] >= 0.995
