"""Kernel-level parity (GPU): each CUDA kernel against a plain fp32 torch statement of the same op, through the C ABI."""
import math

import pytest
import torch

from helpers import assert_close_bf16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from padt_b200 import _capi
    return _capi.load()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _gemm(lib, epi, A, W, bias=None, resid=None, row_map=None, n_out=None, lda=None, ldw=None):
    from padt_b200 import _capi
    M, K = A.shape[0], (A.shape[1] if lda is None else lda[1])
    N = W.shape[0]
    n_out = N // 2 if epi == 3 else N
    C = torch.zeros((M, n_out), dtype=torch.bfloat16, device="cuda")
    _capi.check(lib.padt_op_gemm(epi, A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), None if bias is None else bias.data_ptr(),
                                 None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0), C.data_ptr(), C.stride(0),
                                 None if row_map is None else row_map.data_ptr(), M, N, K, _st()), "gemm")
    torch.cuda.synchronize()
    return C


def _ref_linear(A, W, bias):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias.float()
    return y.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(1, 8, 8), (37, 50, 72), (128, 128, 32), (200, 130, 1176), (513, 384, 160), (64, 2, 216)])
def test_gemm_bias(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + N)
    A = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    W = (0.05 * torch.randn((N, K), generator=g, device="cuda")).to(torch.bfloat16)
    b = torch.randn((N,), generator=g, device="cuda").to(torch.bfloat16)
    C = _gemm(lib, 0, A, W, b)
    ref = _ref_linear(A, W, b)
    r = assert_close_bf16(C, ref, f"gemm {M}x{N}x{K}", rel_to_max=0.01)
    assert r["frac_exact"] > 0.97, r


def test_gemm_k_tail_padded_ld(lib):
    # K = 3420-like: not a multiple of 8, operands stored with padded leading dimension (zero tail)
    M, N, K, Kp = 70, 48, 212, 216
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.zeros((M, Kp), device="cuda", dtype=torch.bfloat16)
    W = torch.zeros((N, Kp), device="cuda", dtype=torch.bfloat16)
    A[:, :K] = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    W[:, :K] = (0.1 * torch.randn((N, K), generator=g, device="cuda")).to(torch.bfloat16)
    A[:, K:] = 7.0   # garbage in the pad of A must be ignored because the kernel stops at K
    from padt_b200 import _capi
    C = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    _capi.check(lib.padt_op_gemm(0, A.data_ptr(), Kp, W.data_ptr(), Kp, None, None, 0, C.data_ptr(), N, None, M, N, K, _st()), "gemm")
    torch.cuda.synchronize()
    ref = (A[:, :K].float() @ W[:, :K].float().t()).to(torch.bfloat16)
    assert_close_bf16(C, ref, "gemm k tail", rel_to_max=0.01)


def test_gemm_epilogues(lib):
    M, N, K = 150, 96, 160
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    W = (0.1 * torch.randn((N, K), generator=g, device="cuda")).to(torch.bfloat16)
    b = torch.randn((N,), generator=g, device="cuda").to(torch.bfloat16)
    R = torch.randn((M, N), generator=g, device="cuda").to(torch.bfloat16)
    lin = _ref_linear(A, W, b)
    # GELU (erf) on the rounded linear output
    assert_close_bf16(_gemm(lib, 1, A, W, b), torch.nn.functional.gelu(lin.float()).to(torch.bfloat16), "gelu", rel_to_max=0.01)
    # residual: bf16(resid + bf16(linear))
    assert_close_bf16(_gemm(lib, 2, A, W, b, resid=R), (R.float() + lin.float()).to(torch.bfloat16), "resid", rel_to_max=0.01)
    # interleaved gate/up -> silu(g) * u
    gate, up = lin[:, 0::2].float(), lin[:, 1::2].float()
    ref = (torch.nn.functional.silu(gate).to(torch.bfloat16).float() * up).to(torch.bfloat16)
    assert_close_bf16(_gemm(lib, 3, A, W, b), ref, "swiglu", rel_to_max=0.01)
    # output row permutation
    perm = torch.randperm(M, device="cuda").to(torch.int32)
    out = _gemm(lib, 0, A, W, b, row_map=perm)
    ref = torch.zeros_like(lin)
    ref[perm.long()] = lin
    assert_close_bf16(out, ref, "row map", rel_to_max=0.01)


@pytest.mark.parametrize("M,N,K,epi", [
    (2350, 2048, 5504, 2),     # down_proj-like: 152 tiles on 148 SMs -> the four tiles past the full round are K-split five ways
    (2350, 2048, 11008, 2),    # llm down_proj at 3B: eight parts
    (2350, 2560, 3072, 0),     # 42 tiles past the round, three parts each, no residual
    (4608, 1280, 3424, 2),     # vit down-like: 180 tiles, 32 past the round, two parts (K padded to 3424, as the engine stores it)
    (2350, 4096, 3072, 3),     # SwiGLU epilogue after the reduction
    (2350, 2048, 2048, 2),     # K too short to split (stays whole-tile): same shape class without the split
])
def test_gemm_split_k_round(lib, M, N, K, epi):
    """The persistent tcgen05 GEMM K-splits the tiles of a short last round (gemm_tc.cu TcSplit); fp32 partials are summed in part order,
    then the usual epilogue with the HF rounding points runs.  Two launches must agree bit for bit (deterministic reduction)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    W = (0.03 * torch.randn((N, K), generator=g, device="cuda")).to(torch.bfloat16)
    b = torch.randn((N,), generator=g, device="cuda").to(torch.bfloat16)
    n_out = N // 2 if epi == 3 else N
    R = torch.randn((M, n_out), generator=g, device="cuda").to(torch.bfloat16) if epi == 2 else None
    lin = _ref_linear(A, W, b)
    if epi == 2:
        ref = (R.float() + lin.float()).to(torch.bfloat16)
    elif epi == 3:
        ref = (torch.nn.functional.silu(lin[:, 0::2].float()).to(torch.bfloat16).float() * lin[:, 1::2].float()).to(torch.bfloat16)
    else:
        ref = lin
    C1 = _gemm(lib, epi, A, W, b, resid=R)
    C2 = _gemm(lib, epi, A, W, b, resid=R)
    assert torch.equal(C1, C2)
    r = assert_close_bf16(C1, ref, f"split-k gemm {M}x{N}x{K} epi {epi}", rel_to_max=0.01)
    assert r["frac_exact"] > 0.95, r


def test_rmsnorm(lib):
    from padt_b200 import _capi
    import padt_oracle as O
    for rows, dim in [(5, 160), (33, 256), (9, 2048)]:
        g = torch.Generator(device="cuda").manual_seed(rows)
        x = torch.randn((rows, dim), generator=g, device="cuda").to(torch.bfloat16)
        w = (1 + 0.1 * torch.randn((dim,), generator=g, device="cuda")).to(torch.bfloat16)
        y = torch.empty_like(x)
        _capi.check(lib.padt_op_rmsnorm(x.data_ptr(), w.data_ptr(), y.data_ptr(), rows, dim, 1e-6, _st()), "rmsnorm")
        torch.cuda.synchronize()
        ref = O.rmsnorm(x.cpu(), w.cpu(), 1e-6)
        r = assert_close_bf16(y, ref, "rmsnorm", rel_to_max=0.01)
        assert r["frac_exact"] > 0.99, r


def _attention(lib, q, k, v, segs_q, segs_k, group, causal, k_start=None):
    """q [Tq,H,hd], k/v [Tk,Hkv,hd] packed varlen; segs = list of (off,len)."""
    from padt_b200 import _capi
    H, hd = q.shape[1], q.shape[2]
    Hkv = k.shape[1]
    o = torch.zeros_like(q)
    i32 = lambda xs: torch.tensor(xs, dtype=torch.int32, device="cuda")
    qo, ql = i32([s[0] for s in segs_q]), i32([s[1] for s in segs_q])
    ko, kl = i32([s[0] for s in segs_k]), i32([s[1] for s in segs_k])
    ks = None if k_start is None else i32(k_start)
    _capi.check(lib.padt_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), H * hd, hd, Hkv * hd, hd, Hkv * hd, hd, H * hd, hd,
                                      qo.data_ptr(), ql.data_ptr(), ko.data_ptr(), kl.data_ptr(), None if ks is None else ks.data_ptr(),
                                      len(segs_q), H, group, max(s[1] for s in segs_q), hd, int(causal), 1.0 / math.sqrt(hd), _st()), "attention")
    torch.cuda.synchronize()
    return o


def _ref_attention(q, k, v, segs_q, segs_k, causal, k_start=None):
    import padt_oracle as O
    q, k, v = q.cpu(), k.cpu(), v.cpu()
    out = torch.zeros_like(q)
    scale = 1.0 / math.sqrt(q.shape[-1])
    for i, ((qo, ql), (ko, kl)) in enumerate(zip(segs_q, segs_k)):
        mask = None
        if causal or k_start is not None:
            ii = torch.arange(ql)[:, None]
            jj = torch.arange(kl)[None, :]
            mask = torch.ones((ql, kl), dtype=torch.bool)
            if causal:
                mask &= jj <= ii + (kl - ql)
            if k_start is not None:
                mask &= jj >= k_start[i]
        out[qo:qo + ql] = O.fa2_attention(q[qo:qo + ql].transpose(0, 1), k[ko:ko + kl].transpose(0, 1), v[ko:ko + kl].transpose(0, 1),
                                          scale, mask).transpose(0, 1)
    return out


@pytest.mark.parametrize("hd,H,Hkv", [(80, 2, 2), (128, 4, 1), (128, 8, 2)])
def test_attention_varlen(lib, hd, H, Hkv):
    g = torch.Generator(device="cuda").manual_seed(hd + H)
    lens = [64, 36, 48, 130, 1, 200]
    segs, off = [], 0
    for n in lens:
        segs.append((off, n)); off += n
    q = torch.randn((off, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn((off, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn((off, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    o = _attention(lib, q, k, v, segs, segs, H // Hkv, False)
    assert_close_bf16(o, _ref_attention(q, k, v, segs, segs, False), "varlen attention", rel_to_max=0.02)


def test_attention_causal_leftpad_gqa(lib):
    hd, H, Hkv = 128, 4, 2
    g = torch.Generator(device="cuda").manual_seed(11)
    lens = [150, 97]
    starts = [0, 23]
    segs, off = [], 0
    for n in lens:
        segs.append((off, n)); off += n
    q = torch.randn((off, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn((off, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn((off, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    o = _attention(lib, q, k, v, segs, segs, H // Hkv, True, k_start=starts)
    ref = _ref_attention(q, k, v, segs, segs, True, k_start=starts)
    assert_close_bf16(o, ref, "causal gqa attention", rel_to_max=0.02)
    # rows that are left pads (no visible key) are exactly zero, as FA2 never computes them
    assert float(o[segs[1][0]: segs[1][0] + starts[1]].abs().max()) == 0.0


def test_attention_cross_asymmetric(lib):
    # PaDT decoder shapes: 8 queries against thousands of keys and the reverse
    hd, H = 80, 2
    g = torch.Generator(device="cuda").manual_seed(13)
    q_lens, k_lens = [8, 5], [2304, 600]
    sq, sk, oq, ok = [], [], 0, 0
    for a, b in zip(q_lens, k_lens):
        sq.append((oq, a)); sk.append((ok, b)); oq += a; ok += b
    q = torch.randn((oq, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn((ok, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn((ok, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    assert_close_bf16(_attention(lib, q, k, v, sq, sk, 1, False), _ref_attention(q, k, v, sq, sk, False), "q->img", rel_to_max=0.02)
    q2 = torch.randn((ok, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    k2 = torch.randn((oq, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    v2 = torch.randn((oq, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    assert_close_bf16(_attention(lib, q2, k2, v2, sk, sq, 1, False), _ref_attention(q2, k2, v2, sk, sq, False), "img->q", rel_to_max=0.02)


@pytest.mark.parametrize("hd,H,Hkv,causal,lens,starts", [
    (80, 2, 2, False, [1500, 640], None),              # ViT full-attention shape (two images), ragged last tiles
    (128, 4, 2, True, [700, 515], [0, 37]),            # causal GQA prefill with left pads
    (128, 2, 1, True, [1024], None),                   # exact multiple of the tile
    (80, 1, 1, False, [513, 512, 600, 129, 1], None),  # mixed: short segments ride along in a long-segment launch
])
def test_attention_long_segments_tensor_path(lib, hd, H, Hkv, causal, lens, starts):
    """Segments >= 512 rows take the tcgen05 + TMA kernel (attention_tc.cu); same FA2 numerics as the mma.sync kernel."""
    g = torch.Generator(device="cuda").manual_seed(hd + sum(lens))
    segs, off = [], 0
    for n in lens:
        segs.append((off, n)); off += n
    q = torch.randn((off, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn((off, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn((off, Hkv, hd), generator=g, device="cuda").to(torch.bfloat16)
    o = _attention(lib, q, k, v, segs, segs, H // Hkv, causal, k_start=starts)
    ref = _ref_attention(q, k, v, segs, segs, causal, k_start=starts)
    r = assert_close_bf16(o, ref, "long-segment attention", rel_to_max=0.02)
    assert r["frac_exact"] > 0.5, r
    if starts:
        assert float(o[segs[1][0]: segs[1][0] + starts[1]].abs().max()) == 0.0


def test_attention_kv_cache_layout_never_reads_unwritten_slots(lib):
    """Prefill layout (engine_llm.cu): K/V live in the cache [row][kv head][s_max][d]; the slots past the prompt are not written yet.
    Fill them with NaN: one 0 * NaN in P @ V would poison the output."""
    from padt_b200 import _capi
    hd, H, Hkv, B, L, s_max = 128, 4, 2, 2, 600, 704
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn((B * L, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    kc = torch.full((B, Hkv, s_max, hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    vc = torch.full((B, Hkv, s_max, hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    kc[:, :, :L] = torch.randn((B, Hkv, L, hd), generator=g, device="cuda").to(torch.bfloat16)
    vc[:, :, :L] = torch.randn((B, Hkv, L, hd), generator=g, device="cuda").to(torch.bfloat16)
    o = torch.zeros_like(q)
    i32 = lambda xs: torch.tensor(xs, dtype=torch.int32, device="cuda")
    qo, ql = i32([b * L for b in range(B)]), i32([L] * B)
    ko, kl = i32([b * Hkv * s_max for b in range(B)]), i32([L] * B)
    ks = i32([0, 11])
    _capi.check(lib.padt_op_attention(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), o.data_ptr(), H * hd, hd, hd, s_max * hd, hd, s_max * hd, H * hd, hd,
                                      qo.data_ptr(), ql.data_ptr(), ko.data_ptr(), kl.data_ptr(), ks.data_ptr(), B, H, H // Hkv, L, hd, 1,
                                      1.0 / math.sqrt(hd), _st()), "attention")
    torch.cuda.synchronize()
    assert bool(torch.isfinite(o.float()).all())
    k = kc[:, :, :L].permute(0, 2, 1, 3).reshape(B * L, Hkv, hd).contiguous()
    v = vc[:, :, :L].permute(0, 2, 1, 3).reshape(B * L, Hkv, hd).contiguous()
    segs = [(b * L, L) for b in range(B)]
    ref = _ref_attention(q, k, v, segs, segs, True, k_start=[0, 11])
    assert_close_bf16(o, ref, "kv-cache layout attention", rel_to_max=0.02)


def test_rope_neox_bit_exact(lib):
    from padt_b200 import _capi
    import padt_oracle as O
    rows, H, hd = 50, 3, 80
    g = torch.Generator().manual_seed(17)
    x = torch.randn((rows, H, hd), generator=g).to(torch.bfloat16)
    ang = torch.rand((rows, hd // 2), generator=g) * 6.0
    cos = torch.cat([ang.cos(), ang.cos()], -1).contiguous()
    sin = torch.cat([ang.sin(), ang.sin()], -1).contiguous()
    ref = O.apply_rotary_emb_fa(x[None].float(), cos[:, : hd // 2], sin[:, : hd // 2]).to(torch.bfloat16)[0]
    xd, cd, sd = x.cuda().contiguous(), cos.cuda(), sin.cuda()
    _capi.check(lib.padt_op_rope_neox(xd.data_ptr(), H * hd, H, hd, cd.data_ptr(), sd.data_ptr(), hd, rows, _st()), "rope")
    torch.cuda.synchronize()
    assert torch.equal(xd.cpu(), ref), "fp32 NeoX rotary must be bit-exact (separate mul/mul/add roundings, no FMA)"


@pytest.mark.parametrize("M,N,K,Kp", [(8, 1280, 1280, 1280), (5, 160, 256, 256), (13, 50, 212, 216), (40, 96, 3420, 3424), (1, 4, 1280, 1280)])
def test_gemm_skinny_path_epilogues(lib, M, N, K, Kp):
    """M <= 64 rows take the skinny (warp-per-two-weight-rows) path: bias / GELU / residual epilogues with the same rounding points as the
    tiled kernel, K tails with padded leading dimensions, in-place residual (C == resid) as the decoder uses it."""
    from padt_b200 import _capi
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.full((M, Kp), 7.0, device="cuda", dtype=torch.bfloat16)            # garbage in the pad of A must be ignored
    W = torch.zeros((N, Kp), device="cuda", dtype=torch.bfloat16)
    A[:, :K] = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    W[:, :K] = (0.05 * torch.randn((N, K), generator=g, device="cuda")).to(torch.bfloat16)
    b = torch.randn((N,), generator=g, device="cuda").to(torch.bfloat16)
    R = torch.randn((M, N), generator=g, device="cuda").to(torch.bfloat16)
    lin = (A[:, :K].float() @ W[:, :K].float().t() + b.float()).to(torch.bfloat16)

    def run(epi, resid=None, out=None):
        C = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda") if out is None else out
        _capi.check(lib.padt_op_gemm(epi, A.data_ptr(), Kp, W.data_ptr(), Kp, b.data_ptr(), None if resid is None else resid.data_ptr(),
                                     0 if resid is None else N, C.data_ptr(), N, None, M, N, K, _st()), "gemm")
        torch.cuda.synchronize()
        return C

    assert_close_bf16(run(0), lin, "skinny bias", rel_to_max=0.01)
    assert_close_bf16(run(1), torch.nn.functional.gelu(lin.float()).to(torch.bfloat16), "skinny gelu", rel_to_max=0.01)
    assert_close_bf16(run(2, resid=R), (R.float() + lin.float()).to(torch.bfloat16), "skinny resid", rel_to_max=0.01)
    inplace = R.clone()
    assert_close_bf16(run(2, resid=inplace, out=inplace), (R.float() + lin.float()).to(torch.bfloat16), "skinny resid in place", rel_to_max=0.01)


@pytest.mark.parametrize("n_split,lens", [(4, [(8, 2304)]), (16, [(8, 9216)]), (7, [(5, 700), (8, 64), (3, 1500)])])
def test_attention_split_kv(lib, n_split, lens):
    """Small-q / large-kv cross attention with the key range split across CTAs + merge: same result as the single-pass kernel and as the
    oracle's FA2 restatement (head_dim 80, 2 heads), including segments whose key range has fewer tiles than parts."""
    from padt_b200 import _capi
    H, hd = 2, 80
    g = torch.Generator(device="cuda").manual_seed(n_split)
    segs_q, segs_k, qo, ko = [], [], 0, 0
    for ql, kl in lens:
        segs_q.append((qo, ql)); segs_k.append((ko, kl)); qo += ql; ko += kl
    q = torch.randn((qo, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn((ko, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn((ko, H, hd), generator=g, device="cuda").to(torch.bfloat16)
    single = _attention(lib, q, k, v, segs_q, segs_k, 1, False)
    o = torch.zeros_like(q)
    i32 = lambda xs: torch.tensor(xs, dtype=torch.int32, device="cuda")
    qoff, qlen, koff, klen = i32([s[0] for s in segs_q]), i32([s[1] for s in segs_q]), i32([s[0] for s in segs_k]), i32([s[1] for s in segs_k])
    n_seg = len(lens)
    part_o = torch.empty((n_seg * H * n_split * 64 * hd,), dtype=torch.float32, device="cuda")
    part_ml = torch.empty((n_seg * H * n_split * 64 * 2,), dtype=torch.float32, device="cuda")
    _capi.check(lib.padt_op_attention_split(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), H * hd, hd, H * hd, hd, H * hd, hd, H * hd, hd,
                                            qoff.data_ptr(), qlen.data_ptr(), koff.data_ptr(), klen.data_ptr(), n_seg, H, max(s[1] for s in segs_q), hd,
                                            1.0 / math.sqrt(hd), n_split, part_o.data_ptr(), part_ml.data_ptr(), _st()), "attention_split")
    torch.cuda.synchronize()
    assert_close_bf16(o, single, "split-KV vs single pass", rel_to_max=0.01, cos=0.9999)
    assert_close_bf16(o, _ref_attention(q, k, v, segs_q, segs_k, False), "split-KV vs FA2 restatement", rel_to_max=0.02)
