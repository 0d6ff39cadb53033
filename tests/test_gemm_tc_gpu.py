"""tcgen05 + TMA GEMM (csrc/gemm_tc.cu) against an fp32 torch statement and against the mma.sync kernel, through the C ABI."""
import pytest
import torch

from helpers import assert_close_bf16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from padt_b200 import _capi
    return _capi.load()


def _run(lib, epi, A, W, bias, resid, row_map, M, N, K, lda, ldw):
    from padt_b200 import _capi
    n_out = N // 2 if epi == 3 else N
    C = torch.zeros((M, n_out), dtype=torch.bfloat16, device="cuda")
    _capi.check(lib.padt_op_gemm(epi, A.data_ptr(), lda, W.data_ptr(), ldw, None if bias is None else bias.data_ptr(),
                                 None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0), C.data_ptr(), C.stride(0),
                                 None if row_map is None else row_map.data_ptr(), M, N, K, torch.cuda.current_stream().cuda_stream), "gemm")
    torch.cuda.synchronize()
    return C


def _ref(epi, A, W, bias, resid, row_map):
    lin = (A.float() @ W.float().t() + bias.float()).to(torch.bfloat16)
    if epi == 1:
        out = torch.nn.functional.gelu(lin.float()).to(torch.bfloat16)
    elif epi == 2:
        out = (resid.float() + lin.float()).to(torch.bfloat16)
    elif epi == 3:
        out = (torch.nn.functional.silu(lin[:, 0::2].float()).to(torch.bfloat16).float() * lin[:, 1::2].float()).to(torch.bfloat16)
    else:
        out = lin
    if row_map is not None:
        o2 = torch.zeros_like(out)
        o2[row_map.long()] = out
        out = o2
    return out


@pytest.mark.parametrize("M,N,K,Kpad", [(256, 256, 64, 64), (300, 200, 160, 160), (1024, 1280, 1176, 1176), (640, 512, 212, 216), (2350, 2560, 2048, 2048),
                                        (512, 3840, 1280, 1280)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_tc_matches_reference_and_mma_sync(lib, M, N, K, Kpad, epi):
    from padt_b200 import _capi
    g = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    A = torch.zeros((M, Kpad), device="cuda", dtype=torch.bfloat16)
    W = torch.zeros((N, Kpad), device="cuda", dtype=torch.bfloat16)
    A[:, :K] = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    W[:, :K] = (0.05 * torch.randn((N, K), generator=g, device="cuda")).to(torch.bfloat16)
    if Kpad > K:
        A[:, K:] = 3.0      # must be ignored: the tensor map stops at K
    bias = torch.randn((N,), generator=g, device="cuda").to(torch.bfloat16)
    n_out = N // 2 if epi == 3 else N
    resid = torch.randn((M, n_out), generator=g, device="cuda").to(torch.bfloat16) if epi == 2 else None
    row_map = torch.randperm(M, device="cuda").to(torch.int32) if (M % 2 == 0 and epi in (0, 3)) else None
    ref = _ref(epi, A[:, :K], W[:, :K], bias, resid, row_map)
    # a dummy engine-independent option switch: the global flag lives in the library
    lib.padt_set_option.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_char_p, __import__("ctypes").c_int]
    out_tc = _run(lib, epi, A, W, bias, resid, row_map, M, N, K, Kpad, Kpad)
    r = assert_close_bf16(out_tc, ref, f"tcgen05 gemm {M}x{N}x{K} epi {epi}", rel_to_max=0.01)
    assert r["frac_exact"] > 0.95, r
