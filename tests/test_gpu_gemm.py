"""GPU: the tcgen05 / TMEM / TMA GEMM building block (ofx_gemm_bf16) against torch fp32 matmul
of the same bf16-rounded operands, at the encoder's layer shapes."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(a, w, bias=None, mish=False, residual=None, out_f32=True):
    from outfitx_b200 import _lib
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty(m, n, dtype=torch.float32 if out_f32 else torch.bfloat16, device=a.device)
    _lib.check(_lib.lib().ofx_gemm_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), m, n, k,
        bias.data_ptr() if bias is not None else None, int(mish),
        residual.data_ptr() if residual is not None else None,
        residual.stride(0) if residual is not None else 0,
        out.data_ptr(), out.stride(0), int(out_f32), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out


# (M, N, K): QKV / out-proj / FFN1 / FFN2 at d_model 512 and 1024, ragged M, tiny M
SHAPES = [(17 * 64, 1536, 512), (1000, 512, 512), (300, 2048, 512), (129, 512, 2048),
          (17 * 40, 3072, 1024), (5, 1024, 1024), (4096, 2048, 1024), (777, 1024, 2048),
          (128 * 200 + 3, 256, 64)]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_gemm_matches_fp32_matmul(m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).to(torch.bfloat16)
    want = a.float() @ w.float().T
    got = _gemm(a, w)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_gemm_epilogue_bias_mish_residual_bf16_out():
    m, n, k = 1234, 2048, 512
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda", generator=g)
    res = torch.randn(m, n, device="cuda", generator=g)
    lin = a.float() @ w.float().T + bias
    torch.testing.assert_close(_gemm(a, w, bias), lin, rtol=1e-4, atol=1e-4)
    mish = torch.nn.functional.mish(lin)
    torch.testing.assert_close(_gemm(a, w, bias, mish=True), mish, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(_gemm(a, w, bias, residual=res), lin + res, rtol=1e-4, atol=1e-4)
    out_bf = _gemm(a, w, bias, mish=True, out_f32=False)
    torch.testing.assert_close(out_bf.float(), mish.to(torch.bfloat16).float(), rtol=2e-2, atol=2e-2)
    # residual may alias the output (the encoder updates the residual stream in place)
    x = res.clone()
    from outfitx_b200 import _lib
    _lib.check(_lib.lib().ofx_gemm_bf16(a.data_ptr(), k, w.data_ptr(), k, m, n, k, bias.data_ptr(), 0,
                                        x.data_ptr(), n, x.data_ptr(), n, 1,
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    torch.testing.assert_close(x, lin + res, rtol=1e-4, atol=1e-4)


def test_gemm_rejects_bad_shapes():
    from outfitx_b200 import _lib
    a = torch.zeros(128, 96, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(128, 96, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.OfxError):
        _gemm(a, w)  # K % 64 != 0


@pytest.mark.parametrize("m,n,k", [(17 * 64, 1536, 512), (1000, 512, 2048), (5, 1024, 1024), (4099, 2048, 1024), (300, 4608, 1536)])
def test_fp32_operands_on_the_tensor_cores(m, n, k):
    """ofx_gemm_f32_tc: fp32 A and W as bf16 hi / lo pieces, three products in one bf16 GEMM with K' = 3K.  Against
    an fp64 matmul the error must be that of 16-bit mantissas (a few 1e-5 of the row scale), two orders below a
    plain bf16 GEMM -- and bias / mish / in-place residual epilogues must work on it."""
    from outfitx_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    bias = torch.randn(n, device="cuda", generator=g)
    ws = torch.empty(L.ofx_gemm_f32_tc_workspace_bytes(m, n, k), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run(act, res):
        out = torch.empty(m, n, device="cuda") if res is None else res
        _lib.check(L.ofx_gemm_f32_tc(a.data_ptr(), k, w.data_ptr(), k, m, n, k, bias.data_ptr(), act,
                                     res.data_ptr() if res is not None else None, n, out.data_ptr(), n,
                                     ws.data_ptr(), ws.numel(), st))
        torch.cuda.synchronize()
        return out

    lin = (a.double() @ w.double().T + bias.double())
    got = run(0, None)
    err = (got.double() - lin).abs().max().item()
    assert err <= 2e-4, err                                           # ~5 sigma of sqrt(K) products of ~2^-17 each
    bf = (a.to(torch.bfloat16).double() @ w.to(torch.bfloat16).double().T + bias.double())
    assert err * 30 < (bf - lin).abs().max().item()                    # the plain bf16 GEMM is >= 30x worse
    torch.testing.assert_close(run(1, None).double(), torch.nn.functional.mish(lin), rtol=1e-4, atol=2e-4)
    x = torch.randn(m, n, device="cuda", generator=g)
    want = x.double() + lin
    torch.testing.assert_close(run(0, x).double(), want, rtol=1e-4, atol=2e-4)
    assert L.ofx_gemm_f32_tc(a.data_ptr(), k, w.data_ptr(), k, m, n, k, None, 0, None, 0, got.data_ptr(), n,
                             ws.data_ptr(), 16, st) == -5
