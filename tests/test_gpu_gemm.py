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
