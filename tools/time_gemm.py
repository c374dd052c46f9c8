"""Times ofx_gemm_bf16 at the encoder's layer shapes (d_model 512, ~82k token rows)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from outfitx_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
M = int(sys.argv[1]) if len(sys.argv) > 1 else 82000
for name, n, k, out_f32, res, mish in [("qkv", 1536, 512, 0, 0, 0), ("outproj", 512, 512, 1, 1, 0),
                                       ("ffn1", 2048, 512, 0, 0, 1), ("ffn2", 512, 2048, 1, 1, 0)]:
    a = torch.randn(M, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    out = torch.zeros(M, n, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    def run():
        _lib.check(L.ofx_gemm_bf16(a.data_ptr(), k, w.data_ptr(), k, M, n, k, bias.data_ptr(), mish,
                                   out.data_ptr() if res else None, n, out.data_ptr(), n, out_f32, st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:8s} M={M} N={n} K={k}: {ms*1e3:7.1f} us  {2.0*M*n*k/ms/1e9:7.1f} TFLOP/s")
