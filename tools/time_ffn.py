"""Times ofx_ffn_block_bf16 on the CP token count of configs[1] (about 82k rows, d_model 512)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_ffn_block import _params, _want, DM
from outfitx_b200 import _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 82000
LN = "--ln" in sys.argv      # ofx_ffn_block_ln_bf16: also emits the next layer's norm1 (staged-H mode of the kernel)
ln_w, ln_b, w1, b1, w2, b2, g = _params(1)
x = torch.randn(rows, DM, device="cuda", generator=g)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
hn = torch.empty(rows, DM, device="cuda", dtype=torch.bfloat16)
ws = torch.zeros(L.ofx_ffn_block_workspace_bytes(rows, DM, 2048), dtype=torch.uint8, device="cuda")
def run(t):
    if LN:
        _lib.check(L.ofx_ffn_block_ln_bf16(t.data_ptr(), rows, DM, 2048, ln_w.data_ptr(), ln_b.data_ptr(), w1.data_ptr(),
                                           b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), hn.data_ptr(), ln_w.data_ptr(),
                                           ln_b.data_ptr(), ws.data_ptr(), ws.numel(), st))
        return
    _lib.check(L.ofx_ffn_block_bf16(t.data_ptr(), rows, DM, 2048, ln_w.data_ptr(), ln_b.data_ptr(), w1.data_ptr(),
                                    b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), ws.data_ptr(), ws.numel(), st))
y = x.clone(); run(y); torch.cuda.synchronize()
err = (y - _want(x, ln_w, ln_b, w1, b1, w2, b2)).abs().max().item()
bufs = [x.clone() for _ in range(4)]
for b in bufs: run(b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 20
for i in range(n): run(bufs[i % 4])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 2.0 * rows * DM * 2048 * 2
print(f"rows {rows} max|err| {err:.3e}  {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (padded 2048)")
