"""CPU: the C-ABI library builds, loads and exports exactly what include/ofx.h declares; the
size queries work without a GPU; compute entry points refuse to run without an sm_100 device
(there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ofx.h")).read()
    return sorted(set(re.findall(r"OFX_API[^;(]*?\b(ofx_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree():
    from outfitx_b200 import _lib
    names = _declared()
    assert len(names) >= 14
    assert sorted(_lib.EXPORTS) == names
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n)
    assert L.ofx_version() == 100


def test_size_queries_without_gpu():
    from outfitx_b200 import _lib
    L = _lib.lib()
    s = _lib.Shape(1024, 1024, 16, 6, 2024, 16, _lib.PREC_BF16)
    n_bf16 = L.ofx_packed_weights_bytes(C.byref(s))
    s.precision = _lib.PREC_FP32
    n_f32 = L.ofx_packed_weights_bytes(C.byref(s))
    # 6 layers x (3+1) Dm^2 + 2 Dm*2048 weights dominate; fp32 packs the fp32 matrices (4 B) plus their bf16 hi / lo
    # pieces for the tensor-core form (3 x 2 B): five times the bf16 bytes
    assert 100e6 < n_bf16 < 110e6 and 4.9 < n_f32 / n_bf16 < 5.0
    assert L.ofx_encoder_workspace_bytes(C.byref(s), 64) > 64 * 17 * 1024 * 4
    bad = _lib.Shape(1000, 1024, 16, 6, 2024, 16, 0)
    assert L.ofx_packed_weights_bytes(C.byref(bad)) == 0
    assert b"d_model" in L.ofx_last_error()
    assert L.ofx_gallery_packed_bytes(1000, 1024) >= 1000 * 1024 * 2 + 4000
    assert L.ofx_search_workspace_bytes(10_000_000, 1024, 8192, 10) > 8192 * 1024 * 2
    assert L.ofx_search_workspace_bytes(1000, 1024, 8, 1000) == 0   # k out of range
    assert L.ofx_search_workspace_bytes(1000, 1024, 8, 64) > 0      # the reference ranks k = 50
    assert L.ofx_exact_search_workspace_bytes(10_000_000, 8, 10) >= 128 * 8 * 10 * 16


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from outfitx_b200 import _lib
    L = _lib.lib()
    buf = (C.c_float * 2048)()
    rc = L.ofx_fuse(C.addressof(buf), C.addressof(buf), 1, 512, 0, 1, C.addressof(buf), None)
    assert rc == -3 and L.ofx_last_error()           # OFX_E_ARCH, with a message
    assert L.ofx_device_ok(0) == -3
    import outfitx_b200 as o
    m = o.OutfitX(o.OutfitXConfig(item_encoder=o.ItemEncoderConfig(type="clip")))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(o.OutfitCompatibilityPredictionTask, outfit_embedding=torch.zeros(1, 16, 1024),
          outfit_mask=torch.zeros(1, 16, dtype=torch.bool))
    from outfitx_b200.search import Gallery
    with pytest.raises(RuntimeError, match="no CPU path"):
        Gallery.build(torch.zeros(4, 1024))
