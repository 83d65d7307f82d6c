"""The C-ABI library loads on a CPU-only box and exports every symbol include/vavae_b200.h declares
(no compute calls here); the product path fails loudly without a GPU instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    header = open(os.path.join(ROOT, "include", "vavae_b200.h")).read()
    return sorted(set(re.findall(r"\b(vavae_[a-z0-9_]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    from va_vae_b200 import engine
    lib = engine.load_library()
    declared = _declared()
    assert len(declared) >= 15
    for sym in declared:
        assert hasattr(lib, sym), f"libvavae_b200.so does not export {sym}"
    assert set(engine.EXPORTS) == set(declared)
    assert lib.vavae_abi_version() == engine.ABI_VERSION


def test_header_cites_reference_interfaces():
    header = open(os.path.join(ROOT, "include", "vavae_b200.h")).read()
    for cite in ("lightningdit.py:391-418", "lightningdit.py:420-442", "transport.py:398-443", "integrators.py:107-126"):
        assert cite in header


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_create_fails_loudly():
    from va_vae_b200 import engine
    lib = engine.load_library()
    cfg = engine.VavaeConfig(abi_version=engine.ABI_VERSION, input_size=16, patch_size=1, in_channels=32, hidden_size=64,
                             depth=1, num_heads=2, mlp_hidden=170, num_classes=10, label_rows=11, learn_sigma=0,
                             use_qknorm=0, use_swiglu=1, use_rope=1, use_rmsnorm=1, wo_shift=0, precision=0,
                             max_chunk_rows=0)
    h = ctypes.c_void_p()
    rc = lib.vavae_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.vavae_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_module_refuses_cpu_execution():
    import va_vae_b200 as V
    m = V.LightningDiT(input_size=4, patch_size=1, in_channels=4, hidden_size=32, depth=1, num_heads=2, num_classes=3,
                       use_swiglu=True, use_rope=True, use_rmsnorm=True).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 4, 4, 4), torch.zeros(2), torch.zeros(2, dtype=torch.long))


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from va_vae_b200 import engine
    monkeypatch.setattr(engine, "_lib", None)
    monkeypatch.setattr(engine, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        engine.load_library()
