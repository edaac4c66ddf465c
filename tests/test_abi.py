"""The C-ABI library builds, loads, and exports every symbol include/pngb200.h declares.
No GPU needed (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT, product


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pngb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pngb200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    p = product()
    path = p.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_size_helpers_match_oracle(orc):
    p = product()
    for (w, h, vol, il) in [(1, 1, 1, 0), (32, 32, 32, 0), (33, 7, 24, 1), (1920, 1080, 32, 0),
                            (7680, 4320, 64, 0), (5, 5, 4, 1), (1, 9, 16, 1)]:
        assert p.filtered_size(w, h, vol, bool(il)) == orc.filtered_size(w, h, vol, bool(il))
        assert p.storage_size(w, h, vol) == orc.storage_size(w, h, vol)


def test_no_gpu_means_loud_failure():
    """the product path must fail loudly, not fall back, when there is no device"""
    import pytest
    p = product()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(p.PNGB200Error):
        p.Context(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "swift-png_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "oracle" not in text, (dirpath, f)
