"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares (no compute calls: there is no GPU here), and fails loudly instead of falling back."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built():
    from rasterize_b200 import build
    return build.build()


def header_symbols():
    text = (ROOT / "include" / "rasterize_b200.h").read_text()
    return sorted(set(re.findall(r"\b(rgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(str(built))
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rasterize_b200.h but not exported"


def test_ffi_table_matches_header(built):
    from rasterize_b200 import ffi
    assert sorted(ffi.SYMBOLS) == header_symbols()
    ffi.lib()


def test_no_cpu_fallback(built):
    """Without a usable CUDA device the context cannot be created and nothing computes on the CPU."""
    import rasterize_b200 as rb
    from rasterize_b200 import ffi
    if ffi.lib().rgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rb.RgpuError) as e:
        rb.GpuRasterizer()
    assert e.value.code == ffi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product package must not reference the oracle (SURVEY/DESIGN: oracle is test infrastructure)."""
    for f in (ROOT / "rasterize_b200").rglob("*"):
        if f.suffix in {".py", ".cu", ".cuh", ".h", ".cpp"}:
            t = f.read_text()
            assert "oracle" not in t.lower(), f
            assert "golden" not in t.lower() and "tests/" not in t, f  # nor the fixtures of the test tree
