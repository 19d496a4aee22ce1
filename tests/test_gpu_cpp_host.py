"""Builds tests/cpp/test_host.cpp (the reference's test_rasterizer / test_fill_rule restated in C++ over
include/rasterize_b200.hpp) with g++, links it to the in-tree C-ABI library and runs it on the GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def build_cpp(tmp_path):
    from rasterize_b200 import build
    lib = build.build()
    exe = tmp_path / "test_host"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_host.cpp"), "-o", str(exe),
           f"-L{lib.parent}", "-lrasterize_b200", f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_header_compiles(tmp_path):
    """CPU: the C++ mirror compiles and links against the C-ABI library (no device calls)."""
    build_cpp(tmp_path)


@pytest.mark.gpu
def test_cpp_host_reference_tests(tmp_path):
    exe = build_cpp(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp host tests ok" in r.stdout
