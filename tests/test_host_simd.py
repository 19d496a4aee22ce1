"""Host-side streaming loops (rasterize_b200/csrc/host_simd.cpp): every vector variant the CPU offers gives the bytes of the
scalar expression — `colour * alpha` per component, `(double)f32`, rows rebuilt from class bytes and literals (run-coded download) — for unaligned heads, ragged tails and empty ranges, and
writes nothing beyond its range.  CPU-only: the file is compiled on its own with g++ (tests/cpp/test_host_simd.cpp)."""
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_host_simd_variants(tmp_path):
    exe = tmp_path / "host_simd"
    subprocess.run(["g++", "-O2", "-std=c++17", str(ROOT / "tests" / "cpp" / "test_host_simd.cpp"),
                    str(ROOT / "rasterize_b200" / "csrc" / "host_simd.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad 0" in r.stdout
