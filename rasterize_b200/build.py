"""Builds librasterize_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "librasterize_b200.so"
SOURCES = ["flatten.cu", "scan.cu", "raster.cu", "scene.cu", "small.cu", "compose.cu", "compact.cu", "stroke.cu", "parse.cu", "context.cu", "host_simd.cpp"]
# stroke.cu / parse.cu restate f64 expressions of the reference with plain operators: no multiply-add contraction there
SOURCE_FLAGS = {"stroke.cu": ["--fmad=false"], "parse.cu": ["--fmad=false"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: rasterize_b200 needs the CUDA toolkit (there is no CPU fallback)")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inl")) + list(CSRC.glob("*.hpp")) + list(CSRC.glob("*.cpp")) + [PKG.parent / "include" / "rasterize_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    extra = os.environ.get("RGPU_NVCC_EXTRA", "").split()  # tuning builds, e.g. -DRGPU_FLAT_MINB=9
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *SOURCE_FLAGS.get(src, []), *extra, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    objs = [str(objdir / (s + ".o")) for s in SOURCES]
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-Xcompiler", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
