"""CPU: the reference-side Rust files (rust/) against the C header.  There is no Rust toolchain in this image, so the
`extern "C"` block of rust/src/gpu.rs is checked mechanically: every function it declares exists in
include/rasterize_b200.h with the same number of arguments and compatible pointer / scalar kinds, and rust/build.rs
compiles exactly the sources rasterize_b200/build.py compiles."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def split_args(s: str):
    s = s.strip()
    if not s or s == "void":
        return []
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([<":
            depth += 1
        elif ch in ")]>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def header_functions():
    text = (ROOT / "include" / "rasterize_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    fns = {}
    for m in re.finditer(r"\b([A-Za-z_][\w \*]*?)\b(rgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        fns[m.group(2)] = split_args(" ".join(m.group(3).split()))
    return fns


def rust_externs():
    text = (ROOT / "rust" / "src" / "gpu.rs").read_text()
    blocks = re.findall(r'extern "C" \{(.*?)\n\}', text, flags=re.S)
    fns = {}
    for b in blocks:
        for m in re.finditer(r"fn (rgpu_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", b, flags=re.S):
            fns[m.group(1)] = split_args(" ".join(m.group(2).split()))
    return fns


def kind_c(arg: str) -> str:
    if "*" in arg or "[" in arg:
        return "ptr"
    if re.search(r"\brgpu_shape\b", arg):
        return "shape"
    if re.search(r"\b(double|float)\b", arg):
        return "float"
    return "int"


def kind_rust(arg: str) -> str:
    ty = arg.split(":", 1)[1].strip()
    if ty.startswith("*"):
        return "ptr"
    if ty == "RgpuShape":
        return "shape"
    if ty in ("f64", "f32"):
        return "float"
    return "int"


def test_extern_block_matches_header():
    c, r = header_functions(), rust_externs()
    assert len(r) >= 15
    for name, rargs in r.items():
        assert name in c, f"rust/src/gpu.rs declares {name}, which include/rasterize_b200.h does not"
        cargs = c[name]
        assert len(cargs) == len(rargs), f"{name}: {len(rargs)} arguments in gpu.rs, {len(cargs)} in the header"
        for i, (ca, ra) in enumerate(zip(cargs, rargs)):
            assert kind_c(ca) == kind_rust(ra), f"{name} argument {i}: `{ca}` vs `{ra}`"


def test_trait_entry_points_are_bound():
    r = rust_externs()
    for name in ("rgpu_create", "rgpu_destroy", "rgpu_last_error", "rgpu_mask", "rgpu_mask_iter", "rgpu_fill", "rgpu_flatten",
                 "rgpu_fill_batch_host", "rgpu_multi_create", "rgpu_multi_fill_batch_host", "rgpu_multi_mask_banded_host"):
        assert name in r


def test_build_rs_lists_the_sources_build_py_compiles():
    from rasterize_b200 import build
    text = (ROOT / "rust" / "build.rs").read_text()
    m = re.search(r"const SOURCES: &\[&str\] = &\[(.*?)\];", text, flags=re.S)
    assert sorted(re.findall(r'"([^"]+)"', m.group(1))) == sorted(build.SOURCES)
    assert "arch=compute_100a,code=sm_100a" in text


def test_patch_fixes_the_round1_slips():
    """`Transform` has a private field and no conversion to an array, and `fill_with_paint` did not exist."""
    patch = (ROOT / "rust" / "paint_gpu_desc.patch").read_text()
    assert "impl From<Transform> for [Scalar; 6]" in patch
    assert "pub(crate) fn fill_with_paint(" in patch and "fn gpu_desc(&self) -> Option<crate::PaintDesc>" in patch
    gpu = (ROOT / "rust" / "src" / "gpu.rs").read_text()
    assert "rasterize::fill_with_paint" in gpu and "tr.into()" in gpu
