"""CPU check of the device stroke's host logic: the unit table (rasterize_b200/csrc/stroke_units.hpp) and the per-unit
code the kernels call (stroke_device.cuh), compiled for the host by tests/cpp/stroke_host_check.cpp and run unit by unit,
against the oracle's `Path::stroke` (reference src/path.rs:374-415).  On the host the C library's sin / cos are the
oracle's own, so here every style must match bit for bit — which pins the decomposition (tables, look-behind, counts,
offsets, order).  The GPU run of the same comparison is tests/test_gpu_stroke.py."""
import ctypes as C
import subprocess
from pathlib import Path as FsPath

import numpy as np
import pytest

import assets
from stroke_common import CAPS, JOINS, STYLES, compare, oracle_stroke, random_paths, synthetic_paths

ROOT = FsPath(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("stroke_check") / "libstroke_check.so"
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", str(ROOT / "tests" / "cpp" / "stroke_host_check.cpp"), "-o", str(so)]
    subprocess.run(cmd, check=True, timeout=300)
    lib = C.CDLL(str(so))
    u32p, u8p, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    lib.stroke_check_run.restype = C.c_int
    lib.stroke_check_run.argtypes = [dp, u8p, C.c_uint32, u32p, u8p, C.c_uint32, C.c_double, C.c_double, C.c_int, C.c_int, u32p, u32p, u32p]
    lib.stroke_check_fetch.argtypes = [dp, u8p, u32p, u8p]
    lib.stroke_check_hypot.restype = C.c_double
    lib.stroke_check_hypot.argtypes = [C.c_double, C.c_double]
    return lib


def run_harness(lib, p, width, join, miter_limit, cap):
    u32p, u8p, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    pts = np.ascontiguousarray(p.points, dtype=np.float64)
    kinds = np.ascontiguousarray(p.kinds, dtype=np.uint8)
    sp = np.ascontiguousarray(p.subpath_offsets, dtype=np.uint32)
    closed = np.ascontiguousarray(p.closed, dtype=np.uint8)
    n = [C.c_uint32() for _ in range(3)]
    rc = lib.stroke_check_run(pts.ctypes.data_as(dp), kinds.ctypes.data_as(u8p), len(kinds), sp.ctypes.data_as(u32p), closed.ctypes.data_as(u8p),
                              len(closed), width, miter_limit, JOINS[join], CAPS[cap], *[C.byref(v) for v in n])
    assert rc == 0, f"harness self-check failed: {rc}"
    n_pts, n_seg, n_sub = (v.value for v in n)
    o_pts = np.zeros((n_pts, 2))
    o_kinds = np.zeros(n_seg, dtype=np.uint8)
    o_sp = np.zeros(n_sub + 1, dtype=np.uint32)
    o_closed = np.zeros(n_sub, dtype=np.uint8)
    lib.stroke_check_fetch(o_pts.ctypes.data_as(dp), o_kinds.ctypes.data_as(u8p), o_sp.ctypes.data_as(u32p), o_closed.ctypes.data_as(u8p))
    return o_pts, o_kinds, o_sp, o_closed


def test_hypot_is_the_c_librarys(harness):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-50, 50, 200000), rng.standard_normal(100000) * 10.0 ** rng.integers(-300, 300, 100000), [0.0, 3.0, np.inf, 1e-320]])
    y = np.concatenate([rng.uniform(-50, 50, 200000), rng.standard_normal(100000) * 10.0 ** rng.integers(-300, 300, 100000), [0.0, 4.0, 1.0, 1e-320]])
    want = np.hypot(x, y)
    got = np.array([harness.stroke_check_hypot(a, b) for a, b in zip(x.tolist(), y.tolist())])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("name", ["squirrel", "tv", "rust", "ava", "huyak"])
def test_assets_match_oracle(harness, name):
    p = assets.load_path(name)
    for width, join, ml, cap in STYLES:
        got = run_harness(harness, p, width, join, ml, cap)
        want = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, width, join, ml, cap)
        compare(got, want, exact=True)


def test_material_matches_oracle(harness):
    p = assets.load_path("material")
    for width, join, ml, cap in STYLES[:2]:
        got = run_harness(harness, p, width, join, ml, cap)
        want = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, width, join, ml, cap)
        compare(got, want, exact=True)


def test_stroked_fixture_is_config5(harness):
    """tests/golden/paths/tv_stroked.npz (config 5's outline, made by the oracle) = stroke(tv, 0.5, round, round)."""
    p, q = assets.load_path("tv"), assets.load_path("tv_stroked")
    got = run_harness(harness, p, *STYLES[0])
    compare(got, (q.points, q.kinds, q.subpath_offsets, q.closed), exact=True)


def test_corner_cases_match_oracle(harness):
    for name, p in synthetic_paths().items():
        for width, join, ml, cap in STYLES:
            got = run_harness(harness, p, width, join, ml, cap)
            want = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, width, join, ml, cap)
            try:
                compare(got, want, exact=True)
            except AssertionError as e:
                raise AssertionError(f"{name} {width} {join} {cap}: {e}") from None


def test_random_paths_match_oracle(harness):
    """400 random paths with degenerate pieces, every style: the unit decomposition follows the reference's serial walk bit for bit."""
    for i, p in enumerate(random_paths(17, 400)):
        for width, join, ml, cap in STYLES[:4] if i % 2 else STYLES[2:]:
            w = width * (1.0 if i % 3 else 0.01)
            got = run_harness(harness, p, w, join, ml, cap)
            want = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, w, join, ml, cap)
            try:
                compare(got, want, exact=True)
            except AssertionError as e:
                raise AssertionError(f"path {i} {w} {join} {cap}: {e}") from None
