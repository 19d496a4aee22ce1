"""CPU check of the device SVG parser's code (rasterize_b200/csrc/parse_device.cuh, compiled for the host by
tests/cpp/parse_host_check.cpp and run through the same count -> scan -> emit passes) against the oracle's parser, `Path::bbox`
and `fit_size` (reference src/svg.rs:62-421, src/path.rs:428-451, 832-972, src/geometry.rs:470-516).  On the host the C
library's sin / cos are the oracle's own, so arcs must match bit for bit here too.  The GPU run of the same comparison is
tests/test_gpu_parse.py."""
import ctypes as C
import subprocess
from pathlib import Path as FsPath

import numpy as np
import pytest

import assets
from parse_common import CORNER_STRINGS, DEGENERATE_ARCS, ERROR_STRINGS, INFO_DTYPE, REFERENCE_STRINGS, check_batch, garbage_strings, pack, random_arcs, random_svg, svg_of

ROOT = FsPath(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("parse_check") / "libparse_check.so"
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", str(ROOT / "tests" / "cpp" / "parse_host_check.cpp"), "-o", str(so)]
    subprocess.run(cmd, check=True, timeout=300)
    lib = C.CDLL(str(so))
    u32p, u8p, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    lib.parse_check_run.restype = C.c_int
    lib.parse_check_run.argtypes = [C.c_char_p, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, u32p, u32p, u32p, u32p]
    lib.parse_check_fetch.argtypes = [dp, u8p, u32p, u8p, u32p]
    return lib


def run_harness(lib, strings, fit=None):
    u32p, u8p, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    text, off = pack(strings)
    n = len(strings)
    info = np.zeros(max(n, 1), dtype=INFO_DTYPE)
    cnt = [C.c_uint32() for _ in range(4)]
    fw, fh, fa = fit if fit is not None else (0, 0, -1)
    rc = lib.parse_check_run(text, off.ctypes.data_as(u32p), n, fw, fh, fa, info.ctypes.data, *[C.byref(v) for v in cnt])
    assert rc == 0, f"harness self-check failed: {rc}"
    n_pts, n_seg, n_sub, run_harness.last_chunks = (v.value for v in cnt)
    pts = np.zeros((n_pts, 2))
    kinds = np.zeros(n_seg, dtype=np.uint8)
    sp = np.zeros(n_sub + 1, dtype=np.uint32)
    closed = np.zeros(n_sub, dtype=np.uint8)
    psp = np.zeros(n + 1, dtype=np.uint32)
    lib.parse_check_fetch(pts.ctypes.data_as(dp), kinds.ctypes.data_as(u8p), sp.ctypes.data_as(u32p), closed.ctypes.data_as(u8p), psp.ctypes.data_as(u32p))
    return (pts, kinds, sp, closed, psp), info[:n]


def test_reference_and_corner_strings(harness):
    strings = REFERENCE_STRINGS + CORNER_STRINGS + ERROR_STRINGS
    for fit in ((64, 64, 1), (0, 300, 0), (512, 0, 2), (0, 0, 1)):
        got, info = run_harness(harness, strings, fit)
        check_batch(strings, got, info, fit=fit, exact_arcs=True)


def test_degenerate_arcs_become_lines(harness):
    got, info = run_harness(harness, [s for s, _ in DEGENERATE_ARCS])
    pts, kinds, sp, closed, psp = got
    for i, (_, want) in enumerate(DEGENERATE_ARCS):
        assert int(info["status"][i]) == 0
        assert list(kinds[sp[psp[i]]:sp[psp[i + 1]]]) == want
    assert np.isfinite(pts).all()


def test_random_grammar(harness):
    rng = np.random.default_rng(7)
    strings = [random_svg(rng, int(rng.integers(1, 60))) for _ in range(1500)]
    got, info = run_harness(harness, strings, (128, 128, 1))
    check_batch(strings, got, info, fit=(128, 128, 1), exact_arcs=True)
    assert (info["status"] == 0).mean() > 0.9
    strings = [random_arcs(rng, int(rng.integers(1, 8))) for _ in range(300)]
    got, info = run_harness(harness, strings, (128, 128, 1))
    check_batch(strings, got, info, fit=(128, 128, 1), exact_arcs=True)


def test_serialised_assets(harness):
    """Long strings are cut at absolute movetos and parsed chunk by chunk (material: 0.58 MB, > 1 000 chunks): same segments,
    same bbox, same fit as the oracle's serial parse."""
    strings = [svg_of(assets.load_path(n)) for n in ("squirrel", "tv", "rust", "ava", "huyak", "material")]
    got, info = run_harness(harness, strings, (0, 2048, 1))
    assert run_harness.last_chunks > 1000
    check_batch(strings, got, info, fit=(0, 2048, 1), exact_arcs=True)
    assert int(info["n_segments"][5]) == assets.load_path("material").segments_count()


def test_long_strings_with_errors_and_relative_groups(harness):
    """Chunked strings: an error deep inside reports the offset in the whole string and empties the path; relative movetos and
    text before the first `M` stay with their chunk; a long string without any `M` is one chunk."""
    rng = np.random.default_rng(3)
    good = " ".join(random_svg(rng, 30).replace("A", "L").replace("a", "l") for _ in range(40))
    body = "M1,2 " + " ".join(f"l{i % 7 - 3},{i % 5 - 2} q1,1 2,{i % 3} z m3,4 l1,1" for i in range(600)) + " M5,5 L6,6 7,7Z"
    strings = [good, body, body[:9000] + " L 1 x " + body[9000:], "L1,1 2,2 " + body, "m0,0 " + " ".join("l1,0 0,1" for _ in range(2000)),
               body + " M9,9 L", "M1 1" + " ".join(f"M{i},{i} L{i + 1},{i} {i + 1},{i + 1}z" for i in range(1500))]
    got, info = run_harness(harness, strings, (256, 256, 1))
    assert run_harness.last_chunks > len(strings) + 20
    check_batch(strings, got, info, fit=(256, 256, 1), exact_arcs=True)
    assert int(info["status"][2]) == 2 and int(info["status"][5]) == 2 and int(info["status"][1]) == 0


def test_garbage_input_reports_the_references_errors(harness):
    strings = garbage_strings(123, 3000)
    got, info = run_harness(harness, strings, (64, 64, 1))
    check_batch(strings, got, info, fit=(64, 64, 1), exact_arcs=True)
    assert (info["status"] != 0).mean() > 0.9
