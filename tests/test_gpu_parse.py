"""Batch SVG path parse + `Path::bbox` + `fit_size` on the device (rgpu_parse_svg_batch; reference src/svg.rs:62-421,
src/path.rs:428-451, 832-972, src/geometry.rs:470-516) against the oracle.  Bars: segment kinds, subpaths, closed flags, parse
status and error offset identical; control points, bbox and the fit transform bit for bit — except in strings with arc
commands (sin / cos / tan / acos: CUDA against the C library), where points and bbox agree within 64 ulp of the coordinates'
magnitude (1e-9 relative for random well-conditioned arcs: see test_random_arcs).  The parsed batch then feeds the fill kernels as it lies in HBM."""
import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb
import assets
from rasterize_b200 import Align, ffi, synth
from parse_common import (CORNER_STRINGS, DEGENERATE_ARCS, ERROR_STRINGS, REFERENCE_STRINGS, check_batch, garbage_strings, oracle_parse, random_arcs, random_svg,
                          svg_of)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def device_parse(rast, strings, fit=None):
    dpb, info = rast.parse_svg_batch(strings, fit=fit)
    b = dpb.download()
    got = (b.flat.points, b.flat.kinds, b.flat.subpath_offsets if len(b.flat.closed) else np.zeros(1, dtype=np.uint32), b.flat.closed,
           b.path_subpath_offsets)
    dpb.free()
    return got, info


def test_reference_and_corner_strings(rast):
    strings = REFERENCE_STRINGS + CORNER_STRINGS + ERROR_STRINGS
    for fit in ((64, 64, Align.Mid), (0, 300, Align.Min), (512, 0, Align.Max), (0, 0, Align.Mid), None):
        got, info = device_parse(rast, strings, fit)
        check_batch(strings, got, info, fit=None if fit is None else (fit[0], fit[1], int(fit[2])))


def test_reference_bbox_known_answer(rast):
    """src/path.rs:1098-1105 test_bbox: the squirrel's box is (0.25, 1.0) + (15.75, 14.0)"""
    _, info = device_parse(rast, [REFERENCE_STRINGS[0]])
    bb = info["bbox"][0]
    assert abs(bb[0] - 0.25) < 1e-12 and abs(bb[1] - 1.0) < 1e-12 and abs(bb[2] - bb[0] - 15.75) < 1e-12 and abs(bb[3] - bb[1] - 14.0) < 1e-12


def test_random_grammar(rast):
    """3 000 random walks over the grammar without arc commands: bit for bit.  (Random arcs: next test.)"""
    rng = np.random.default_rng(7)
    strings = [random_svg(rng, int(rng.integers(1, 60)), arcs=False) for _ in range(3000)]
    got, info = device_parse(rast, strings, (128, 128, Align.Mid))
    check_batch(strings, got, info, fit=(128, 128, 1))
    assert (info["status"] == 0).mean() > 0.9


def test_garbage_input_reports_the_references_errors(rast):
    strings = garbage_strings(321, 3000)
    got, info = device_parse(rast, strings, (64, 64, Align.Mid))
    check_batch(strings, got, info, fit=(64, 64, 1))
    assert (info["status"] != 0).mean() > 0.9


def test_random_arcs(rast):
    """Arcs: sin / cos / tan / acos come from CUDA instead of the C library.  Well-conditioned arcs agree within 1e-9 of the
    coordinates' magnitude, structure identical; the ill-conditioned ones are noise in the reference itself (see
    parse_common.random_arcs) and are pinned on the host, where the libm is shared (tests/test_parse_units.py)."""
    rng = np.random.default_rng(9)
    strings = [random_arcs(rng, int(rng.integers(1, 8))) for _ in range(1500)]
    got, info = device_parse(rast, strings, (128, 128, Align.Mid))
    check_batch.arc_splits = 0
    check_batch(strings, got, info, arc_rtol=1e-9)
    assert (info["status"] == 0).all() and check_batch.arc_splits == 0


def test_serialised_assets(rast):
    strings = [svg_of(assets.load_path(n)) for n in ("squirrel", "tv", "rust", "ava", "huyak", "material")]
    got, info = device_parse(rast, strings, (0, 2048, Align.Mid))
    check_batch(strings, got, info, fit=(0, 2048, 1))
    assert int(info["n_segments"][5]) == assets.load_path("material").segments_count()


def test_long_strings_with_errors_and_relative_groups(rast):
    """Long strings are cut at absolute movetos and parsed by many threads: an error deep inside reports the offset in the whole
    string and empties the path; relative movetos and text before the first `M` stay with their chunk."""
    rng = np.random.default_rng(3)
    good = " ".join(random_svg(rng, 30).replace("A", "L").replace("a", "l") for _ in range(40))
    body = "M1,2 " + " ".join(f"l{i % 7 - 3},{i % 5 - 2} q1,1 2,{i % 3} z m3,4 l1,1" for i in range(600)) + " M5,5 L6,6 7,7Z"
    strings = [good, body, body[:9000] + " L 1 x " + body[9000:], "L1,1 2,2 " + body, "m0,0 " + " ".join("l1,0 0,1" for _ in range(2000)),
               body + " M9,9 L", "M1 1" + " ".join(f"M{i},{i} L{i + 1},{i} {i + 1},{i + 1}z" for i in range(1500))]
    got, info = device_parse(rast, strings, (256, 256, Align.Mid))
    check_batch(strings, got, info, fit=(256, 256, 1))
    assert int(info["status"][2]) == 2 and int(info["status"][5]) == 2 and int(info["status"][1]) == 0


def test_degenerate_arcs_become_lines(rast):
    got, info = device_parse(rast, [s for s, _ in DEGENERATE_ARCS])
    pts, kinds, sp, closed, psp = got
    for i, (_, want) in enumerate(DEGENERATE_ARCS):
        assert int(info["status"][i]) == 0
        assert list(kinds[sp[psp[i]]:sp[psp[i + 1]]]) == want
    assert np.isfinite(pts).all()


def test_strict_mode_and_bad_arguments(rast):
    with pytest.raises(rb.RgpuError, match="path 1: InvalidScalar at offset 6"):
        rast.parse_svg_batch(["M0,0L1,1", "M0,0 L"], strict=True)
    dpb, info = rast.parse_svg_batch([])
    assert dpb.counts() == (0, 0, 0, 0)
    with pytest.raises(rb.RgpuError):
        rast.parse_svg_batch((b"M0,0", np.array([0, 9, 4], dtype=np.uint32)))  # decreasing offsets


def test_glyph_batch_from_text_to_pixels(rast):
    """The whole glyph pipeline from text: 2 000 glyph outlines as SVG strings are parsed on the device, fitted into 64 x 64
    by the transforms the parse returns, and filled straight from the parsed batch — against the oracle (parse, fit_size,
    mask) on a sample, and bit for bit against the same jobs over an uploaded copy of the parsed paths."""
    n = 2000
    pb = synth.glyph_batch(21, n)
    strings = [svg_of(pb.path(i)) for i in range(n)]
    dpb, info = rast.parse_svg_batch(strings, fit=(64, 64, Align.Mid))
    assert (info["status"] == 0).all() and (info["fit_width"] == 64).all() and (info["fit_height"] == 64).all()
    slab = rast.device_alloc(n * 4096 * 4)

    def render(handles):
        t = np.zeros(n, dtype=rb.JOB_DTYPE)
        t["path"] = handles
        t["tr"] = info["fit_tr"]
        t["fill_rule"] = int(rb.FillRule.NonZero)
        t["mode"] = ffi.JOB_MASK
        t["canvas"] = slab
        t["origin"] = np.arange(n, dtype=np.uint64) * np.uint64(4096)
        t["row_stride"] = 64
        t["width"] = 64
        t["height"] = 64
        rast.device_zero(slab, n * 4096 * 4)
        prepared = rast.prepare_job_table(t)
        prepared.render()
        rast.batch_status()
        out = rast.to_host(slab, (n, 64, 64), np.float32)
        prepared.free()
        return out

    got = render(dpb.handles())
    host = dpb.download()
    up = rast.upload_batch(host)
    assert np.array_equal(got, render(up.handles()))
    for i in range(0, n, 97):
        want = oracle_parse(strings[i])
        op = O.OraclePath.from_flat(want["points"], want["kinds"], want["subpath_offsets"], want["closed"])
        (ow, oh), tr = O.fit_size(want["bbox"], 64, 64, 1)
        assert np.array_equal(tr, info["fit_tr"][i])
        ref = np.zeros((64, 64))
        op.mask(tr, O.NONZERO, ref)
        assert np.abs(got[i] - ref).max() <= 1e-4, i
    assert got.max() > 0.99
    rast.device_free(slab)
    dpb.free()
    up.free()
