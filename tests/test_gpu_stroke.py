"""`Path::stroke` on the device (rgpu_path_stroke; reference src/path.rs:374-415, 692-732, src/curve.rs:978-1078, 1283-1433)
against the oracle.  Bars: the segment list (kinds, subpaths, closed flags) identical for every style; control points bit
for bit where no arc is involved (Miter / Bevel joins, Butt / Square caps, no cubic in the source — the pieces of an offset
cubic are joined by arcs whatever the style); arcs go through sin / cos / tan / acos (CUDA against the C library): their
control points within 64 ulp of the coordinates' magnitude, and more than 99 % of all coordinates bit-identical.  The stroked device path then goes straight into K1..K4: the mask of
config 5 rendered from the device-stroked outline equals the one rendered from the oracle-stroked fixture."""
import numpy as np
import pytest

import rasterize_b200 as rb
import assets
from rasterize_b200 import LineCap, LineJoin, StrokeStyle, ffi, sharding
from stroke_common import STYLES, compare, exact_case, oracle_stroke, random_paths, synthetic_paths

pytestmark = pytest.mark.gpu

JOIN = {"miter": LineJoin.Miter, "bevel": LineJoin.Bevel, "round": LineJoin.Round}
CAP = {"butt": LineCap.Butt, "square": LineCap.Square, "round": LineCap.Round}


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def device_stroke(rast, p, width, join, ml, cap):
    dp = rast.stroke(p, StrokeStyle(width, JOIN[join], ml, CAP[cap]))
    q = dp.download()
    dp.free()
    return q.points, q.kinds, q.subpath_offsets, q.closed


@pytest.mark.parametrize("name", ["squirrel", "tv", "rust", "ava", "huyak", "material"])
def test_assets_match_oracle(rast, name):
    p = assets.load_path(name)
    for width, join, ml, cap in (STYLES if name != "material" else STYLES[:3]):
        got = device_stroke(rast, p, width, join, ml, cap)
        want = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, width, join, ml, cap)
        same = compare(got, want, exact=exact_case(p.kinds, join, cap))
        # the offsets of lines and curves are exact, only some arcs differ in the last bit (a handful of coordinates)
        assert same > (0.99 if len(want[0]) > 10000 else 0.9)


def test_reference_known_answers(rast):
    """The reference's own `test_stroke` (src/path.rs:1225-1269): expected outlines given as SVG strings at 5-6 digits."""
    import oracle as O

    def host(svg):
        pts, kinds, sp, closed = O.OraclePath.parse(svg).export()
        return rb.Path(pts, kinds, sp, closed)

    def check(p, style, ref_svg):
        q = rast.stroke(p, style).download()
        r = host(ref_svg)
        assert list(q.kinds) == list(r.kinds)
        assert list(q.subpath_offsets) == list(r.subpath_offsets) and list(q.closed) == list(r.closed)
        assert np.abs(np.asarray(q.points) - np.asarray(r.points)).max() < 1e-4

    rr = StrokeStyle(1.0, LineJoin.Round, 4.0, LineCap.Round)
    path = host("M2,2L8,2C11,2 11,8 8,8L5,4")
    check(path, StrokeStyle(1.0), """
        M2,1.5 L8,1.5 C9.80902,1.5 10.75,3.38197 10.75,5 10.75,6.61803 9.80902,8.5 8,8.5 L7.75,8.5 7.6,8.3 4.6,4.3
        5.4,3.7 8.4,7.7 8,7.5 C9.19098,7.5 9.75,6.38197 9.75,5 9.75,3.61803 9.19098,2.5 8,2.5 L2,2.5 2,1.5 Z
    """)
    check(path, rr, """
        M2,1.5 L8,1.5 C9.80902,1.5 10.75,3.38197 10.75,5 10.75,6.61803 9.80902,8.5 8,8.5 7.84274,8.5 7.69436,8.42581
        7.6,8.3 L4.6,4.3 C4.43542,4.08057 4.48057,3.76458 4.7,3.6 4.91943,3.43542 5.23542,3.48057 5.4,3.7 L8.4,7.7 8,7.5
        C9.19098,7.5 9.75,6.38197 9.75,5 9.75,3.61803 9.19098,2.5 8,2.5 L2,2.5 C1.72571,2.5 1.5,2.27429 1.5,2
        1.5,1.72571 1.72571,1.5 2,1.5 Z
    """)
    check(host("M2,2L8,2C11,2 11,8 8,8L5,4Z"), rr, """
        M2,1.5 L8,1.5 C9.80902,1.5 10.75,3.38197 10.75,5 10.75,6.61803 9.80902,8.5 8,8.5 7.84274,8.5 7.69436,8.42581
        7.6,8.3 L4.6,4.3 4.72265,4.41603 1.72265,2.41603 C1.53984,2.29415 1.45778,2.06539 1.52145,1.85511 1.58512,1.64482
        1.78029,1.5 2,1.5 ZM5.4,3.7 L8.4,7.7 8,7.5 C9.19098,7.5 9.75,6.38197 9.75,5 9.75,3.61803 9.19098,2.5 8,2.5
        L2,2.5 2.27735,1.58397 5.27735,3.58397 C5.32451,3.61542 5.36599,3.65465 5.4,3.7 Z
    """)


def test_corner_cases_match_oracle(rast):
    for name, p in synthetic_paths().items():
        for width, join, ml, cap in STYLES:
            got = device_stroke(rast, p, width, join, ml, cap)
            want = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, width, join, ml, cap)
            try:
                compare(got, want, exact=exact_case(p.kinds, join, cap))
            except AssertionError as e:
                raise AssertionError(f"{name} {width} {join} {cap}: {e}") from None


def test_random_paths_match_oracle(rast):
    """300 random paths with degenerate pieces (repeated points, segments shorter than EPSILON, collinear runs, cusps, tiny and
    huge scales): same structure as the oracle in every style; points bit for bit where no arc is involved."""
    for i, p in enumerate(random_paths(23, 300)):
        for width, join, ml, cap in STYLES[:3] if i % 2 else STYLES[3:]:
            w = width * (1.0 if i % 3 else 0.01)
            got = device_stroke(rast, p, w, join, ml, cap)
            want = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, w, join, ml, cap)
            try:
                compare(got, want, exact=exact_case(p.kinds, join, cap), scale=float(np.abs(p.points).max()))
            except AssertionError as e:
                raise AssertionError(f"path {i} {w} {join} {cap}: {e}") from None


def test_empty_and_errors(rast):
    dp = rast.stroke(rb.Path.empty(), StrokeStyle(1.0))
    assert dp.counts() == (0, 0, 0)
    q = dp.download()
    assert q.segments_count() == 0
    b = rb.PathBuilder(); b.move_to((1, 1)); b.line_to((1, 1 + 1e-18)); p = b.build()  # nothing survives the offset
    if p.segments_count():
        assert rast.stroke(p, StrokeStyle(1.0)).counts() == (0, 0, 0)
    p = assets.load_path("squirrel")
    with pytest.raises(rb.RgpuError):
        st = ffi.CStrokeStyle(1.0, 4.0, 7, 0)
        import ctypes as C
        h = C.c_void_p()
        c = p._c()
        rast._check(ffi.lib().rgpu_path_stroke(rast.ctx, C.byref(c), C.byref(st), C.byref(h)))


def test_stroke_of_device_resident_paths(rast):
    """rgpu_dpath_stroke: the source is already on the device — an uploaded path, an element of an uploaded batch, an element of
    a batch parsed from text — and gives the same outline as the host-path entry point, bit for bit."""
    paths = [assets.load_path(n) for n in ("squirrel", "tv", "rust")]
    style = StrokeStyle(0.7, LineJoin.Miter, 4.0, LineCap.Square)
    want = [rast.stroke(p, style).download() for p in paths]
    up = rast.upload(paths[1])
    got = rast.stroke(up, style).download()
    assert np.array_equal(got.points, want[1].points) and np.array_equal(got.kinds, want[1].kinds) and np.array_equal(got.closed, want[1].closed)
    dpb = rast.upload_batch(rb.PathBatch.from_paths(paths))
    for i in range(3):
        g = rast.stroke(dpb.handle(i), style).download()
        assert np.array_equal(g.points, want[i].points) and np.array_equal(g.subpath_offsets, want[i].subpath_offsets)
    # text -> parse -> stroke -> mask without the geometry visiting the host, against the oracle's parse + stroke + mask
    import oracle as O
    svg = paths[1].to_svg_path()
    parsed, info = rast.parse_svg_batch([svg], fit=(512, 512, rb.Align.Mid))
    outline = rast.stroke(parsed.handle(0), StrokeStyle(0.5, LineJoin.Round, 4.0, LineCap.Round))
    tr = info["fit_tr"][0]
    canvas = rast.device_alloc(512 * 512 * 4)
    rast.render_batch([rb.Job(outline, tr, rb.FillRule.NonZero, ffi.JOB_MASK, canvas, 512, 512, 512)], independent=True)
    img = rast.to_host(canvas, (512, 512), np.float32)
    rast.device_free(canvas)
    ref = np.zeros((512, 512))
    O.OraclePath.parse(svg).stroke(0.5, "round", 4.0, "round").mask(tr, O.NONZERO, ref)
    assert img.max() > 0.9 and np.abs(img - ref).max() <= 1e-4


def test_config5_from_device_stroke(rast):
    """config 5 without the host round trip: tv.path stroked on the device and the outline rasterized as it lies in HBM,
    against the same canvas rendered from the oracle-stroked fixture (4096 rows of the 32768-wide canvas)."""
    p, fixture = assets.load_path("tv"), assets.load_path("tv_stroked")
    c5 = assets.expected()["paths"]["tv_stroked"]["c5"]
    w, hfull = c5["size"]
    tr = np.array(c5["tr"])
    y0, y1 = sharding.band_rows(hfull, 3, 8)
    h = y1 - y0
    dp = rast.stroke(p, StrokeStyle(0.5, LineJoin.Round, 4.0, LineCap.Round))
    assert dp.counts() == (len(fixture.points), len(fixture.kinds), len(fixture.closed))
    df = rast.upload(fixture)
    outs = []
    for d in (dp, df):
        canvas = rast.device_alloc(w * h * 4)
        rast.render_batch([rb.Job(d, sharding.band_transform(tr, y0), rb.FillRule.NonZero, ffi.JOB_MASK, canvas, w, h, w)], independent=True)
        outs.append(rast.to_host(canvas, (h, w), np.float32))
        rast.device_free(canvas)
    assert outs[1].max() > 0.5
    assert np.abs(outs[0] - outs[1]).max() <= 1e-6  # a few ulp on the arcs' control points
