"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run on a B200: pytest -m gpu.

Bars (BASELINE.json north_star): flattened line counts identical (we require bit-identical lines), f32
coverage within 1e-4 absolute per pixel, 8-bit RGBA within 1 LSB.
"""
import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb
import assets
from rasterize_b200 import ffi

pytestmark = pytest.mark.gpu

COV_TOL = 1e-4   # north_star: f32 coverage within 1e-4 absolute per pixel
ASSETS = ["squirrel", "tv", "rust", "material", "ava", "huyak", "tv_stroked", "squirrel_stroked"]


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def opath(p: rb.Path) -> O.OraclePath:
    return O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)


@pytest.mark.parametrize("name", ASSETS)
def test_flatten_identical(rast, name):
    p = assets.load_path(name)
    e = assets.expected()["paths"][name]
    tr = np.array(e["size_tr"])
    got = rast.flatten(p, tr, True)
    ref = opath(p).flatten(tr)
    assert len(got) == len(ref) == e["lines_at_size"]
    assert np.array_equal(got, ref)  # bit-identical endpoints, reference order
    got_open = rast.flatten(p, tr, False)
    ref_open = opath(p).flatten(tr, close=False)
    assert np.array_equal(got_open, ref_open)


def test_flatten_rotated_and_configs(rast):
    """src/path.rs:1192-1211 test_flatten transform + the BASELINE config transforms"""
    import math
    p = assets.load_path("squirrel")
    tr = O.transform_mul(O.rotate(math.pi / 3.0), O.translate(-10.0, -20.0))
    assert np.array_equal(rast.flatten(p, tr), opath(p).flatten(tr))
    ex = assets.expected()["paths"]
    for name, key in (("squirrel", "c1"), ("material", "c2"), ("tv_stroked", "c5")):
        p = assets.load_path(name)
        tr = np.array(ex[name][key]["tr"])
        got = rast.flatten(p, tr)
        assert len(got) == ex[name][key]["lines"]
        assert np.array_equal(got, opath(p).flatten(tr))


def test_flatten_quads_and_flatness(rast):
    b = rb.Path.builder()
    b.move_to((1, 1)).quad_to((40, 3), (50, 60)).quad_to((60, 120), (5, 90)).line_to((3, 3)).close()
    b.move_to((10, 10)).cubic_to((300, 20), (-200, 200), (90, 90))
    p = b.build()
    for fl in (0.05, 0.5, 1e-3):
        r = rb.GpuRasterizer(flatness=fl)
        for close in (True, False):
            got = r.flatten(p, rb.Transform.new_scale(3.0, 2.0), close)
            ref = opath(p).flatten(rb.Transform.new_scale(3.0, 2.0).array(), fl, close)
            assert np.array_equal(got, ref), (fl, close, len(got), len(ref))
        r.close()


def test_nan_is_an_error(rast):
    """reference panics: "cannot flatten segment with NaN" (src/path.rs:765-767)"""
    p = rb.Path([[0, 0], [float("nan"), 1], [2, 2]], [3], [0, 1], [0])
    with pytest.raises(rb.RgpuError) as e:
        rast.flatten(p)
    assert e.value.code == ffi.ERR_NAN
    # the context stays usable
    assert len(rast.flatten(assets.load_path("squirrel"))) > 0


@pytest.mark.parametrize("name", ["squirrel", "tv", "rust", "ava", "huyak", "material", "squirrel_stroked"])
@pytest.mark.parametrize("rule", [rb.FillRule.NonZero, rb.FillRule.EvenOdd])
def test_mask_matches_oracle(rast, name, rule):
    """`Rasterizer::mask` on the `Path::size` canvas (the reference bench's workload, benches/rasterize_bench.rs:99-108)"""
    p = assets.load_path(name)
    e = assets.expected()["paths"][name]
    w, h = e["size"]
    tr = np.array(e["size_tr"])
    img = np.zeros((h, w))
    rast.mask(p, tr, img, rule)
    ref = np.zeros((h, w))
    opath(p).mask(tr, int(rule), ref)
    assert np.abs(img - ref).max() <= COV_TOL
    key = "mask_sum_nonzero" if rule == rb.FillRule.NonZero else "mask_sum_evenodd"
    assert abs(img.sum() - e[key]) <= 1e-4 * w * h
    # f32 entry point agrees with the f64 one
    img32 = np.zeros((h, w), dtype=np.float32)
    rast.mask(p, tr, img32, rule)
    assert np.array_equal(img32.astype(np.float64), img)


def test_reference_kats_through_gpu(rast):
    """src/rasterize.rs:1065-1161 test_rasterizer + test_fill_rule with the GpuRasterizer in the loop"""
    b = rb.Path.builder()
    b.move_to((1, 0)).line_to((1, 1)).line_to((2, 2)).line_to((4, 3)).line_to((5, 3)).line_to((7, 2)).line_to((8, 1)).line_to((8, 0)).close()
    img = np.zeros((4, 9))
    b.build().mask(rast, rb.Transform.identity(), rb.FillRule.EvenOdd, img)
    expected = np.array([
        0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0,
        0.0, 0.5, 1.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.0,
        0.0, 0.0, 0.25, 0.75, 1.0, 0.75, 0.25, 0.0, 0.0,
        0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]).reshape(4, 9)
    assert np.abs(img - expected).max() <= 1e-6
    # star + boxes
    op = O.OraclePath.parse("M50,0 21,90 98,35 2,35 79,90z M110,0 h90 v90 h-90z M130,20 h50 v50 h-50 z M210,0  h90 v90 h-90 z M230,20 v50 h50 v-50 z")
    p = rb.Path(*op.export())
    (w, h), tr, _ = op.size()
    img = np.zeros((h, w))
    p.mask(rast, tr, rb.FillRule.EvenOdd, img)
    assert abs(img[50, 50]) < 1e-6 and abs(img[50, 150]) < 1e-6 and abs(img[50, 250]) < 1e-6
    assert abs(img.sum() - 13130.0) < 1.0
    img[:] = 0
    p.mask(rast, tr, rb.FillRule.NonZero, img)
    assert abs(img[50, 50] - 1) < 1e-6 and abs(img[50, 150] - 1) < 1e-6 and abs(img[50, 250]) < 1e-6
    assert abs(img.sum() - 16492.5) < 1.0


def test_mask_edges_and_clipping(rast):
    """H8 edge semantics: x<0 fold, right-edge clamp, y clipping, overflow column, ragged sizes"""
    p = assets.load_path("squirrel")
    op = opath(p)
    for (w, h, tr) in [(37, 23, O.translate(-30.0, -20.0)), (130, 9, O.translate(5.0, -40.0)), (1, 50, O.IDENTITY),
                       (1030, 70, O.transform_mul(O.translate(-40.0, -300.0), O.scale(12.0, 5.0))), (64, 64, O.scale(0.2, 0.2)),
                       (3, 3, O.translate(-50.0, -50.0))]:
        for rule in (rb.FillRule.NonZero, rb.FillRule.EvenOdd):
            img = np.zeros((h, w))
            rast.mask(p, tr, img, rule)
            ref = np.zeros((h, w))
            op.mask(tr, int(rule), ref)
            assert np.abs(img - ref).max() <= COV_TOL, (w, h, rule)


def test_mask_strided_view(rast):
    p = assets.load_path("tv")
    e = assets.expected()["paths"]["tv"]
    w, h = e["size"]
    tr = np.array(e["size_tr"])
    big = np.zeros((h + 4, 2 * (w + 3)))
    view = big[2:2 + h, 1:1 + 2 * w:2]  # row and column strides
    rast.mask(p, tr, view, rb.FillRule.NonZero)
    ref = np.zeros((h, w))
    opath(p).mask(tr, O.NONZERO, ref)
    assert np.abs(view - ref).max() <= COV_TOL
    mask = np.ones_like(big, dtype=bool)
    mask[2:2 + h, 1:1 + 2 * w:2] = False
    assert (big[mask] == 0).all()


def test_mask_iter_matches_oracle(rast):
    p = assets.load_path("rust")
    e = assets.expected()["paths"]["rust"]
    w, h = e["size"]
    tr = np.array(e["size_tr"])
    got = rast.mask_iter(p, tr, rb.Size(w, h), rb.FillRule.NonZero)
    ref = opath(p).mask_iter(tr, w, h, O.NONZERO)
    gd = {(x, y): a for x, y, a in got}
    rd = {(x, y): a for x, y, a in ref}
    # pixels may differ only where alpha is at the 1e-6 drop threshold
    for k in set(gd) | set(rd):
        assert abs(gd.get(k, 0.0) - rd.get(k, 0.0)) <= COV_TOL
    assert got == sorted(got, key=lambda t: (t[1], t[0]))
    assert rast.mask_iter(p, tr, rb.Size(0, 10), rb.FillRule.NonZero) == []


def test_empty_path_and_zero_size(rast):
    img = np.zeros((5, 7))
    rast.mask(rb.Path.empty(), rb.Transform.identity(), img, rb.FillRule.NonZero)
    assert (img == 0).all()
    assert len(rast.flatten(rb.Path.empty())) == 0
    rast.mask(assets.load_path("squirrel"), rb.Transform.identity(), np.zeros((0, 0)), rb.FillRule.NonZero)


def test_deterministic(rast):
    p = assets.load_path("ava")
    e = assets.expected()["paths"]["ava"]
    w, h = e["size"]
    tr = np.array(e["size_tr"])
    a = np.zeros((h, w), dtype=np.float32)
    b = np.zeros((h, w), dtype=np.float32)
    rast.mask(p, tr, a, rb.FillRule.NonZero)
    for _ in range(3):
        rast.mask(p, tr, b, rb.FillRule.NonZero)
        assert np.array_equal(a, b)  # bit-identical across runs (fixed-point atomics)


def test_c2_material_4096(rast):
    """BASELINE config 2: material.path fitted to 4096x4096, nonzero and even-odd"""
    p = assets.load_path("material")
    c2 = assets.expected()["paths"]["material"]["c2"]
    tr = np.array(c2["tr"])
    w, h = c2["size"]
    op = opath(p)
    for rule in (rb.FillRule.NonZero, rb.FillRule.EvenOdd):
        img = np.zeros((h, w), dtype=np.float32)
        rast.mask(p, tr, img, rule)
        ref = np.zeros((h, w))
        op.mask_threads(tr, int(rule), ref, threads=8)
        assert np.abs(img - ref).max() <= COV_TOL
    assert rast.last_counts()["lines"] == c2["lines"]


def test_mask_pinned_f64_matches_pageable(rast):
    """rgpu_mask into a PINNED f64 image takes the split path (top rows widened on the device and DMA'd as f64, bottom
    rows widened by host threads): same pixels as the pageable path, for dense rows and for a row-strided view, over
    several calls (the split moves from call to call)."""
    p = assets.load_path("rust")
    tr = np.array(assets.expected()["paths"]["rust"]["size_tr"]) * 3.0
    w, h = 700, 900
    ref = np.zeros((h, w))
    rast.mask(p, tr, ref, rb.FillRule.NonZero)  # pageable: host widening only
    pinned = rast.host_alloc((h, w + 16), np.float64)
    for _ in range(6):
        pinned[:] = -1.0
        rast.mask(p, tr, pinned[:, :w], rb.FillRule.NonZero)
        assert np.array_equal(pinned[:, :w], ref)
        assert (pinned[:, w:] == -1.0).all()
    oref = np.zeros((h, w))
    opath(p).mask(tr, O.NONZERO, oref)
    assert np.abs(ref - oref).max() <= COV_TOL


def test_two_pass_fallback_matches(monkeypatch):
    """The exact count -> scan -> emit binning (taken when fixed-capacity bins would exceed their memory budget) gives the
    same pixels as the single-pass scheme: forced through its A/B switch on a fresh context."""
    p = assets.load_path("rust")
    tr = np.array(assets.expected()["paths"]["rust"]["size_tr"]) * 2.0
    w, h = 1500, 700  # two column chunks: the carry look-back runs too
    a = np.zeros((h, w))
    r1 = rb.GpuRasterizer()
    r1.mask(p, tr, a, rb.FillRule.EvenOdd)
    r1.close()
    monkeypatch.setenv("RGPU_TWO_PASS", "1")
    r2 = rb.GpuRasterizer()
    b = np.zeros((h, w))
    r2.mask(p, tr, b, rb.FillRule.EvenOdd)
    launches = r2.last_counts()["launches"]
    r2.close()
    assert np.array_equal(a, b)
    assert launches >= 4  # count, scan, emit, raster
    ref = np.zeros((h, w))
    opath(p).mask(tr, O.EVENODD, ref)
    assert np.abs(a - ref).max() <= COV_TOL


def test_staged_item_lists_follow_the_path_structure(rast):
    """The host-buffer entry points keep the item lists of the previous call on the device when the path's structure is
    unchanged (only the control points are uploaded again).  Paths with the same counts but different structure, and the
    same structure with different points, must each give their own result."""
    pts = np.array([[2.0, 2.0], [30.0, 4.0], [30.0, 4.0], [40.0, 30.0], [8.0, 36.0], [8.0, 36.0], [2.0, 2.0]])
    a = rb.Path(pts, np.array([2, 3, 2], dtype=np.uint8), np.array([0, 3], dtype=np.uint32), np.array([1], dtype=np.uint8))
    b = rb.Path(pts, np.array([3, 2, 2], dtype=np.uint8), np.array([0, 3], dtype=np.uint32), np.array([1], dtype=np.uint8))   # kinds permuted
    c = rb.Path(pts, np.array([2, 3, 2], dtype=np.uint8), np.array([0, 1, 3], dtype=np.uint32)[:3], np.array([1, 0], dtype=np.uint8))  # two subpaths
    a2 = rb.Path(pts * 1.2 + 1.0, a.kinds, a.subpath_offsets, a.closed)  # same structure, other points
    tr = rb.Transform.identity()
    for p in (a, b, a, c, a2, a, b, c, a2):
        for close in (True, False):
            assert np.array_equal(rast.flatten(p, tr, close), opath(p).flatten(np.array(O.IDENTITY), close=close))
        img = np.zeros((48, 52))
        rast.mask(p, tr, img, rb.FillRule.NonZero)
        ref = np.zeros((48, 52))
        opath(p).mask(O.IDENTITY, O.NONZERO, ref)
        assert np.abs(img - ref).max() <= COV_TOL


def test_shallow_lines_next_to_a_chunk_boundary_at_large_y(rast):
    """the binning used to estimate a line's columns per band in f32 with 1.5 px of slack; at y ~ 32000 an f32 row
    coordinate is 2e-3 off, which a shallow slope turns into many pixels, so a chunk the line reaches could miss it.  Long,
    nearly horizontal edges that end a fraction of a pixel before / after the 1024-column boundary, far down a tall canvas."""
    w, h = 1100, 32768
    b = rb.Path.builder()
    rng = np.random.default_rng(5)
    for k in range(40):
        y = 31000.0 + 40.0 * k + float(rng.uniform(0, 1))
        xe = 1024.0 + float(rng.uniform(-1.6, 1.6))       # the shallow edge ends next to the chunk boundary
        x0 = float(rng.uniform(3.0, 200.0))
        dy = float(rng.uniform(0.002, 0.9))               # slopes dx/dy between ~1e3 and ~5e5
        b.move_to((x0, y)).line_to((xe, y + dy)).line_to((x0 + 5.0, y + 12.0)).close()
        b.move_to((xe + 30.0, y + 20.0)).line_to((1024.0 - float(rng.uniform(0.0, 3.0)), y + 20.0 + dy)).line_to((1090.0, y + 31.0)).close()
    p = b.build()
    img = np.zeros((h, w), dtype=np.float32)
    rast.mask(p, rb.Transform.identity(), img, rb.FillRule.NonZero)
    ref = np.zeros((h, w))
    opath(p).mask_threads(O.IDENTITY, O.NONZERO, ref, threads=8)
    assert np.abs(img[30900:] - ref[30900:]).max() <= COV_TOL
    assert not img[:30900].any()


def test_too_many_gradient_stops_is_an_error_everywhere(rast):
    """rgpu_render_batch used to truncate a gradient to RGPU_MAX_STOPS silently; every entry point now rejects it."""
    stops = [(i / 40.0, [1.0, 0.0, 0.0, 1.0]) for i in range(41)]
    grad = rb.GradLinear(stops, rb.Units.UserSpaceOnUse, True, rb.GradSpread.Pad, rb.Transform.identity(), (0, 0), (10, 10))
    p = assets.load_path("squirrel")
    dp = rast.upload(p)
    canvas = rast.device_alloc(100 * 100 * 16)
    with pytest.raises(rb.RgpuError) as e:
        rast.render_batch([rb.Job(dp, rb.Transform.identity(), rb.FillRule.NonZero, ffi.JOB_FILL, canvas, 100, 100, 100, paint=grad)])
    assert e.value.code == ffi.ERR_INVALID
    rast.device_free(canvas)


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (7, 64), (130, 33), (257, 129), (1000, 700), (4099, 53)])
def test_mask_iter_device_compaction(rast, w, h):
    """rgpu_mask_iter compacts the yielded pixels on the device (compact.cu): the list must be exactly the non-zero pixels of the
    dense coverage (same launch path, so bit-identical alphas) in row-major order — for widths on both sides of the float4 /
    128-pixel / 4096-pixel block boundaries — and must match the oracle's iterator within the coverage tolerance."""
    p = assets.load_path("squirrel")
    e = assets.expected()["paths"]["squirrel"]["c1"]
    w0, h0 = e["size"]
    tr = np.array(e["tr"]).copy()
    # scale the asset's own fit transform to this canvas
    sx, sy = w / w0, h / h0
    tr = np.array([tr[0] * sx, tr[1] * sx, tr[2] * sx, tr[3] * sy, tr[4] * sy, tr[5] * sy])
    for rule, orule in ((rb.FillRule.NonZero, O.NONZERO), (rb.FillRule.EvenOdd, O.EVENODD)):
        got = rast.mask_iter_array(p, tr, rb.Size(w, h), rule)
        dense = rast.coverage(p, tr, rb.Size(w, h), rule)
        ys, xs = np.nonzero(dense)
        assert len(got) == len(xs)
        assert np.array_equal(got["x"], xs.astype(np.uint64)) and np.array_equal(got["y"], ys.astype(np.uint64))
        assert np.array_equal(got["alpha"], dense[ys, xs].astype(np.float64))
        ref = np.zeros((h, w))
        for x, y, a in opath(p).mask_iter(tr, w, h, orule):
            ref[y, x] = a
        mine = np.zeros((h, w))
        mine[got["y"].astype(np.int64), got["x"].astype(np.int64)] = got["alpha"]
        assert np.abs(mine - ref).max() <= COV_TOL


def test_mask_iter_capacity_and_empty(rast):
    """A buffer that is too small takes the first `cap` pixels and reports the full count (RGPU_ERR_CAPACITY); an empty path
    and an off-canvas path yield nothing."""
    p = assets.load_path("squirrel")
    e = assets.expected()["paths"]["squirrel"]["c1"]
    w, h = e["size"]
    tr = np.array(e["tr"])
    full = rast.mask_iter_array(p, tr, rb.Size(w, h), rb.FillRule.NonZero)
    assert len(full) > 100
    L = ffi.lib()
    import ctypes as C
    c = p._c()
    n = C.c_size_t()
    cap = 37
    buf = np.zeros(cap + 3, dtype=rast.PIXEL_DTYPE)
    rc = L.rgpu_mask_iter(rast.ctx, C.byref(c), tr.ctypes.data_as(C.POINTER(C.c_double)), w, h, int(rb.FillRule.NonZero),
                          C.cast(buf.ctypes.data, C.POINTER(ffi.CPixel)), cap, C.byref(n))
    assert rc == ffi.ERR_CAPACITY and n.value == len(full)
    assert np.array_equal(buf[:cap], full[:cap]) and (buf["alpha"][cap:] == 0).all()
    rc = L.rgpu_mask_iter(rast.ctx, C.byref(c), tr.ctypes.data_as(C.POINTER(C.c_double)), w, h, int(rb.FillRule.NonZero), None, 0, C.byref(n))
    assert rc == ffi.ERR_CAPACITY and n.value == len(full)
    assert len(rast.mask_iter_array(p, tr, rb.Size(w, h), rb.FillRule.NonZero, cap=5)) == len(full)  # retried with the exact count
    assert len(rast.mask_iter_array(rb.Path.empty(), tr, rb.Size(w, h), rb.FillRule.NonZero)) == 0
    far = np.array([tr[0], tr[1], tr[2] + 1e6, tr[3], tr[4], tr[5]])
    assert len(rast.mask_iter_array(p, far, rb.Size(w, h), rb.FillRule.NonZero)) == 0


def test_mask_f64_run_coded_strided_rows(rast):
    """rgpu_mask on a canvas of 4 Mpixel and more: the f64 rows are rebuilt by host threads from the run-coded download
    (download_runcoded).  A row-strided view (rows of 2300 + 13 doubles) must get the bits of the f32 device mask widened
    exactly — the reference point is the dense f32 download assembled from small band calls — the padding stays untouched,
    and the result is within the coverage tolerance of the oracle."""
    p = assets.load_path("material")
    c2 = assets.expected()["paths"]["material"]["c2"]
    w, h = 2300, 1900
    tr = np.array(c2["tr"]) * np.array([w / c2["size"][0]] * 3 + [h / c2["size"][1]] * 3)
    big = np.full((h, w + 13), -2.0)
    rast.mask(p, tr, big[:, :w], rb.FillRule.NonZero)
    assert (big[:, w:] == -2.0).all()
    dense = np.zeros((h, w), dtype=np.float32)
    for b in range(8):
        rast.mask_banded(p, tr, dense, rb.FillRule.NonZero, n_bands=8, band_first=b, band_count=1)  # 0.5 Mpixel each: dense copies
    assert np.array_equal(big[:, :w], dense.astype(np.float64))
    ref = np.zeros((h, w))
    opath(p).mask_threads(tr, O.NONZERO, ref, threads=8)
    assert np.abs(big[:, :w] - ref).max() <= COV_TOL
