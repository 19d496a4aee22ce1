"""Batch (BASELINE config 4) and band-sharded (config 5) forms of the fill path on the GPU, through the C ABI,
against the CPU oracle; plus size-independent properties at the full BASELINE sizes.  Run: pytest -m gpu."""
import numpy as np
import pytest

import bench
import oracle as O
import rasterize_b200 as rb
import assets
from rasterize_b200 import ffi, sharding

pytestmark = pytest.mark.gpu
COV_TOL = 1e-4


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def opath(p):
    return O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)


def test_c4_glyph_batch_mask_and_fill(rast):
    """config 4 (reduced count): independent 64x64 glyphs in ONE batch launch, mask and solid-black fill"""
    n = 257
    glyphs = [bench.glyph_path(rb, i + 1) for i in range(n)]
    dps = [rast.upload(g) for g in glyphs]
    ident = rb.Transform.identity()
    # masks
    slab = rast.device_alloc(n * 4096 * 4)
    jobs = [rb.Job(dps[i], ident, rb.FillRule.NonZero, ffi.JOB_MASK, slab, 64, 64, 64, origin=i * 4096) for i in range(n)]
    rast.render_batch(jobs, independent=True)
    got = rast.to_host(slab, (n, 64, 64), np.float32)
    lines = 0
    for i in range(n):
        ref = np.zeros((64, 64))
        og = opath(glyphs[i])
        og.mask(O.IDENTITY, O.NONZERO, ref)
        lines += len(og.flatten())
        assert np.abs(got[i] - ref).max() <= COV_TOL, i
    assert rast.last_counts()["lines"] == lines
    rast.device_free(slab)
    # fills (solid black over transparent, even-odd) into one [n,64,64,4] slab
    slab = rast.device_alloc(n * 4096 * 16)
    rast.device_zero(slab, n * 4096 * 16)
    black = rb.LinColor(0.0, 0.0, 0.0, 1.0)
    jobs = [rb.Job(dps[i], ident, rb.FillRule.EvenOdd, ffi.JOB_FILL, slab, 64, 64, 64, origin=i * 4096, paint=black) for i in range(n)]
    rast.render_batch(jobs, independent=True)
    got = rast.to_host(slab, (n, 64, 64, 4), np.float32)
    paint = O.OraclePaint.solid([0, 0, 0, 1])
    for i in range(0, n, 7):
        ref = np.zeros((64, 64, 4), dtype=np.float32)
        opath(glyphs[i]).fill(O.IDENTITY, O.EVENODD, paint, ref)
        assert np.abs(got[i] - ref).max() <= 2e-4, i
    rast.device_free(slab)


def test_batch_mixed_sizes_and_ragged(rast):
    """one independent batch mixing canvas sizes, empty paths, zero-sized jobs and both rules"""
    names = ["squirrel", "tv", "rust"]
    ex = assets.expected()["paths"]
    jobs, metas, keep = [], [], []
    total = 0
    for k, name in enumerate(names * 2):
        e = ex[name]
        w, h = e["size"]
        metas.append((name, w, h, total, k % 2))
        total += w * h
    slab = rast.device_alloc(total * 4)
    rast.device_zero(slab, total * 4)
    for name, w, h, off, rule in metas:
        dp = rast.upload(assets.load_path(name))
        keep.append(dp)
        jobs.append(rb.Job(dp, np.array(ex[name]["size_tr"]), rb.FillRule(rule), ffi.JOB_MASK, slab, w, h, w, origin=off))
    empty = rast.upload(rb.Path.empty())
    jobs.insert(2, rb.Job(empty, rb.Transform.identity(), rb.FillRule.NonZero, ffi.JOB_MASK, slab, 0, 0, 0))
    rast.render_batch(jobs, independent=True)
    flat = rast.to_host(slab, (total,), np.float32)
    for name, w, h, off, rule in metas:
        ref = np.zeros((h, w))
        opath(assets.load_path(name)).mask(np.array(ex[name]["size_tr"]), rule, ref)
        assert np.abs(flat[off:off + w * h].reshape(h, w) - ref).max() <= COV_TOL, (name, rule)
    rast.device_free(slab)


def test_c5_band_sharding_reduced(rast):
    """config 5 logic at 8192 px: each band rendered on its own with a band-local translate equals the rows of the
    oracle's full-canvas mask (what 4 GPUs would produce), and equals the GPU's own full-canvas mask bit for bit"""
    p = assets.load_path("tv_stroked")
    op = opath(p)
    (w, h), tr = O.fit_size(op.bbox(), 8192, 8192)
    ref = np.zeros((h, w))
    op.mask_threads(tr, O.NONZERO, ref, threads=8)
    full = np.zeros((h, w), dtype=np.float32)
    rast.mask(p, tr, full, rb.FillRule.NonZero)
    assert np.abs(full - ref).max() <= COV_TOL
    world = 4
    for r in range(world):
        y0, y1 = sharding.band_rows(h, r, world)
        band = np.zeros((y1 - y0, w), dtype=np.float32)
        rast.mask(p, sharding.band_transform(tr, y0), band, rb.FillRule.NonZero)
        assert np.abs(band - ref[y0:y1]).max() <= COV_TOL
        assert np.array_equal(band, full[y0:y1])


def test_c5_full_size_band_properties(rast):
    """config 5 at the BASELINE size (32768 wide): one 4096-row band on the device; size-independent checks —
    line count, rows outside the glyph are exactly zero, coverage in [0,1], band == oracle on sampled rows"""
    p = assets.load_path("tv_stroked")
    c5 = assets.expected()["paths"]["tv_stroked"]["c5"]
    w, hfull = c5["size"]
    tr = np.array(c5["tr"])
    y0, y1 = sharding.band_rows(hfull, 2, 8)
    h = y1 - y0
    dp = rast.upload(p)
    canvas = rast.device_alloc(w * h * 4)
    rast.render_batch([rb.Job(dp, sharding.band_transform(tr, y0), rb.FillRule.NonZero, ffi.JOB_MASK, canvas, w, h, w)], independent=True)
    assert rast.last_counts()["lines"] == c5["lines"]
    band = rast.to_host(canvas, (h, w), np.float32)
    rast.device_free(canvas)
    assert band.min() >= 0.0 and band.max() <= 1.0
    # oracle on a thin slice of the band (64 rows) via the same band-translate trick
    op = opath(p)
    for off in (0, 1777, h - 64):
        ref = np.zeros((64, w))
        op.mask(sharding.band_transform(tr, y0 + off), O.NONZERO, ref)
        assert np.abs(band[off:off + 64] - ref).max() <= COV_TOL
    # band 0 of 8 holds no geometry at all: exactly zero
    b0 = sharding.band_rows(hfull, 0, 8)
    z = np.ones((64, w), dtype=np.float32)
    rast.mask(p, sharding.band_transform(tr, b0[0]), z, rb.FillRule.NonZero)
    assert (z == 0).all()


def test_prepared_batch_resubmission_is_stable(rast):
    """bench.py's loop: the same prepared batch submitted repeatedly (async) gives bit-identical results"""
    p = assets.load_path("ava")
    e = assets.expected()["paths"]["ava"]
    w, h = e["size"]
    dp = rast.upload(p)
    canvas = rast.device_alloc(w * h * 4)
    prep = rast.prepare_batch([rb.Job(dp, np.array(e["size_tr"]), rb.FillRule.NonZero, ffi.JOB_MASK, canvas, w, h, w)])
    rast.submit_prepared(prep, independent=True, sync=True)
    a = rast.to_host(canvas, (h, w), np.float32)
    for _ in range(5):
        rast.submit_prepared(prep, independent=True, sync=False)
    rast.batch_status()
    b = rast.to_host(canvas, (h, w), np.float32)
    assert np.array_equal(a, b)
    rast.device_free(canvas)


def test_render_jobs_create_their_canvas(rast):
    """RGPU_JOB_RENDER (`ImageOwned::new_default` + `Rasterizer::fill`) is bit-identical to RGPU_JOB_FILL on a zeroed
    canvas, whatever the canvas held before: fused small-canvas kernel (64x64 glyphs) and tiled path (300x200)."""
    n = 65
    glyphs = [bench.glyph_path(rb, i + 1) for i in range(n)]
    dps = [rast.upload(g) for g in glyphs]
    ident = rb.Transform.identity()
    grad = rb.GradRadial([(0.0, [0.8, 0.1, 0.1, 0.9]), (0.5, [0.1, 0.7, 0.2, 1.0]), (1.0, [0.0, 0.0, 0.3, 0.3])], rb.Units.UserSpaceOnUse, False,
                         rb.GradSpread.Reflect, rb.Transform.identity(), (30.0, 34.0), 25.0, (26.0, 30.0), 3.0)
    for (w, h, tr), paint in (((64, 64, ident), rb.LinColor(0.1, 0.2, 0.3, 0.6)), ((300, 200, rb.Transform.new_scale(300 / 64.0, 200 / 64.0)), rb.LinColor(0.1, 0.2, 0.3, 0.6)),
                              ((64, 64, ident), grad), ((61, 47, ident), grad)):
        nb = n * w * h * 16
        slab = rast.device_alloc(nb)
        rast.device_zero(slab, nb)
        jobs = [rb.Job(dps[i], tr, rb.FillRule(i % 2), ffi.JOB_FILL, slab, w, h, w, origin=i * w * h, paint=paint) for i in range(n)]
        rast.render_batch(jobs, independent=True)
        want = rast.to_host(slab, (n, h, w, 4), np.float32)
        rast.to_device(slab, np.full((n, h, w, 4), 3.0, dtype=np.float32))  # garbage the job must not read
        jobs = [rb.Job(dps[i], tr, rb.FillRule(i % 2), ffi.JOB_RENDER, slab, w, h, w, origin=i * w * h, paint=paint) for i in range(n)]
        rast.render_batch(jobs, independent=True)
        got = rast.to_host(slab, (n, h, w, 4), np.float32)
        assert np.array_equal(got, want)
        assert got.any()
        rast.device_free(slab)
    # an empty path still creates (clears) its canvas
    slab = rast.device_alloc(64 * 64 * 16)
    rast.to_device(slab, np.full((64, 64, 4), 3.0, dtype=np.float32))
    empty = rast.upload(rb.Path.empty())
    rast.render_batch([rb.Job(empty, ident, rb.FillRule.NonZero, ffi.JOB_RENDER, slab, 64, 64, 64, paint=grad)], independent=True)
    assert not rast.to_host(slab, (64, 64, 4), np.float32).any()
    rast.device_free(slab)


def test_c5_whole_canvas_properties(rast, monkeypatch):
    """config 5 at the BASELINE size, the WHOLE 32768 x 32768 canvas through the host-buffer call bench.py times
    (rgpu_mask_banded_host, 4.29 GB of f32): the run-coded download and dense copies (RGPU_E2E_RUNCODE=0, read per call) give the
    same bytes for the same jobs; coverage in [0, 1]; rows of the empty top band are exactly zero; sampled rows match the oracle;
    the run-coded call moves a fraction of the bytes.  Two blocks of bands rendered separately (what two GPUs do) stitch to the
    whole-canvas image BIT FOR BIT: a band job flattens the path with the canvas transform and subtracts its integer row origin
    from the finished lines (JobDev::y_org).  (With a translate(0, -y0) folded into the transform — what a caller of the plain
    trait call would do — the transformed control points round differently and 6 of the 1.07e9 pixels differ by one f32 ulp.)"""
    p = assets.load_path("tv_stroked")
    c5 = assets.expected()["paths"]["tv_stroked"]["c5"]
    w, h = c5["size"]
    tr = np.array(c5["tr"])
    whole = np.empty((h, w), dtype=np.float32)
    whole[:] = -1.0
    rast.mask_banded(p, tr, whole, rb.FillRule.NonZero, n_bands=64)
    _, d2h = rast.last_transfer_bytes()
    assert d2h < 0.25 * w * h * 4
    assert whole.min() == 0.0 and whole.max() == 1.0
    assert not whole[: h // 8].any()  # band 0 of 8 holds no geometry
    other = np.empty((h, w), dtype=np.float32)
    other[:] = -1.0
    monkeypatch.setenv("RGPU_E2E_RUNCODE", "0")
    rast.mask_banded(p, tr, other, rb.FillRule.NonZero, n_bands=64)
    _, d2h_dense = rast.last_transfer_bytes()
    monkeypatch.delenv("RGPU_E2E_RUNCODE")
    assert d2h_dense == w * h * 4
    assert np.array_equal(whole.view(np.uint32), other.view(np.uint32))
    # two blocks of bands, the first with dense copies and the second run-coded, into one image
    other[:] = -1.0
    monkeypatch.setenv("RGPU_E2E_RUNCODE", "0")
    rast.mask_banded(p, tr, other, rb.FillRule.NonZero, n_bands=64, band_first=0, band_count=40)
    monkeypatch.delenv("RGPU_E2E_RUNCODE")
    rast.mask_banded(p, tr, other, rb.FillRule.NonZero, n_bands=64, band_first=40, band_count=24)
    assert np.array_equal(whole.view(np.uint32), other.view(np.uint32))
    # the same block through the plain mask call with a translated transform: within one f32 ulp, in a handful of pixels
    y0, y1 = sharding.band_rows(h, 5, 8)
    band = np.zeros((y1 - y0, w), dtype=np.float32)
    rast.mask(p, sharding.band_transform(tr, y0), band, rb.FillRule.NonZero)
    ys, xs = np.nonzero(band.view(np.uint32) != whole[y0:y1].view(np.uint32))
    assert len(ys) <= 64 and (len(ys) == 0 or np.abs(band[ys, xs] - whole[y0:y1][ys, xs]).max() <= 1.2e-7)
    op = opath(p)
    for y0 in (h // 2 - 32, 11111, h - 5000):
        ref = np.zeros((64, w))
        op.mask(sharding.band_transform(tr, y0), O.NONZERO, ref)
        assert np.abs(whole[y0:y0 + 64] - ref).max() <= COV_TOL
