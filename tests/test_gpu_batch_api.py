"""Batches of independent paths through the round-2 entry points (rgpu_path_upload_batch, rgpu_batch_*,
rgpu_fill_batch_host, rgpu_mask_banded_host, rgpu_multi_*) and the rare paths of the fused small-canvas kernel, against
the CPU oracle and against the single-call results.  Run: pytest -m gpu."""
import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb
import assets
from rasterize_b200 import ffi, synth

pytestmark = pytest.mark.gpu
COV_TOL = 1e-4


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def opath(p):
    return O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)


def job_table(handles, canvas, mode, rule, n, w, h, paint_ptr=0):
    t = np.zeros(n, dtype=rb.JOB_DTYPE)
    t["path"] = handles
    t["tr"] = np.array([1.0, 0, 0, 0, 1.0, 0])
    t["fill_rule"] = int(rule)
    t["mode"] = mode
    t["paint"] = paint_ptr
    t["canvas"] = canvas
    t["origin"] = np.arange(n, dtype=np.uint64) * np.uint64(w * h)
    t["row_stride"] = w
    t["width"] = w
    t["height"] = h
    return t


def test_upload_batch_and_prepared_batch_match_per_path_jobs(rast):
    """One upload + a device-resident job table give the same bits as per-path uploads through rgpu_render_batch."""
    n = 300
    pb = synth.glyph_batch(1, n)
    dpb = rast.upload_batch(pb)
    slab = rast.device_alloc(n * 4096 * 4)
    prepared = rast.prepare_job_table(job_table(dpb.handles(), slab, ffi.JOB_MASK, rb.FillRule.NonZero, n, 64, 64))
    prepared.render()
    rast.batch_status()
    got = rast.to_host(slab, (n, 64, 64), np.float32)
    lines = rast.last_counts()["lines"]
    prepared.render()  # a second submission of the same handle
    rast.batch_status()
    assert np.array_equal(got, rast.to_host(slab, (n, 64, 64), np.float32))
    dps = [rast.upload(pb.path(i)) for i in range(n)]
    rast.device_zero(slab, n * 4096 * 4)
    rast.render_batch([rb.Job(dps[i], rb.Transform.identity(), rb.FillRule.NonZero, ffi.JOB_MASK, slab, 64, 64, 64, origin=i * 4096)
                       for i in range(n)], independent=True)
    assert np.array_equal(got, rast.to_host(slab, (n, 64, 64), np.float32))
    assert rast.last_counts()["lines"] == lines
    for i in range(0, n, 13):
        ref = np.zeros((64, 64))
        opath(pb.path(i)).mask(O.IDENTITY, O.NONZERO, ref)
        assert np.abs(got[i] - ref).max() <= COV_TOL, i
    prepared.free()
    rast.device_free(slab)


@pytest.mark.parametrize("n", [1, 37, 5000])
def test_fill_batch_host_lincolor_rgba_coverage(rast, n):
    """rgpu_fill_batch_host (chunked, downloads overlapped): LinColor == per-glyph Path::fill of the oracle, RGBA8 ==
    conversion of that image, coverage == dense mask_iter.  5000 glyphs span several chunks of the slab ring."""
    pb = synth.glyph_batch(11, n)
    black = rb.LinColor(0.0, 0.0, 0.0, 1.0)
    lin = np.full((n, 64, 64, 4), 7.0, dtype=np.float32)
    rast.fill_batch_host(pb, rb.FillRule.NonZero, black, 64, 64, lin)
    rgba = np.zeros((n, 64, 64, 4), dtype=np.uint8)
    rast.fill_batch_host(pb, rb.FillRule.NonZero, black, 64, 64, rgba)
    cov = np.zeros((n, 64, 64), dtype=np.float32)
    rast.fill_batch_host(pb, rb.FillRule.NonZero, None, 64, 64, cov)
    paint = O.OraclePaint.solid([0, 0, 0, 1])
    for i in sorted(set([0, n - 1] + list(range(0, n, max(1, n // 23))))):
        g = opath(pb.path(i))
        ref = np.zeros((64, 64, 4), dtype=np.float32)
        g.fill(O.IDENTITY, O.NONZERO, paint, ref)
        assert np.abs(lin[i] - ref).max() <= 2e-4, i
        assert np.abs(rgba[i].astype(int) - O.lin_to_rgba(lin[i]).astype(int)).max() == 0, i
        assert np.abs(cov[i] - lin[i][:, :, 3]).max() == 0.0, i
    # every image of the batch was written (no stale 7.0 left) and identical glyph data gives identical bits per call
    assert lin.max() <= 1.0 + 1e-6
    lin2 = np.empty_like(lin)
    rast.fill_batch_host(pb, rb.FillRule.NonZero, black, 64, 64, lin2)
    assert np.array_equal(lin, lin2)


def test_fill_batch_host_transforms_gradient_and_large_canvas(rast):
    """Per-path transforms, a gradient paint, even-odd, and canvases beyond the fused kernel (tiled path, 200 x 150)."""
    n = 9
    pb = synth.glyph_batch(3, n)
    trs = np.array([[1.5 + 0.1 * i, 0.2, 3.0 * i, -0.1, 2.0, 5.0] for i in range(n)])
    stops = [(0.0, [1.0, 0.0, 0.0, 1.0]), (1.0, [0.0, 0.0, 0.5, 0.5])]
    grad = rb.GradLinear(stops, rb.Units.UserSpaceOnUse, True, rb.GradSpread.Reflect, rb.Transform.identity(), (0, 0), (40, 30))
    out = np.zeros((n, 150, 200, 4), dtype=np.float32)
    rast.fill_batch_host(pb, rb.FillRule.EvenOdd, grad, 200, 150, out, trs=trs)
    og = O.OraclePaint.linear(stops, (0, 0), (40, 30), units=0, linear_colors=True, spread=2)
    for i in range(n):
        ref = np.zeros((150, 200, 4), dtype=np.float32)
        opath(pb.path(i)).fill(trs[i], O.EVENODD, og, ref)
        assert np.abs(out[i] - ref).max() <= 3e-4, i


def test_small_canvas_rare_paths(rast):
    """The fused kernel's f64 / deep / overflow paths: glyphs scaled far beyond the canvas (curves leave the columns and
    subdivide 8+ levels), shifted off every edge, quads, open subpaths, line-only paths, tiny canvases."""
    pb = synth.glyph_batch(101, 12)
    cases = []
    for i in range(12):
        p = pb.path(i)
        cases.append((p, [9.0, 0, -200.0 - 10 * i, 0, 9.0, -150.0], 64, 64))      # deep curves, mostly outside
        cases.append((p, [1.0, 0, -30.0, 0, 1.0, 17.5], 64, 64))                  # crosses x < 0 and the bottom
        cases.append((p, [1.0, 0, 25.0, 0, 1.0, -40.25], 64, 64))                 # crosses x > width and the top
        cases.append((p, [0.45, 0, 1.0, 0, 0.3, 2.0], 33, 21))                    # odd small canvas
        cases.append((p, [40.0, 0, -900.0, 0, 40.0, -1200.0], 64, 64))            # a few huge pieces, far coordinates
    b = rb.Path.builder()
    b.move_to((5, 5)).quad_to((60, 0), (58, 40)).quad_to((30, 70), (8, 50)).line_to((2, 30))   # open: closed by the rasterizer
    b.move_to((20, 20)).line_to((40, 22)).line_to((38, 44)).line_to((18, 40)).close()
    quads = b.build()
    cases += [(quads, [1.0, 0, 0, 0, 1.0, 0], 64, 64), (quads, [3.0, 0, -50.0, 0, 3.0, -40.0], 64, 64), (quads, [1.0, 0, 0, 0, 1.0, 0], 17, 64)]
    for rule, orule in ((rb.FillRule.NonZero, O.NONZERO), (rb.FillRule.EvenOdd, O.EVENODD)):
        dps = [rast.upload(c[0]) for c in cases]
        slab = rast.device_alloc(len(cases) * 4096 * 4)
        rast.device_zero(slab, len(cases) * 4096 * 4)
        jobs = [rb.Job(dps[k], c[1], rule, ffi.JOB_MASK, slab, c[2], c[3], c[2], origin=k * 4096) for k, c in enumerate(cases)]
        rast.render_batch(jobs, independent=True)
        got = rast.to_host(slab, (len(cases), 4096), np.float32)
        lines = 0
        for k, (p, tr, w, h) in enumerate(cases):
            ref = np.zeros((h, w))
            og = opath(p)
            og.mask(np.array(tr), orule, ref)
            lines += len(og.flatten(np.array(tr)))
            err = np.abs(got[k][: w * h].reshape(h, w) - ref).max()
            assert err <= COV_TOL, (k, tr, w, h, err)
        assert rast.last_counts()["lines"] == lines
        rast.device_free(slab)


def test_small_canvas_errors(rast):
    """NaN control points and unbounded subdivision are reported by the fused kernel like by the tiled path."""
    b = rb.Path.builder()
    b.move_to((1, 1)).cubic_to((float("nan"), 3), (5, 6), (7, 8)).close()
    img = np.zeros((32, 32), dtype=np.float32)
    with pytest.raises(rb.RgpuError) as e:
        rast.mask(b.build(), rb.Transform.identity(), img, rb.FillRule.NonZero)
    assert e.value.code == ffi.ERR_NAN
    b = rb.Path.builder()
    b.move_to((1, 1)).cubic_to((1e40, 3), (-1e40, 6), (7, 8)).close()
    with pytest.raises(rb.RgpuError) as e:
        rast.mask(b.build(), rb.Transform.identity(), img, rb.FillRule.NonZero)
    assert e.value.code == ffi.ERR_DEPTH


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mask_banded_is_bit_identical_to_single_mask(rast, dtype):
    """rgpu_mask_banded_host: any band decomposition stitches to the bits of the unsharded mask (rows are independent,
    reference src/rasterize.rs:421-469), for the stroked tv outline of config 5 at 4096 x 4096 and material at 1500 x 1100."""
    ex = assets.expected()["paths"]
    for name, (w, h) in (("tv_stroked", (4096, 4096)), ("material", (1500, 1100))):
        p = assets.load_path(name)
        c = ex[name]["c5" if name == "tv_stroked" else "c2"]
        tr = np.array(c["tr"]) * np.array([w / c["size"][0]] * 3 + [h / c["size"][1]] * 3)
        whole = np.zeros((h, w), dtype=dtype)
        rast.mask(p, tr, whole, rb.FillRule.NonZero)
        for n_bands in (1, 5, 64):
            img = np.full((h, w), -3.0, dtype=dtype)
            rast.mask_banded(p, tr, img, rb.FillRule.NonZero, n_bands=n_bands)
            assert np.array_equal(img, whole), (name, n_bands)
        # three "devices" with blocks of the 16 bands, into one image
        img = np.full((h, w), -3.0, dtype=dtype)
        rast.mask_banded(p, tr, img, rb.FillRule.NonZero, n_bands=16, band_first=5, band_count=6)
        assert (img == -3.0).any()
        rast.mask_banded(p, tr, img, rb.FillRule.NonZero, n_bands=16, band_first=0, band_count=5)
        rast.mask_banded(p, tr, img, rb.FillRule.NonZero, n_bands=16, band_first=11, band_count=5)
        assert np.array_equal(img, whole), name


def test_multi_gpu_matches_single_gpu():
    """rgpu_multi_*: shards by path / by scanline band over every device of the box (also exercised with one device, and
    with the same device listed twice when only one exists: two contexts, two host threads, disjoint output regions)."""
    n_dev = ffi.lib().rgpu_device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    single = rb.GpuRasterizer()
    multi = rb.MultiGpuRasterizer(devices)
    try:
        n = 1000
        pb = synth.glyph_batch(5, n)
        black = rb.LinColor(0.0, 0.0, 0.0, 1.0)
        a = np.zeros((n, 64, 64, 4), dtype=np.float32)
        b = np.ones((n, 64, 64, 4), dtype=np.float32)
        single.fill_batch_host(pb, rb.FillRule.NonZero, black, 64, 64, a)
        multi.fill_batch_host(pb, rb.FillRule.NonZero, black, 64, 64, b)
        assert np.array_equal(a, b)
        p = assets.load_path("tv_stroked")
        c = assets.expected()["paths"]["tv_stroked"]["c5"]
        w = h = 8192
        tr = np.array(c["tr"]) * (w / c["size"][0])
        whole = np.zeros((h, w), dtype=np.float32)
        single.mask(p, tr, whole, rb.FillRule.NonZero)
        img = np.full((h, w), -1.0, dtype=np.float32)
        multi.mask_banded(p, tr, img, rb.FillRule.NonZero)
        assert np.array_equal(img, whole)
    finally:
        multi.close()
        single.close()


def test_fill_batch_host_split_download_is_bit_identical(rast):
    """LinColor output of a plain solid paint: a share of every chunk crosses PCIe as coverage and is expanded to colour * alpha
    by host threads (rgpu_fill_batch_host, split download).  The share adapts from call to call, so repeated calls route
    different glyphs through the host expansion: every call must give the same bytes, and they must be the bytes of the
    device-side RENDER path (a prepared batch rendering into a device slab, downloaded whole)."""
    n = 6000  # three chunks of 2048 glyphs
    pb = synth.glyph_batch(77, n)
    colour = rb.LinColor(0.25, 0.5, 0.125, 0.75)
    outs = []
    for _ in range(4):
        out = np.full((n, 64, 64, 4), 9.0, dtype=np.float32)
        rast.fill_batch_host(pb, rb.FillRule.NonZero, colour, 64, 64, out)
        outs.append(out)
    for o in outs[1:]:
        assert np.array_equal(outs[0].view(np.uint32), o.view(np.uint32))
    dpb = rast.upload_batch(pb)
    slab = rast.device_alloc(n * 4096 * 16)
    keep = []
    paint_ptr = colour._c(keep)
    import ctypes as C
    t = job_table(dpb.handles(), slab, ffi.JOB_RENDER, rb.FillRule.NonZero, n, 64, 64, C.addressof(paint_ptr))
    prepared = rast.prepare_job_table(t, keep=[keep, paint_ptr])
    prepared.render()
    rast.batch_status()
    dev = rast.to_host(slab, (n, 64, 64, 4), np.float32)
    assert np.array_equal(outs[0].view(np.uint32), dev.view(np.uint32))
    prepared.free()
    rast.device_free(slab)
    dpb.free()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_run_coded_download_matches_dense_copies(rast, dtype):
    """Masks of 4 Mpixel and more come down run-coded (compact.cu: class byte per 64-pixel segment, literals for the edge
    segments, rows rebuilt by host threads); smaller calls copy the dense rows.  Sixteen small band calls (dense) and one call
    for the whole canvas (run-coded) must give the same bytes — for rows that are 64-byte aligned (4096), rows with a ragged
    last segment that are only 16-byte aligned (4100) and rows with no alignment to speak of (5001, 5003 odd) — and the
    whole-canvas call must move a fraction of the dense bytes."""
    ex = assets.expected()["paths"]
    p = assets.load_path("tv_stroked")
    c = ex["tv_stroked"]["c5"]
    for (w, h) in ((4096, 2048), (4100, 1500), (5001, 1000), (5003, 999)):
        tr = np.array(c["tr"]) * np.array([w / c["size"][0]] * 3 + [h / c["size"][1]] * 3)
        dense = np.full((h, w), -3.0, dtype=dtype)
        for b in range(16):
            rast.mask_banded(p, tr, dense, rb.FillRule.NonZero, n_bands=16, band_first=b, band_count=1)
        whole = np.full((h, w), -5.0, dtype=dtype)
        rast.mask_banded(p, tr, whole, rb.FillRule.NonZero, n_bands=1)
        _, d2h = rast.last_transfer_bytes()
        assert np.array_equal(whole.view(np.uint32 if dtype == np.float32 else np.uint64), dense.view(np.uint32 if dtype == np.float32 else np.uint64)), (w, h)
        assert 0 < whole.max() <= 1.0 and whole.min() == 0.0
        assert d2h < 0.5 * w * h * 4, (w, h, d2h)
    # rgpu_mask_f32 takes the same way
    w, h = 4096, 1024
    tr = np.array(c["tr"]) * np.array([w / c["size"][0]] * 3 + [h / c["size"][1]] * 3)
    a = np.zeros((h, w), dtype=np.float32)
    rast.mask(p, tr, a, rb.FillRule.EvenOdd)
    b = np.zeros((h, w), dtype=np.float32)
    for k in range(8):
        rast.mask_banded(p, tr, b, rb.FillRule.EvenOdd, n_bands=8, band_first=k, band_count=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_config4_full_size_properties(rast):
    """BASELINE config 4 at its full size (100 000 glyphs at 64 x 64, what bench.py times), through size-independent properties:
    * LinColor output of solid black = (0, 0, 0, coverage): the alpha plane of the whole batch equals the COVERAGE output
      of the same call bit for bit (two kernel modes, and the LinColor call mixes DMA'd and host-expanded glyphs);
    * a glyph does not depend on its batch: 61 glyphs picked across the batch and rendered as their own small batch give the
      bits they have inside the 100 000;
    * the same 61 glyphs against the oracle's Path::fill (<= 1e-4), and the oracle's coverage sums (a checksum of checksums);
    * nothing of the output buffer is left unwritten."""
    n = 100000
    pb = synth.glyph_batch(1, n)
    black = rb.LinColor(0.0, 0.0, 0.0, 1.0)
    lin = np.empty((n, 64, 64, 4), dtype=np.float32)
    lin[:] = 7.0
    rast.fill_batch_host(pb, rb.FillRule.NonZero, black, 64, 64, lin)
    cov = np.empty((n, 64, 64), dtype=np.float32)
    cov[:] = 7.0
    rast.fill_batch_host(pb, rb.FillRule.NonZero, None, 64, 64, cov)
    assert cov.max() <= 1.0 and cov.min() >= 0.0
    assert np.array_equal(lin[..., 3].view(np.uint32), cov.view(np.uint32))
    assert not lin[..., :3].any()
    picks = sorted(set(int(i) for i in np.linspace(0, n - 1, 61)))
    small = rb.PathBatch.from_paths([pb.path(i) for i in picks])
    alone = np.empty((len(picks), 64, 64), dtype=np.float32)
    rast.fill_batch_host(small, rb.FillRule.NonZero, None, 64, 64, alone)
    assert np.array_equal(alone.view(np.uint32), cov[picks].view(np.uint32))
    paint = O.OraclePaint.solid([0, 0, 0, 1])
    sums_gpu, sums_ref = [], []
    for k, i in enumerate(picks):
        ref = np.zeros((64, 64, 4), dtype=np.float32)
        opath(pb.path(i)).fill(O.IDENTITY, O.NONZERO, paint, ref)
        assert np.abs(lin[i] - ref).max() <= COV_TOL, i
        sums_gpu.append(float(cov[i].astype(np.float64).sum()))
        sums_ref.append(float(ref[..., 3].astype(np.float64).sum()))
    assert abs(sum(sums_gpu) - sum(sums_ref)) <= 1e-2 and max(abs(a - b) for a, b in zip(sums_gpu, sums_ref)) <= 5e-3
