"""GPU: winding numbers beyond the integer range of the Q7.24 cells (VERDICT r1 missing #6).  The 32-bit cells
wrap modulo 256 windings: EvenOdd cannot see that, NonZero goes wrong near a non-zero multiple of 256.  The row scans flag a
NonZero winding >= 120 and the *_sync / host-buffer entry points repeat the batch in Q13.18.  Checked against the f64 oracle
(reference src/rasterize.rs:478-506) on stacks of same-direction rectangles whose common core has winding 130 / 256 / 300."""
import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb
from helpers import opath
from rasterize_b200 import ffi

pytestmark = pytest.mark.gpu
COV_TOL = 1e-4


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def stacked_rects(n, w, h):
    """n same-direction rectangles, every edge at its own position, all containing the core of the canvas"""
    b = rb.Path.builder()
    for i in range(n):
        x0, y0 = 2.0 + 0.031 * i, 1.5 + 0.023 * i
        x1, y1 = w - 2.5 - 0.027 * i, h - 2.0 - 0.019 * i
        b.move_to((x0, y0)).line_to((x1, y0)).line_to((x1, y1)).line_to((x0, y1)).close()
    return b.build()


@pytest.mark.parametrize("size", [(64, 64), (200, 90), (1500, 40)])  # fused small-canvas kernel, 128-wide tiles, 1024 x 8 tiles with carries
@pytest.mark.parametrize("n", [130, 256, 300])
def test_mask_with_high_windings_matches_oracle(rast, size, n):
    w, h = size
    p = stacked_rects(n, w, h)
    for rule, orule in ((rb.FillRule.NonZero, O.NONZERO), (rb.FillRule.EvenOdd, O.EVENODD)):
        img = np.zeros((h, w))
        rast.mask(p, rb.Transform.identity(), img, rule)
        ref = np.zeros((h, w))
        opath(p).mask(O.IDENTITY, orule, ref)
        assert np.abs(img - ref).max() <= COV_TOL, (size, n, rule, float(np.abs(img - ref).max()))
        if rule == rb.FillRule.NonZero:
            assert img[h // 2, w // 2] == 1.0  # the core: winding n, which Q7.24 alone would wrap to n - 256


def test_fill_and_scene_with_high_windings(rast):
    """`Rasterizer::fill` over an existing host image is repeated from the caller's copy, never blended twice."""
    w, h = 180, 70
    p = stacked_rects(256, w, h)
    paint, opaint = rb.LinColor(0.2, 0.1, 0.05, 0.5), O.OraclePaint.solid([0.2, 0.1, 0.05, 0.5])
    rng = np.random.default_rng(3)
    base = rng.uniform(0.0, 0.4, size=(h, w, 4)).astype(np.float32)
    img = base.copy()
    rast.fill(p, rb.Transform.identity(), rb.FillRule.NonZero, paint, img)
    ref = base.copy()
    opath(p).fill(O.IDENTITY, O.NONZERO, opaint, ref)
    assert np.abs(img - ref).max() <= 2e-4
    # a glyph-style batch (RENDER jobs overwrite their canvases: repeated transparently)
    pb = rb.PathBatch.from_paths([stacked_rects(256, 64, 64), stacked_rects(10, 64, 64)])
    out = np.zeros((2, 64, 64, 4), dtype=np.float32)
    rast.fill_batch_host(pb, rb.FillRule.NonZero, paint, 64, 64, out)
    for i in range(2):
        ref = np.zeros((64, 64, 4), dtype=np.float32)
        opath(pb.path(i)).fill(O.IDENTITY, O.NONZERO, opaint, ref)
        assert np.abs(out[i] - ref).max() <= 2e-4, i


def test_asynchronous_submission_reports_the_guard(rast):
    """rgpu_render_batch + rgpu_batch_status cannot repeat a batch: they report RGPU_ERR_WINDING; Q13.18 selected up front works."""
    w, h = 200, 90
    p = stacked_rects(256, w, h)
    dp = rast.upload(p)
    canvas = rast.device_alloc(w * h * 4)
    jobs = [rb.Job(dp, rb.Transform.identity(), rb.FillRule.NonZero, ffi.JOB_MASK, canvas, w, h, w)]
    rast.render_batch(jobs, independent=True, sync=False)
    with pytest.raises(rb.RgpuError) as e:
        rast.batch_status()
    assert e.value.code == ffi.ERR_WINDING
    rast.set_winding_bits(14)
    try:
        rast.render_batch(jobs, independent=True, sync=False)
        rast.batch_status()
        got = rast.to_host(canvas, (h, w), np.float32)
    finally:
        rast.set_winding_bits(8)
    ref = np.zeros((h, w))
    opath(p).mask(O.IDENTITY, O.NONZERO, ref)
    assert np.abs(got - ref).max() <= COV_TOL
    # even-odd never trips the guard (a wrap by 256 windings keeps the parity)
    jobs = [rb.Job(dp, rb.Transform.identity(), rb.FillRule.EvenOdd, ffi.JOB_MASK, canvas, w, h, w)]
    rast.render_batch(jobs, independent=True, sync=False)
    rast.batch_status()
    ref = np.zeros((h, w))
    opath(p).mask(O.IDENTITY, O.EVENODD, ref)
    assert np.abs(rast.to_host(canvas, (h, w), np.float32) - ref).max() <= COV_TOL
    rast.device_free(canvas)
