"""Parity of paint/composite (K4) through the C ABI against the CPU oracle: `Rasterizer::fill` with solid,
linear- and radial-gradient paints, scene rendering (Fill nodes composited in order on a device-resident
layer) and the LinColor -> RGBA8 conversion.  Bars: LinColor within 2e-4 (coverage 1e-4 x colour), 8-bit RGBA
within 1 LSB of the oracle's x86-simd colour variant (north_star).  Run on a B200: pytest -m gpu."""
import math

import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb
import assets
from rasterize_b200 import ffi

pytestmark = pytest.mark.gpu

LIN_TOL = 2e-4


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def opath(p):
    return O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)


from helpers import oracle_paint, render_scene_gpu, render_scene_oracle  # noqa: E402


def rgba_close(a, b):
    return np.abs(a.astype(np.int16) - b.astype(np.int16)).max()


def test_fill_solid_host_image(rast):
    """trait-level rgpu_fill on a strided host LinColor image over a non-trivial background"""
    p = assets.load_path("squirrel")
    e = assets.expected()["paths"]["squirrel"]
    w, h = e["size"]
    tr = np.array(e["size_tr"])
    rng = np.random.default_rng(7)
    bg = rng.random((h + 3, w + 5, 4), dtype=np.float32) * 0.5
    bg[..., 3] = 0.5 + bg[..., 3]
    bg[..., :3] *= bg[..., 3:4]
    got = bg.copy()
    ref = bg.copy()
    color = np.float32([0.2 * 0.7, 0.5 * 0.7, 0.1 * 0.7, 0.7])
    view = got[1:1 + h, 2:2 + w]
    rast.fill(p, tr, rb.FillRule.NonZero, rb.LinColor(*color), view)
    opath(p).fill(tr, O.NONZERO, O.OraclePaint.solid(color), ref, shape=O.Shape((w + 5) + 2, w, h, w + 5, 1))
    assert np.abs(got - ref).max() <= LIN_TOL
    untouched = np.ones(bg.shape[:2], dtype=bool)
    untouched[1:1 + h, 2:2 + w] = False
    assert np.array_equal(got[untouched], bg[untouched])


def test_fill_column_strided_view(rast):
    """rgpu_fill on a view with a column stride (every second pixel of every second row of a larger image): only the bounding
    rectangle of the path is gathered / scattered; the result matches the oracle's fill through the same Shape and the pixels
    between the view's pixels stay untouched."""
    p = assets.load_path("squirrel")
    e = assets.expected()["paths"]["squirrel"]
    w, h = e["size"]
    # the path shrunk into the middle of the view: the moved rectangle is a proper part of it
    tr = np.array(e["size_tr"]) * 0.5
    tr[2] += w * 0.25
    tr[5] += h * 0.25
    rng = np.random.default_rng(11)
    bg = rng.random((2 * h + 3, 2 * w + 4, 4), dtype=np.float32) * 0.5
    got = bg.copy()
    ref = bg.copy()
    color = np.float32([0.1, 0.3, 0.2, 0.6])
    view = got[1:1 + 2 * h:2, 3:3 + 2 * w:2]
    assert view.shape == (h, w, 4)
    rast.fill(p, tr, rb.FillRule.EvenOdd, rb.LinColor(*color), view)
    W = 2 * w + 4
    opath(p).fill(tr, O.EVENODD, O.OraclePaint.solid(color), ref, shape=O.Shape(W + 3, w, h, 2 * W, 2))
    assert np.abs(got - ref).max() <= LIN_TOL
    assert not np.array_equal(got, bg)
    untouched = np.ones(bg.shape[:2], dtype=bool)
    untouched[1:1 + 2 * h:2, 3:3 + 2 * w:2] = False
    assert np.array_equal(got[untouched], bg[untouched])


@pytest.mark.parametrize("kind", ["linear", "radial"])
@pytest.mark.parametrize("spread", [0, 1, 2])
@pytest.mark.parametrize("linear_colors", [True, False])
def test_fill_gradients(rast, kind, spread, linear_colors):
    p = assets.load_path("rust")
    e = assets.expected()["paths"]["rust"]
    w, h = e["size"]
    tr = np.array(e["size_tr"])
    stops_lin = [(0.0, O.parse_color("#ffbd4f")), (0.4, O.parse_color("#ff0000c0")), (0.7, O.parse_color("#ff00ff40")), (1.0, O.parse_color("#000000"))]
    ptr = O.transform_mul(O.rotate(0.3), O.scale(1.2, 0.8))
    if kind == "linear":
        op = O.OraclePaint.linear(stops_lin, (10, 20), (70, 60), units=0, linear_colors=linear_colors, spread=spread, tr=ptr)
    else:
        op = O.OraclePaint.radial(stops_lin, (50, 50), 35.0, fcenter=(40, 45), fradius=3.0, units=0, linear_colors=linear_colors,
                                  spread=spread, tr=ptr)
    gp = rb.paint_from_desc(op.describe())
    ref = np.zeros((h, w, 4), dtype=np.float32)
    ref[..., :] = np.float32([0.1, 0.1, 0.1, 1.0])
    got = ref.copy()
    opath(p).fill(tr, O.EVENODD, op, ref)
    rast.fill(p, tr, rb.FillRule.EvenOdd, gp, got)
    assert np.abs(got - ref).max() <= LIN_TOL
    assert rgba_close(O.lin_to_rgba(got), O.lin_to_rgba(ref)) <= 1


def test_fill_bounding_box_units(rast):
    p = assets.load_path("tv")
    op_ = opath(p)
    e = assets.expected()["paths"]["tv"]
    w, h = e["size"]
    tr = O.transform_mul(O.scale(4.0, 4.0), np.array(e["size_tr"]))
    w, h = w * 4, h * 4
    stops = [(0.0, O.parse_color("#ff0000")), (1.0, O.parse_color("#00ff00"))]
    op = O.OraclePaint.linear(stops, (0, 0), (1, 0), units=1)
    gp = rb.paint_from_desc(op.describe())
    ref = np.zeros((h, w, 4), dtype=np.float32)
    got = ref.copy()
    op_.fill(tr, O.NONZERO, op, ref)
    rast.fill(p, tr, rb.FillRule.NonZero, gp, got, bbox=op_.bbox())
    assert ref[..., 3].max() > 0.9
    assert np.abs(got - ref).max() <= LIN_TOL
    # bounding-box units without a bbox, and a singular paint transform, are silent no-ops (src/rasterize.rs:87-92)
    z = np.zeros((h, w, 4), dtype=np.float32)
    rast.fill(p, tr, rb.FillRule.NonZero, gp, z, bbox=None)
    assert (z == 0).all()
    sing = rb.GradLinear([(0.0, (1, 0, 0, 1)), (1.0, (0, 1, 0, 1))], 0, True, 0, (0, 0, 0, 0, 0, 0), (0, 0), (1, 0))
    rast.fill(p, tr, rb.FillRule.NonZero, sing, z)
    assert (z == 0).all()


@pytest.mark.parametrize("name", ["squirrel_cli_512", "linear_colors", "firefox_256", "many_circles_64", "firefox_2048"])
def test_scene_render(rast, name):
    """Scene::render (Fill nodes, device-resident layer): C1 squirrel CLI scene, linear-colors.scene,
    C3 firefox.scene at 256 and 2048, the many-circles bench scene"""
    sc = assets.load_scene(name)
    lin, rgba = render_scene_gpu(rast, sc)
    ref = render_scene_oracle(sc)
    assert lin.shape == ref.shape
    assert np.abs(lin - ref).max() <= LIN_TOL, np.abs(lin - ref).max()
    ref_rgba = O.lin_to_rgba(ref)
    assert rgba_close(rgba, ref_rgba) <= 1
    # device RGBA8 conversion of identical input is bit-exact
    assert np.array_equal(O.lin_to_rgba(lin), rgba)
    ex = assets.expected()["scenes"][name]
    assert [0, 0, lin.shape[1], lin.shape[0]][2:] == ex["layer"][2:]
    assert np.allclose(lin.reshape(-1, 4).sum(0, dtype=np.float64), ex["lin_sum"], rtol=2e-4, atol=1.0)


def test_to_rgba8_edge_values(rast):
    vals = np.float32([[0, 0, 0, 0], [1, 1, 1, 1], [0.5, 0.25, 0.125, 0.5], [1e-7, 1e-7, 1e-7, 1e-7], [2, 2, 2, 1], [-0.1, 0.3, 0.2, 1.0],
                       [0.003, 0.0031308, 0.0031309, 1.0], [0.2, 0.2, 0.2, 1.0008736]])
    n = len(vals)
    d_in = rast.device_alloc(n * 16)
    d_out = rast.device_alloc(n * 4)
    rast.to_device(d_in, vals)
    rast.to_rgba8(d_in, d_out, n)
    got = rast.to_host(d_out, (n, 4), np.uint8)
    assert np.array_equal(got, O.lin_to_rgba(vals))
    rast.device_free(d_in)
    rast.device_free(d_out)


@pytest.mark.parametrize("offset", [(300.0, 180.0), (-40.0, -25.0), (830.0, 520.0), (2000.0, 2000.0)])
def test_fill_small_shape_on_big_image(rast, offset):
    """rgpu_fill moves only the bounding rectangle of the transformed path over PCIe: a small shape placed inside a large
    image, hanging over its left/top edge, over its right/bottom edge, and entirely outside must give the oracle's
    pixels everywhere (the rest of the image is left untouched)."""
    p = assets.load_path("squirrel")
    e = assets.expected()["paths"]["squirrel"]
    tr = O.transform_mul(O.translate(*offset), np.array(e["size_tr"]) * np.array([0.25, 0.25, 0.25, 0.25, 0.25, 0.25]))
    W, H = 900, 600
    rng = np.random.default_rng(3)
    bg = rng.random((H, W, 4), dtype=np.float32) * 0.4
    bg[..., 3] += 0.5
    got, ref = bg.copy(), bg.copy()
    color = np.float32([0.1, 0.3, 0.6, 0.8])
    rast.fill(p, tr, rb.FillRule.EvenOdd, rb.LinColor(*color), got)
    opath(p).fill(tr, O.EVENODD, O.OraclePaint.solid(color), ref, shape=O.Shape(0, W, H, W, 1))
    assert np.abs(got - ref).max() <= LIN_TOL
    changed = np.abs(ref - bg).max(axis=2) > 0
    assert np.array_equal(got[~changed], bg[~changed])


def test_ordered_fills_are_reproducible(rast):
    """The fills of a scene are chained with dependent launches (a fill accumulates its tiles while the previous one is
    still compositing): twenty renders of firefox.scene at 2048 x 2048 must be bit-identical."""
    sc = assets.load_scene("firefox_2048")
    first, _ = render_scene_gpu(rast, sc, to_rgba=False)
    for _ in range(19):
        again, _ = render_scene_gpu(rast, sc, to_rgba=False)
        assert np.array_equal(first, again)
