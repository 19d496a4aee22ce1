"""GPU: the scene compositor (`rgpu_render_scene`, csrc/scene.cu — all fills of a layer in ONE raster launch) against

* the ordered batch (`RGPU_BATCH_ORDERED`, one launch per fill): **bit-identical** LinColor and RGBA8 — both paths round
  the same per-pixel coverages to the same fixed-point cells, so tiling cannot change a bit;
* the CPU oracle's `Scene::render` (LinColor within 2e-4, RGBA8 within 1 LSB).

Covers fresh layers with and without a background, fills over an existing layer, windows that are not aligned with the
256 x 8 layer tiles (carry chain across chunks, rows above / columns left of a window), empty batches, and the
two-pass fallback.  Run on a B200: pytest -m gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest

import bench
import oracle as O
import rasterize_b200 as rb
from helpers import render_scene_oracle
import assets
from rasterize_b200 import ffi, scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def both_ways(rast, make_jobs, W, H, bg, initial=None):
    """Render the same job list with the ordered batch and with the scene kernel; returns (lin_a, rgba_a, lin_b, rgba_b)."""
    n = W * H
    out = []
    for use_scene in (False, True):
        layer = rast.device_alloc(n * 16)
        rgba = rast.device_alloc(n * 4)
        jobs, keep = make_jobs(layer)
        if initial is not None:
            rast.to_device(layer, initial)
        if use_scene:
            rast.render_scene(jobs, layer, W, H, fresh=initial is None, bg=bg, rgba_ptr=rgba)
        else:
            if initial is None:
                if bg is not None:
                    rast.fill_color(layer, n, bg)
                else:
                    rast.device_zero(layer, n * 16)
            rast.render_batch(jobs, independent=False)
            rast.to_rgba8(layer, rgba, n)
        out.append(rast.to_host(layer, (H, W, 4), np.float32))
        out.append(rast.to_host(rgba, (H, W, 4), np.uint8))
        rast.device_free(layer)
        rast.device_free(rgba)
        del keep
    return out


@pytest.mark.parametrize("name", ["squirrel_cli_512", "firefox_256", "firefox_2048", "linear_colors", "many_circles_64"])
def test_scene_kernel_matches_ordered_batch_and_oracle(rast, name):
    sc = assets.load_scene(name)

    def make(layer):
        jobs, keep, _, _, _ = scene.fixture_jobs(rast, sc, layer)
        return jobs, keep

    _, _, W, H, _ = scene.fixture_jobs(rast, sc, 1)
    lin_a, rgba_a, lin_b, rgba_b = both_ways(rast, make, W, H, sc.bg)
    assert np.array_equal(lin_a, lin_b)
    assert np.array_equal(rgba_a, rgba_b)
    if name != "firefox_2048":  # the oracle takes 0.4 s per 2048^2 render; the smaller scenes pin the values
        ref = render_scene_oracle(sc)
        assert ref.shape == lin_b.shape
        assert np.abs(lin_b - ref).max() <= 2e-4
        assert np.abs(rgba_b.astype(np.int32) - O.lin_to_rgba(ref).astype(np.int32)).max() <= 1


def synthetic_jobs(rast, W, H, n_jobs, seed, max_w, max_h):
    """Random-cubic glyphs (bench.glyph_path, 64 x 64 units) scaled into windows at arbitrary layer offsets, with solid,
    linear- and radial-gradient paints and both fill rules."""
    rng = np.random.default_rng(seed)
    specs = []
    for k in range(n_jobs):
        w = int(rng.integers(1, max_w + 1))
        h = int(rng.integers(1, max_h + 1))
        x = int(rng.integers(0, W - w + 1))
        y = int(rng.integers(0, H - h + 1))
        a = float(rng.uniform(0.3, 1.0))
        rgb = rng.uniform(0.0, 1.0, 3) * a
        kind = k % 3
        bbox = None
        stops = [(0.0, [*(rng.uniform(0, 1, 3) * 0.8), 0.8]), (0.6, [0.1, 0.9, 0.3, 1.0]), (1.0, [*rng.uniform(0, 1, 3), 1.0])]
        if kind == 0:
            paint = rb.LinColor(float(rgb[0]), float(rgb[1]), float(rgb[2]), a)
        elif kind == 1 and k % 2:
            # objectBoundingBox units: resolved through the path's bbox (src/rasterize.rs:85-91)
            paint = rb.GradLinear(stops, rb.Units.BoundingBox, bool(k & 2), rb.GradSpread(k % 3), rb.Transform.identity(), (0.1, 0.2), (0.9, 0.7))
            bbox = np.array([4.0, 4.0, 60.0, 60.0])
        elif kind == 1:
            paint = rb.GradLinear(stops, rb.Units.UserSpaceOnUse, bool(k & 1), rb.GradSpread(k % 3), rb.Transform.identity(), (4.0, 8.0), (60.0, 50.0))
        else:
            paint = rb.GradRadial(stops, rb.Units.UserSpaceOnUse, bool(k & 1), rb.GradSpread((k + 1) % 3), rb.Transform.identity(), (32.0, 32.0), 30.0,
                                  (28.0, 30.0), 2.0)
        # the glyph overhangs the window on every side for some jobs (exercises the x < 0 / x > width / y clipping)
        s = float(rng.uniform(0.8, 1.4))
        tr = rb.Transform.new_translate(float(rng.uniform(-0.2, 0.1)) * w, float(rng.uniform(-0.2, 0.1)) * h) * rb.Transform.new_scale(s * w / 64.0, s * h / 64.0)
        specs.append((bench.glyph_path(rb, 1000 * seed + k + 1), tr, rb.FillRule(k % 2), paint, x, y, w, h, bbox))

    def make(layer):
        jobs, keep = [], []
        for path, tr, rule, paint, x, y, w, h, bbox in specs:
            dp = rast.upload(path)
            keep.append(dp)
            jobs.append(rb.Job(dp, tr, rule, ffi.JOB_FILL, layer, w, h, W, origin=y * W + x, paint=paint, path_bbox=bbox))
        return jobs, keep

    return make


@pytest.mark.parametrize("W,H,n_jobs,max_w,max_h", [(1300, 70, 24, 1300, 70), (517, 33, 40, 200, 33), (2049, 19, 12, 2049, 19), (64, 64, 6, 64, 64),
                                                   (1536, 24, 30, 1100, 24), (900, 120, 600, 80, 60)])
def test_unaligned_windows_bit_identical(rast, W, H, n_jobs, max_w, max_h):
    make = synthetic_jobs(rast, W, H, n_jobs, seed=W + H, max_w=max_w, max_h=max_h)
    lin_a, rgba_a, lin_b, rgba_b = both_ways(rast, make, W, H, bg=[0.1, 0.2, 0.3, 0.5])
    if max_w <= 64 and max_h <= 64:
        # every window fits the fused small-canvas kernel, which the ordered batch then uses: it evaluates a line's rows in
        # f32 (end points below 128), the scene compositor in f64 — same coverage within the 1e-4 budget, not the same bits
        assert np.abs(lin_a - lin_b).max() <= 4e-4
        assert np.abs(rgba_a.astype(int) - rgba_b.astype(int)).max() <= 1
    else:
        assert np.array_equal(lin_a, lin_b)
        assert np.array_equal(rgba_a, rgba_b)
    assert np.abs(lin_b - np.float32([0.1, 0.2, 0.3, 0.5])).max() > 0.1  # something was drawn


def test_fills_over_existing_layer(rast):
    W, H = 1100, 41
    make = synthetic_jobs(rast, W, H, 16, seed=5, max_w=900, max_h=41)
    rng = np.random.default_rng(11)
    init = rng.random((H, W, 4), dtype=np.float32)
    init[..., :3] *= init[..., 3:4]
    lin_a, rgba_a, lin_b, rgba_b = both_ways(rast, make, W, H, bg=None, initial=init)
    assert np.array_equal(lin_a, lin_b)
    assert np.array_equal(rgba_a, rgba_b)


def test_empty_scene_and_transparent_background(rast):
    W, H = 700, 13
    n = W * H
    layer = rast.device_alloc(n * 16)
    rgba = rast.device_alloc(n * 4)
    rast.to_device(layer, np.full((H, W, 4), 7.0, dtype=np.float32))
    rast.render_scene([], layer, W, H, fresh=True, bg=None, rgba_ptr=rgba)
    assert not rast.to_host(layer, (H, W, 4), np.float32).any()
    assert not rast.to_host(rgba, (H, W, 4), np.uint8).any()
    rast.render_scene([], layer, W, H, fresh=True, bg=[0.25, 0.25, 0.25, 1.0], rgba_ptr=0)
    assert np.array_equal(rast.to_host(layer, (H, W, 4), np.float32), np.broadcast_to(np.float32([0.25, 0.25, 0.25, 1.0]), (H, W, 4)))
    # a job whose window lies outside every line of its path still materialises the background
    make = synthetic_jobs(rast, W, H, 1, seed=3, max_w=5, max_h=5)
    jobs, keep = make(layer)
    jobs[0].tr = rb.Transform.new_translate(1e6, 1e6)
    rast.render_scene(jobs, layer, W, H, fresh=True, bg=[0.5, 0.0, 0.0, 0.5])
    assert np.array_equal(rast.to_host(layer, (H, W, 4), np.float32), np.broadcast_to(np.float32([0.5, 0.0, 0.0, 0.5]), (H, W, 4)))
    rast.device_free(layer)
    rast.device_free(rgba)


def test_scene_rejects_foreign_jobs(rast):
    W, H = 100, 10
    layer = rast.device_alloc(W * H * 16)
    other = rast.device_alloc(W * H * 16)
    dp = rast.upload(bench.glyph_path(rb, 1))
    black = rb.LinColor(0, 0, 0, 1)
    with pytest.raises(rb.RgpuError):  # not a window of the layer
        rast.render_scene([rb.Job(dp, rb.Transform.identity(), rb.FillRule.NonZero, ffi.JOB_FILL, other, 50, 10, W, paint=black)], layer, W, H)
    with pytest.raises(rb.RgpuError):  # window leaves the layer
        rast.render_scene([rb.Job(dp, rb.Transform.identity(), rb.FillRule.NonZero, ffi.JOB_FILL, layer, 60, 10, W, origin=50, paint=black)], layer, W, H)
    with pytest.raises(rb.RgpuError):  # masks do not belong in a scene batch
        rast.render_scene([rb.Job(dp, rb.Transform.identity(), rb.FillRule.NonZero, ffi.JOB_MASK, layer, 50, 10, W)], layer, W, H)
    rast.device_free(layer)
    rast.device_free(other)


def test_scene_two_pass_fallback_bit_identical():
    """RGPU_TWO_PASS=1 (exact count -> scan -> emit bins) routes scene batches through the ordered launches: same bits."""
    code = """
import numpy as np, rasterize_b200 as rb
import assets
from rasterize_b200 import scene
r = rb.GpuRasterizer()
sc = assets.load_scene('firefox_256')
_, _, W, H, _ = scene.fixture_jobs(r, sc, 1)
layer = r.device_alloc(W * H * 16)
jobs, keep, _, _, _ = scene.fixture_jobs(r, sc, layer)
r.render_scene(jobs, layer, W, H, fresh=True, bg=sc.bg)
np.save('%s', r.to_host(layer, (H, W, 4), np.float32))
"""
    import tempfile
    outs = []
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env_extra in ({}, {"RGPU_TWO_PASS": "1"}):
        with tempfile.NamedTemporaryFile(suffix=".npy", delete=False) as f:
            name = f.name
        env = dict(os.environ, **env_extra)
        env["PYTHONPATH"] = os.path.join(root, "tests") + os.pathsep + root + os.pathsep + env.get("PYTHONPATH", "")
        subprocess.run([sys.executable, "-c", code % name], check=True, env=env, cwd=root, timeout=300)
        outs.append(np.load(name))
        os.unlink(name)
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("name", ["squirrel_cli_512", "firefox_256", "linear_colors", "many_circles_64"])
def test_render_scene_host_matches_device_path(rast, name):
    """`rgpu_render_scene_host` (host paths in, host images out) is the device-resident scene render, bit for bit."""
    sc = assets.load_scene(name)
    fills, W, H = scene.fixture_fills_host(sc)
    lin = np.empty((H, W, 4), dtype=np.float32)
    rgba = np.empty((H, W, 4), dtype=np.uint8)
    rast.render_scene_host(fills, W, H, bg=sc.bg, lin_out=lin, rgba_out=rgba)
    h2d, d2h = rast.last_transfer_bytes()
    assert d2h == W * H * 20 and h2d >= sum(f[0].input_bytes() for f in fills)

    def make(layer):
        jobs, keep, _, _, _ = scene.fixture_jobs(rast, sc, layer)
        return jobs, keep

    _, _, lin_b, rgba_b = both_ways(rast, make, W, H, sc.bg)
    assert np.array_equal(lin, lin_b)
    assert np.array_equal(rgba, rgba_b)
    # RGBA8 only (the CLI's output): same bytes, a quarter of the download
    _, rgba2 = rast.render_scene_host(fills, W, H, bg=sc.bg)
    assert np.array_equal(rgba2, rgba)
    assert rast.last_transfer_bytes()[1] == W * H * 4


def test_render_scene_host_edge_cases(rast):
    # no fills: the background alone
    _, rgba = rast.render_scene_host([], 33, 9, bg=[1.0, 1.0, 1.0, 1.0])
    assert (rgba == 255).all()
    # empty path and zero-sized window are no-ops; a NaN control point is reported like the reference's panic
    black = rb.LinColor(0, 0, 0, 1)
    ident = rb.Transform.identity()
    g = bench.glyph_path(rb, 3)
    fills = [(rb.Path.empty(), ident, rb.FillRule.NonZero, black, None, 0, 0, 20, 9), (g, ident, rb.FillRule.NonZero, black, None, 5, 2, 0, 0)]
    _, rgba = rast.render_scene_host(fills, 33, 9, bg=None)
    assert not rgba.any()
    bad = rb.Path(np.array([[0.0, 0.0], [np.nan, 1.0]]), np.array([2], dtype=np.uint8), np.array([0, 1], dtype=np.uint32), np.array([1], dtype=np.uint8))
    with pytest.raises(rb.RgpuError) as e:
        rast.render_scene_host([(bad, ident, rb.FillRule.NonZero, black, None, 0, 0, 33, 9)], 33, 9)
    assert e.value.code == ffi.ERR_NAN
    with pytest.raises(rb.RgpuError):  # window outside the layer
        rast.render_scene_host([(g, ident, rb.FillRule.NonZero, black, None, 30, 0, 10, 9)], 33, 9)
