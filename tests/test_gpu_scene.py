"""GPU: device-resident `Scene::render` with clip and opacity nodes against the oracle.

Tolerances are the north star's: LinColor within 2e-4 per channel (coverage 1e-4 through the paint), RGBA8 within 1 LSB."""
import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb
from helpers import render_pipeline_oracle
from rasterize_b200 import assets, scene

pytestmark = pytest.mark.gpu

LIN_TOL = 2e-4


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


@pytest.mark.parametrize("name", ["grad", "grad_1024", "nested", "nested_900", "firefox_512"])
def test_pipeline_render(rast, name):
    pl = assets.load_pipeline(name)
    x, y, lin = scene.render(rast, pl)
    ox, oy, ref = render_pipeline_oracle(pl)
    assert (x, y) == (ox, oy) and lin.shape == ref.shape
    assert np.abs(lin - ref).max() <= LIN_TOL, np.abs(lin - ref).max()
    _, _, rgba = scene.render(rast, pl, rgba=True)
    ref8 = O.lin_to_rgba(ref)
    assert np.abs(rgba.astype(np.int32) - ref8.astype(np.int32)).max() <= 1


def test_pipeline_render_is_deterministic(rast):
    pl = assets.load_pipeline("nested_900")
    a = scene.render(rast, pl)[2]
    b = scene.render(rast, pl)[2]
    assert np.array_equal(a, b)


def test_compose_primitives(rast):
    """rgpu_layer_scale_by_mask_dev / rgpu_layer_blend_over_dev on offset rectangles against numpy f32."""
    rng = np.random.default_rng(7)
    H, W, h, w = 37, 53, 20, 31
    dst = rng.random((H, W, 4), dtype=np.float32)
    src = rng.random((h + 3, w + 5, 4), dtype=np.float32)
    mask = rng.random((h + 2, w + 1), dtype=np.float32)
    d_dst, d_src, d_mask = (rast.device_alloc(a.nbytes) for a in (dst, src, mask))
    rast.to_device(d_dst, dst); rast.to_device(d_src, src); rast.to_device(d_mask, mask)
    # src[1:1+h, 2:2+w] *= mask[2:2+h, 1:1+w]
    rast.layer_scale_by_mask(d_src, 1 * (w + 5) + 2, w + 5, d_mask, 2 * (w + 1) + 1, w + 1, w, h)
    src[1:1 + h, 2:2 + w] *= mask[2:2 + h, 1:1 + w, None]
    assert np.array_equal(rast.to_host(d_src, src.shape, np.float32), src)
    for opacity in (None, 0.37):
        rast.layer_blend_over(d_dst, 5 * W + 7, W, d_src, 1 * (w + 5) + 2, w + 5, w, h, opacity=opacity)
        s = src[1:1 + h, 2:2 + w] * (np.float32(opacity) if opacity is not None else np.float32(1.0))
        dst[5:5 + h, 7:7 + w] = s + dst[5:5 + h, 7:7 + w] * (np.float32(1.0) - s[..., 3:4])
        assert np.array_equal(rast.to_host(d_dst, dst.shape, np.float32), dst)
    for p in (d_dst, d_src, d_mask):
        rast.device_free(p)
