"""Shared test helpers: Fill-only `Scene::render` on the GPU (device-resident layer) and on the oracle."""
import math

import numpy as np

import oracle as O
import rasterize_b200 as rb
from rasterize_b200 import ffi


def opath(p):
    return O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)


def render_scene_gpu(rast, sc, to_rgba=True):
    """Scene::render for Fill-only pipelines (reference src/scene.rs:384-430): layer = view, fills in order."""
    x0, y0, x1, y1 = sc.view
    lx, ly = math.floor(x0), math.floor(y0)
    W, H = math.ceil(x1) - lx, math.ceil(y1) - ly
    layer = rast.device_alloc(W * H * 16)
    if sc.bg is not None:
        rast.fill_color(layer, W * H, sc.bg)
    else:
        rast.device_zero(layer, W * H * 16)
    jobs, keep = [], []
    for f in sc.fills:
        bx0, by0, bx1, by1 = f.bbox
        col_min = max(0, min(math.floor(bx0) - lx, W))
        col_max = max(col_min, min(math.ceil(bx1) - lx + 1, W))
        row_min = max(0, min(math.floor(by0) - ly, H))
        row_max = max(row_min, min(math.ceil(by1) - ly + 1, H))
        align = rb.Transform.new_translate(-math.floor(bx0), -math.floor(by0))
        tr = align * rb.Transform.from_array(f.tr)
        dp = rast.upload(f.path)
        keep.append(dp)
        jobs.append(rb.Job(dp, tr, f.fill_rule, ffi.JOB_FILL, layer, col_max - col_min, row_max - row_min, W,
                           origin=row_min * W + col_min, paint=f.paint, path_bbox=f.path_bbox))
    rast.render_batch(jobs, independent=False)
    lin = rast.to_host(layer, (H, W, 4), np.float32)
    rgba = None
    if to_rgba:
        out = rast.device_alloc(W * H * 4)
        rast.to_rgba8(layer, out, W * H)
        rgba = rast.to_host(out, (H, W, 4), np.uint8)
        rast.device_free(out)
    rast.device_free(layer)
    return lin, rgba


def render_scene_oracle(sc):
    """Same pipeline on the oracle: fills applied in order with the oracle's `fill` on sub-views."""
    x0, y0, x1, y1 = sc.view
    lx, ly = math.floor(x0), math.floor(y0)
    W, H = math.ceil(x1) - lx, math.ceil(y1) - ly
    img = np.zeros((H, W, 4), dtype=np.float32)
    if sc.bg is not None:
        img[:] = sc.bg
    for f in sc.fills:
        bx0, by0, bx1, by1 = f.bbox
        col_min = max(0, min(math.floor(bx0) - lx, W))
        col_max = max(col_min, min(math.ceil(bx1) - lx + 1, W))
        row_min = max(0, min(math.floor(by0) - ly, H))
        row_max = max(row_min, min(math.ceil(by1) - ly + 1, H))
        align = O.translate(-math.floor(bx0), -math.floor(by0))
        tr = O.transform_mul(align, f.tr)
        shape = O.Shape(row_min * W + col_min, col_max - col_min, row_max - row_min, W, 1)
        opath(f.path).fill(tr, int(f.fill_rule), oracle_paint(f.paint_desc), img, shape=shape)
    return img


def oracle_paint(d):
    """Oracle paint from a fixture description whose stop colours are already in STORED space."""
    import ctypes as C
    L = O.lib()
    if d["kind"] == 0:
        return O.OraclePaint.solid(d["solid"])
    pos = np.ascontiguousarray(d["stop_pos"], dtype=np.float64)
    col = np.ascontiguousarray(d["stop_colors"], dtype=np.float32).reshape(-1, 4)
    tr = np.ascontiguousarray(d["tr"], dtype=np.float64)
    p0 = np.ascontiguousarray(d["p0"], dtype=np.float64)
    p1 = np.ascontiguousarray(d["p1"], dtype=np.float64)
    pd, pf = C.POINTER(C.c_double), C.POINTER(C.c_float)
    a = lambda x, t: x.ctypes.data_as(t)  # noqa: E731
    if d["kind"] == 1:
        h = L.orc_paint_linear_stored(a(pos, pd), a(col, pf), len(pos), d["units"], int(d["linear_colors"]), d["spread"], a(tr, pd),
                                      a(p0, pd), a(p1, pd))
    else:
        h = L.orc_paint_radial_stored(a(pos, pd), a(col, pf), len(pos), d["units"], int(d["linear_colors"]), d["spread"], a(tr, pd),
                                      a(p0, pd), float(d["r0"]), a(p1, pd), float(d["r1"]))
    return O.OraclePaint(h)


