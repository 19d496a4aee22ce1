"""Shared test helpers: Fill-only `Scene::render` on the GPU (device-resident layer) and on the oracle."""
import math

import numpy as np

import oracle as O
import rasterize_b200 as rb
from rasterize_b200 import ffi
from rasterize_b200.scene import fill_window


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
        bx0, by0 = f.bbox[0], f.bbox[1]
        col_min, row_min, ww, wh = fill_window(f.bbox, lx, ly, W, H)  # the reference's window, incl. its `as usize` wrap
        align = rb.Transform.new_translate(-math.floor(bx0), -math.floor(by0))
        tr = align * rb.Transform.from_array(f.tr)
        dp = rast.upload(f.path)
        keep.append(dp)
        jobs.append(rb.Job(dp, tr, f.fill_rule, ffi.JOB_FILL, layer, ww, wh, W, origin=row_min * W + col_min, paint=f.paint,
                           path_bbox=f.path_bbox))
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
        bx0, by0 = f.bbox[0], f.bbox[1]
        col_min, row_min, ww, wh = fill_window(f.bbox, lx, ly, W, H)
        align = O.translate(-math.floor(bx0), -math.floor(by0))
        tr = O.transform_mul(align, f.tr)
        shape = O.Shape(row_min * W + col_min, ww, wh, W, 1)
        if ww and wh:
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




# ---- full pipelines (clip / opacity) -------------------------------------------------------------------------
def _layer_geom(bbox):
    x0, y0 = math.floor(bbox[0]), math.floor(bbox[1])
    return int(x0), int(y0), max(0, int(math.ceil(bbox[2])) - int(x0)), max(0, int(math.ceil(bbox[3])) - int(y0))


def _intersect(a, b):
    """(ax, ay, aw, ah), (bx, by, bw, bh) -> slices into a and b of Layer::compose's rectangle (src/scene.rs:540-549)."""
    x0, x1 = max(a[0], b[0]), min(a[0] + a[2], b[0] + b[2])
    y0, y1 = max(a[1], b[1]), min(a[1] + a[3], b[1] + b[3])
    if x1 <= x0 or y1 <= y0:
        return None
    return (slice(y0 - a[1], y1 - a[1]), slice(x0 - a[0], x1 - a[0])), (slice(y0 - b[1], y1 - b[1]), slice(x0 - b[0], x1 - b[0]))


def _blend_over(dst, src):
    """f32, unfused: other + self * (1 - other.alpha), src/color.rs:342-344"""
    k = np.float32(1.0) - src[..., 3:4]
    return src + dst * k


def render_pipeline_oracle(pl):
    """`Pipeline::render_rec` (src/scene.rs:397-459) with the oracle's `fill` / `mask` and numpy f32 layer composition."""
    nodes = pl.nodes

    def render(node_id, view, bg):
        geom = _layer_geom(view if view is not None else nodes[node_id].bbox)
        img = np.zeros((geom[3], geom[2], 4), dtype=np.float32)
        if bg is not None:
            img[:] = bg
        rec(node_id, img, geom)
        return img, geom

    def rec(node_id, img, geom):
        n = nodes[node_id]
        lx, ly, W, H = geom
        if n.kind == 0:
            col_min, row_min, ww, wh = fill_window(n.bbox, lx, ly, W, H)
            tr = O.transform_mul(O.translate(-math.floor(n.bbox[0]), -math.floor(n.bbox[1])), n.tr)
            shape = O.Shape(row_min * W + col_min, ww, wh, W, 1)
            if shape.width and shape.height:
                opath(n.path).fill(tr, int(n.fill_rule), oracle_paint(n.paint_desc), img, shape=shape)
        elif n.kind == 1:
            for c in n.children:
                rec(c, img, geom)
        elif n.kind == 2:
            child, cg = render(n.child, None, None)
            r = _intersect(geom, cg)
            if r:
                img[r[0]] = _blend_over(img[r[0]], child[r[1]] * np.float32(n.opacity))
        elif n.kind == 3:
            mg = _layer_geom(n.bbox)
            mask = np.zeros((mg[3], mg[2]), dtype=np.float64)
            child, cg = render(n.child, None, None)
            if mask.size:
                opath(n.path).mask(O.transform_mul(O.translate(-float(mg[0]), -float(mg[1])), n.tr), int(n.fill_rule), mask)
            r = _intersect(cg, mg)
            if r:
                child[r[0]] = child[r[0]] * mask[r[1]].astype(np.float32)[..., None]
            r = _intersect(geom, cg)
            if r:
                img[r[0]] = _blend_over(img[r[0]], child[r[1]])

    if not nodes:
        return 0, 0, np.zeros((0, 0, 4), dtype=np.float32)
    img, geom = render(len(nodes) - 1, pl.view, pl.bg)
    return geom[0], geom[1], img
