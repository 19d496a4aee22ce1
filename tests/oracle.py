"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package rasterize_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path as _P

import numpy as np

ROOT = _P(__file__).resolve().parent.parent
LIB_PATH = ROOT / "oracle" / "liboracle.so"


def build(force: bool = False) -> None:
    """Compile oracle/liboracle.so with the committed Makefile (gcc only, seconds)."""
    srcs = list((ROOT / "oracle").glob("*.hpp")) + [ROOT / "oracle" / "oracle_capi.cpp", ROOT / "oracle" / "oracle.h"]
    if not force and LIB_PATH.exists() and all(LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "-B", "liboracle.so"], check=True, capture_output=True)


class Shape(C.Structure):
    _fields_ = [("start", C.c_size_t), ("width", C.c_size_t), ("height", C.c_size_t), ("row_stride", C.c_size_t),
                ("col_stride", C.c_size_t)]

    @staticmethod
    def simple(height: int, width: int) -> "Shape":
        return Shape(0, width, height, width, 1)


class PixelS(C.Structure):
    _fields_ = [("x", C.c_size_t), ("y", C.c_size_t), ("alpha", C.c_double)]


class PaintDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("units", C.c_int), ("linear_colors", C.c_int), ("spread", C.c_int),
                ("tr", C.c_double * 6), ("p0", C.c_double * 2), ("p1", C.c_double * 2), ("dir", C.c_double * 2),
                ("r0", C.c_double), ("r1", C.c_double), ("solid", C.c_float * 4), ("n_stops", C.c_size_t)]


class FillJob(C.Structure):
    _fields_ = [("path", C.c_void_p), ("paint", C.c_void_p), ("fill_rule", C.c_int), ("tr", C.c_double * 6),
                ("bbox", C.c_double * 4)]


class PipeNode(C.Structure):
    _fields_ = [("kind", C.c_int), ("path", C.c_void_p), ("paint", C.c_void_p), ("fill_rule", C.c_int), ("tr", C.c_double * 6),
                ("bbox", C.c_double * 4), ("opacity", C.c_double), ("child", C.c_long), ("child_begin", C.c_long),
                ("child_count", C.c_long)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(LIB_PATH))
    vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    pd = C.POINTER(C.c_double)
    pf = C.POINTER(C.c_float)
    pu8 = C.POINTER(C.c_uint8)
    pu32 = C.POINTER(C.c_uint32)
    psz = C.POINTER(C.c_size_t)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("orc_last_error", C.c_char_p)
    sig("orc_set_simd_x86", None, i32)
    sig("orc_path_parse", vp, C.c_char_p, sz)
    sig("orc_path_from_flat", vp, pd, pu8, sz, pu32, sz, pu8)
    sig("orc_path_free", None, vp)
    sig("orc_path_counts", None, vp, psz, psz, psz)
    sig("orc_path_export", None, vp, pd, pu8, pu32, pu8)
    sig("orc_path_bbox", i32, vp, pd, pd)
    sig("orc_path_size", i32, vp, pd, psz, psz, pd, pd)
    sig("orc_path_stroke", vp, vp, dbl, i32, dbl, i32)
    sig("orc_path_transformed", vp, vp, pd)
    sig("orc_path_checkerboard", vp, pd, dbl)
    sig("orc_path_circle", vp, dbl, dbl, dbl)
    sig("orc_fit_size", None, pd, sz, sz, i32, psz, psz, pd)
    sig("orc_transform_parse", i32, C.c_char_p, pd)
    sig("orc_transform_mul", None, pd, pd, pd)
    sig("orc_transform_invert", i32, pd, pd)
    sig("orc_parse_scalar", dbl, C.c_char_p, sz, psz)
    sig("orc_flatten", C.c_long, vp, pd, dbl, i32, pd, sz)
    sig("orc_signed_difference_line", None, pd, sz, Shape, pd)
    sig("orc_signed_difference_to_mask", None, pd, Shape, i32)
    sig("orc_mask", i32, vp, pd, dbl, i32, pd, sz, Shape)
    sig("orc_mask_threads", i32, vp, pd, dbl, i32, pd, sz, sz, i32)
    sig("orc_mask_iter", C.c_long, vp, pd, dbl, sz, sz, i32, C.POINTER(PixelS), sz)
    sig("orc_batch_threads", i32, C.POINTER(vp), sz, pd, dbl, i32, vp, sz, sz, i32)
    sig("orc_fill", i32, vp, pd, dbl, i32, vp, pf, Shape)
    sig("orc_rgba_to_lin", None, pu8, pf)
    sig("orc_lin_to_rgba", None, pf, pu8)
    sig("orc_lin_to_rgba_image", None, pf, sz, pu8)
    sig("orc_parse_color", i32, C.c_char_p, pf)
    sig("orc_l2s", None, pf, pf)
    sig("orc_s2l", None, pf, pf)
    sig("orc_linear_to_srgb", C.c_float, C.c_float)
    sig("orc_srgb_to_linear", C.c_float, C.c_float)
    sig("orc_spread_at", dbl, i32, dbl)
    sig("orc_paint_solid", vp, pf)
    sig("orc_paint_linear", vp, pd, pf, sz, i32, i32, i32, i32, pd, pd, pd)
    sig("orc_paint_radial", vp, pd, pf, sz, i32, i32, i32, i32, pd, pd, dbl, pd, dbl)
    sig("orc_paint_linear_stored", vp, pd, pf, sz, i32, i32, i32, pd, pd, pd)
    sig("orc_paint_radial_stored", vp, pd, pf, sz, i32, i32, i32, pd, pd, dbl, pd, dbl)
    sig("orc_paint_free", None, vp)
    sig("orc_paint_at", None, vp, dbl, dbl, pf)
    sig("orc_paint_radial_offset", i32, vp, dbl, dbl, pd)
    sig("orc_paint_describe", None, vp, C.POINTER(PaintDesc))
    sig("orc_paint_stops", None, vp, pd, pf)
    sig("orc_scene_load_json", vp, C.c_char_p, sz)
    sig("orc_scene_cli_rasterize", vp, vp, pd, sz, sz)
    sig("orc_scene_many_circles", vp, C.c_uint32, sz, sz)
    sig("orc_scene_free", None, vp)
    sig("orc_scene_bbox", i32, vp, pd, pd)
    sig("orc_scene_render", vp, vp, dbl, pd, pd, pf)
    sig("orc_layer_info", None, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), psz, psz)
    sig("orc_layer_data", pf, vp)
    sig("orc_layer_free", None, vp)
    sig("orc_scene_fill_jobs", C.c_long, vp, pd, pd, C.POINTER(FillJob), sz)
    sig("orc_scene_pipeline", C.c_long, vp, pd, pd, C.POINTER(PipeNode), sz, C.POINTER(C.c_long), sz, psz)
    sig("orc_lcg_uniform", dbl, C.POINTER(C.c_uint32))
    sig("orc_glyph", vp, C.c_uint32)
    _lib = L
    return L


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pf(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _tr(tr):
    a = np.ascontiguousarray(np.asarray(tr, dtype=np.float64).reshape(6))
    return a


IDENTITY = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
NONZERO, EVENODD = 0, 1
DEFAULT_FLATNESS = 0.05


class OraclePath:
    """Handle to an oracle-side `Path` (src/path.rs:227-233)."""

    def __init__(self, handle, owned=True):
        if not handle:
            raise ValueError("oracle: " + lib().orc_last_error().decode())
        self.h = handle
        self.owned = owned

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            lib().orc_path_free(self.h)
            self.h = None

    @staticmethod
    def parse(svg: str | bytes) -> "OraclePath":
        b = svg.encode() if isinstance(svg, str) else svg
        return OraclePath(lib().orc_path_parse(b, len(b)))

    @staticmethod
    def from_flat(pts, kinds, sub_off, closed) -> "OraclePath":
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        sub_off = np.ascontiguousarray(sub_off, dtype=np.uint32)
        closed = np.ascontiguousarray(closed, dtype=np.uint8)
        return OraclePath(lib().orc_path_from_flat(_pd(pts), kinds.ctypes.data_as(C.POINTER(C.c_uint8)), len(kinds),
                                                   sub_off.ctypes.data_as(C.POINTER(C.c_uint32)), len(closed),
                                                   closed.ctypes.data_as(C.POINTER(C.c_uint8))))

    @staticmethod
    def glyph(seed: int) -> "OraclePath":
        return OraclePath(lib().orc_glyph(seed))

    @staticmethod
    def checkerboard(bbox, cell) -> "OraclePath":
        b = np.asarray(bbox, dtype=np.float64)
        return OraclePath(lib().orc_path_checkerboard(_pd(b), cell))

    @staticmethod
    def circle(cx, cy, r) -> "OraclePath":
        return OraclePath(lib().orc_path_circle(cx, cy, r))

    def counts(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib().orc_path_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export(self):
        """-> (pts[n_points,2] f64, kinds[n_segs] u8, sub_off[n_sub+1] u32, closed[n_sub] u8)"""
        ns, npnt, nsub = self.counts()
        pts = np.zeros((npnt, 2), dtype=np.float64)
        kinds = np.zeros(ns, dtype=np.uint8)
        sub = np.zeros(nsub + 1 if nsub else 0, dtype=np.uint32)
        closed = np.zeros(nsub, dtype=np.uint8)
        lib().orc_path_export(self.h, _pd(pts), kinds.ctypes.data_as(C.POINTER(C.c_uint8)),
                              sub.ctypes.data_as(C.POINTER(C.c_uint32)), closed.ctypes.data_as(C.POINTER(C.c_uint8)))
        return pts, kinds, sub, closed

    def bbox(self, tr=IDENTITY):
        out = np.zeros(4)
        if not lib().orc_path_bbox(self.h, _pd(_tr(tr)), _pd(out)):
            return None
        return out

    def size(self, tr=IDENTITY):
        w, h = C.c_size_t(), C.c_size_t()
        tro, mn = np.zeros(6), np.zeros(2)
        if not lib().orc_path_size(self.h, _pd(_tr(tr)), C.byref(w), C.byref(h), _pd(tro), _pd(mn)):
            return None
        return (w.value, h.value), tro, mn

    def stroke(self, width, join="miter", miter_limit=4.0, cap="butt") -> "OraclePath":
        j = {"miter": 0, "bevel": 1, "round": 2}[join]
        c = {"butt": 0, "square": 1, "round": 2}[cap]
        return OraclePath(lib().orc_path_stroke(self.h, width, j, miter_limit, c))

    def transformed(self, tr) -> "OraclePath":
        return OraclePath(lib().orc_path_transformed(self.h, _pd(_tr(tr))))

    def flatten(self, tr=IDENTITY, flatness=DEFAULT_FLATNESS, close=True) -> np.ndarray:
        t = _tr(tr)
        n = lib().orc_flatten(self.h, _pd(t), flatness, int(close), None, 0)
        if n < 0:
            raise ValueError(lib().orc_last_error().decode())
        out = np.zeros((n, 4), dtype=np.float64)
        lib().orc_flatten(self.h, _pd(t), flatness, int(close), _pd(out), n)
        return out

    def mask(self, tr, rule, img: np.ndarray, flatness=DEFAULT_FLATNESS, shape: Shape | None = None) -> np.ndarray:
        """`SignedDifferenceRasterizer::mask` into a zeroed f64 image (in place)."""
        assert img.dtype == np.float64 and img.flags.c_contiguous
        if shape is None:
            shape = Shape.simple(img.shape[0], img.shape[1])
        rc = lib().orc_mask(self.h, _pd(_tr(tr)), flatness, rule, _pd(img), img.size, shape)
        if rc:
            raise ValueError(lib().orc_last_error().decode())
        return img

    def mask_threads(self, tr, rule, img: np.ndarray, threads: int, flatness=DEFAULT_FLATNESS) -> np.ndarray:
        assert img.dtype == np.float64 and img.flags.c_contiguous and img.ndim == 2
        rc = lib().orc_mask_threads(self.h, _pd(_tr(tr)), flatness, rule, _pd(img), img.shape[1], img.shape[0], threads)
        if rc:
            raise ValueError(lib().orc_last_error().decode())
        return img

    def mask_iter(self, tr, w, h, rule, flatness=DEFAULT_FLATNESS):
        t = _tr(tr)
        n = lib().orc_mask_iter(self.h, _pd(t), flatness, w, h, rule, None, 0)
        buf = (PixelS * max(n, 1))()
        lib().orc_mask_iter(self.h, _pd(t), flatness, w, h, rule, buf, n)
        return [(buf[i].x, buf[i].y, buf[i].alpha) for i in range(n)]

    def fill(self, tr, rule, paint: "OraclePaint", img: np.ndarray, flatness=DEFAULT_FLATNESS, shape: Shape | None = None):
        """Default `Rasterizer::fill` into an f32 [H,W,4] LinColor image (in place)."""
        assert img.dtype == np.float32 and img.flags.c_contiguous
        if shape is None:
            shape = Shape.simple(img.shape[0], img.shape[1])
        rc = lib().orc_fill(self.h, _pd(_tr(tr)), flatness, rule, paint.h, _pf(img), shape)
        if rc:
            raise ValueError(lib().orc_last_error().decode())
        return img


def batch_threads(paths, tr, rule, paint, w: int, h: int, threads: int, flatness=DEFAULT_FLATNESS) -> None:
    """n independent paths on `threads` host threads (std::thread inside the oracle), each rendered into a private
    w x h image: clear + `Rasterizer::mask` (paint None) or clear + `Rasterizer::fill`.  Timing helper of bench.py."""
    arr = (C.c_void_p * len(paths))(*[p.h for p in paths])
    rc = lib().orc_batch_threads(arr, len(paths), _pd(_tr(tr)), flatness, rule, paint.h if paint is not None else None, w, h, threads)
    if rc:
        raise ValueError(lib().orc_last_error().decode())


class OraclePaint:
    def __init__(self, handle, owned=True):
        self.h = handle
        self.owned = owned

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            lib().orc_paint_free(self.h)
            self.h = None

    @staticmethod
    def solid(lin) -> "OraclePaint":
        a = np.asarray(lin, dtype=np.float32)
        return OraclePaint(lib().orc_paint_solid(_pf(a)))

    @staticmethod
    def linear(stops, start, end, units=0, linear_colors=False, spread=0, tr=IDENTITY, sort_stops=True) -> "OraclePaint":
        pos = np.ascontiguousarray([s[0] for s in stops], dtype=np.float64)
        col = np.ascontiguousarray([s[1] for s in stops], dtype=np.float32).reshape(-1, 4)
        s, e = np.asarray(start, dtype=np.float64), np.asarray(end, dtype=np.float64)
        return OraclePaint(lib().orc_paint_linear(_pd(pos), _pf(col), len(pos), int(sort_stops), units, int(linear_colors),
                                                  spread, _pd(_tr(tr)), _pd(s), _pd(e)))

    @staticmethod
    def radial(stops, center, radius, fcenter=None, fradius=0.0, units=0, linear_colors=False, spread=0, tr=IDENTITY,
               sort_stops=True) -> "OraclePaint":
        pos = np.ascontiguousarray([s[0] for s in stops], dtype=np.float64)
        col = np.ascontiguousarray([s[1] for s in stops], dtype=np.float32).reshape(-1, 4)
        c = np.asarray(center, dtype=np.float64)
        f = np.asarray(center if fcenter is None else fcenter, dtype=np.float64)
        return OraclePaint(lib().orc_paint_radial(_pd(pos), _pf(col), len(pos), int(sort_stops), units, int(linear_colors),
                                                  spread, _pd(_tr(tr)), _pd(c), radius, _pd(f), fradius))

    def at(self, x, y) -> np.ndarray:
        out = np.zeros(4, dtype=np.float32)
        lib().orc_paint_at(self.h, x, y, _pf(out))
        return out

    def radial_offset(self, x, y):
        o = C.c_double()
        if not lib().orc_paint_radial_offset(self.h, x, y, C.byref(o)):
            return None
        return o.value

    def describe(self) -> dict:
        d = PaintDesc()
        lib().orc_paint_describe(self.h, C.byref(d))
        n = d.n_stops
        pos = np.zeros(n, dtype=np.float64)
        col = np.zeros((n, 4), dtype=np.float32)
        if n:
            lib().orc_paint_stops(self.h, _pd(pos), _pf(col))
        return dict(kind=d.kind, units=d.units, linear_colors=d.linear_colors, spread=d.spread, tr=np.array(d.tr[:]),
                    p0=np.array(d.p0[:]), p1=np.array(d.p1[:]), dir=np.array(d.dir[:]), r0=d.r0, r1=d.r1,
                    solid=np.array(d.solid[:], dtype=np.float32), stop_pos=pos, stop_colors=col)


class OracleScene:
    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle: " + lib().orc_last_error().decode())
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_free(self.h)
            self.h = None

    @staticmethod
    def load_json(text: str | bytes) -> "OracleScene":
        b = text.encode() if isinstance(text, str) else text
        return OracleScene(lib().orc_scene_load_json(b, len(b)))

    @staticmethod
    def cli_rasterize(path: OraclePath, tr, w, h) -> "OracleScene":
        return OracleScene(lib().orc_scene_cli_rasterize(path.h, _pd(_tr(tr)), w, h))

    @staticmethod
    def many_circles(seed=0, count=1024, size=1024) -> "OracleScene":
        return OracleScene(lib().orc_scene_many_circles(seed, count, size))

    def bbox(self, tr=IDENTITY):
        out = np.zeros(4)
        if not lib().orc_scene_bbox(self.h, _pd(_tr(tr)), _pd(out)):
            return None
        return out

    def render(self, tr=IDENTITY, view=None, bg=None, flatness=DEFAULT_FLATNESS):
        """`Scene::render` -> (x, y, LinColor image [H,W,4] f32)"""
        v = None if view is None else np.asarray(view, dtype=np.float64)
        b = None if bg is None else np.asarray(bg, dtype=np.float32)
        lay = lib().orc_scene_render(self.h, flatness, _pd(_tr(tr)), None if v is None else _pd(v), None if b is None else _pf(b))
        if not lay:
            raise ValueError(lib().orc_last_error().decode())
        x, y, w, h = C.c_int32(), C.c_int32(), C.c_size_t(), C.c_size_t()
        lib().orc_layer_info(lay, C.byref(x), C.byref(y), C.byref(w), C.byref(h))
        n = w.value * h.value * 4
        if n:
            img = np.ctypeslib.as_array(lib().orc_layer_data(lay), shape=(n,)).reshape(h.value, w.value, 4).copy()
        else:
            img = np.zeros((h.value, w.value, 4), dtype=np.float32)
        lib().orc_layer_free(lay)
        return x.value, y.value, img

    def fill_jobs(self, tr=IDENTITY, view=None):
        """Fill nodes of Pipeline::build in render order: list of dict(path, paint, fill_rule, tr, bbox)."""
        v = None if view is None else np.asarray(view, dtype=np.float64)
        vp = None if v is None else _pd(v)
        n = lib().orc_scene_fill_jobs(self.h, _pd(_tr(tr)), vp, None, 0)
        if n < 0:
            raise ValueError("scene has clip/opacity nodes")
        buf = (FillJob * max(n, 1))()
        lib().orc_scene_fill_jobs(self.h, _pd(_tr(tr)), vp, buf, n)
        jobs = []
        for i in range(n):
            j = buf[i]
            jobs.append(dict(path=OraclePath(j.path, owned=False), paint=OraclePaint(j.paint, owned=False), fill_rule=j.fill_rule,
                             tr=np.array(j.tr[:]), bbox=np.array(j.bbox[:]), _keep=self))
        return jobs


def _pipeline(self, tr=IDENTITY, view=None):
    """Node table of Pipeline::build (children before parents, root last): list of dict(kind, path, paint, fill_rule, tr,
    bbox, opacity, child, children)."""
    v = None if view is None else np.asarray(view, dtype=np.float64)
    vp = None if v is None else _pd(v)
    nch = C.c_size_t()
    n = lib().orc_scene_pipeline(self.h, _pd(_tr(tr)), vp, None, 0, None, 0, C.byref(nch))
    buf = (PipeNode * max(n, 1))()
    ch = (C.c_long * max(nch.value, 1))()
    lib().orc_scene_pipeline(self.h, _pd(_tr(tr)), vp, buf, n, ch, nch.value, C.byref(nch))
    nodes = []
    for i in range(n):
        j = buf[i]
        nodes.append(dict(kind=j.kind, path=OraclePath(j.path, owned=False) if j.path else None,
                          paint=OraclePaint(j.paint, owned=False) if j.paint else None, fill_rule=j.fill_rule, tr=np.array(j.tr[:]),
                          bbox=np.array(j.bbox[:]), opacity=j.opacity, child=j.child,
                          children=[ch[k] for k in range(j.child_begin, j.child_begin + j.child_count)], _keep=self))
    return nodes


OracleScene.pipeline = _pipeline


def fit_size(bbox, w, h, align=1):
    ow, oh = C.c_size_t(), C.c_size_t()
    tr = np.zeros(6)
    b = np.asarray(bbox, dtype=np.float64)
    lib().orc_fit_size(_pd(b), w, h, align, C.byref(ow), C.byref(oh), _pd(tr))
    return (ow.value, oh.value), tr


def transform_parse(text: str) -> np.ndarray:
    out = np.zeros(6)
    if not lib().orc_transform_parse(text.encode(), _pd(out)):
        raise ValueError(lib().orc_last_error().decode())
    return out


def transform_mul(a, b) -> np.ndarray:
    out = np.zeros(6)
    lib().orc_transform_mul(_pd(_tr(a)), _pd(_tr(b)), _pd(out))
    return out


def transform_invert(a):
    out = np.zeros(6)
    if not lib().orc_transform_invert(_pd(_tr(a)), _pd(out)):
        return None
    return out


def translate(tx, ty):
    return np.array([1.0, 0.0, tx, 0.0, 1.0, ty])


def scale(sx, sy):
    return np.array([sx, 0.0, 0.0, 0.0, sy, 0.0])


def rotate(a):
    import math
    s, c = math.sin(a), math.cos(a)
    return np.array([c, -s, 0.0, s, c, 0.0])


def parse_scalar(text: str):
    b = text.encode()
    n = C.c_size_t()
    v = lib().orc_parse_scalar(b, len(b), C.byref(n))
    return v, n.value


def signed_difference_line(img: np.ndarray, line, shape: Shape | None = None):
    assert img.dtype == np.float64 and img.flags.c_contiguous
    if shape is None:
        shape = Shape.simple(img.shape[0], img.shape[1])
    l = np.asarray(line, dtype=np.float64).reshape(4)
    lib().orc_signed_difference_line(_pd(img), img.size, shape, _pd(l))


def parse_color(text: str) -> np.ndarray:
    out = np.zeros(4, dtype=np.float32)
    if not lib().orc_parse_color(text.encode(), _pf(out)):
        raise ValueError("bad color " + text)
    return out


def lin_to_rgba(lin) -> np.ndarray:
    a = np.ascontiguousarray(lin, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((a.shape[0], 4), dtype=np.uint8)
    lib().orc_lin_to_rgba_image(_pf(a), a.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.reshape(np.asarray(lin).shape)


def rgba_to_lin(rgba) -> np.ndarray:
    a = np.asarray(rgba, dtype=np.uint8)
    out = np.zeros(4, dtype=np.float32)
    lib().orc_rgba_to_lin(a.ctypes.data_as(C.POINTER(C.c_uint8)), _pf(out))
    return out


def l2s(v):
    a = np.asarray(v, dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    lib().orc_l2s(_pf(a), _pf(out))
    return out


def s2l(v):
    a = np.asarray(v, dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    lib().orc_s2l(_pf(a), _pf(out))
    return out


def lcg_uniform(state: int):
    s = C.c_uint32(state)
    v = lib().orc_lcg_uniform(C.byref(s))
    return v, s.value


def set_simd_x86(on: bool):
    lib().orc_set_simd_x86(int(on))
