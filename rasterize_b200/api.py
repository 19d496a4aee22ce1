"""Host-side mirror of the reference's interface for the fill path, on top of the C ABI.

Names, argument meaning and error behaviour follow aslpavel/rasterize v0.6.7 (paths relative to the reference):
`Path` / `PathBuilder` (src/path.rs:227-233, 800-1056), `Transform` (src/geometry.rs:317-539), `FillRule`
(src/path.rs:21-29), `Size`, `Rasterizer::{name, mask, mask_iter, fill}` (src/rasterize.rs:44-101), `Units`,
`LinColor` / `GradLinear` / `GradRadial` paints (src/color.rs:357-374, src/grad.rs:150-226, 307-426).

Everything that computes goes through librasterize_b200.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import enum
import math
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

from . import ffi
from .ffi import RgpuError

DEFAULT_FLATNESS = 0.05  # src/path.rs:16
EPSILON = float(np.finfo(np.float64).eps)


class FillRule(enum.IntEnum):
    NonZero = 0
    EvenOdd = 1


class LineJoin(enum.IntEnum):
    """`LineJoin` (src/path.rs:78-93); the miter limit of `Miter(limit)` lives in `StrokeStyle.miter_limit`."""
    Miter = 0
    Bevel = 1
    Round = 2


class LineCap(enum.IntEnum):
    """`LineCap` (src/path.rs:103-110)"""
    Butt = 0
    Square = 1
    Round = 2


@dataclass
class StrokeStyle:
    """`StrokeStyle` (src/path.rs:121-136); defaults as `LineJoin::default()` = `Miter(4.0)` and `LineCap::default()` = `Butt`."""
    width: float
    line_join: LineJoin = LineJoin.Miter
    miter_limit: float = 4.0
    line_cap: LineCap = LineCap.Butt

    def _c(self) -> "ffi.CStrokeStyle":
        return ffi.CStrokeStyle(float(self.width), float(self.miter_limit), int(self.line_join), int(self.line_cap))


class Units(enum.IntEnum):
    UserSpaceOnUse = 0
    BoundingBox = 1


class GradSpread(enum.IntEnum):
    Pad = 0
    Repeat = 1
    Reflect = 2


@dataclass(frozen=True)
class Size:
    width: int
    height: int


class Transform:
    """2x3 affine `[m00, m01, m02, m10, m11, m12]` (src/geometry.rs:317)."""

    __slots__ = ("m",)

    def __init__(self, m00=1.0, m01=0.0, m02=0.0, m10=0.0, m11=1.0, m12=0.0):
        self.m = (float(m00), float(m01), float(m02), float(m10), float(m11), float(m12))

    @staticmethod
    def identity() -> "Transform":
        return Transform()

    @staticmethod
    def from_array(a) -> "Transform":
        return Transform(*[float(v) for v in np.asarray(a).reshape(6)])

    @staticmethod
    def new_translate(tx, ty) -> "Transform":
        return Transform(1.0, 0.0, tx, 0.0, 1.0, ty)

    @staticmethod
    def new_scale(sx, sy) -> "Transform":
        return Transform(sx, 0.0, 0.0, 0.0, sy, 0.0)

    @staticmethod
    def new_rotate(a) -> "Transform":
        s, c = math.sin(a), math.cos(a)
        return Transform(c, -s, 0.0, s, c, 0.0)

    def __mul__(self, o: "Transform") -> "Transform":  # src/geometry.rs:519-539
        s, t = self.m, o.m
        return Transform(s[0] * t[0] + s[1] * t[3], s[0] * t[1] + s[1] * t[4], s[0] * t[2] + s[1] * t[5] + s[2],
                         s[3] * t[0] + s[4] * t[3], s[3] * t[1] + s[4] * t[4], s[3] * t[2] + s[4] * t[5] + s[5])

    def pre_translate(self, tx, ty) -> "Transform":
        return self * Transform.new_translate(tx, ty)

    def pre_scale(self, sx, sy) -> "Transform":
        return self * Transform.new_scale(sx, sy)

    def pre_rotate(self, a) -> "Transform":
        return self * Transform.new_rotate(a)

    def apply(self, p):  # src/geometry.rs:363-367
        m = self.m
        x, y = p
        return (x * m[0] + y * m[1] + m[2], x * m[3] + y * m[4] + m[5])

    def array(self) -> np.ndarray:
        return np.array(self.m, dtype=np.float64)

    def __repr__(self):
        return "Transform(%r, %r, %r, %r, %r, %r)" % self.m


def _as_tr(tr) -> np.ndarray:
    if isinstance(tr, Transform):
        return tr.array()
    return np.ascontiguousarray(np.asarray(tr, dtype=np.float64).reshape(6))


class Path:
    """Flat `Path { segments, subpaths, closed }` (src/path.rs:227-233): the encoding that crosses the FFI."""

    def __init__(self, points, kinds, subpath_offsets, closed):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        self.kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        self.closed = np.ascontiguousarray(closed, dtype=np.uint8)
        so = np.ascontiguousarray(subpath_offsets, dtype=np.uint32)
        if len(self.closed) == 0:
            so = np.zeros(0, dtype=np.uint32)
        self.subpath_offsets = so

    @staticmethod
    def empty() -> "Path":
        return Path(np.zeros((0, 2)), [], [], [])

    @staticmethod
    def builder() -> "PathBuilder":
        return PathBuilder()

    @staticmethod
    def load_npz(file) -> "Path":
        z = np.load(file)
        return Path(z["points"], z["kinds"], z["subpath_offsets"], z["closed"])

    def save_npz(self, file) -> None:
        np.savez_compressed(file, points=self.points, kinds=self.kinds, subpath_offsets=self.subpath_offsets, closed=self.closed)

    def segments_count(self) -> int:  # src/path.rs:366-368
        return int(len(self.kinds))

    def is_empty(self) -> bool:
        return len(self.closed) == 0

    def transform(self, tr) -> "Path":
        """`Path::transform` (src/path.rs:359-363) — returns a new path"""
        m = _as_tr(tr)
        x, y = self.points[:, 0], self.points[:, 1]
        pts = np.stack([x * m[0] + y * m[1] + m[2], x * m[3] + y * m[4] + m[5]], axis=1)
        return Path(pts, self.kinds, self.subpath_offsets, self.closed)

    def to_svg_path(self) -> str:
        """The path as an SVG path string (absolute M / L / Q / C / Z, shortest round-trip decimals).  `Path::write_svg_path`
        (src/path.rs:537-541) prints 4 significant digits; this one keeps every digit.  NOTE: the reference's scanner is not
        correctly rounded ((i64 mantissa) * powi(10, e)), so parsing the text again may differ from `self` in the last bit."""
        pts = self.points
        out, at = [], 0
        for s in range(len(self.closed)):
            for k in range(int(self.subpath_offsets[s]), int(self.subpath_offsets[s + 1])):
                n = int(self.kinds[k])
                if k == self.subpath_offsets[s]:
                    out.append(f"M{float(pts[at][0])!r},{float(pts[at][1])!r}")
                out.append("LQC"[n - 2] + " ".join(f"{float(x)!r},{float(y)!r}" for x, y in pts[at + 1:at + n]))
                at += n
            if self.closed[s]:
                out.append("Z")
        return "".join(out)

    def input_bytes(self) -> int:
        """Algorithmic input bytes: 16 B per control point."""
        return int(self.points.shape[0]) * 16

    def _c(self) -> ffi.CPath:
        c = ffi.CPath()
        c.points = self.points.ctypes.data_as(C.POINTER(C.c_double))
        c.kinds = self.kinds.ctypes.data_as(C.POINTER(C.c_uint8))
        c.subpath_offsets = self.subpath_offsets.ctypes.data_as(C.POINTER(C.c_uint32))
        c.closed = self.closed.ctypes.data_as(C.POINTER(C.c_uint8))
        c.n_points = self.points.shape[0]
        c.n_segments = len(self.kinds)
        c.n_subpaths = len(self.closed)
        return c

    # reference convenience methods: `Path::flatten/mask/fill` delegate to the rasterizer
    def flatten(self, rasterizer: "GpuRasterizer", tr=Transform(), close: bool = True) -> np.ndarray:
        return rasterizer.flatten(self, tr, close)

    def mask(self, rasterizer: "GpuRasterizer", tr, fill_rule: FillRule, img: np.ndarray) -> np.ndarray:
        rasterizer.mask(self, tr, img, fill_rule)
        return img

    def fill(self, rasterizer: "GpuRasterizer", tr, fill_rule: FillRule, paint, img: np.ndarray, bbox=None) -> np.ndarray:
        rasterizer.fill(self, tr, fill_rule, paint, img, bbox=bbox)
        return img


class PathBatch:
    """Many independent paths in one flat encoding: the subpaths of every path back to back plus, per path, the range of
    subpaths it owns (`rgpu_path` + `path_subpath_offsets`, include/rasterize_b200.h)."""

    def __init__(self, points, kinds, subpath_offsets, closed, path_subpath_offsets):
        self.flat = Path(points, kinds, subpath_offsets, closed)
        self.path_subpath_offsets = np.ascontiguousarray(path_subpath_offsets, dtype=np.uint32)
        if len(self.path_subpath_offsets) == 0:
            self.path_subpath_offsets = np.zeros(1, dtype=np.uint32)

    @staticmethod
    def from_paths(paths) -> "PathBatch":
        paths = list(paths)
        seg = np.cumsum([0] + [p.segments_count() for p in paths])
        sub = np.cumsum([0] + [len(p.closed) for p in paths]).astype(np.uint32)
        so = [np.zeros(1, dtype=np.uint32)]
        for p, s0 in zip(paths, seg[:-1]):
            if len(p.closed):
                so.append(p.subpath_offsets[1:].astype(np.uint32) + np.uint32(s0))
        return PathBatch(np.concatenate([p.points for p in paths]) if paths else np.zeros((0, 2)),
                         np.concatenate([p.kinds for p in paths]) if paths else [], np.concatenate(so),
                         np.concatenate([p.closed for p in paths]) if paths else [], sub)

    def __len__(self) -> int:
        return len(self.path_subpath_offsets) - 1

    def _seg_range(self, a: int, b: int):
        so = self.flat.subpath_offsets
        pso = self.path_subpath_offsets
        n_seg = len(self.flat.kinds)
        s0 = int(so[pso[a]]) if len(so) and pso[a] < len(so) else n_seg
        s1 = int(so[pso[b]]) if len(so) and pso[b] < len(so) else n_seg
        return s0, s1

    def slice(self, a: int, b: int) -> "PathBatch":
        """Paths [a, b) as a batch of their own (rebased offsets; arrays are copies)."""
        pso = self.path_subpath_offsets
        s0, s1 = self._seg_range(a, b)
        pt = np.concatenate([[0], np.cumsum(self.flat.kinds, dtype=np.int64)])
        sub0, sub1 = int(pso[a]), int(pso[b])
        so = self.flat.subpath_offsets[sub0:sub1 + 1].astype(np.int64) - s0 if sub1 > sub0 else np.zeros(0, dtype=np.int64)
        return PathBatch(self.flat.points[pt[s0]:pt[s1]], self.flat.kinds[s0:s1], so, self.flat.closed[sub0:sub1], pso[a:b + 1] - pso[a])

    def path(self, i: int) -> Path:
        return self.slice(i, i + 1).flat

    def segments_per_path(self) -> np.ndarray:
        so = self.flat.subpath_offsets.astype(np.int64)
        if len(so) == 0:
            return np.zeros(len(self), dtype=np.int64)
        b = so[self.path_subpath_offsets]
        return b[1:] - b[:-1]

    def input_bytes(self) -> int:
        return self.flat.input_bytes()


def _angle_between(a, b):
    """`Point::angle_between` (src/geometry.rs:186-203); None when a vector has (almost) no length."""
    lengths = math.hypot(a[0], a[1]) * math.hypot(b[0], b[1])
    if lengths < EPSILON:
        return None
    c = np.float64(a[0] * b[0] + a[1] * b[1]) / np.float64(lengths)
    if c != c:
        return float("nan")
    angle = math.acos(min(max(float(c), -1.0), 1.0))
    return -angle if (a[0] * b[1] - a[1] * b[0]) < 0.0 else angle


def _rotate(phi, p):
    """`Transform::new_rotate(phi).apply(p)` (src/geometry.rs:363-367, 409-412)"""
    s, c = math.sin(phi), math.cos(phi)
    return (p[0] * c + p[1] * -s + 0.0, p[0] * s + p[1] * c + 0.0)


def ellip_arc_cubics(src, dst, rx, ry, x_axis_rot, large_flag, sweep_flag):
    """`EllipArc::new_param(..).to_cubics()` (src/ellipse.rs:40-96, 167-214): the cubics (4 points each) of an SVG endpoint arc,
    or None when the arc is degenerate (`arc_to` then draws a line).  IEEE semantics as in Rust (numpy scalars: x / 0 = inf,
    0 / 0 = NaN): a zero sweep gives a NaN step and NO cubic."""
    f = np.float64
    with np.errstate(all="ignore"):
        rx, ry = f(abs(rx)), f(abs(ry))
        phi = x_axis_rot * math.pi / 180.0
        x1, y1 = _rotate(-phi, (0.5 * (src[0] - dst[0]), 0.5 * (src[1] - dst[1])))
        x1, y1 = f(x1), f(y1)
        ax, ay = x1 / rx, y1 / ry
        s = ax * ax + ay * ay
        if s > 1.0:
            sq = np.sqrt(s)
            rx, ry = rx * sq, ry * sq
        rxry, rxy1, ryx1 = rx * ry, rx * y1, ry * x1
        q = rxry * rxry / (rxy1 * rxy1 + ryx1 * ryx1) - 1.0
        q = q if q > 0.0 else f(0.0)  # f64::max(0.0); NaN.max(0.0) == 0.0
        sq = np.sqrt(q)
        sq = -sq if large_flag == sweep_flag else sq
        cx, cy = sq * (rx * y1 / ry), sq * (-ry * x1 / rx)
        rc = _rotate(phi, (float(cx), float(cy)))
        center = (rc[0] + 0.5 * (dst[0] + src[0]), rc[1] + 0.5 * (dst[1] + src[1]))
        v1 = (float((x1 - cx) / rx), float((y1 - cy) / ry))
        v2 = (float((-x1 - cx) / rx), float((-y1 - cy) / ry))
        eta = _angle_between((1.0, 0.0), v1)
        if eta is None:
            return None
        ed = _angle_between(v1, v2)
        if ed is None:
            return None
        if eta != eta or ed != ed:
            # NaN angles (coincident end points, a zero radius): the reference's cubic iterator then never terminates
            # (`segment_index > NaN` is false for ever, src/ellipse.rs:198-201).  Treated like the degenerate arcs it does detect.
            return None
        two_pi = 2.0 * math.pi
        eta_delta = math.fmod(ed, two_pi)
        if eta_delta < 0.0:
            eta_delta += two_pi
        if not sweep_flag and eta_delta > 0.0:
            eta_delta -= two_pi
        elif sweep_flag and eta_delta < 0.0:
            eta_delta += two_pi
        rx, ry = float(rx), float(ry)
        count = f(math.ceil(abs(eta_delta) / (math.pi / 2.0)))
        delta = f(eta_delta) / count
        index, count = 0.0, float(count) - 1.0

        def at(alpha):
            sn, cs = math.sin(alpha), math.cos(alpha)
            a = _rotate(phi, (rx * cs, ry * sn))
            return (a[0] + center[0], a[1] + center[1]), _rotate(phi, (-rx * sn, ry * cs))

        out = []
        delta = float(delta)
        while not (index > count):
            eta_1 = eta + delta * index
            eta_2 = eta_1 + delta
            index += 1.0
            tn = math.tan((eta_2 - eta_1) / 2.0)
            sq = math.sqrt(4.0 + 3.0 * (tn * tn))
            alpha = math.sin(eta_2 - eta_1) * (sq - 1.0) / 3.0
            p0, d0 = at(eta_1)
            p3, d3 = at(eta_2)
            out.append((p0, (p0[0] + alpha * d0[0], p0[1] + alpha * d0[1]), (p3[0] - alpha * d3[0], p3[1] - alpha * d3[1]), p3))
        return out


class PathBuilder:
    """`PathBuilder` (src/path.rs:800-1056).  Arcs are converted to cubics here, on the host, exactly as the reference's
    `arc_to` does (a `Path` stores only lines, quads and cubics): the device never sees an arc."""

    def __init__(self):
        self.position = (0.0, 0.0)
        self._pts: list = []
        self._kinds: list = []
        self._sub: list = []
        self._closed: list = []

    def _finish(self, close: bool):  # src/path.rs:849-868
        n = len(self._kinds)
        if n == 0 or (self._sub and self._sub[-1] == n):
            return
        if not self._sub:
            self._sub.append(0)
        if close:
            first = self._sub[-1]
            off = sum(self._kinds[:first])
            self.position = self._pts[off]
        self._sub.append(n)
        self._closed.append(1 if close else 0)

    def move_to(self, p) -> "PathBuilder":
        self._finish(False)
        self.position = (float(p[0]), float(p[1]))
        return self

    def close(self) -> "PathBuilder":
        self._finish(True)
        return self

    def line_to(self, p) -> "PathBuilder":  # src/path.rs:895-903
        p = (float(p[0]), float(p[1]))
        if not (abs(self.position[0] - p[0]) < EPSILON and abs(self.position[1] - p[1]) < EPSILON):
            self._pts += [self.position, p]
            self._kinds.append(2)
            self.position = p
        return self

    def quad_to(self, p1, p2) -> "PathBuilder":
        p1, p2 = (float(p1[0]), float(p1[1])), (float(p2[0]), float(p2[1]))
        self._pts += [self.position, p1, p2]
        self._kinds.append(3)
        self.position = p2
        return self

    def cubic_to(self, p1, p2, p3) -> "PathBuilder":
        p1, p2, p3 = (float(p1[0]), float(p1[1])), (float(p2[0]), float(p2[1])), (float(p3[0]), float(p3[1]))
        self._pts += [self.position, p1, p2, p3]
        self._kinds.append(4)
        self.position = p3
        return self

    def arc_to(self, radii, x_axis_rot: float, large: bool, sweep: bool, p) -> "PathBuilder":  # src/path.rs:945-972
        p = (float(p[0]), float(p[1]))
        cubics = ellip_arc_cubics(self.position, p, float(radii[0]), float(radii[1]), float(x_axis_rot), bool(large), bool(sweep))
        if cubics is None:
            return self.line_to(p)
        for c in cubics:
            self._pts += list(c)
            self._kinds.append(4)
        self.position = p
        return self

    def build(self) -> Path:
        self._finish(False)
        p = Path(np.array(self._pts, dtype=np.float64).reshape(-1, 2), self._kinds, self._sub, self._closed)
        self.__init__()
        return p


# ---- paints -------------------------------------------------------------------------------------------
class LinColor:
    """Premultiplied linear RGBA f32x4; as a paint: `impl Paint for LinColor` (src/color.rs:357-374)."""

    def __init__(self, r, g, b, a):
        self.c = np.array([r, g, b, a], dtype=np.float32)

    def _c(self, keep: list) -> ffi.CPaint:
        p = ffi.CPaint()
        p.kind = 0
        p.tr[:] = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
        p.solid[:] = [float(v) for v in self.c]
        return p


@dataclass
class GradStop:
    position: float
    color: Sequence[float]  # stored-space premultiplied f32x4


class _Grad:
    kind = 1

    def __init__(self, stops, units, linear_colors, spread, tr):
        self.stop_pos = np.ascontiguousarray([s.position if isinstance(s, GradStop) else s[0] for s in stops], dtype=np.float64)
        self.stop_colors = np.ascontiguousarray([s.color if isinstance(s, GradStop) else s[1] for s in stops],
                                                dtype=np.float32).reshape(-1, 4)
        self.units = Units(units)
        self.linear_colors = bool(linear_colors)
        self.spread = GradSpread(spread)
        self.tr = _as_tr(tr)

    def _base(self, keep: list) -> ffi.CPaint:
        p = ffi.CPaint()
        p.kind = self.kind
        p.units = int(self.units)
        p.linear_colors = int(self.linear_colors)
        p.spread = int(self.spread)
        p.tr[:] = [float(v) for v in self.tr]
        p.n_stops = len(self.stop_pos)
        p.stop_pos = self.stop_pos.ctypes.data_as(C.POINTER(C.c_double))
        p.stop_colors = self.stop_colors.ctypes.data_as(C.POINTER(C.c_float))
        keep.append(self)
        return p


class GradLinear(_Grad):
    """`GradLinear` (src/grad.rs:150-226).  `stops` carry the colours in the space the reference stores them
    (after `convert_to_srgb` when `linear_colors` is false)."""
    kind = 1

    def __init__(self, stops, units, linear_colors, spread, tr, start, end):
        super().__init__(stops, units, linear_colors, spread, tr)
        self.start = (float(start[0]), float(start[1]))
        self.end = (float(end[0]), float(end[1]))

    def _c(self, keep: list) -> ffi.CPaint:
        p = self._base(keep)
        p.p0[:] = self.start
        p.p1[:] = self.end
        return p


class GradRadial(_Grad):
    """`GradRadial` (src/grad.rs:307-426)."""
    kind = 2

    def __init__(self, stops, units, linear_colors, spread, tr, center, radius, fcenter=None, fradius=0.0):
        super().__init__(stops, units, linear_colors, spread, tr)
        self.center = (float(center[0]), float(center[1]))
        self.fcenter = self.center if fcenter is None else (float(fcenter[0]), float(fcenter[1]))
        self.radius, self.fradius = float(radius), float(fradius)

    def _c(self, keep: list) -> ffi.CPaint:
        p = self._base(keep)
        p.p0[:] = self.center
        p.p1[:] = self.fcenter
        p.r0, p.r1 = self.radius, self.fradius
        return p


def paint_from_desc(d: dict):
    """Paint from a flat description dict (kind, units, ..., stop_pos, stop_colors)."""
    kind = int(d["kind"])
    if kind == 0:
        return LinColor(*[float(v) for v in d["solid"]])
    stops = list(zip([float(v) for v in d["stop_pos"]], np.asarray(d["stop_colors"], dtype=np.float32).reshape(-1, 4)))
    if kind == 1:
        return GradLinear(stops, int(d["units"]), bool(d["linear_colors"]), int(d["spread"]), d["tr"], d["p0"], d["p1"])
    return GradRadial(stops, int(d["units"]), bool(d["linear_colors"]), int(d["spread"]), d["tr"], d["p0"], float(d["r0"]),
                      d["p1"], float(d["r1"]))


# ---- device objects -------------------------------------------------------------------------------------
class DevicePath:
    """Device-resident path (`rgpu_dpath`): an uploaded host path, or (`stroke=`) the outline of its stroke computed on the
    device (`Path::stroke`, src/path.rs:374-415), which never visits the host unless `download()` is called."""

    def __init__(self, rast: "GpuRasterizer", path, stroke: "StrokeStyle | None" = None):
        self.rast = rast
        self.path = path if stroke is None else None
        h = C.c_void_p()
        if not isinstance(path, Path):  # a device path (DevicePath or a raw rgpu_dpath handle, e.g. an element of a batch): stroke it in place
            st = stroke._c()
            src = path.h if isinstance(path, DevicePath) else C.c_void_p(int(path))
            rast._check(ffi.lib().rgpu_dpath_stroke(rast.ctx, src, C.byref(st), C.byref(h)))
            self.h = h
            return
        c = path._c()
        if stroke is None:
            rast._check(ffi.lib().rgpu_path_upload(rast.ctx, C.byref(c), C.byref(h)))
        else:
            st = stroke._c()
            rast._check(ffi.lib().rgpu_path_stroke(rast.ctx, C.byref(c), C.byref(st), C.byref(h)))
        self.h = h

    def counts(self) -> tuple[int, int, int]:
        """(n_points, n_segments, n_subpaths)"""
        n = [C.c_uint32() for _ in range(3)]
        self.rast._check(ffi.lib().rgpu_dpath_info(self.h, *[C.byref(v) for v in n]))
        return tuple(v.value for v in n)

    def download(self) -> Path:
        """The device path as a host `Path` (for a stroked path: what `Path::stroke` returns)."""
        n_pts, n_seg, n_sub = self.counts()
        pts = np.zeros((n_pts, 2), dtype=np.float64)
        kinds = np.zeros(n_seg, dtype=np.uint8)
        sp = np.zeros(n_sub + 1, dtype=np.uint32)
        closed = np.zeros(n_sub, dtype=np.uint8)
        self.rast._check(ffi.lib().rgpu_dpath_download(self.rast.ctx, self.h, pts.ctypes.data_as(C.POINTER(C.c_double)),
                                                       kinds.ctypes.data_as(C.POINTER(C.c_uint8)), sp.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                       closed.ctypes.data_as(C.POINTER(C.c_uint8))))
        return Path(pts, kinds, sp, closed)

    def free(self):
        if self.h:
            ffi.lib().rgpu_path_free(self.rast.ctx, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


#: numpy view of `rgpu_parse_info` (include/rasterize_b200.h)
PARSE_INFO_DTYPE = np.dtype([("bbox", "<f8", 4), ("fit_tr", "<f8", 6), ("fit_width", "<u4"), ("fit_height", "<u4"), ("n_points", "<u4"),
                             ("n_segments", "<u4"), ("n_subpaths", "<u4"), ("status", "<i4"), ("error_offset", "<u4"), ("has_bbox", "<i4"),
                             ("n_curves", "<u4"), ("reserved", "<u4")])


class Align(enum.IntEnum):
    """`Align` (src/geometry.rs:298-305)"""
    Min = 0
    Mid = 1
    Max = 2


class DevicePathBatch:
    """Device-resident `PathBatch` (`rgpu_dpath_batch`): element i is an ordinary device path.  Made from a host batch
    (upload) or, with `handle=`, wrapped around a batch the library built itself (`GpuRasterizer.parse_svg_batch`)."""

    def __init__(self, rast: "GpuRasterizer", batch: PathBatch | None, handle=None, n_paths: int | None = None):
        self.rast = rast
        self.batch = batch
        if handle is not None:
            self.h = handle
            self.n_paths = int(n_paths)
            return
        h = C.c_void_p()
        c = batch.flat._c()
        rast._check(ffi.lib().rgpu_path_upload_batch(rast.ctx, C.byref(c), batch.path_subpath_offsets.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                     len(batch), C.byref(h)))
        self.h = h
        self.n_paths = len(batch)

    def __len__(self) -> int:
        return self.n_paths

    def counts(self) -> tuple[int, int, int, int]:
        """(n_paths, n_points, n_segments, n_subpaths)"""
        n = C.c_size_t()
        c = [C.c_uint32() for _ in range(3)]
        self.rast._check(ffi.lib().rgpu_path_batch_info(self.h, C.byref(n), *[C.byref(v) for v in c]))
        return (int(n.value),) + tuple(v.value for v in c)

    def download(self) -> PathBatch:
        """The batch as a host `PathBatch` (for a parsed batch: what `str::parse::<Path>` gives, path by path)."""
        n, n_pts, n_seg, n_sub = self.counts()
        pts = np.zeros((n_pts, 2), dtype=np.float64)
        kinds = np.zeros(n_seg, dtype=np.uint8)
        sp = np.zeros(n_sub + 1, dtype=np.uint32)
        closed = np.zeros(n_sub, dtype=np.uint8)
        psp = np.zeros(n + 1, dtype=np.uint32)
        u32p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        self.rast._check(ffi.lib().rgpu_path_batch_download(self.rast.ctx, self.h, pts.ctypes.data_as(C.POINTER(C.c_double)), kinds.ctypes.data_as(u8p),
                                                            sp.ctypes.data_as(u32p), closed.ctypes.data_as(u8p), psp.ctypes.data_as(u32p)))
        return PathBatch(pts, kinds, sp, closed, psp)

    def handle(self, i: int) -> int:
        return int(ffi.lib().rgpu_path_batch_get(self.h, i) or 0)

    def handles(self) -> np.ndarray:
        """Device-path handles of all elements as u64 (they are elements of one array)."""
        n = self.n_paths
        if n == 0:
            return np.zeros(0, dtype=np.uint64)
        h0 = self.handle(0)
        stride = self.handle(1) - h0 if n > 1 else 0
        return (np.uint64(h0) + np.arange(n, dtype=np.uint64) * np.uint64(stride)).astype(np.uint64)

    def free(self):
        if self.h:
            ffi.lib().rgpu_path_batch_free(self.rast.ctx, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


#: numpy view of `rgpu_job` (include/rasterize_b200.h) for job tables built without a Python loop
JOB_DTYPE = np.dtype([("path", np.uint64), ("tr", np.float64, 6), ("fill_rule", np.int32), ("mode", np.int32), ("paint", np.uint64),
                      ("path_bbox", np.uint64), ("canvas", np.uint64), ("origin", np.uint64), ("row_stride", np.uint64),
                      ("width", np.uint32), ("height", np.uint32)], align=True)


class PreparedBatch:
    """`rgpu_batch`: a job table marshalled once and kept on the device."""

    def __init__(self, rast: "GpuRasterizer", table: np.ndarray, independent: bool, keep=None):
        assert table.dtype == JOB_DTYPE and table.flags.c_contiguous and JOB_DTYPE.itemsize == C.sizeof(ffi.CJob)
        self.rast, self.table, self.keep = rast, table, keep
        h = C.c_void_p()
        flags = ffi.BATCH_INDEPENDENT if independent else ffi.BATCH_ORDERED
        rast._check(ffi.lib().rgpu_batch_create(rast.ctx, C.c_void_p(table.ctypes.data), len(table), flags, C.byref(h)))
        self.h = h

    def render(self) -> None:
        """Asynchronous; `GpuRasterizer.batch_status()` waits and reports device-side errors."""
        self.rast._check(ffi.lib().rgpu_batch_render(self.rast.ctx, self.h))

    def free(self):
        if self.h:
            ffi.lib().rgpu_batch_free(self.rast.ctx, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class Job:
    """One entry of `GpuRasterizer.render_batch` (`rgpu_job`)."""
    path: DevicePath
    tr: object
    fill_rule: FillRule
    mode: int           # ffi.JOB_MASK / JOB_COVERAGE / JOB_FILL / JOB_RENDER
    canvas: int         # device pointer
    width: int
    height: int
    row_stride: int
    origin: int = 0
    paint: object = None
    path_bbox: object = None


class GpuRasterizer:
    """`impl Rasterizer` on a B200 (`GpuRasterizer` of the north star); one CUDA context/stream per instance."""

    def __init__(self, flatness: float = DEFAULT_FLATNESS, device: int = 0):
        self.ctx = None
        L = ffi.lib()
        h = C.c_void_p()
        rc = L.rgpu_create(device, float(flatness), C.byref(h))
        if rc != 0:
            raise RgpuError(rc, L.rgpu_last_error(None).decode())
        self.ctx = h
        self.flatness = float(flatness)
        self.device = device

    def close(self):
        if self.ctx:
            ffi.lib().rgpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise RgpuError(rc, ffi.lib().rgpu_last_error(self.ctx).decode())

    # -- Rasterizer trait ------------------------------------------------------------------------------
    def name(self) -> str:
        return ffi.lib().rgpu_name().decode()

    def flatten(self, path: Path, tr=Transform(), close: bool = True) -> np.ndarray:
        """`Path::flatten(tr, flatness, close)` -> [n,4] f64 lines (x0,y0,x1,y1) in the reference's order."""
        L = ffi.lib()
        c = path._c()
        t = _as_tr(tr)
        n = C.c_size_t()
        cap = max(64, path.segments_count() * 32)
        while True:
            out = np.empty((cap, 4), dtype=np.float64)
            rc = L.rgpu_flatten(self.ctx, C.byref(c), t.ctypes.data_as(C.POINTER(C.c_double)), int(close),
                                out.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(n))
            if rc == ffi.ERR_CAPACITY and n.value > cap:
                cap = n.value
                continue
            self._check(rc)
            return out[: n.value].copy()

    @staticmethod
    def _shape_of(img: np.ndarray, elem: int):
        """(base array pointer, CShape) of a 2-D (or [H,W,4]) possibly strided numpy view."""
        h, w = img.shape[0], img.shape[1]
        item = img.itemsize * elem
        rs, cs = img.strides[0], img.strides[1]
        if rs % item or cs % item or rs < 0 or cs < 0:
            raise ValueError("image strides must be non-negative multiples of the pixel size")
        if elem == 4 and (img.shape[2] != 4 or img.strides[2] != img.itemsize):
            raise ValueError("LinColor images must be [H,W,4] with contiguous channels")
        return CShapeOf(0, w, h, rs // item, cs // item)

    def mask(self, path: Path, tr, img: np.ndarray, fill_rule: FillRule) -> None:
        """`Rasterizer::mask`: img is an f64 (trait-faithful) or f32 (device-native) 2-D host image, zero on entry."""
        L = ffi.lib()
        c = path._c()
        t = _as_tr(tr)
        tp = t.ctypes.data_as(C.POINTER(C.c_double))
        if img.dtype == np.float64:
            shape = self._shape_of(img, 1)
            self._check(L.rgpu_mask(self.ctx, C.byref(c), tp, int(fill_rule), img.ctypes.data, shape))
        elif img.dtype == np.float32:
            if not img.flags.c_contiguous:
                raise ValueError("f32 masks must be dense row-major")
            self._check(L.rgpu_mask_f32(self.ctx, C.byref(c), tp, int(fill_rule), img.ctypes.data, img.shape[1], img.shape[0]))
        else:
            raise TypeError("mask image must be float64 or float32")

    PIXEL_DTYPE = np.dtype([("x", "<u8"), ("y", "<u8"), ("alpha", "<f8")])  # = rgpu_pixel

    def mask_iter_array(self, path: Path, tr, size: Size, fill_rule: FillRule, cap: int | None = None) -> np.ndarray:
        """`Rasterizer::mask_iter` as a structured array (x, y, alpha) in row-major order.  The list is compacted on the
        device and only the yielded pixels are downloaded; `cap` is the first guess of their number (the call is repeated
        with the exact count when it was too small)."""
        L = ffi.lib()
        c = path._c()
        t = _as_tr(tr)
        n = C.c_size_t()
        total = size.width * size.height
        cap = min(total, 1 << 20) if cap is None else min(total, max(0, int(cap)))
        while True:
            buf = np.empty(max(cap, 1), dtype=self.PIXEL_DTYPE)
            rc = L.rgpu_mask_iter(self.ctx, C.byref(c), t.ctypes.data_as(C.POINTER(C.c_double)), size.width, size.height, int(fill_rule),
                                  C.cast(buf.ctypes.data, C.POINTER(ffi.CPixel)), cap, C.byref(n))
            if rc == ffi.ERR_CAPACITY and n.value > cap:
                cap = n.value
                continue
            self._check(rc)
            return buf[:n.value]

    def mask_iter(self, path: Path, tr, size: Size, fill_rule: FillRule):
        """`Rasterizer::mask_iter`: list of (x, y, alpha) with abs(alpha) >= 1e-6, row-major order."""
        return [(int(x), int(y), float(a)) for x, y, a in self.mask_iter_array(path, tr, size, fill_rule).tolist()]

    def coverage(self, path: Path, tr, size: Size, fill_rule: FillRule) -> np.ndarray:
        """Dense form of mask_iter: f32 [H,W]."""
        L = ffi.lib()
        c = path._c()
        t = _as_tr(tr)
        out = np.zeros((size.height, size.width), dtype=np.float32)
        self._check(L.rgpu_coverage_f32(self.ctx, C.byref(c), t.ctypes.data_as(C.POINTER(C.c_double)), int(fill_rule),
                                        out.ctypes.data, size.width, size.height))
        return out

    def fill(self, path: Path, tr, fill_rule: FillRule, paint, img: np.ndarray, bbox=None) -> None:
        """Default `Rasterizer::fill`: blend `paint` over the f32 [H,W,4] LinColor host image in place.
        `bbox` = `path.bbox(identity)` (minx,miny,maxx,maxy) for bounding-box units."""
        if img.dtype != np.float32 or img.ndim != 3:
            raise TypeError("fill image must be float32 [H,W,4]")
        L = ffi.lib()
        c = path._c()
        t = _as_tr(tr)
        keep: list = []
        p = paint._c(keep)
        bb = None
        bbp = None
        if bbox is not None:
            bb = np.ascontiguousarray(bbox, dtype=np.float64)
            bbp = bb.ctypes.data_as(C.POINTER(C.c_double))
        shape = self._shape_of(img, 4)
        self._check(L.rgpu_fill(self.ctx, C.byref(c), t.ctypes.data_as(C.POINTER(C.c_double)), int(fill_rule), C.byref(p), bbp,
                                img.ctypes.data, shape))

    # -- device-resident API ---------------------------------------------------------------------------
    def upload(self, path: Path) -> DevicePath:
        return DevicePath(self, path)

    def stroke(self, path, style: StrokeStyle) -> DevicePath:
        """`Path::stroke` (src/path.rs:374-415) on the device: the outline as a device-resident path, usable wherever an
        uploaded path is (`Job.path`); `.download()` gives the host `Path` the reference returns.  `path` = a host `Path`, or a
        path that is already on the device (`DevicePath`, or the handle of a batch element: `DevicePathBatch.handle(i)`)."""
        return DevicePath(self, path, stroke=style)

    def _cjobs(self, jobs: Iterable[Job]):
        jobs = list(jobs)
        arr = (ffi.CJob * max(len(jobs), 1))()
        keep: list = []
        for i, j in enumerate(jobs):
            cj = arr[i]
            cj.path = j.path.h
            cj.tr[:] = [float(v) for v in _as_tr(j.tr)]
            cj.fill_rule = int(j.fill_rule)
            cj.mode = int(j.mode)
            if j.paint is not None:
                cp = j.paint._c(keep)
                keep.append(cp)
                cj.paint = C.pointer(cp)
            if j.path_bbox is not None:
                bb = np.ascontiguousarray(j.path_bbox, dtype=np.float64)
                keep.append(bb)
                cj.path_bbox = bb.ctypes.data_as(C.POINTER(C.c_double))
            cj.canvas = int(j.canvas)
            cj.origin = int(j.origin)
            cj.row_stride = int(j.row_stride)
            cj.width = int(j.width)
            cj.height = int(j.height)
        return arr, len(jobs), keep

    def prepare_batch(self, jobs: Iterable[Job]):
        """Marshal a job list once; the result can be submitted repeatedly with `submit_prepared`."""
        return self._cjobs(jobs)

    def submit_prepared(self, prepared, independent: bool = False, sync: bool = True) -> None:
        arr, n, _ = prepared
        flags = ffi.BATCH_INDEPENDENT if independent else ffi.BATCH_ORDERED
        fn = ffi.lib().rgpu_render_batch_sync if sync else ffi.lib().rgpu_render_batch
        self._check(fn(self.ctx, arr, n, flags))

    def render_batch(self, jobs: Iterable[Job], independent: bool = False, sync: bool = True) -> None:
        self.submit_prepared(self._cjobs(jobs), independent, sync)

    def submit_scene_prepared(self, prepared, layer_ptr: int, width: int, height: int, fresh: bool = True, bg=None, rgba_ptr: int = 0,
                              sync: bool = True) -> None:
        """All FILL jobs of one dense device layer in one raster launch (`rgpu_render_scene`): the Fill arm of
        `Pipeline::render_rec` (src/scene.rs:397-435) fused with `Layer::new` and, with `rgba_ptr`, the RGBA8 export."""
        arr, n, _ = prepared
        cbg = None
        if fresh and bg is not None:
            cbg = (C.c_float * 4)(*[float(v) for v in bg])
        fn = ffi.lib().rgpu_render_scene_sync if sync else ffi.lib().rgpu_render_scene
        self._check(fn(self.ctx, arr, n, layer_ptr, int(width), int(height), 1 if fresh else 0, cbg, rgba_ptr or None))

    def render_scene(self, jobs: Iterable[Job], layer_ptr: int, width: int, height: int, fresh: bool = True, bg=None, rgba_ptr: int = 0,
                     sync: bool = True) -> None:
        self.submit_scene_prepared(self._cjobs(jobs), layer_ptr, width, height, fresh, bg, rgba_ptr, sync)

    def prepare_scene_host(self, fills):
        """Marshal host-side scene fills once: `fills` = iterable of (Path, tr, FillRule, paint, path_bbox | None, x, y, width, height)."""
        fills = list(fills)
        arr = (ffi.CSceneFill * max(len(fills), 1))()
        keep: list = []
        for i, (path, tr, rule, paint, bbox, x, y, w, h) in enumerate(fills):
            cf = arr[i]
            cp = path._c()
            keep.append((cp, path))
            cf.path = C.pointer(cp)
            cf.tr[:] = [float(v) for v in _as_tr(tr)]
            cf.fill_rule = int(rule)
            pp = paint._c(keep)
            keep.append(pp)
            cf.paint = C.pointer(pp)
            if bbox is not None:
                bb = np.ascontiguousarray(bbox, dtype=np.float64)
                keep.append(bb)
                cf.path_bbox = bb.ctypes.data_as(C.POINTER(C.c_double))
            cf.x, cf.y, cf.width, cf.height = int(x), int(y), int(w), int(h)
        return arr, len(fills), keep

    def render_scene_host(self, fills, width: int, height: int, bg=None, lin_out: np.ndarray | None = None, rgba_out: np.ndarray | None = None):
        """`Scene::render` of a Fill-only pipeline + export, host buffers in and out (`rgpu_render_scene_host`).  `fills` is an
        iterable as for `prepare_scene_host`, or its result.  Returns (lin_out, rgba_out)."""
        arr, n, _ = fills if isinstance(fills, tuple) and len(fills) == 3 and isinstance(fills[1], int) else self.prepare_scene_host(fills)
        cbg = (C.c_float * 4)(*[float(v) for v in bg]) if bg is not None else None
        if lin_out is None and rgba_out is None:
            rgba_out = np.empty((height, width, 4), dtype=np.uint8)
        for a, dt in ((lin_out, np.float32), (rgba_out, np.uint8)):
            assert a is None or (a.dtype == dt and a.flags.c_contiguous and a.shape == (height, width, 4))
        self._check(ffi.lib().rgpu_render_scene_host(self.ctx, arr, n, int(width), int(height), cbg,
                                                     lin_out.ctypes.data if lin_out is not None else None,
                                                     rgba_out.ctypes.data if rgba_out is not None else None))
        return lin_out, rgba_out

    def batch_status(self) -> None:
        self._check(ffi.lib().rgpu_batch_status(self.ctx))

    def set_winding_bits(self, integer_bits: int) -> None:
        """8 = Q7.24 winding cells (default), 14 = Q13.18 (`rgpu_set_winding_bits`)."""
        self._check(ffi.lib().rgpu_set_winding_bits(self.ctx, integer_bits))

    # -- batches of independent paths (BASELINE config 4) / band-sharded masks (config 5) ---------------------------
    def upload_batch(self, batch: PathBatch) -> DevicePathBatch:
        return DevicePathBatch(self, batch)

    def parse_svg_batch(self, strings, fit: tuple[int, int, Align] | None = None, strict: bool = False):
        """Batch `str::parse::<Path>` + `Path::bbox` (+ `fit_size`) on the device (src/svg.rs:241-421,
        src/path.rs:428-451, src/geometry.rs:490-516).  `strings` = a sequence of str / bytes, or (text, offsets u32[n + 1]) with
        `text` = bytes or a uint8 array (e.g. pinned, from `host_alloc`).
        -> (DevicePathBatch, info: ndarray of PARSE_INFO_DTYPE).  A string that does not parse gives an empty path and its
        `status` / `error_offset`; with `strict` the call raises instead."""
        if isinstance(strings, tuple):
            text, off = strings
            off = np.ascontiguousarray(off, dtype=np.uint32)
        else:
            enc = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
            off = np.zeros(len(enc) + 1, dtype=np.uint32)
            off[1:] = np.cumsum([len(e) for e in enc])
            text = b"".join(enc)
        n = len(off) - 1
        info = np.zeros(max(n, 1), dtype=PARSE_INFO_DTYPE)
        opt = ffi.CParseOptions(int(fit[0]), int(fit[1]), int(fit[2])) if fit is not None else ffi.CParseOptions(0, 0, -1)
        h = C.c_void_p()
        if isinstance(text, np.ndarray):
            assert text.dtype == np.uint8 and text.flags.c_contiguous
            text = C.cast(text.ctypes.data, C.c_char_p)
        self._check(ffi.lib().rgpu_parse_svg_batch(self.ctx, text, off.ctypes.data_as(C.POINTER(C.c_uint32)), n, C.byref(opt), C.byref(h),
                                                   None if strict else info.ctypes.data))
        batch = DevicePathBatch(self, None, handle=h, n_paths=n)
        return batch, (None if strict else info[:n])

    def prepare_job_table(self, table: np.ndarray, independent: bool = True, keep=None) -> PreparedBatch:
        return PreparedBatch(self, table, independent, keep)

    def fill_batch_host(self, batch: PathBatch, fill_rule: FillRule, paint, width: int, height: int, out: np.ndarray, trs=None) -> np.ndarray:
        """`ImageOwned::new_default` + `Path::fill` for every path of the batch, host buffers in and out
        (`rgpu_fill_batch_host`).  `out` selects the format: f32 [n,H,W,4] LinColor, u8 [n,H,W,4] RGBA8, f32 [n,H,W] coverage."""
        return _fill_batch_host(ffi.lib().rgpu_fill_batch_host, self.ctx, self._check, batch, fill_rule, paint, width, height, out, trs)

    def mask_banded(self, path: Path, tr, img: np.ndarray, fill_rule: FillRule, n_bands: int = 8, band_first: int = 0, band_count: int | None = None) -> None:
        """`Rasterizer::mask` of bands [band_first, band_first + band_count) of n_bands scanline bands (`rgpu_mask_banded_host`);
        img is the dense f32 or f64 [H, W] image of the whole canvas."""
        if img.dtype not in (np.float32, np.float64) or not img.flags.c_contiguous or img.ndim != 2:
            raise TypeError("banded mask image must be dense float32 / float64 [H, W]")
        c = path._c()
        t = _as_tr(tr)
        self._check(ffi.lib().rgpu_mask_banded_host(self.ctx, C.byref(c), t.ctypes.data_as(C.POINTER(C.c_double)), int(fill_rule), img.ctypes.data,
                                                    img.itemsize, img.shape[1], img.shape[0], n_bands, band_first, n_bands if band_count is None else band_count))

    def last_counts(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        ffi.lib().rgpu_last_counts(self.ctx, C.byref(a), C.byref(b), C.byref(c))
        return dict(lines=a.value, line_refs=b.value, launches=c.value)

    def last_transfer_bytes(self):
        """(h2d, d2h) bytes the last `mask` call moved over PCIe."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(ffi.lib().rgpu_last_transfer_bytes(self.ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_profiling(self, on: bool) -> None:
        self._check(ffi.lib().rgpu_set_profiling(self.ctx, int(on)))

    def last_stage_ms(self):
        """(flatten_ms, bin_ms, raster_ms) of the last profiled batch, from CUDA events on the context's stream."""
        out = (C.c_float * 3)()
        self._check(ffi.lib().rgpu_last_stage_ms(self.ctx, out))
        return float(out[0]), float(out[1]), float(out[2])

    def stream(self) -> int:
        return int(ffi.lib().rgpu_stream(self.ctx) or 0)

    def sync(self) -> None:
        self._check(ffi.lib().rgpu_sync(self.ctx))

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(ffi.lib().rgpu_device_alloc(self.ctx, nbytes, C.byref(p)))
        return int(p.value)

    def device_free(self, ptr: int) -> None:
        self._check(ffi.lib().rgpu_device_free(self.ctx, C.c_void_p(ptr)))

    def device_zero(self, ptr: int, nbytes: int) -> None:
        self._check(ffi.lib().rgpu_device_zero(self.ctx, C.c_void_p(ptr), nbytes))

    def to_host(self, ptr: int, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        self._check(ffi.lib().rgpu_memcpy_d2h(self.ctx, out.ctypes.data, C.c_void_p(ptr), out.nbytes))
        return out

    def to_device(self, ptr: int, arr: np.ndarray) -> None:
        a = np.ascontiguousarray(arr)
        self._check(ffi.lib().rgpu_memcpy_h2d(self.ctx, C.c_void_p(ptr), a.ctypes.data, a.nbytes))

    def to_rgba8(self, lin_ptr: int, rgba_ptr: int, n_pixels: int) -> None:
        self._check(ffi.lib().rgpu_to_rgba8_dev(self.ctx, C.c_void_p(lin_ptr), C.c_void_p(rgba_ptr), n_pixels))

    def layer_scale_by_mask(self, lin_ptr, lin_origin, lin_stride, mask_ptr, mask_origin, mask_stride, width, height) -> None:
        """`child_layer.compose(mask_layer, |dst, src| dst * src)` on the intersection rectangle (src/scene.rs:453-455)."""
        self._check(ffi.lib().rgpu_layer_scale_by_mask_dev(self.ctx, C.c_void_p(lin_ptr), lin_origin, lin_stride, C.c_void_p(mask_ptr),
                                                            mask_origin, mask_stride, width, height))

    def layer_blend_over(self, dst_ptr, dst_origin, dst_stride, src_ptr, src_origin, src_stride, width, height, opacity=None) -> None:
        """`layer.compose(child_layer, |dst, src| dst.blend_over(src [* opacity]))` (src/scene.rs:436-457)."""
        self._check(ffi.lib().rgpu_layer_blend_over_dev(self.ctx, C.c_void_p(dst_ptr), dst_origin, dst_stride, C.c_void_p(src_ptr), src_origin,
                                                         src_stride, width, height, int(opacity is not None),
                                                         float(opacity if opacity is not None else 1.0)))

    def download_rgba8(self, lin_ptr: int, shape) -> np.ndarray:
        """LinColor device image -> RGBA8 host image [H, W, 4] u8 (conversion on device, 4 B/pixel over PCIe)."""
        h, w = shape
        out = np.empty((h, w, 4), dtype=np.uint8)
        self._check(ffi.lib().rgpu_download_rgba8(self.ctx, C.c_void_p(lin_ptr), h * w, out.ctypes.data))
        return out

    def fill_color(self, lin_ptr: int, n_pixels: int, color) -> None:
        c = np.asarray(color, dtype=np.float32)
        self._check(ffi.lib().rgpu_fill_color_dev(self.ctx, C.c_void_p(lin_ptr), n_pixels, c.ctypes.data_as(C.POINTER(C.c_float))))

    def host_alloc(self, shape, dtype) -> np.ndarray:
        """Pinned host array (freed with the process; small helper for benchmarks)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._check(ffi.lib().rgpu_host_alloc(self.ctx, n, C.byref(p)))
        buf = (C.c_byte * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _fill_batch_host(fn, handle, check, batch: PathBatch, fill_rule, paint, width, height, out: np.ndarray, trs):
    n = len(batch)
    if out.dtype == np.float32 and out.shape == (n, height, width, 4):
        fmt = ffi.OUT_LINCOLOR
    elif out.dtype == np.uint8 and out.shape == (n, height, width, 4):
        fmt = ffi.OUT_RGBA8
    elif out.dtype == np.float32 and out.shape == (n, height, width):
        fmt = ffi.OUT_COVERAGE
    else:
        raise TypeError("out must be f32 [n,H,W,4], u8 [n,H,W,4] or f32 [n,H,W]")
    if not out.flags.c_contiguous:
        raise ValueError("out must be dense")
    keep: list = []
    cp = paint._c(keep) if paint is not None else None
    c = batch.flat._c()
    t = None
    if trs is not None:
        t = np.ascontiguousarray(trs, dtype=np.float64).reshape(n, 6)
    check(fn(handle, C.byref(c), batch.path_subpath_offsets.ctypes.data_as(C.POINTER(C.c_uint32)), n,
             t.ctypes.data_as(C.POINTER(C.c_double)) if t is not None else None, int(fill_rule), C.byref(cp) if cp is not None else None,
             width, height, fmt, out.ctypes.data))
    return out


class MultiGpuRasterizer:
    """`rgpu_multi`: one context + worker thread per device of the box; batches shard by path, huge canvases by
    scanline bands, no collective."""

    def __init__(self, devices=None, flatness: float = DEFAULT_FLATNESS):
        L = ffi.lib()
        self.h = None
        if devices is None:
            devices = list(range(L.rgpu_device_count()))
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = L.rgpu_multi_create(arr, len(devices), float(flatness), C.byref(h))
        if rc != 0:
            raise RgpuError(rc, L.rgpu_multi_last_error(None).decode())
        self.h = h
        self.devices = list(devices)

    def close(self):
        if self.h:
            ffi.lib().rgpu_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise RgpuError(rc, ffi.lib().rgpu_multi_last_error(self.h).decode())

    def fill_batch_host(self, batch: PathBatch, fill_rule: FillRule, paint, width: int, height: int, out: np.ndarray, trs=None) -> np.ndarray:
        return _fill_batch_host(ffi.lib().rgpu_multi_fill_batch_host, self.h, self._check, batch, fill_rule, paint, width, height, out, trs)

    def mask_banded(self, path: Path, tr, img: np.ndarray, fill_rule: FillRule, n_bands: int = 0) -> None:
        if img.dtype not in (np.float32, np.float64) or not img.flags.c_contiguous or img.ndim != 2:
            raise TypeError("banded mask image must be dense float32 / float64 [H, W]")
        c = path._c()
        t = _as_tr(tr)
        self._check(ffi.lib().rgpu_multi_mask_banded_host(self.h, C.byref(c), t.ctypes.data_as(C.POINTER(C.c_double)), int(fill_rule),
                                                          img.ctypes.data, img.itemsize, img.shape[1], img.shape[0], n_bands))


def CShapeOf(start, width, height, row_stride, col_stride) -> ffi.CShape:
    return ffi.CShape(start, width, height, row_stride, col_stride)
