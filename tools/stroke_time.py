"""Wall-clock of rgpu_path_stroke (host path in, device-resident outline out) beside the oracle's `Path::stroke` on the host."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rasterize_b200 as rb
import assets
from rasterize_b200 import LineCap, LineJoin, StrokeStyle
import oracle as O

rast = rb.GpuRasterizer()
rast.set_profiling(True)
for name in ("tv", "squirrel", "material"):
    p = assets.load_path(name)
    for style, label in ((StrokeStyle(0.5, LineJoin.Round, 4.0, LineCap.Round), "round/round"), (StrokeStyle(1.0), "miter/butt")):
        for _ in range(3):
            rast.stroke(p, style).free()
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            dp = rast.stroke(p, style)
            ts.append(time.perf_counter() - t0)
            st = rast.last_stage_ms()
            n = dp.counts()
            dp.free()
        op = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)
        j = {LineJoin.Round: "round", LineJoin.Miter: "miter"}[style.line_join]
        c = {LineCap.Round: "round", LineCap.Butt: "butt"}[style.line_cap]
        tc = []
        for _ in range(5):
            t0 = time.perf_counter()
            op.stroke(style.width, j, style.miter_limit, c)
            tc.append(time.perf_counter() - t0)
        print(f"{name:9s} {label:12s} segments {p.segments_count():7d} -> {n[1]:8d}  device call {np.median(ts) * 1e3:8.3f} ms (min {min(ts) * 1e3:.3f})"
              f"  kernels: pieces+count+scan {st[0]:.3f} ms, emit {st[2]:.3f} ms;  oracle {np.median(tc) * 1e3:8.3f} ms")
