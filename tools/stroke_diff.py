"""Where does the device stroke differ from the oracle's?  (diagnostic)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rasterize_b200 as rb
import assets
from rasterize_b200 import LineCap, LineJoin, StrokeStyle
from stroke_common import STYLES, oracle_stroke

JOIN = {"miter": LineJoin.Miter, "bevel": LineJoin.Bevel, "round": LineJoin.Round}
CAP = {"butt": LineCap.Butt, "square": LineCap.Square, "round": LineCap.Round}
rast = rb.GpuRasterizer()
for name in sys.argv[1:] or ["huyak"]:
    p = assets.load_path(name)
    for width, join, ml, cap in STYLES:
        q = rast.stroke(p, StrokeStyle(width, JOIN[join], ml, CAP[cap])).download()
        wp, wk, ws, wc = oracle_stroke(p.points, p.kinds, p.subpath_offsets, p.closed, width, join, ml, cap)
        gp = np.asarray(q.points).reshape(-1, 2)
        same_struct = np.array_equal(q.kinds, wk) and np.array_equal(q.subpath_offsets, ws)
        bad = np.nonzero((gp.view(np.uint64) != wp.view(np.uint64)).any(axis=1))[0] if same_struct else []
        print(f"{name} {width} {join} {cap}: structure {'same' if same_struct else 'DIFFERS'}, {len(bad)} of {len(gp)} points differ")
        if same_struct and len(bad):
            pt_off = np.concatenate([[0], np.cumsum(wk)])
            for b in bad[:12]:
                seg = int(np.searchsorted(pt_off, b, side="right") - 1)
                print(f"   point {b} = control {b - pt_off[seg]} of segment {seg} (kind {wk[seg]}; neighbours {wk[max(seg - 2, 0):seg + 3].tolist()}) "
                      f"got {gp[b].tolist()} want {wp[b].tolist()}")
