"""Wall-clock of rgpu_parse_svg_batch (host text in, device-resident path batch + bbox + fit transforms out) beside the
oracle's parser + bbox + fit_size on one host thread."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rasterize_b200 as rb
import assets
from rasterize_b200 import Align, synth
from parse_common import pack, svg_of
import oracle as O

rast = rb.GpuRasterizer()
rast.set_profiling(True)
for n in (1000, 20000, 100000):
    pb = synth.glyph_batch(1, n)
    strings = [svg_of(pb.path(i)) for i in range(min(n, 2000))]
    strings = (strings * (n // len(strings) + 1))[:n]  # the text of 2 000 glyphs repeated: the parser's work is the same
    text, off = pack(strings)
    for _ in range(2):
        rast.parse_svg_batch((text, off), fit=(64, 64, Align.Mid))[0].free()
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        dpb, info = rast.parse_svg_batch((text, off), fit=(64, 64, Align.Mid))
        ts.append(time.perf_counter() - t0)
        st = rast.last_stage_ms()
        dpb.free()
    m = min(400, n)
    t0 = time.perf_counter()
    for s in strings[:m]:
        op = O.OraclePath.parse(s)
        O.fit_size(op.bbox(), 64, 64, 1)
    tc = (time.perf_counter() - t0) / m
    print(f"glyph batch n={n:7d}  text {len(text) / 1e6:7.2f} MB  segments {int(info['n_segments'].sum()):9d}  device call {np.median(ts) * 1e3:8.3f} ms "
          f"(min {min(ts) * 1e3:.3f}) = {len(text) / np.median(ts) / 1e9:.2f} GB/s of text, {n / np.median(ts) / 1e6:.2f} Mpaths/s;  "
          f"kernels: count+scan {st[0]:.3f} ms, emit {st[2]:.3f} ms;  oracle {tc * 1e6:.1f} us per path = {tc * n * 1e3:.1f} ms for the batch on one thread")
s = svg_of(assets.load_path("material"))
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    dpb, info = rast.parse_svg_batch([s])
    ts.append(time.perf_counter() - t0)
    dpb.free()
t0 = time.perf_counter()
O.OraclePath.parse(s).bbox()
tc = time.perf_counter() - t0
print(f"material as ONE string ({len(s) / 1e6:.2f} MB, {int(info['n_segments'][0])} segments): device call {np.median(ts) * 1e3:.2f} ms (cut at its "
      f"absolute movetos), oracle {tc * 1e3:.2f} ms")
