"""Timing breakdown of the host-buffer mask call (RGPU_E2E_TRACE=1 python tools/e2e_trace.py)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import rasterize_b200 as rb
import assets
r = rb.GpuRasterizer()
p = assets.load_path("material")
c2 = assets.expected()["paths"]["material"]["c2"]
w, h = c2["size"]; tr = np.array(c2["tr"])
img = r.host_alloc((h, w), np.float64)
for i in range(24):
    t0 = time.perf_counter(); r.mask(p, tr, img, rb.FillRule.NonZero); dt = time.perf_counter() - t0
    print(f"call {i}: {dt * 1e3:.3f} ms", file=sys.stderr)
