"""Split download of rgpu_fill_batch_host (LinColor, plain solid paint): call time against the share of the images that cross
PCIe as coverage and are expanded on the host.  RGPU_E2E_TRACE=1 prints the library's own breakdown per call.
    python tools/e2e_split.py [n_glyphs]            (RGPU_E2E_EXPAND_FRAC=<share> fixes the share; unset = adaptive)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rasterize_b200 as rb
from rasterize_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
r = rb.GpuRasterizer()
pb = synth.glyph_batch(1, n)
out = r.host_alloc((n, 64, 64, 4), np.float32)
black = rb.LinColor(0.0, 0.0, 0.0, 1.0)
ts = []
for i in range(calls):
    t0 = time.perf_counter()
    r.fill_batch_host(pb, rb.FillRule.NonZero, black, 64, 64, out)
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"share={os.environ.get('RGPU_E2E_EXPAND_FRAC', 'adaptive')} expand={os.environ.get('RGPU_E2E_EXPAND', '1')} threads={os.environ.get('RGPU_HOST_THREADS', 'all')}"
      f" n={n}: " + " ".join(f"{t:.1f}" for t in ts) + f"  min {min(ts):.1f} ms")
