"""A/B timing of the host-buffer mask call: pipelined f32 D2H + threaded widening (default) vs device widening +
f64 D2H (RGPU_MASK_DEVICE_WIDEN=1), plus the f32 entry point and a raw pinned D2H of the same bytes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rasterize_b200 as rb
from rasterize_b200 import assets

r = rb.GpuRasterizer()
p = assets.load_path("material")
c2 = assets.expected()["paths"]["material"]["c2"]
w, h = c2["size"]; tr = np.array(c2["tr"])
img64 = r.host_alloc((h, w), np.float64)
img32 = r.host_alloc((h, w), np.float32)
pageable = np.zeros((h, w))
def t(fn, n=10):
    fn(); fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return f"min {min(ts)*1e3:.3f} ms  med {sorted(ts)[len(ts)//2]*1e3:.3f} ms"
print("mode:", "device-widen f64 D2H" if os.environ.get("RGPU_MASK_DEVICE_WIDEN") else "pipelined f32 D2H + host widen", "cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
print("rgpu_mask f64 pinned  :", t(lambda: r.mask(p, tr, img64, rb.FillRule.NonZero)))
print("rgpu_mask f64 pageable:", t(lambda: r.mask(p, tr, pageable, rb.FillRule.NonZero)))
print("rgpu_mask_f32 pinned  :", t(lambda: r.mask(p, tr, img32, rb.FillRule.NonZero)))
d = torch.empty(h * w, dtype=torch.float64, device="cuda")
hp = torch.empty(h * w, dtype=torch.float64).pin_memory()
def raw():
    hp.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("raw D2H 134 MB pinned :", t(raw))
a = np.zeros((h, w), dtype=np.float32)
def widen1():
    pageable[:] = a
print("numpy widen 1 thread  :", t(widen1, 5))
