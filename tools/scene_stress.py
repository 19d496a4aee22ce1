"""Randomised stress of the scene compositor against the ordered batch (bit-identity), many layer shapes / fill counts.

    python tools/scene_stress.py [n_cases] [seed]        # on a B200; prints one line per case and a summary
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import rasterize_b200 as rb  # noqa: E402
from test_gpu_scene_kernel import both_ways, synthetic_jobs  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
rast = rb.GpuRasterizer()
bad = 0
for case in range(n_cases):
    W = int(rng.choice([1, 7, 63, 64, 255, 256, 257, 511, 512, 513, 700, 1024, 1500, 2048, 2100]))
    H = int(rng.choice([1, 3, 7, 8, 9, 15, 16, 17, 40, 100]))
    n_jobs = int(rng.choice([1, 2, 5, 17, 60]))
    max_w = int(rng.integers(1, W + 1))
    max_h = int(rng.integers(1, H + 1))
    make = synthetic_jobs(rast, W, H, n_jobs, seed=int(rng.integers(1, 1 << 20)), max_w=max_w, max_h=max_h)
    fresh = bool(rng.integers(0, 2))
    init = None
    bg = [0.1, 0.2, 0.3, 0.5] if rng.integers(0, 2) else None
    if not fresh:
        init = rng.random((H, W, 4), dtype=np.float32)
        init[..., :3] *= init[..., 3:4]
    lin_a, rgba_a, lin_b, rgba_b = both_ways(rast, make, W, H, bg=bg, initial=init)
    ok = np.array_equal(lin_a, lin_b) and np.array_equal(rgba_a, rgba_b)
    bad += not ok
    print(f"case {case:3d}: layer {W}x{H}, {n_jobs} fills (windows <= {max_w}x{max_h}), fresh={fresh}, bg={'yes' if bg else 'no'}: "
          f"{'identical' if ok else 'MISMATCH max |d| = %g' % np.abs(lin_a - lin_b).max()}")
print(f"scene stress: {n_cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
