"""Config 5 at full size: run-coded download against dense copies, and one whole-canvas job against two blocks of bands (where the
band-local translate shows up: 6 pixels of 1.07e9 differ by one f32 ulp).  python tools/diag_c5.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import assets
import rasterize_b200 as rb
rast = rb.GpuRasterizer()
p = assets.load_path("tv_stroked")
c5 = assets.expected()["paths"]["tv_stroked"]["c5"]
w, h = c5["size"]; tr = np.array(c5["tr"])
def run(env, parts):
    if env: os.environ["RGPU_E2E_RUNCODE"] = "0"
    img = np.empty((h, w), dtype=np.float32); img[:] = -1.0
    for (a, n) in parts: rast.mask_banded(p, tr, img, rb.FillRule.NonZero, n_bands=64, band_first=a, band_count=n)
    if env: del os.environ["RGPU_E2E_RUNCODE"]
    return img
A = run(False, [(0, 64)])
B = run(True, [(0, 64)])
print("runcoded vs dense, whole:", np.array_equal(A.view(np.uint32), B.view(np.uint32)))
C = run(True, [(0, 40), (40, 24)])
d = np.nonzero(B.view(np.uint32) != C.view(np.uint32))
print("dense whole vs dense 40+24:", len(d[0]), "pixels differ", (d[0][:5], d[1][:5]) if len(d[0]) else "")
if len(d[0]): print("max abs diff", np.abs(B[d] - C[d]).max(), "rows", d[0].min(), d[0].max())
D = run(False, [(0, 40), (40, 24)])
print("runcoded 40+24 vs dense 40+24:", np.array_equal(C.view(np.uint32), D.view(np.uint32)))
