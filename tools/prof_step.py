"""Minimal driver for ncu captures: builds a bench workload and runs a few steps of the hot path (no e2e / CPU legs).

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/x \
        python tools/prof_step.py c2 6
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import rasterize_b200 as rb  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.cuda.set_device(0)
rast = rb.GpuRasterizer(device=0)
step_fn, info = bench.build_workload(workload, rb, rast, 0, 1, torch)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def step():
    step_fn(sync=True)


step()
for _ in range(steps):
    flush.zero_()
    torch.cuda.synchronize()
    step()
print("ok", rast.last_counts())
