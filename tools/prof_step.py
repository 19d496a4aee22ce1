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
jobs, independent, info = bench.build_workload(workload, rb, rast, 0, 1, torch)
prepared = rast.prepare_batch(jobs)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
scn = info.get("scene")


def step():
    if scn:  # c1 / c3: the scene compositor
        rast.submit_scene_prepared(prepared, scn["layer"], scn["W"], scn["H"], fresh=True, bg=scn["bg"], rgba_ptr=scn["rgba"], sync=True)
        return
    if info.get("pre_step"):
        info["pre_step"]()
    rast.submit_prepared(prepared, independent=independent, sync=True)
    if info.get("post_step"):
        info["post_step"]()


step()
for _ in range(steps):
    flush.zero_()
    torch.cuda.synchronize()
    step()
print("ok", rast.last_counts())
