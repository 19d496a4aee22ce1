#!/usr/bin/env python
"""Benchmark of the fill hot path (BASELINE.json: Mpix/s, flattened lines/s, nonzero fill) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]

A step is one pass of the hot path (flatten -> bin -> signed-difference raster) over one batch:
  c2 (default, BASELINE configs[1]): data/material.path fitted to 4096x4096, `Rasterizer::mask`, non-zero.
  c4: a batch of synthetic random-cubic glyphs at 64x64 (SURVEY §8d generator), solid fill onto a fresh LinColor canvas per glyph
      (RB_C4_MASK=1: mask per glyph).
  c5: tv.path stroked on the 32768 x 32768 canvas as 8 scanline bands; rank r renders bands r, r + N, ... (band sharding,
      SURVEY §8e; fixed total work => strong scaling).
  c1 / c3: Scene::render of the squirrel CLI scene (512 px) / firefox.scene (2048 x 2048) on a device-resident layer + RGBA8.
N > 1 (under torchrun): every rank runs the same per-GPU workload on its own device with no data-path collective
(independent paths of a batch / bands of a canvas) => weak scaling; value = units of all ranks / max-over-ranks time.

Timing: CUDA events on the rasterizer's own stream around every step, L2 flushed (256 MiB memset) between steps
outside the event pairs; torch is used for device buffers, events, the barrier and the max-over-ranks reduction only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload: str):
    """dram bytes per raster launch from the committed ncu --set full capture, if one exists for this workload."""
    p = os.path.join(ROOT, "profiles", "raster_traffic.json")
    try:
        return float(json.load(open(p))[workload]["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- workloads -------------------------------------------------------------------------------------------
def glyph_path(rb, seed: int):
    """Synthetic glyph of SURVEY §8d (LCG of the reference's benches/scene_bench.rs:53-88): 3 closed contours of
    6 cubics, every coordinate uniform()*56+4."""
    state = seed & 0xffffffff

    def step():
        nonlocal state
        state = (state * 214013 + 2531011) & 0x7fffffff
        return state >> 16

    def u32():
        hi = step() & 0xffff
        lo = step() & 0xffff
        return (hi << 16) | lo

    def uniform():
        hi = u32()
        lo = u32()
        return (((hi << 32) | lo) >> 10) * 2.0 ** -53

    def pt():
        x = uniform() * 56.0 + 4.0
        y = uniform() * 56.0 + 4.0
        return (x, y)

    b = rb.Path.builder()
    for _ in range(3):
        b.move_to(pt())
        for _ in range(6):
            q1 = pt(); q2 = pt(); q3 = pt()
            b.cubic_to(q1, q2, q3)
        b.close()
    return b.build()


def build_workload(name: str, rb, rast, rank: int, world: int, torch):
    from rasterize_b200 import assets, ffi, sharding
    ex = assets.expected()["paths"]
    dev = torch.device("cuda", torch.cuda.current_device())
    if name == "c2":
        path = assets.load_path("material")
        c2 = ex["material"]["c2"]
        w, h = c2["size"]
        tr = np.array(c2["tr"])
        canvases = [torch.empty((h, w), dtype=torch.float32, device=dev)]
        dp = rast.upload(path)
        jobs = [rb.Job(dp, tr, rb.FillRule.NonZero, ffi.JOB_MASK, canvases[0].data_ptr(), w, h, w)]
        info = dict(workload="c2: data/material.path (21106 segments) fitted to 4096x4096, Rasterizer::mask, nonzero",
                    canvas=[w, h], items_per_gpu=1, pixels_per_step=w * h, in_bytes=path.input_bytes(), out_bytes=4 * w * h,
                    host_path=path, host_tr=tr, host_size=(w, h), keep=[dp, canvases])
        return jobs, True, info
    if name == "c4":
        n = int(os.environ.get("RB_GLYPHS", "20000"))
        first = rank * n
        paths = [glyph_path(rb, first + i + 1) for i in range(n)]
        dps = [rast.upload(p) for p in paths]
        ident = np.array([1.0, 0, 0, 0, 1.0, 0])
        if os.environ.get("RB_C4_MASK"):  # mask-only variant (SURVEY §8d C4 "(m)": 4 B per pixel)
            slab = torch.empty((n, 64, 64), dtype=torch.float32, device=dev)
            jobs = [rb.Job(dps[i], ident, rb.FillRule.NonZero, ffi.JOB_MASK, slab.data_ptr(), 64, 64, 64, origin=i * 4096) for i in range(n)]
            what, px_bytes = "Rasterizer::mask per glyph (f32 coverage)", 4
        else:  # SURVEY §8d C4 "(s)": every glyph filled with solid black onto its own fresh LinColor canvas, 16 B per pixel
            slab = torch.empty((n, 64, 64, 4), dtype=torch.float32, device=dev)
            black = rb.LinColor(0.0, 0.0, 0.0, 1.0)
            jobs = [rb.Job(dps[i], ident, rb.FillRule.NonZero, ffi.JOB_RENDER, slab.data_ptr(), 64, 64, 64, origin=i * 4096, paint=black)
                    for i in range(n)]
            what, px_bytes = "Rasterizer::fill with solid black onto a fresh LinColor canvas per glyph (RGPU_JOB_RENDER)", 16
        info = dict(workload=f"c4: {n} synthetic random-cubic glyphs per GPU at 64x64, {what}, nonzero", canvas=[64, 64],
                    items_per_gpu=n, pixels_per_step=n * 4096, in_bytes=sum(p.input_bytes() for p in paths), out_bytes=px_bytes * n * 4096,
                    keep=[dps, slab], metric="fill throughput (Rasterizer::fill, solid paint, nonzero), pixels rasterized per second" if px_bytes == 16 else None)
        return jobs, True, info
    if name == "c5":
        # the whole 32768 x 32768 canvas as 8 scanline bands: rank r renders bands r, r + N, ... as independent
        # jobs of one batch (band-local translate(0, -y0): the reference's own y clipping crops exactly), so the total work is
        # fixed and N ranks split it => strong scaling.  Bands 0 and 7 hold no lines, bands 1..6 between 1387 and 3593.
        path = assets.load_path("tv_stroked")
        c5 = ex["tv_stroked"]["c5"]
        w, hfull = c5["size"]
        bands = int(os.environ.get("RB_BANDS", "8"))
        mine = [b for b in range(bands) if b % world == rank % bands] if world <= bands else [rank % bands]
        dp = rast.upload(path)
        jobs, canvases, rows = [], [], 0
        for b in mine:
            y0, y1 = sharding.band_rows(hfull, b, bands)
            canvas = torch.empty((y1 - y0, w), dtype=torch.float32, device=dev)
            canvases.append(canvas)
            rows += y1 - y0
            jobs.append(rb.Job(dp, sharding.band_transform(c5["tr"], y0), rb.FillRule.NonZero, ffi.JOB_MASK, canvas.data_ptr(), w, y1 - y0, w))
        info = dict(workload=f"c5: tv.path stroked (w=0.5 round/round) on a {w}x{hfull} canvas as {bands} scanline bands, bands {mine} on this rank "
                             f"({rows} rows), mask, nonzero",
                    canvas=[w, rows], items_per_gpu=len(mine), pixels_per_step=w * rows, in_bytes=path.input_bytes() * len(mine), out_bytes=4 * w * rows,
                    keep=[dp, canvases], scaling="strong" if world <= bands else "weak")
        return jobs, True, info
    if name in ("c1", "c3"):
        # Scene::render of a Fill-only scene on a device-resident LinColor layer + RGBA8 export (SURVEY §8d "(s)" bytes):
        # c1 = examples/rasterize default scene for squirrel.path -w 512; c3 = firefox.scene at 2048 x 2048 (14 gradient fills)
        from rasterize_b200 import scene as rscene
        sc = assets.load_scene("squirrel_cli_512" if name == "c1" else "firefox_2048")
        _, _, W, H, _ = rscene.fixture_jobs(rast, sc, 1)
        layer = torch.empty((H, W, 4), dtype=torch.float32, device=dev)
        rgba = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
        jobs, keep, _, _, in_bytes = rscene.fixture_jobs(rast, sc, layer.data_ptr())
        bg = sc.bg
        what = ("c1: examples/rasterize scene of data/squirrel.path at 512 px (checkerboard + fill over #f0f0f0)" if name == "c1"
                else "c3: data/firefox.scene Scene::render at 2048x2048, 14 linear/radial gradient fills")
        info = dict(canvas=[W, H], items_per_gpu=len(jobs), pixels_per_step=W * H, in_bytes=in_bytes, out_bytes=(16 + 4) * W * H,
                    keep=[keep, layer, rgba])
        if os.environ.get("RB_SCENE_ORDERED"):
            # A/B: Layer::new kernel, one raster launch per fill (ordered batch), export kernel
            def pre():
                if bg is not None:
                    rast.fill_color(layer.data_ptr(), W * H, bg)
                else:
                    rast.device_zero(layer.data_ptr(), W * H * 16)

            def post():
                rast.to_rgba8(layer.data_ptr(), rgba.data_ptr(), W * H)

            info.update(workload=what + ", device-resident LinColor layer + RGBA8 export, one launch per fill (RB_SCENE_ORDERED)", pre_step=pre,
                        post_step=post)
        else:
            info.update(workload=what + ", scene compositor: Layer::new + all fills + RGBA8 export in one raster launch, LinColor layer and RGBA8 image left in HBM",
                        scene=dict(layer=layer.data_ptr(), W=W, H=H, bg=bg, rgba=rgba.data_ptr()))
        return jobs, False, info
    raise SystemExit(f"unknown workload {name}")


def run_ours(args):
    import torch

    from rasterize_b200 import build as rb_build
    rb_build.build()  # no-op when the in-tree .so is up to date
    import rasterize_b200 as rb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (rasterize_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # the ranks of a node share its host cores: split them between the widening pools of the e2e leg
    os.environ.setdefault("RGPU_HOST_THREADS", str(max(2, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))))
    rast = rb.GpuRasterizer(device=local_rank)
    jobs, independent, info = build_workload(args.workload, rb, rast, rank, world, torch)
    prepared = rast.prepare_batch(jobs)
    stream = torch.cuda.ExternalStream(rast.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    pre_step, post_step = info.get("pre_step"), info.get("post_step")

    scn = info.get("scene")

    def step(sync=False):
        if scn:
            rast.submit_scene_prepared(prepared, scn["layer"], scn["W"], scn["H"], fresh=True, bg=scn["bg"], rgba_ptr=scn["rgba"], sync=sync)
            return
        if pre_step:
            pre_step()
        rast.submit_prepared(prepared, independent=independent, sync=sync)
        if post_step:
            post_step()

    # first call sizes the scratch buffers (and re-runs on overflow); then untimed warm-up
    step(sync=True)
    counts0 = rast.last_counts()
    for _ in range(max(args.warmup, 3)):
        step()
    rast.batch_status()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    sampler.start()
    l0 = rast.last_counts()["launches"]
    with torch.cuda.stream(stream):
        for i in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations, outside the event pair
            starts[i].record(stream)
            step()
            stops[i].record(stream)
            if (i & 31) == 31 or i == args.steps - 1:
                rast.batch_status()  # surfaces device-side errors; also bounds the launch queue depth
    barrier()
    clocks = sampler.stop()
    launches = rast.last_counts()["launches"] - l0
    per_step = np.array([a.elapsed_time(b) for a, b in zip(starts, stops)])
    total_ms = float(per_step.sum())
    # per-stage split (flatten / bin / raster) from the library's own events, sampled on a few extra steps
    rast.set_profiling(True)
    stage_samples = []
    for _ in range(min(20, args.steps)):
        with torch.cuda.stream(stream):
            flush.zero_()
        step(sync=True)
        stage_samples.append(rast.last_stage_ms())
    stage_ms = np.median(np.array(stage_samples), axis=0)
    rast.set_profiling(False)

    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    counts = rast.last_counts()
    pixels = info["pixels_per_step"] * world
    value = pixels / (ms_per_step * 1e-3) / 1e6
    lines_per_s = counts["lines"] * world / (ms_per_step * 1e-3)

    peak, peak_src = measured_peak()
    raster_s = float(stage_ms[2]) * 1e-3
    achieved = info["out_bytes"] / raster_s / 1e9 if raster_s > 0 else 0.0
    step_alg = info["in_bytes"] + info["out_bytes"]
    roofline = {
        "bound": "hbm", "kernel": ("scene_kernel (K3 + K4 for every fill of the layer + Layer::new + RGBA8 export)" if scn else
                                   "raster_kernel (K3: accumulate + row scan + fill rule + store)"),
        "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "peak_source": peak_src,
        "traffic": recorded_traffic(args.workload + ("_mask" if args.workload == "c4" and os.environ.get("RB_C4_MASK") else "")),
        "algorithmic_bytes_per_launch": info["out_bytes"], "kernel_ms": round(float(stage_ms[2]), 5),
        "stage_ms": {"flatten_and_bin": round(float(stage_ms[0]), 5), "two_pass_only_scan_and_emit": round(float(stage_ms[1]), 5),
                     "raster": round(float(stage_ms[2]), 5)},
        "step_algorithmic_bytes": step_alg, "step_frac": round(step_alg / (ms_per_step * 1e-3) / 1e9 / peak, 4),
    }

    # ---- e2e: the trait-level C-ABI call with HOST buffers (H2D of the path, kernels, D2H of the f64 mask) ----
    e2e = None
    cpu_baseline = None
    if args.workload == "c2":
        path, tr, (w, h) = info["host_path"], info["host_tr"], info["host_size"]
        img = rast.host_alloc((h, w), np.float64)  # pinned host image, as the contract asks
        for _ in range(16):  # also lets the device/host widening split of rgpu_mask settle
            rast.mask(path, tr, img, rb.FillRule.NonZero)
        n_e2e = max(5, min(20, args.steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            rast.mask(path, tr, img, rb.FillRule.NonZero)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d, d2h = rast.last_transfer_bytes()  # what the last call really moved (path points + items up; f32 and f64 rows down)
        e2e = {"value": round(w * h * world / dt / 1e6, 1), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_call": round(dt * 1e3, 4),
               "call": "rgpu_mask: host path in, f64 pinned host image out (bottom rows cross PCIe as f32 and are widened by host threads, top rows are widened on the device and DMA'd as f64; the split adapts)"}
        # f32 variant of the same call, for context
        img32 = rast.host_alloc((h, w), np.float32)
        rast.mask(path, tr, img32, rb.FillRule.NonZero)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            rast.mask(path, tr, img32, rb.FillRule.NonZero)
        e2e["f32_ms_per_call"] = round((time.perf_counter() - t0) / n_e2e * 1e3, 4)
        if rank == 0 and world == 1:
            cpu_baseline = cpu_reference_c2(threads=1, budget_s=12.0)
    elif args.workload in ("c1", "c3"):
        # Scene::render + RGBA8 export through the host-buffer entry point: host paths in (H2D), pinned RGBA8 image out (D2H)
        from rasterize_b200 import assets as _assets, scene as rscene
        sc = _assets.load_scene("squirrel_cli_512" if args.workload == "c1" else "firefox_2048")
        fills, W, H = rscene.fixture_fills_host(sc)
        prepared_host = rast.prepare_scene_host(fills)
        img = rast.host_alloc((H, W, 4), np.uint8)
        for _ in range(5):
            rast.render_scene_host(prepared_host, W, H, bg=sc.bg, rgba_out=img)
        n_e2e = max(5, min(50, args.steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            rast.render_scene_host(prepared_host, W, H, bg=sc.bg, rgba_out=img)
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d, d2h = rast.last_transfer_bytes()
        e2e = {"value": round(W * H * world / dt / 1e6, 1), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_call": round(dt * 1e3, 4),
               "call": "rgpu_render_scene_host: host paths + paints in, Layer::new + all fills + RGBA8 export on the device, pinned RGBA8 host image out"}
        if rank == 0 and world == 1:
            cpu_baseline = cpu_reference_other(args.workload, budget_s=10.0)
    elif rank == 0 and world == 1:
        cpu_baseline = cpu_reference_other(args.workload, budget_s=10.0)

    if rank == 0:
        out = {
            "metric": info.get("metric") or ("scene render throughput (Scene::render fills + RGBA8 export), pixels per second" if args.workload in ("c1", "c3")
                                             else "fill throughput (Rasterizer::mask, nonzero), pixels rasterized per second"),
            "value": round(value, 1), "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5),
            "higher_is_better": True, "scaling": info.get("scaling", "weak"), "vs_baseline": None, "dtype": "f64 geometry / Q7.24 fixed-point accumulation / f32 coverage",
            "data": "synthetic" if args.workload == "c4" else "reference asset (flat fixture of data/*.path), random-free",
            "config": {"workload": info["workload"], "canvas": info["canvas"], "items_per_gpu": info["items_per_gpu"],
                       "flatness": 0.05, "l2": "flushed between timed steps (256 MiB memset outside the event pairs)",
                       "parallelism": f"{world} independent replicas/shards, no collective"},
            "lines_per_s": round(lines_per_s, 1), "lines_per_step_per_gpu": counts["lines"], "line_refs_per_step_per_gpu": counts["line_refs"],
            "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
            "roofline": roofline, "clocks": clocks,
            "step_ms_min_med_max": [round(float(per_step.min()), 5), round(float(np.median(per_step)), 5), round(float(per_step.max()), 5)],
        }
        if e2e is not None:
            out["e2e"] = e2e
        if cpu_baseline is not None:
            out["cpu_baseline"] = cpu_baseline
        _emit(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_reference_c2(threads: int, budget_s: float):
    """Times the CPU oracle (restatement of SignedDifferenceRasterizer::mask; the Rust reference cannot be built
    here) on config 2: img.clear() + mask per iteration as benches/rasterize_bench.rs:99-108 does."""
    import oracle as O
    from rasterize_b200 import assets
    p = assets.load_path("material")
    c2 = assets.expected()["paths"]["material"]["c2"]
    w, h = c2["size"]
    tr = np.array(c2["tr"])
    op = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)
    img = np.zeros((h, w))
    op.mask_threads(tr, O.NONZERO, img, threads)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        img[:] = 0
        op.mask_threads(tr, O.NONZERO, img, threads)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 200:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": round(w * h / dt / 1e6, 1), "unit": "Mpix/s", "cores": threads, "kind": "port",
            "sample": f"{n} x (clear + mask) of material.path at {w}x{h}, nonzero, {threads} thread(s), {dt * 1e3:.2f} ms each",
            "host_cores": os.cpu_count()}


def cpu_reference_other(workload: str, budget_s: float):
    """The CPU oracle (one thread, like the single-threaded reference) on a bounded sample of the other workloads."""
    import math

    import oracle as O
    from rasterize_b200 import assets, sharding
    t_end = time.perf_counter() + budget_s
    if workload in ("c1", "c3"):
        from helpers import render_scene_oracle
        sc = assets.load_scene("squirrel_cli_512" if workload == "c1" else "firefox_2048")
        render_scene_oracle(sc)  # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            img = render_scene_oracle(sc)
            O.lin_to_rgba(img)
            n += 1
            if time.perf_counter() > t_end or n >= 50:
                break
        dt = (time.perf_counter() - t0) / n
        px = img.shape[0] * img.shape[1]
        sample = f"{n} x Scene::render + RGBA8 of the {workload} scene ({img.shape[1]}x{img.shape[0]}), fills through the oracle's Rasterizer::fill, {dt * 1e3:.2f} ms each"
    elif workload == "c4":
        n_glyphs = 400
        glyphs = [O.OraclePath.glyph(i + 1) for i in range(n_glyphs)]
        as_mask = bool(os.environ.get("RB_C4_MASK"))
        img = np.zeros((64, 64)) if as_mask else np.zeros((64, 64, 4), dtype=np.float32)
        black = O.OraclePaint.solid([0.0, 0.0, 0.0, 1.0])
        n, t0 = 0, time.perf_counter()
        while True:
            for g in glyphs:
                img[:] = 0
                if as_mask:
                    g.mask(O.IDENTITY, O.NONZERO, img)
                else:
                    g.fill(O.IDENTITY, O.NONZERO, black, img)
            n += 1
            if time.perf_counter() > t_end or n >= 20:
                break
        dt = (time.perf_counter() - t0) / (n * n_glyphs)
        px = 4096
        sample = f"{n} x {n_glyphs} glyph {'masks' if as_mask else 'solid fills (clear + Rasterizer::fill)'} at 64x64, {dt * 1e6:.1f} us per glyph"
    else:  # c5
        p = assets.load_path("tv_stroked")
        c5 = assets.expected()["paths"]["tv_stroked"]["c5"]
        w, hfull = c5["size"]
        y0, y1 = sharding.band_rows(hfull, 0 + 2, 8)  # a band that holds lines (band 0 is empty)
        op = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)
        img = np.zeros((y1 - y0, w))
        tr = sharding.band_transform(c5["tr"], y0)
        n, t0 = 0, time.perf_counter()
        while True:
            img[:] = 0
            op.mask(tr, O.NONZERO, img)
            n += 1
            if time.perf_counter() > t_end or n >= 10:
                break
        dt = (time.perf_counter() - t0) / n
        px = w * (y1 - y0)
        sample = f"{n} x (clear + mask) of band 2 of 8 ({w}x{y1 - y0}) of tv.path stroked, {dt * 1e3:.1f} ms each"
    return {"value": round(px / dt / 1e6, 1), "unit": "Mpix/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": os.cpu_count()}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.  The Rust crate cannot
    be compiled in this image (no cargo/rustc), so this is the oracle port with all the host threads the workload can
    use: rows are independent (c2, c5: one band of rows per thread), glyphs are independent (c4: glyphs dealt out over
    std::threads, one private image each); a scene's fills blend in order onto one layer, so c1 / c3 run on one thread
    exactly like the single-threaded reference.  Every step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import oracle as O
    from rasterize_b200 import assets, sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = args.workload
    metric = "fill throughput (Rasterizer::mask, nonzero), pixels rasterized per second"
    lines = None
    if wl in ("c2", "c5"):
        name, key = ("material", "c2") if wl == "c2" else ("tv_stroked", "c5")
        p = assets.load_path(name)
        e = assets.expected()["paths"][name][key]
        w, h = e["size"]
        tr = np.array(e["tr"])
        what = "c2: data/material.path (21106 segments) fitted to 4096x4096, Rasterizer::mask, nonzero"
        sample_of = f"material.path at {w}x{h}"
        if wl == "c5":
            y0, y1 = sharding.band_rows(h, 2, 8)  # a band that holds lines (band 0 is empty)
            tr = sharding.band_transform(e["tr"], y0)
            what = f"c5: tv.path stroked (w=0.5 round/round) on a {w}x{h} canvas, band 2 of 8 ({y1 - y0} rows), mask, nonzero"
            sample_of = f"band 2 of 8 ({w}x{y1 - y0}) of tv.path stroked"
            h = y1 - y0
        else:
            lines = e["lines"]
        op = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)
        img = np.zeros((h, w))
        canvas, items = [w, h], 1

        def one_step():
            img[:] = 0
            op.mask_threads(tr, O.NONZERO, img, threads)

        px, max_steps = w * h, (60 if wl == "c2" else 6)
        sample = f"(clear + mask) of {sample_of}, {threads} threads (one band of rows each)"
    elif wl == "c4":
        n_glyphs = 4000
        as_mask = bool(os.environ.get("RB_C4_MASK"))
        glyphs = [O.OraclePath.glyph(i + 1) for i in range(n_glyphs)]
        black = O.OraclePaint.solid([0.0, 0.0, 0.0, 1.0])

        def one_step():
            O.batch_threads(glyphs, O.IDENTITY, O.NONZERO, None if as_mask else black, 64, 64, threads)

        metric = metric if as_mask else "fill throughput (Rasterizer::fill, solid paint, nonzero), pixels rasterized per second"
        what = (f"c4: synthetic random-cubic glyphs at 64x64, {'Rasterizer::mask' if as_mask else 'Rasterizer::fill with solid black onto a fresh LinColor canvas'} "
                "per glyph, nonzero")
        canvas, items = [64, 64], n_glyphs
        px, max_steps = n_glyphs * 4096, 20
        sample = f"{n_glyphs} glyph {'masks' if as_mask else 'solid fills'} (clear + call) dealt out over {threads} std::threads (one private image per thread)"
    else:  # c1 / c3: order-dependent fills on one layer -> one thread, like the reference
        from helpers import render_scene_oracle
        sc = assets.load_scene("squirrel_cli_512" if wl == "c1" else "firefox_2048")
        threads = 1
        shape = []

        def one_step():
            img = render_scene_oracle(sc)
            O.lin_to_rgba(img)
            shape[:] = [img.shape[1], img.shape[0]]

        one_step()
        metric = "scene render throughput (Scene::render fills + RGBA8 export), pixels per second"
        what = ("c1: examples/rasterize scene of data/squirrel.path at 512 px (checkerboard + fill over #f0f0f0)" if wl == "c1"
                else "c3: data/firefox.scene Scene::render at 2048x2048, 14 linear/radial gradient fills") + ", Scene::render + RGBA8 export"
        canvas, items = list(shape), len(sc.fills)
        px, max_steps = shape[0] * shape[1], (100 if wl == "c1" else 12)
        sample = "Scene::render + RGBA8 export through the oracle's Rasterizer::fill, 1 thread (fills blend in order)"
    steps = max(1, min(args.steps, max_steps))
    warm = min(max(args.warmup, 1), 3)
    for _ in range(warm):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    value = round(px / dt / 1e6, 1)
    out = {
        "impl": "reference", "metric": metric, "value": value,
        "unit": "Mpix/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic" if wl == "c4" else "reference asset (flat fixture of data/*.path)",
        "config": {"workload": what, "canvas": canvas, "items_per_gpu": items, "flatness": 0.05},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": threads, "kind": "port", "sample": f"{steps} x {sample}", "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if lines is not None:
        out["lines_per_s"] = round(lines / dt, 1)
    _emit(json.dumps(out))


def _emit(line: str) -> None:
    """The ONE JSON line goes to the real stdout; everything libraries print (NCCL banners ...) was sent to stderr."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints of native libraries (e.g. "NCCL version ...") must not pollute the result line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
