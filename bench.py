#!/usr/bin/env python
"""Benchmark of the fill hot path (BASELINE.json: Mpix/s, flattened lines/s, nonzero fill) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]

A step is one pass of the hot path (flatten -> signed-difference raster -> composite) over one batch:
  c4 (default, BASELINE configs[3], the config the metric's 1/2/4/8-GPU sweep and the north-star target are stated on):
      a batch of 100 000 synthetic random-cubic glyphs at 64x64 (SURVEY §8d generator), each filled with solid black onto
      its own fresh LinColor canvas (RB_C4_MASK=1: mask per glyph); rank r of N takes a contiguous range of glyphs
      (sharded by path, fixed total work => strong scaling).
  c5: tv.path stroked on the 32768 x 32768 canvas as 64 scanline bands; rank r renders bands r, r + N, ... (band sharding,
      SURVEY §8e; fixed total work => strong scaling).
  c2: data/material.path fitted to 4096x4096, `Rasterizer::mask`, non-zero (one outline: replicas only at N > 1).
  c1 / c3: Scene::render of the squirrel CLI scene (512 px) / firefox.scene (2048 x 2048) on a device-resident layer + RGBA8.
The default run also measures c5 (every N) and c2 / c3 (N = 1) briefly and reports them under "other_configs".
No data-path collective anywhere (independent paths of a batch / bands of a canvas); value = units of the whole job /
max-over-ranks device time.

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


def recorded_traffic(workload: str, out_bytes=None):
    """dram bytes per raster launch from the committed ncu --set full capture, if one exists for this workload.  The capture is of
    the one-GPU launch; a rank that renders a shard of it (N > 1) gets the capture scaled by its share of the output bytes."""
    p = os.path.join(ROOT, "profiles", "raster_traffic.json")
    try:
        e = json.load(open(p))[workload]
        t = float(e["dram_bytes_per_launch"])
        if out_bytes and e.get("out_bytes_per_launch"):
            t *= out_bytes / float(e["out_bytes_per_launch"])
        return t
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


C4_TOTAL_GLYPHS = 100_000  # BASELINE configs[3]
C5_BANDS = 256             # scanline bands of the canvas (128 rows each): rank r renders a contiguous block of them as one job; the cut
                           # points balance the blocks' cost (rows + spans), see build_workload


def c4_total() -> int:
    return int(os.environ.get("RB_GLYPHS", str(C4_TOTAL_GLYPHS)))


def workload_label(name: str) -> str:
    """config.workload of both arms (ours / --impl reference): the whole job, independent of the number of GPUs."""
    if name == "c4":
        what = ("Rasterizer::mask per glyph (f32 coverage)" if os.environ.get("RB_C4_MASK") else
                "Path::fill with solid black onto a fresh LinColor canvas per glyph (ImageOwned::new_default + Rasterizer::fill)")
        return f"c4: batch of {c4_total()} synthetic random-cubic glyph paths at 64x64, {what}, nonzero, sharded by path over the GPUs"
    if name == "c2":
        return "c2: data/material.path (21106 segments) fitted to 4096x4096, Rasterizer::mask, nonzero"
    if name == "c5":
        return ("c5: data/tv.path stroked (w=0.5 round/round) on a 32768x32768 canvas, Rasterizer::mask, nonzero, split into "
                f"{int(os.environ.get('RB_BANDS', C5_BANDS))} scanline bands, a contiguous cost-balanced block of bands per GPU")
    if name == "c1":
        return "c1: examples/rasterize scene of data/squirrel.path at 512 px (checkerboard + fill over #f0f0f0), Scene::render + RGBA8 export"
    return "c3: data/firefox.scene Scene::render at 2048x2048, 14 linear/radial gradient fills, + RGBA8 export"


def workload_metric(name: str) -> str:
    if name in ("c1", "c3"):
        return "scene render throughput (Scene::render fills + RGBA8 export), pixels per second"
    if name == "c4" and not os.environ.get("RB_C4_MASK"):
        return "fill throughput (Path::fill, solid paint, nonzero), pixels rasterized per second"
    return "fill throughput (Rasterizer::mask, nonzero), pixels rasterized per second"


def build_workload(name: str, rb, rast, rank: int, world: int, torch):
    """Device-resident form of a workload for rank `rank` of `world`.  Returns (step(sync) callable, info)."""
    import assets
    from rasterize_b200 import ffi, sharding, synth
    ex = assets.expected()["paths"]
    dev = torch.device("cuda", torch.cuda.current_device())
    if name == "c2":
        path = assets.load_path("material")
        c2 = ex["material"]["c2"]
        w, h = c2["size"]
        tr = np.array(c2["tr"])
        canvas = torch.empty((h, w), dtype=torch.float32, device=dev)
        dp = rast.upload(path)
        prepared = rast.prepare_batch([rb.Job(dp, tr, rb.FillRule.NonZero, ffi.JOB_MASK, canvas.data_ptr(), w, h, w)])
        info = dict(canvas=[w, h], items=1, total_items=1, pixels_per_step=w * h, total_pixels=w * h, in_bytes=path.input_bytes(), out_bytes=4 * w * h,
                    host_path=path, host_tr=tr, host_size=(w, h), keep=[dp, canvas, prepared], scaling="weak",
                    parallelism=f"{world} independent replicas of the whole job (one outline does not shard below a band; see c5), no collective")
        return (lambda sync=False: rast.submit_prepared(prepared, independent=True, sync=sync)), info
    if name == "c4":
        # strong scaling: the batch is fixed, rank r takes a contiguous range of glyphs cut by segment count
        total = c4_total()
        a, b = sharding.shard_range(total, rank, world, weights=np.full(total, 18.0))
        n = b - a
        batch = synth.glyph_batch(a + 1, n)  # glyph i of the batch has seed i + 1
        dpb = rast.upload_batch(batch)
        as_mask = bool(os.environ.get("RB_C4_MASK"))
        px_bytes = 4 if as_mask else 16
        slab = torch.empty((max(n, 1), 64, 64) if as_mask else (max(n, 1), 64, 64, 4), dtype=torch.float32, device=dev)
        table = np.zeros(n, dtype=rb.JOB_DTYPE)
        table["path"] = dpb.handles()
        table["tr"] = np.array([1.0, 0, 0, 0, 1.0, 0])
        table["fill_rule"] = int(rb.FillRule.NonZero)
        table["mode"] = ffi.JOB_MASK if as_mask else ffi.JOB_RENDER
        table["canvas"] = slab.data_ptr()
        table["origin"] = np.arange(n, dtype=np.uint64) * np.uint64(4096)
        table["row_stride"] = 64
        table["width"] = 64
        table["height"] = 64
        keep: list = []
        if not as_mask:
            cp = rb.LinColor(0.0, 0.0, 0.0, 1.0)._c(keep)
            keep.append(cp)
            import ctypes
            table["paint"] = ctypes.addressof(cp)
        prepared = rast.prepare_job_table(table, independent=True, keep=keep)
        info = dict(canvas=[64, 64], items=n, total_items=total, pixels_per_step=n * 4096, total_pixels=total * 4096,
                    in_bytes=batch.input_bytes(), out_bytes=px_bytes * n * 4096, host_batch=batch, keep=[dpb, slab, prepared], scaling="strong",
                    parallelism=f"glyphs [{a}, {b}) of {total} on this rank, {world} ranks, no collective")
        return (lambda sync=False: (prepared.render(), rast.batch_status() if sync else None)), info
    if name == "c5":
        # strong scaling: the whole 32768 x 32768 canvas as scanline bands, rank r renders a contiguous block of them
        # as one job (band-local translate(0, -y0): the reference's own y clipping crops exactly)
        path = assets.load_path("tv_stroked")
        c5 = ex["tv_stroked"]["c5"]
        w, hfull = c5["size"]
        bands = int(os.environ.get("RB_BANDS", C5_BANDS))
        # Blocks of bands with equal COST, not equal rows: a band costs its rows (empty tiles are bound by the bytes written) plus
        # its (line, row) spans (tiles with geometry are bound by issue) — with equal rows the 8 ranks took 0.115 .. 0.152 ms, the
        # glyph sits in the middle of the canvas (sharding.band_costs; RB_C5_EQUAL_ROWS=1 for the A/B).
        if os.environ.get("RB_C5_EQUAL_ROWS") or world == 1:
            cuts = [bands * k // world for k in range(world + 1)]
        else:
            costs = sharding.band_costs(rast.flatten(path, c5["tr"], True), hfull, bands, w)
            cuts = [sharding.shard_range(bands, k, world, costs)[0] for k in range(world)] + [bands]
        if os.environ.get("RB_C5_BLOCK"):  # diagnosis: time one block of bands on one GPU, e.g. RB_C5_BLOCK=24:32
            a, b = (int(v) for v in os.environ["RB_C5_BLOCK"].split(":"))
            cuts = [a if k <= rank else b for k in range(world + 1)]
        mine = list(range(cuts[rank], cuts[rank + 1]))
        dp = rast.upload(path)
        jobs, canvases, rows = [], [], 0
        runs = []  # consecutive bands of this rank are one job (flattened once): at N = 1 the whole canvas
        for b in mine:
            y0, y1 = sharding.band_rows(hfull, b, bands)
            if runs and runs[-1][1] == y0:
                runs[-1][1] = y1
            elif y1 > y0:
                runs.append([y0, y1])
        for y0, y1 in runs:
            canvas = torch.empty((y1 - y0, w), dtype=torch.float32, device=dev)
            canvases.append(canvas)
            rows += y1 - y0
            jobs.append(rb.Job(dp, sharding.band_transform(c5["tr"], y0), rb.FillRule.NonZero, ffi.JOB_MASK, canvas.data_ptr(), w, y1 - y0, w))
        prepared = rast.prepare_batch(jobs)
        info = dict(canvas=[w, hfull], items=len(mine), total_items=bands, pixels_per_step=w * rows, total_pixels=w * hfull,
                    in_bytes=path.input_bytes() * len(mine), out_bytes=4 * w * rows, host_path=path, host_tr=np.array(c5["tr"]), host_size=(w, hfull),
                    bands=bands, band_block=(cuts[rank], cuts[rank + 1]), keep=[dp, canvases, prepared], scaling="strong",
                    parallelism=f"{len(mine)} of {bands} scanline bands ({rows} rows) on this rank, {world} ranks, no collective")
        return (lambda sync=False: rast.submit_prepared(prepared, independent=True, sync=sync)), info
    if name in ("c1", "c3"):
        # Scene::render of a Fill-only scene on a device-resident LinColor layer + RGBA8 export (SURVEY §8d "(s)" bytes):
        # c1 = examples/rasterize default scene for squirrel.path -w 512; c3 = firefox.scene at 2048 x 2048 (14 gradient fills)
        from rasterize_b200 import scene as rscene
        sc = assets.load_scene("squirrel_cli_512" if name == "c1" else "firefox_2048")
        _, _, W, H, _ = rscene.fixture_jobs(rast, sc, 1)
        layer = torch.empty((H, W, 4), dtype=torch.float32, device=dev)
        rgba = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
        jobs, keep, _, _, in_bytes = rscene.fixture_jobs(rast, sc, layer.data_ptr())
        prepared = rast.prepare_batch(jobs)
        bg = sc.bg
        info = dict(canvas=[W, H], items=len(jobs), total_items=len(jobs), pixels_per_step=W * H, total_pixels=W * H, in_bytes=in_bytes,
                    out_bytes=(16 + 4) * W * H, keep=[keep, layer, rgba, prepared], scene_name="squirrel_cli_512" if name == "c1" else "firefox_2048",
                    scaling="weak", parallelism=f"{world} independent replicas of the scene (fills blend in order on one layer), no collective")
        if os.environ.get("RB_SCENE_ORDERED"):
            # A/B: Layer::new kernel, one raster launch per fill (ordered batch), export kernel
            def step(sync=False):
                if bg is not None:
                    rast.fill_color(layer.data_ptr(), W * H, bg)
                else:
                    rast.device_zero(layer.data_ptr(), W * H * 16)
                rast.submit_prepared(prepared, independent=False, sync=sync)
                rast.to_rgba8(layer.data_ptr(), rgba.data_ptr(), W * H)
            info["variant"] = "one launch per fill (RB_SCENE_ORDERED)"
        else:
            def step(sync=False):
                rast.submit_scene_prepared(prepared, layer.data_ptr(), W, H, fresh=True, bg=bg, rgba_ptr=rgba.data_ptr(), sync=sync)
            info["variant"] = "scene compositor: Layer::new + all fills + RGBA8 export in one raster launch"
            info["scene"] = True
        return step, info
    raise SystemExit(f"unknown workload {name}")


class Harness:
    """Process-wide state of one bench run: device, rasterizer, distributed plumbing, L2 flush buffer."""

    def __init__(self):
        import torch

        from rasterize_b200 import build as rb_build
        rb_build.build()  # no-op when the in-tree .so is up to date
        import rasterize_b200 as rb
        self.torch, self.rb = torch, rb
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (rasterize_b200 has no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        # Run this rank (and the pinned buffers and host threads it creates from here on) on the CPUs next to its GPU: the
        # e2e legs are bound by device writes into host memory, and a pinned buffer on the far socket costs PCIe bandwidth.
        self.affinity = None
        if not os.environ.get("RB_NO_AFFINITY"):
            try:
                import pynvml
                pynvml.nvmlInit()
                h = pynvml.nvmlDeviceGetHandleByIndex(self.local_rank)
                pynvml.nvmlDeviceSetCpuAffinity(h)
                self.affinity = len(os.sched_getaffinity(0))
            except Exception as e:  # no NVML, a container without the topology: keep the inherited affinity
                self.affinity = f"unchanged ({type(e).__name__})"
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        # the ranks of a node share its host cores: split them between the widening pools of the e2e legs
        os.environ.setdefault("RGPU_HOST_THREADS", str(max(2, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", self.world))))))
        self.rast = rb.GpuRasterizer(device=self.local_rank)
        self.stream = torch.cuda.ExternalStream(self.rast.stream(), device=torch.device("cuda", self.local_rank))
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_over_ranks(self, v: float) -> list:
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.dist is None:
            return [float(t.item())]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def sum_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def time_calls(hx: Harness, call, n: int, warm: int) -> float:
    """Seconds per call of a synchronous host-buffer entry point: max over ranks of the wall time of n calls, bracketed by barriers."""
    for _ in range(warm):
        call()
    hx.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    hx.torch.cuda.synchronize()
    return hx.max_over_ranks((time.perf_counter() - t0) / n)


def measure(hx: Harness, name: str, steps: int, warmup: int, with_cpu: bool, sample_clocks: bool):
    """One workload on this rank's GPU: device-timed steps (inputs resident), roofline of the dominant kernel, e2e through
    the host-buffer C-ABI call, CPU oracle beside it (rank 0)."""
    torch, rb, rast = hx.torch, hx.rb, hx.rast
    step, info = build_workload(name, rb, rast, hx.rank, hx.world, torch)
    has_work = info["pixels_per_step"] > 0
    # first call sizes the scratch buffers (and re-runs on overflow); then untimed warm-up
    if has_work:
        step(sync=True)
    for _ in range(max(warmup, 3)):
        if has_work:
            step()
    rast.batch_status()
    sampler = ClockSampler(hx.local_rank) if sample_clocks else None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    hx.barrier()
    if sampler:
        sampler.start()
    l0 = rast.last_counts()["launches"]
    with torch.cuda.stream(hx.stream):
        for i in range(steps):
            hx.flush.zero_()  # L2 flush between timed iterations, outside the event pair
            starts[i].record(hx.stream)
            if has_work:
                step()
            stops[i].record(hx.stream)
            if (i & 31) == 31 or i == steps - 1:
                rast.batch_status()  # surfaces device-side errors; also bounds the launch queue depth
    hx.barrier()
    clocks = sampler.stop() if sampler else None
    launches = rast.last_counts()["launches"] - l0
    per_step = np.array([a.elapsed_time(b) for a, b in zip(starts, stops)])
    rank_ms = [v / steps for v in hx.gather_over_ranks(float(per_step.sum()))]  # device time of every rank's steps
    ms_per_step = max(rank_ms)
    counts = rast.last_counts() if has_work else dict(lines=0, line_refs=0, launches=0)
    # per-stage split (flatten / bin / raster) from the library's own events, sampled on a few extra steps
    stage_ms = np.zeros(3)
    if has_work:
        rast.set_profiling(True)
        samples = []
        for _ in range(min(20, steps)):
            with torch.cuda.stream(hx.stream):
                hx.flush.zero_()
            step(sync=True)
            samples.append(rast.last_stage_ms())
        stage_ms = np.median(np.array(samples), axis=0)
        rast.set_profiling(False)
    total_lines = hx.sum_over_ranks(float(counts["lines"]))
    value = info["total_pixels"] / (ms_per_step * 1e-3) / 1e6
    if info["scaling"] == "weak":
        value *= hx.world
    peak, peak_src = measured_peak()
    raster_s = float(stage_ms[2]) * 1e-3
    achieved = info["out_bytes"] / raster_s / 1e9 if raster_s > 0 else 0.0
    step_alg = info["in_bytes"] + info["out_bytes"]
    kernel = {"c4": "small_canvas_kernel (K1..K4 fused: flatten + accumulate + row scan + fill rule + composite + store, one CTA per glyph)",
              "c1": "scene_kernel (K3 + K4 for every fill of the layer + Layer::new + RGBA8 export)",
              "c3": "scene_kernel (K3 + K4 for every fill of the layer + Layer::new + RGBA8 export)"}.get(
                  name, "raster_kernel (K3: accumulate + row scan + fill rule + store)")
    roofline = {
        "bound": "hbm", "kernel": kernel, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "peak_source": peak_src, "traffic": recorded_traffic(name + ("_mask" if name == "c4" and os.environ.get("RB_C4_MASK") else ""), info["out_bytes"]),
        "algorithmic_bytes_per_launch": info["out_bytes"], "kernel_ms": round(float(stage_ms[2]), 5),
        "stage_ms": {"flatten_and_bin": round(float(stage_ms[0]), 5), "two_pass_only_scan_and_emit": round(float(stage_ms[1]), 5),
                     "raster": round(float(stage_ms[2]), 5)},
        "step_algorithmic_bytes": step_alg,
        "step_frac": round(hx.sum_over_ranks(step_alg) / (ms_per_step * 1e-3) / 1e9 / (peak * hx.world), 4),
        "note": "rank 0's launch; achieved = algorithmic output bytes of the launch / its CUDA-event duration",
    }

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (H2D of the paths, kernels, D2H of the result) ----
    e2e = None
    n_e2e = max(3, min(10, steps))
    if name == "c4":
        batch = info["host_batch"]
        n = len(batch)
        as_mask = bool(os.environ.get("RB_C4_MASK"))
        out = rast.host_alloc((max(n, 1), 64, 64) if as_mask else (max(n, 1), 64, 64, 4), np.float32)[:n]  # ONE pinned buffer per rank
        black = None if as_mask else rb.LinColor(0.0, 0.0, 0.0, 1.0)
        call = (lambda: rast.fill_batch_host(batch, rb.FillRule.NonZero, black, 64, 64, out)) if n else (lambda: None)
        dt = time_calls(hx, call, n_e2e, 10)  # warm-up lets the split of the download settle
        h2d, d2h = rast.last_transfer_bytes() if n else (0, 0)
        e2e = {"value": round(info["total_pixels"] / dt / 1e6, 1), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_call": round(dt * 1e3, 3),
               "call": "rgpu_fill_batch_host on this rank's shard: host path arrays in, kernels in chunks, every chunk's images copied with cudaMemcpyAsync "
                       "on a second stream into ONE pinned host buffer while the next chunk renders and the chunk after it is prepared on a helper thread; "
                       "f32 LinColor out, 16 B per pixel: with a plain solid paint an adaptive share of every chunk crosses PCIe as f32 coverage (4 B per "
                       "pixel) and host threads write colour * coverage into the caller's buffer with streaming stores while the rest arrives "
                       "as LinColor by DMA (bit-identical to the device's own multiplication; d2h_bytes_per_step counts what really crossed)"}
        if not as_mask and n:
            # the same call with the split download switched off: every byte of the result crosses PCIe as LinColor
            os.environ["RGPU_E2E_EXPAND"] = "0"
            try:
                e2e["dma_only_ms_per_call"] = round(time_calls(hx, call, 3, 1) * 1e3, 3)
            finally:
                del os.environ["RGPU_E2E_EXPAND"]
            rgba = rast.host_alloc((n, 64, 64, 4), np.uint8)
            dt8 = time_calls(hx, lambda: rast.fill_batch_host(batch, rb.FillRule.NonZero, black, 64, 64, rgba), n_e2e, 1)
            e2e["rgba8_ms_per_call"] = round(dt8 * 1e3, 3)
            e2e["rgba8_value"] = round(info["total_pixels"] / dt8 / 1e6, 1)
    elif name == "c2":
        path, tr, (w, h) = info["host_path"], info["host_tr"], info["host_size"]
        img = rast.host_alloc((h, w), np.float64)  # pinned host image, as the contract asks
        dt = time_calls(hx, lambda: rast.mask(path, tr, img, rb.FillRule.NonZero), max(5, min(20, steps)), 16)  # warm-up lets the widening split settle
        h2d, d2h = rast.last_transfer_bytes()  # what the last call really moved (path points + items up; f32 and f64 rows down)
        e2e = {"value": round(w * h * hx.world / dt / 1e6, 1), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_call": round(dt * 1e3, 4),
               "call": "rgpu_mask: host path in, f64 pinned host image out (bottom rows cross PCIe as f32 and are widened by host threads, top rows are widened on the device and DMA'd as f64; the split adapts)"}
        img32 = rast.host_alloc((h, w), np.float32)
        e2e["f32_ms_per_call"] = round(time_calls(hx, lambda: rast.mask(path, tr, img32, rb.FillRule.NonZero), n_e2e, 1) * 1e3, 4)
    elif name == "c5":
        path, tr, (w, h) = info["host_path"], info["host_tr"], info["host_size"]
        img = rast.host_alloc((h, w), np.float32)  # the whole canvas, pinned; this rank fills the rows of its bands
        b0, b1 = info["band_block"]
        call = lambda: rast.mask_banded(path, tr, img, rb.FillRule.NonZero, n_bands=info["bands"], band_first=b0, band_count=b1 - b0)  # noqa: E731
        dt = time_calls(hx, call, max(2, min(4, steps)), 1)
        h2d, d2h = rast.last_transfer_bytes()
        e2e = {"value": round(info["total_pixels"] / dt / 1e6, 1), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_call": round(dt * 1e3, 3),
               "call": "rgpu_mask_banded_host on this rank's block of bands: host path in, the block rendered as one job and brought into its rows of a "
                       "pinned f32 host image of the whole canvas run-coded (class byte per 64-pixel segment + the literal edge segments cross PCIe, "
                       "host threads rebuild the rows: the same bytes as dense copies; d2h_bytes_per_step counts what really crossed)"}
        os.environ["RGPU_E2E_RUNCODE"] = "0"  # the same call with dense copies
        try:
            e2e["dense_copy_ms_per_call"] = round(time_calls(hx, call, 2, 1) * 1e3, 3)
        finally:
            del os.environ["RGPU_E2E_RUNCODE"]
    elif name in ("c1", "c3"):
        # Scene::render + RGBA8 export through the host-buffer entry point: host paths in (H2D), pinned RGBA8 image out (D2H)
        import assets as _assets
        from rasterize_b200 import scene as rscene
        sc = _assets.load_scene(info["scene_name"])
        fills, W, H = rscene.fixture_fills_host(sc)
        prepared_host = rast.prepare_scene_host(fills)
        img = rast.host_alloc((H, W, 4), np.uint8)
        dt = time_calls(hx, lambda: rast.render_scene_host(prepared_host, W, H, bg=sc.bg, rgba_out=img), max(5, min(50, steps)), 5)
        h2d, d2h = rast.last_transfer_bytes()
        e2e = {"value": round(W * H * hx.world / dt / 1e6, 1), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_call": round(dt * 1e3, 4),
               "call": "rgpu_render_scene_host: host paths + paints in, Layer::new + all fills + RGBA8 export on the device, pinned RGBA8 host image out"}
    cpu_baseline = None
    if with_cpu and hx.rank == 0:
        cpu_baseline = cpu_reference_c2(threads=1, budget_s=10.0) if name == "c2" else cpu_reference_other(name, budget_s=10.0)
    if hx.dist is not None:
        hx.dist.barrier()  # the other ranks wait for rank 0's CPU leg instead of racing ahead into the next workload

    out = {
        "metric": workload_metric(name), "value": round(value, 1), "unit": "Mpix/s", "n_gpus": hx.world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": info["scaling"], "vs_baseline": None,
        "dtype": "f64 geometry / Q7.24 fixed-point accumulation / f32 coverage and colour",
        "data": "synthetic" if name == "c4" else "reference asset (flat fixture of data/*.path), random-free",
        "config": {"workload": workload_label(name), "canvas": info["canvas"], "total_items": info["total_items"], "items_this_rank": info["items"],
                   "flatness": 0.05, "l2": "flushed between timed steps (256 MiB memset outside the event pairs)", "parallelism": info["parallelism"]},
        "lines_per_s": round(total_lines / (ms_per_step * 1e-3), 1),
        "lines_per_step": int(total_lines), "gpu_launches": int(launches), "launches_per_step": launches / steps, "roofline": roofline,
        "rank_ms_per_step": [round(v, 5) for v in rank_ms],
        "host_cpus_of_this_rank": hx.affinity,
        "step_ms_min_med_max": [round(float(per_step.min()), 5), round(float(np.median(per_step)), 5), round(float(per_step.max()), 5)],
    }
    if "variant" in info:
        out["config"]["variant"] = info["variant"]
    if clocks is not None:
        out["clocks"] = clocks
    if e2e is not None:
        out["e2e"] = e2e
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
    del step, info
    return out


def measure_pre_stages():
    """The steps in front of the path that SURVEY §8f ranks next, on the device (N = 1): `Path::stroke` of config 5's / config 2's
    outline and the batch SVG parse + bbox + fit_size of a glyph batch.  Wall clock of the reference-facing call (host data in,
    device-resident path(s) out) with the kernels' CUDA-event time beside it, and the oracle on one host thread as baseline."""
    import oracle as O
    import rasterize_b200 as rb
    import assets
    from rasterize_b200 import Align, LineCap, LineJoin, StrokeStyle, synth
    rast = rb.GpuRasterizer()
    rast.set_profiling(True)
    out = {}

    def timed(call, n, warm=2):
        for _ in range(warm):
            call()
        ts, st = [], None
        for _ in range(n):
            t0 = time.perf_counter()
            r = call(keep=True)
            ts.append(time.perf_counter() - t0)
            st = rast.last_stage_ms()
            r.free()
        return statistics.median(ts), st

    style = StrokeStyle(0.5, LineJoin.Round, 4.0, LineCap.Round)
    for name in ("tv", "material"):
        p = assets.load_path(name)

        def call(keep=False):
            dp = rast.stroke(p, style)
            if keep:
                return dp
            dp.free()

        t, st = timed(call, 10)
        op = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 0.3:
            op.stroke(0.5, "round", 4.0, "round")
            reps += 1
        tc = (time.perf_counter() - t0) / reps
        out[f"stroke_{name}"] = {"call": "rgpu_path_stroke (w = 0.5, round / round): host path in, device-resident outline out",
                                 "segments_in": p.segments_count(), "ms_per_call": round(t * 1e3, 4),
                                 "kernel_ms": {"pieces_count_scan": round(st[0], 4), "emit": round(st[2], 4)},
                                 "cpu_baseline": {"value": round(tc * 1e3, 4), "unit": "ms", "cores": 1, "kind": "port", "sample": f"{reps} x Path::stroke"}}
    n = 100000
    pb = synth.glyph_batch(1, 2000)
    strings = [pb.path(i).to_svg_path().encode() for i in range(2000)]
    strings = (strings * (n // len(strings)))[:n]  # 2 000 distinct outlines, repeated: the parser's work is the same
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in strings])
    text_bytes = b"".join(strings)
    text = rast.host_alloc((len(text_bytes),), np.uint8)  # pinned, as the contract asks of a step's inputs
    text[:] = np.frombuffer(text_bytes, dtype=np.uint8)
    info_box = {}

    def call(keep=False):
        dpb, info = rast.parse_svg_batch((text, off), fit=(64, 64, Align.Mid))
        info_box["info"] = info
        if keep:
            return dpb
        dpb.free()

    t, st = timed(call, 5, warm=1)
    t0 = time.perf_counter()
    m = 300
    for s in strings[:m]:
        op = O.OraclePath.parse(s)
        O.fit_size(op.bbox(), 64, 64, 1)
    tc = (time.perf_counter() - t0) / m
    out["parse_glyph_batch"] = {"call": "rgpu_parse_svg_batch + fit_size(64 x 64, Mid): host text in, device path batch + bbox + fit transforms out",
                                "paths": n, "text_mb": round(len(text) / 1e6, 2), "segments": int(info_box["info"]["n_segments"].sum()),
                                "ms_per_call": round(t * 1e3, 3), "value": round(n / t / 1e6, 3), "unit": "Mpaths/s",
                                "kernel_ms": {"count": round(st[0], 4), "emit": round(st[2], 4)},
                                "cpu_baseline": {"value": round(1e-6 / tc, 4), "unit": "Mpaths/s", "cores": 1, "kind": "port",
                                                 "sample": f"{m} x (parse + bbox + fit_size), {tc * 1e6:.1f} us per path"}}
    s = assets.load_path("material").to_svg_path().encode()

    def call1(keep=False):
        dpb, _ = rast.parse_svg_batch([s])
        if keep:
            return dpb
        dpb.free()

    t, st = timed(call1, 10)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 0.3:
        O.OraclePath.parse(s).bbox()
        reps += 1
    tc = (time.perf_counter() - t0) / reps
    out["parse_material"] = {"call": "rgpu_parse_svg_batch of ONE string (the reference's material-big `parse` + `bbox` bench ids), cut at absolute movetos",
                             "text_mb": round(len(s) / 1e6, 3), "ms_per_call": round(t * 1e3, 4),
                             "kernel_ms": {"count": round(st[0], 4), "emit": round(st[2], 4)},
                             "cpu_baseline": {"value": round(tc * 1e3, 4), "unit": "ms", "cores": 1, "kind": "port", "sample": f"{reps} x (parse + bbox)"}}
    rast.close()
    return out


def run_ours(args):
    hx = Harness()
    main = measure(hx, args.workload, args.steps, args.warmup, with_cpu=True, sample_clocks=True)
    # the other BASELINE configs, measured briefly in the same run (value, ms_per_step, roofline, e2e): c5 shards by scanline band
    # at every N; c2 / c3 do not shard (replicas only), so they are measured at N = 1
    others = {}
    if args.others:
        names = [n for n in (["c5", "c2", "c3"] if hx.world == 1 else ["c5"]) if n != args.workload]
        for n in names:
            try:
                r = measure(hx, n, max(5, min(args.steps, 30)), min(args.warmup, 5), with_cpu=(hx.world == 1), sample_clocks=False)
                others[n] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "rank_ms_per_step", "scaling", "config", "lines_per_s", "gpu_launches", "e2e", "cpu_baseline")
                             if k in r}
                others[n]["roofline"] = {k: r["roofline"][k] for k in ("kernel", "achieved", "frac", "kernel_ms", "stage_ms", "step_frac")}
            except Exception as e:  # a side measurement must not take the headline line down with it
                others[n] = {"error": f"{type(e).__name__}: {e}"}
        if hx.world == 1:
            try:
                others["pre_stages"] = measure_pre_stages()
            except Exception as e:
                others["pre_stages"] = {"error": f"{type(e).__name__}: {e}"}
    if hx.rank == 0:
        if others:
            main["other_configs"] = others
        _emit(json.dumps(main))
    if hx.dist is not None:
        hx.dist.barrier()
        hx.dist.destroy_process_group()


def cpu_reference_c2(threads: int, budget_s: float):
    """Times the CPU oracle (restatement of SignedDifferenceRasterizer::mask; the Rust reference cannot be built
    here) on config 2: img.clear() + mask per iteration as benches/rasterize_bench.rs:99-108 does."""
    import oracle as O
    import assets
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
    import assets
    from rasterize_b200 import sharding
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
    exactly like the single-threaded reference.  Every step is a bounded sample of the workload; config / metric / unit
    are those of our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import oracle as O
    import assets
    from rasterize_b200 import sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = args.workload
    lines = None
    if wl in ("c2", "c5"):
        name, key = ("material", "c2") if wl == "c2" else ("tv_stroked", "c5")
        p = assets.load_path(name)
        e = assets.expected()["paths"][name][key]
        w, h = e["size"]
        tr = np.array(e["tr"])
        canvas, total_items = [w, h], 1
        sample_of = f"material.path at {w}x{h}"
        if wl == "c5":
            y0, y1 = sharding.band_rows(h, 2, 8)  # a band that holds lines (the top and bottom 1/8 of the canvas are empty)
            tr = sharding.band_transform(e["tr"], y0)
            sample_of = f"rows [{y0}, {y1}) of the {w}x{h} canvas of tv.path stroked"
            h = y1 - y0
            total_items = int(os.environ.get("RB_BANDS", C5_BANDS))
        else:
            lines = e["lines"]
        op = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)
        img = np.zeros((h, w))

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

        canvas, total_items = [64, 64], c4_total()
        px, max_steps = n_glyphs * 4096, 20
        sample = (f"the first {n_glyphs} of the {c4_total()} glyphs, {'masks' if as_mask else 'solid fills'} (clear + call) dealt out over {threads} "
                  "std::threads (one private image per thread)")
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
        canvas, total_items = list(shape), len(sc.fills)
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
        "impl": "reference", "metric": workload_metric(wl), "value": value,
        "unit": "Mpix/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 4),
        "higher_is_better": True, "scaling": "strong" if wl in ("c4", "c5") else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic" if wl == "c4" else "reference asset (flat fixture of data/*.path), random-free",
        "config": {"workload": workload_label(wl), "canvas": canvas, "total_items": total_items, "flatness": 0.05},
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
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-others", dest="others", action="store_false", help="skip the brief other_configs measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
