"""Exchange-free sharding of the fill path across the GPUs of one box.

Two modes, both without any data-path collective:
  * a batch of independent paths/scenes  -> contiguous ranges of item index per rank, balanced by a weight
    (segments per item), every rank rasterizes its range into its own output slab;
  * one huge canvas                      -> horizontal bands of rows; rows are independent in the
    signed-difference rasterizer (reference src/rasterize.rs:421-469, 478-503), and a band-local
    `translate(0, -y0)` makes the reference's own y < 0 / y >= H clipping crop exactly.  (The library's own banded call,
    `rgpu_mask_banded_host`, shifts the finished lines by the integer row origin instead of translating the transform: that
    is bit-identical to the unsharded canvas at every size; `band_transform` below, for callers of the plain calls, is
    bit-identical up to 8192^2 and within one f32 ulp in a handful of pixels at 32768^2.)
Results go back to the host over each GPU's own PCIe link (cudaMemcpyAsync); NCCL is not involved.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int, weights: Optional[Sequence[float]] = None) -> Tuple[int, int]:
    """[begin, end) of the items rank `rank` of `world` processes handles.  Ranges are contiguous, disjoint and cover
    [0, n_items).  With `weights` (e.g. segments per item) the cut points balance the summed weight instead of the count."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if weights is None:
        return n_items * rank // world, n_items * (rank + 1) // world
    w = np.asarray(weights, dtype=np.float64)
    if len(w) != n_items:
        raise ValueError("weights must have one entry per item")
    c = np.concatenate([[0.0], np.cumsum(w)])
    total = c[-1]
    cuts = [int(np.searchsorted(c, total * k / world, side="left")) for k in range(world + 1)]
    cuts[0], cuts[-1] = 0, n_items
    for k in range(1, world + 1):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return cuts[rank], cuts[rank + 1]


#: what one (line, row) span costs the tiled raster path, in rows of a 32768-wide f32 canvas: fitted to the eight ranks of config 5
#: on 8 B200s (profiles/bench/r2v_bench_n8.json: rank time = 0.1156 ms per 4096 rows + 0.465 ns per span, residual < 0.003 ms)
SPAN_COST_ROWS = 0.0165


def band_costs(lines, height: int, n_bands: int, width: int = 32768) -> np.ndarray:
    """Relative cost of every scanline band of a canvas for the tiled raster path: its rows (the kernel is bound by the bytes it
    writes where a band is empty) plus its (line, row) spans (tiles that hold geometry are bound by issue, not bytes).  `lines` =
    the flattened outline in canvas space, (n, 4) as `GpuRasterizer.flatten` returns it.  With equal rows the ranks of config 5
    took 0.115 .. 0.152 ms on 8 GPUs: the glyph sits in the middle of the canvas."""
    L = np.asarray(lines, dtype=np.float64).reshape(-1, 4)
    ylo, yhi = np.minimum(L[:, 1], L[:, 3]), np.maximum(L[:, 1], L[:, 3])
    out = np.zeros(n_bands)
    for b in range(n_bands):
        y0, y1 = band_rows(height, b, n_bands)
        m = (yhi > y0) & (ylo < y1)
        spans = float((np.minimum(yhi[m], y1) - np.maximum(ylo[m], y0)).sum())
        out[b] = (y1 - y0) * (width / 32768.0) + SPAN_COST_ROWS * spans
    return out


def band_rows(height: int, rank: int, world: int, align: int = 8) -> Tuple[int, int]:
    """Rows [y0, y1) of the band rank `rank` renders; cut points are multiples of `align` (the raster tile height)
    except the last one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")

    def cut(k):
        if k >= world:
            return height
        y = height * k // world
        return min(height, (y + align - 1) // align * align)

    return cut(rank), max(cut(rank), cut(rank + 1))


def band_transform(tr, y0: int) -> np.ndarray:
    """`translate(0, -y0) * tr` in the reference's row-major 2x3 layout [m00, m01, m02, m10, m11, m12]."""
    t = np.array(tr, dtype=np.float64).reshape(6).copy()
    t[5] -= float(y0)
    return t
