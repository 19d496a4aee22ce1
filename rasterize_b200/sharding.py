"""Exchange-free sharding of the fill path across the GPUs of one box.

Two modes, both without any data-path collective:
  * a batch of independent paths/scenes  -> contiguous ranges of item index per rank, balanced by a weight
    (segments per item), every rank rasterizes its range into its own output slab;
  * one huge canvas                      -> horizontal bands of rows; rows are independent in the
    signed-difference rasterizer (reference src/rasterize.rs:421-469, 478-503), and a band-local
    `translate(0, -y0)` makes the reference's own y < 0 / y >= H clipping crop exactly.
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
