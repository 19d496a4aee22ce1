"""Synthetic glyph batches of BASELINE config 4: glyph i = 3 closed contours of 6 cubics, every coordinate
`uniform() * 56 + 4` drawn in order x, y per point from the LCG of the reference's own bench
(benches/scene_bench.rs:53-88) seeded with i + 1.  Vectorised over glyphs; bit-identical to a glyph-by-glyph build."""
from __future__ import annotations

import numpy as np

from .api import PathBatch

CONTOURS, CUBICS = 3, 6


def glyph_batch(first_seed: int, n: int) -> PathBatch:
    """Glyphs with seeds first_seed .. first_seed + n - 1 as one PathBatch (18 cubics, 72 control points each)."""
    state = (np.arange(n, dtype=np.uint64) + np.uint64(first_seed)) & np.uint64(0xffffffff)

    def step():
        nonlocal state
        state = (state * np.uint64(214013) + np.uint64(2531011)) & np.uint64(0x7fffffff)
        return state >> np.uint64(16)

    def u32():
        hi = step() & np.uint64(0xffff)
        lo = step() & np.uint64(0xffff)
        return (hi << np.uint64(16)) | lo

    def uniform():
        hi = u32()
        lo = u32()
        return (((hi << np.uint64(32)) | lo) >> np.uint64(10)).astype(np.float64) * 2.0 ** -53

    def coord():
        return uniform() * 56.0 + 4.0

    pts = np.empty((n, CONTOURS * CUBICS * 4, 2), dtype=np.float64)
    k = 0
    for _ in range(CONTOURS):
        px = coord()
        py = coord()
        for _ in range(CUBICS):
            pts[:, k, 0], pts[:, k, 1] = px, py
            for j in range(1, 4):
                pts[:, k + j, 0] = coord()
                pts[:, k + j, 1] = coord()
            px, py = pts[:, k + 3, 0], pts[:, k + 3, 1]
            k += 4
    seg_per = CONTOURS * CUBICS
    kinds = np.full(n * seg_per, 4, dtype=np.uint8)
    subpath_offsets = np.arange(n * CONTOURS + 1, dtype=np.uint32) * np.uint32(CUBICS)
    closed = np.ones(n * CONTOURS, dtype=np.uint8)
    path_subpath_offsets = np.arange(n + 1, dtype=np.uint32) * np.uint32(CONTOURS)
    return PathBatch(pts.reshape(-1, 2), kinds, subpath_offsets, closed, path_subpath_offsets)
