"""Shared by the CPU and GPU stroke tests: the style matrix, the oracle's `Path::stroke` as flat arrays and the comparison."""
import numpy as np

import oracle as O

JOINS = {"miter": 0, "bevel": 1, "round": 2}
CAPS = {"butt": 0, "square": 1, "round": 2}
# (width, join, miter_limit, cap); the first is config 5's (`tv.path` stroked w = 0.5 round / round)
STYLES = [(0.5, "round", 4.0, "round"), (1.0, "miter", 4.0, "butt"), (0.7, "bevel", 4.0, "square"), (2.5, "miter", 1.5, "square"),
          (0.3, "round", 4.0, "butt"), (4.0, "miter", 10.0, "round")]


def exact_case(kinds, join, cap):
    """No sin / cos / tan / acos on the way — no round join, no round cap, and no cubic in the source (`cubic_offset_rec`
    puts round joins between the pieces of a cubic whatever the style, src/curve.rs:1395-1404): the segment list must match
    the oracle bit for bit.  Otherwise only the arcs' control points may differ, in the last bits."""
    return join != "round" and cap != "round" and not (np.asarray(kinds) == 4).any()


def oracle_stroke(points, kinds, sp_off, closed, width, join, miter_limit, cap):
    op = O.OraclePath.from_flat(points, kinds, sp_off, closed)
    return op.stroke(width, join, miter_limit, cap).export()


def compare(got, want, exact, scale=1.0):
    """got / want = (points, kinds, subpath_offsets, closed).  The structure must always be identical; the points bit for
    bit when `exact`, else within a few ulp of the coordinates' magnitude (round joins: libm against CUDA trigonometry)."""
    gp, gk, gs, gc = got
    wp, wk, ws, wc = want
    if len(ws) == 0:  # the oracle exports an empty path without the leading 0
        ws = np.zeros(1, dtype=np.uint32)
    if len(gs) == 0:
        gs = np.zeros(1, dtype=np.uint32)
    assert np.array_equal(np.asarray(gk, dtype=np.uint8), np.asarray(wk, dtype=np.uint8)), "segment kinds differ"
    assert np.array_equal(np.asarray(gs, dtype=np.uint32), np.asarray(ws, dtype=np.uint32)), "subpath offsets differ"
    assert np.array_equal(np.asarray(gc, dtype=np.uint8), np.asarray(wc, dtype=np.uint8)), "closed flags differ"
    gp = np.asarray(gp, dtype=np.float64).reshape(-1)
    wp = np.asarray(wp, dtype=np.float64).reshape(-1)
    assert gp.shape == wp.shape
    same = gp.view(np.uint64) == wp.view(np.uint64)
    if exact:
        assert same.all(), f"{(~same).sum()} of {same.size} coordinates differ, max |d| = {np.abs(gp - wp).max():.3e}"
    else:
        tol = 64 * np.finfo(np.float64).eps * max(scale, float(np.abs(wp).max(initial=1.0)))
        assert np.abs(gp - wp).max(initial=0.0) <= tol, f"max |d| = {np.abs(gp - wp).max():.3e} > {tol:.3e}"
    return float(same.mean()) if same.size else 1.0


def synthetic_paths():
    """Small paths for the corner cases of the walk: open / closed, degenerate segments, cusps, coincident control points."""
    from rasterize_b200 import PathBuilder
    out = {}
    b = PathBuilder(); b.move_to((2, 2)); b.line_to((10, 2)); b.line_to((10, 8)); out["open_polyline"] = b.build()
    b = PathBuilder(); b.move_to((2, 2)); b.line_to((10, 2)); b.line_to((10, 8)); b.close(); out["closed_triangle"] = b.build()
    b = PathBuilder(); b.move_to((0, 0)); b.cubic_to((10, 0), (10, 10), (0, 10)); b.quad_to((-5, 5), (0, 0)); b.close(); out["cubic_quad_closed"] = b.build()
    b = PathBuilder(); b.move_to((0, 0)); b.cubic_to((30, 30), (-10, 30), (20, 0)); out["cubic_loop_open"] = b.build()
    b = PathBuilder(); b.move_to((0, 0)); b.quad_to((10, 0.001), (0.5, 0)); out["quad_cusp"] = b.build()
    b = PathBuilder(); b.move_to((1, 1)); b.line_to((5, 1)); b.line_to((5, 1 + 1e-17)); b.line_to((9, 1)); out["tiny_line_inside"] = b.build()
    b = PathBuilder(); b.move_to((1, 1)); b.cubic_to((1, 1), (6, 6), (6, 6)); b.cubic_to((6, 6), (6, 6), (9, 2)); out["coincident_controls"] = b.build()
    b = PathBuilder(); b.move_to((3, 3)); b.line_to((7, 3)); b.close(); b.move_to((20, 20)); b.line_to((24, 27)); out["two_subpaths_mixed"] = b.build()
    b = PathBuilder(); b.move_to((0, 0)); b.line_to((10, 0)); b.line_to((0, 0.5)); b.line_to((10, 1)); b.close(); out["sharp_miters"] = b.build()
    b = PathBuilder(); b.move_to((0, 0)); b.line_to((10, 0)); b.line_to((20, 0)); b.line_to((20, 10)); b.close(); out["collinear"] = b.build()
    rng = np.random.default_rng(11)  # lines and quads only: bit for bit in the non-round styles
    b = PathBuilder()
    for sp in range(12):
        b.move_to(tuple(rng.uniform(0, 200, 2)))
        for _ in range(40):
            if rng.random() < 0.5:
                b.line_to(tuple(rng.uniform(0, 200, 2)))
            else:
                b.quad_to(tuple(rng.uniform(0, 200, 2)), tuple(rng.uniform(0, 200, 2)))
        if sp % 2:
            b.close()
    out["random_lines_quads"] = b.build()
    return out


def random_paths(seed: int, n: int):
    """Random paths for the walk's decomposition: mixed kinds, open and closed subpaths, repeated points, zero-length and nearly
    zero-length segments, cusps, collinear runs, tiny and huge coordinates."""
    from rasterize_b200 import PathBuilder
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        b = PathBuilder()
        scale = float(rng.choice([1.0, 1.0, 30.0, 1e-3, 1e4]))
        for _sp in range(int(rng.integers(1, 4))):
            cur = rng.uniform(-10, 10, 2) * scale
            b.move_to(tuple(cur))

            def nxt():
                mode = rng.integers(0, 10)
                if mode == 0:
                    return cur.copy()                      # repeated point
                if mode == 1:
                    return cur + rng.uniform(-1, 1, 2) * 1e-17 * scale  # closer than EPSILON at scale 1
                if mode == 2:
                    return cur + np.array([rng.uniform(-5, 5) * scale, 0.0])  # collinear runs
                return cur + rng.uniform(-6, 6, 2) * scale

            for _seg in range(int(rng.integers(1, 9))):
                k = rng.integers(0, 3)
                if k == 0:
                    p = nxt(); b.line_to(tuple(p)); cur = p
                elif k == 1:
                    c, p = nxt(), nxt(); b.quad_to(tuple(c), tuple(p)); cur = p
                else:
                    c1, c2, p = nxt(), nxt(), nxt(); b.cubic_to(tuple(c1), tuple(c2), tuple(p)); cur = p
            if rng.random() < 0.5:
                b.close()
        p = b.build()
        if p.segments_count():
            out.append(p)
    return out
