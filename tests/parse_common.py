"""Shared by the CPU and GPU tests of the batch SVG parser: test strings (the reference's own, serialised assets, a random
walk over the grammar, malformed input) and the oracle's answer for each."""
import re

import numpy as np

import oracle as O

# src/path.rs:1107-1114 (test_bbox, test_path_parse, test_flatten ...)
SQUIRREL = """
    M12 1C9.79 1 8 2.31 8 3.92c0 1.94.5 3.03 0 6.08 0-4.5-2.77-6.34-4-6.34.05-.5-.48
    -.66-.48-.66s-.22.11-.3.34c-.27-.31-.56-.27-.56-.27l-.13.58S.7 4.29 .68 6.87c.2.33
    1.53.6 2.47.43.89.05.67.79.47.99C2.78 9.13 2 8 1 8S0 9 1 9s1 1 3 1c-3.09 1.2 0 4 0 4
    H3c-1 0-1 1-1 1h6c3 0 5-1 5-3.47 0-.85-.43-1.79 -1-2.53-1.11-1.46.23-2.68 1-2
    .77.68 3 1 3-2 0-2.21-1.79-4-4-4zM2.5 6 c-.28 0-.5-.22-.5-.5s.22-.5.5-.5.5.22.5.5
    -.22.5-.5.5z
    """
# the other strings of the reference's tests: src/path.rs:1155-1176 (test_path_parse), src/svg.rs:668-677 (test_parse_scalar, as
# an implicit polyline), src/path.rs:1225-1269 (test_stroke sources)
REFERENCE_STRINGS = [SQUIRREL, " M0,0L1-1L1,0ZL0,1 L1,1Z ", "M.5-3-11-.11", " m.5,-3 -11.5\n2.89 ", "M1 .22e0.32 3.21e-3-1.24 1e4",
                     "M2,2L8,2C11,2 11,8 8,8L5,4", "M2,2L8,2C11,2 11,8 8,8L5,4Z", "M50,0 21,90 98,35 2,35 79,90z M110,0 h90 v90 h-90z M130,20 h50 v50 h-50 z"]
# grammar corners: arcs with packed flags (SURVEY §8c gotcha 2), smooth commands with and without a predecessor, relative
# moves after close, a line shorter than EPSILON, empty and whitespace-only strings, exponents, trailing dot, plus signs
CORNER_STRINGS = ["", "   \n\t ", "M1,1", "M1,1Z", "M0 0a1 1 0 00-2 3", "M10 10A5 3 30 1 0 20 20 a4 4 0 0 1 3 0 A1 1 0 0 0 25 25",
                  "M0,0 T5,5 T10,0 S1,1 2,2 s1,1 2,2 Q1,2 3,4 t1,1 C1,1 2,2 3,3 S5,5 6,6", "M1,1 L1,1.0000000000000000001 L2,2", "M1 1 l0 0 l1e-17 0 l1 1",
                  "m1,1 2,2 z m1,1 l1,0 0,1 z l5,5", "M1. 2.e1 L+3,+4e+0 L-.5E1,1E-1", "M1,2 3,4 5,6 H7 8 V1 2 h-1 v-1", "M 0 0 L 1e400 1 L 1e-400 2",
                  "M123456789012345678901234567890 1 L2 2", "M0,0L1,1ZZzZ M5,5 Z L6,6", "L1,2 3,4", "M0 0 Q 1,1 2,2 z T 3,3", "M0 0 C 1,1 2,2 3,3 z S 3,3 4,4"]
# Arcs whose parametrisation yields NaN angles (zero radius, coincident end points): the reference's cubic iterator never
# terminates on these (src/ellipse.rs:198-201), so there is no reference answer; here they degrade to `line_to`, as the SVG
# specification asks and as the host builders do.  (string, expected kinds)
DEGENERATE_ARCS = [("M0 0 A0 0 0 0 0 5 5", [2]), ("M1 1 a4 4 0 0 1 0 0", []), ("M1 1 L2 2 a0 3 0 0 1 1 1 L9 9", [2, 2, 2])]
# malformed input: (string, kind); offsets are compared with the oracle's message
ERROR_STRINGS = ["M0,0 L", "M 1", "X", "1 2", "M0,0 A1 1 0 2 0 3 3", "M0,0 A1 1 0 1 x 3 3", "M0,0 L1,e", "M0 0 Z 1 2", "M0 0 L 1 1e", "M0 0 L .", "M0 0 L 1 -"]


def fmt(v: float) -> str:
    return repr(float(v))


def svg_of(p) -> str:
    """A `Path` as an absolute-command SVG string with shortest round-trip decimals (`Path.to_svg_path`).  NOTE: the
    reference's scanner is not correctly rounded, so parsing this text does not give the path back bit for bit — the oracle
    parses the same text and the comparison is between the two parsers."""
    return p.to_svg_path()


def random_arcs(rng, n_arcs=6) -> str:
    """Well-conditioned arcs: end points taken from an actual ellipse, sweeps away from multiples of 90 degrees.  (For an arc
    whose radii are too small, whose sweep is a multiple of 90 degrees to the last bit or whose end points nearly coincide, the
    number of cubics — or whether the arc is a full turn or nothing — hangs on the last bit of sin / cos in the reference
    itself; those are compared on the host, where the C library is shared: tests/test_parse_units.py.)"""
    parts = []
    for _ in range(n_arcs):
        cx, cy = rng.uniform(-40, 40, 2)
        rx, ry = rng.uniform(5, 30, 2)
        phi = rng.uniform(-180, 180) if rng.random() < 0.7 else 0.0
        t0 = rng.uniform(0, 2 * np.pi)
        while True:
            d = np.radians(rng.uniform(20, 340)) * (1 if rng.random() < 0.5 else -1)
            if min(abs(abs(np.degrees(d)) - m) for m in (90, 180, 270, 360)) > 8:
                break
        c, sn = np.cos(np.radians(phi)), np.sin(np.radians(phi))

        def at(t):
            x, y = rx * np.cos(t), ry * np.sin(t)
            return cx + c * x - sn * y, cy + sn * x + c * y

        (x0, y0), (x1, y1) = at(t0), at(t0 + d)
        x0, y0, x1, y1, rx, ry, phi = (float(v) for v in (x0, y0, x1, y1, rx, ry, phi))
        parts.append(f"M{x0!r},{y0!r} A{rx!r} {ry!r} {phi!r} {int(abs(d) > np.pi)} {int(d > 0)} {x1!r},{y1!r} l1,1")
    return " ".join(parts)


def random_svg(rng, n_cmds=40, arcs=True) -> str:
    """A random walk over the grammar: every command letter in both cases, implicit repeats, all number spellings, all
    separators."""
    def num(signed=True, scale=50.0):
        v = rng.uniform(-scale if signed else 1.0, scale)  # unsigned = arc radii: never 0 (a zero radius hangs the reference)
        style = rng.integers(0, 6)
        if style == 0:
            s = str(int(v))
        elif style == 1:
            s = f"{v:.3f}"
        elif style == 2:
            s = f"{v:.2e}"
        elif style == 3:
            s = f"{v:.4f}".replace("0.", ".", 1) if abs(v) < 1 else f"{v:.1f}"
        elif style == 4:
            s = f"{int(v)}."
        else:
            s = repr(float(v))
        if rng.random() < 0.1 and not s.startswith("-"):
            s = "+" + s
        return s

    def sep():
        return [" ", ",", "  ", "\n", " , ", "\t"][rng.integers(0, 6)]

    def nums(k):
        out = ""
        for i in range(k):
            s = num()
            # a minus sign or a leading dot after a fraction separates numbers by itself
            glue = "" if (i > 0 and s[0] == "-" and rng.random() < 0.5) else (sep() if i > 0 else "")
            out += glue + s
        return out

    parts = ["M" + nums(2)] if rng.random() < 0.9 else []
    for _ in range(n_cmds):
        c = "MmLlHhVvCcSsQqTtAaZz"[rng.integers(0, 20)]
        if not arcs and c in "Aa":
            c = "Ll"[c == "a"]
        if c in "Zz":
            parts.append(c)
            continue
        reps = int(rng.integers(1, 4))
        body = []
        for _ in range(reps):
            if c in "Aa":
                flags = f"{rng.integers(0, 2)}{['', ' ', ','][rng.integers(0, 3)]}{rng.integers(0, 2)}"
                tail = nums(2)
                body.append(f"{num(False, 20.0)}{sep()}{num(False, 20.0)}{sep()}{num()}{sep()}{flags}{'' if tail[0] == '-' else sep()}{tail}")
            else:
                body.append(nums({"M": 2, "L": 2, "H": 1, "V": 1, "C": 6, "S": 4, "Q": 4, "T": 2}[c.upper()]))
        parts.append(c + (sep() if rng.random() < 0.3 else "") + sep().join(body))
        if rng.random() < 0.2:
            parts.append(sep())
    return "".join(parts)


def garbage_strings(seed: int, n: int):
    """Random strings over the grammar's alphabet (no arc commands: degenerate arcs have no reference answer): almost all are
    malformed, and the kind and byte offset of the first error must be the reference's."""
    rng = np.random.default_rng(seed)
    alpha = "MmLlHhVvCcSsQqTtZz0123456789.-+eE ,\n\t"
    out = []
    for _ in range(n):
        s = "".join(alpha[j] for j in rng.integers(0, len(alpha), int(rng.integers(0, 60))))
        out.append("M" + s if rng.random() < 0.6 else s)
    return out


def oracle_parse(text: str):
    """-> dict(points, kinds, subpath_offsets, closed, bbox | None) or dict(error=(kind, offset))"""
    try:
        op = O.OraclePath.parse(text)
    except ValueError as e:
        m = re.search(r"(InvalidCmd|InvalidScalar|InvalidFlag) at offset (\d+)", str(e))
        assert m, str(e)
        return {"error": ({"InvalidCmd": 1, "InvalidScalar": 2, "InvalidFlag": 3}[m.group(1)], int(m.group(2)))}
    pts, kinds, sp, closed = op.export()
    if len(sp) == 0:
        sp = np.zeros(1, dtype=np.uint32)
    return {"points": pts, "kinds": kinds, "subpath_offsets": sp, "closed": closed, "bbox": op.bbox()}


def pack(strings):
    """-> (bytes, offsets u32[n + 1])"""
    enc = [s.encode() for s in strings]
    off = np.zeros(len(enc) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(e) for e in enc])
    return b"".join(enc), off


def has_arc(s: str) -> bool:
    return "A" in s or "a" in s


def check_batch(strings, got, infos, fit=None, exact_arcs=False, arc_rtol=64 * np.finfo(np.float64).eps):
    """got = (points[n,2], kinds, subpath_offsets, closed, path_subpath_offsets) of the whole batch; infos = per-path records
    with fields bbox, has_bbox, status, error_offset, n_segments, n_subpaths, n_points, fit_tr, fit_width, fit_height."""
    pts, kinds, sp, closed, psp = got
    pts = np.asarray(pts).reshape(-1, 2)
    seg_pt = np.concatenate([[0], np.cumsum(np.asarray(kinds, dtype=np.int64))])
    for i, s in enumerate(strings):
        want = oracle_parse(s)
        inf = infos[i]
        s0, s1 = int(psp[i]), int(psp[i + 1])
        k0, k1 = int(sp[s0]), int(sp[s1])
        if "error" in want:
            assert (int(inf["status"]), int(inf["error_offset"])) == want["error"], (s, inf["status"], inf["error_offset"], want["error"])
            assert s1 == s0 and int(inf["n_segments"]) == 0
            continue
        assert int(inf["status"]) == 0, (s, int(inf["status"]), int(inf["error_offset"]))
        g_kinds = np.asarray(kinds[k0:k1])
        if has_arc(s) and not exact_arcs and not np.array_equal(g_kinds, want["kinds"]):
            # An arc is cut into ceil(|sweep| / 90 deg) cubics.  When the sweep is a multiple of 90 deg to the last bit (always
            # so when the radii were too small and got scaled: the sweep is then pi up to rounding noise) and the axis is
            # rotated, the count hangs on the last bit of sin / cos — in the reference too.  Same curve, one cubic more or
            # less: compare the outline's box instead, and count how often it happens.
            check_batch.arc_splits += 1
            assert np.array_equal(closed[s0:s1], want["closed"]) and abs(len(g_kinds) - len(want["kinds"])) <= s.count("A") + s.count("a"), s
            gb, wb = np.asarray(inf["bbox"], dtype=np.float64), want["bbox"]
            assert np.allclose(gb, wb, rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(wb).max()))), (s, gb, wb)
            continue
        assert np.array_equal(g_kinds, want["kinds"]), (s, g_kinds, want["kinds"])
        assert np.array_equal(np.asarray(sp[s0:s1 + 1], dtype=np.int64) - k0, np.asarray(want["subpath_offsets"], dtype=np.int64)), s
        assert np.array_equal(closed[s0:s1], want["closed"]), s
        g_pts = pts[seg_pt[k0]:seg_pt[k1]]
        assert (int(inf["n_segments"]), int(inf["n_subpaths"]), int(inf["n_points"])) == (k1 - k0, s1 - s0, len(g_pts))
        if has_arc(s) and not exact_arcs:
            tol = arc_rtol * max(1.0, float(np.abs(want["points"]).max(initial=1.0)))
            assert np.abs(g_pts - want["points"]).max(initial=0.0) <= tol, (s, np.abs(g_pts - want["points"]).max())
        else:
            assert np.array_equal(g_pts.view(np.uint64), want["points"].view(np.uint64)), (s, np.abs(g_pts - want["points"]).max(initial=0.0))
        if want["bbox"] is None:
            assert int(inf["has_bbox"]) == 0
            continue
        assert int(inf["has_bbox"]) == 1
        gb = np.asarray(inf["bbox"], dtype=np.float64)
        if has_arc(s) and not exact_arcs:
            assert np.allclose(gb, want["bbox"], rtol=1e-13, atol=1e-12), (s, gb, want["bbox"])
        else:
            assert np.array_equal(gb, want["bbox"], equal_nan=True), (s, gb, want["bbox"])
            if fit is not None and np.isfinite(gb).all():
                (ow, oh), tr = O.fit_size(want["bbox"], fit[0], fit[1], fit[2])
                if ow < 2 ** 32 - 1 and oh < 2 ** 32 - 1:
                    assert (int(inf["fit_width"]), int(inf["fit_height"])) == (ow, oh), (s, inf["fit_width"], inf["fit_height"], ow, oh)
                w = np.asarray(inf["fit_tr"], dtype=np.float64)
                assert np.array_equal(w, tr, equal_nan=True), (s, w, tr)


check_batch.arc_splits = 0

INFO_DTYPE = np.dtype([("bbox", "<f8", 4), ("fit_tr", "<f8", 6), ("fit_width", "<u4"), ("fit_height", "<u4"), ("n_points", "<u4"),
                       ("n_segments", "<u4"), ("n_subpaths", "<u4"), ("status", "<i4"), ("error_offset", "<u4"), ("has_bbox", "<i4"),
                       ("n_curves", "<u4"), ("reserved", "<u4")])
assert INFO_DTYPE.itemsize == 120
