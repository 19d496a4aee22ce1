"""CPU: `PathBuilder.arc_to` of the host mirror (arcs are converted to cubics on the host, at path
construction, exactly as the reference does — src/path.rs:945-972, src/ellipse.rs:40-96, 167-214) against the oracle's own
parser + builder on SVG paths with arcs: same segments, same control points."""
import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb

CASES = [
    # (start, [(rx, ry, rot, large, sweep, x, y), ...])
    ((10.0, 20.0), [(30.0, 15.0, 0.0, 0, 1, 60.0, 40.0)]),
    ((10.0, 20.0), [(30.0, 15.0, 35.0, 1, 0, 60.0, 40.0), (8.0, 8.0, 0.0, 1, 1, 20.0, 25.0)]),
    ((0.0, 0.0), [(1.0, 1.0, 0.0, 0, 0, 100.0, 0.0)]),            # radii too small: scaled up (s > 1)
    ((5.0, 5.0), [(25.0, 60.0, -120.0, 1, 1, 45.5, 12.25), (3.0, 9.0, 77.0, 0, 0, 5.0, 5.5)]),
]
# coincident end points / a zero radius make every angle NaN: the reference's cubic iterator never terminates there
# (src/ellipse.rs:198-201: `segment_index > NaN`), the oracle stops after 65 NaN cubics; the host mirrors draw a line like for
# the degenerate arcs `EllipArc::new_param` does detect
DEGENERATE = [
    ((1.0, 1.0), (4.0, 4.0, 0.0, 0, 1, 1.0, 1.0), 0),
    ((2.0, 3.0), (0.0, 5.0, 0.0, 0, 1, 9.0, 3.0), 1),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_arc_to_matches_the_oracle_builder(case):
    start, arcs = CASES[case]
    svg = f"M{start[0]} {start[1]}" + "".join(f" A{rx} {ry} {rot} {l} {s} {x} {y}" for rx, ry, rot, l, s, x, y in arcs) + " Z"
    b = rb.Path.builder().move_to(start)
    for rx, ry, rot, l, s, x, y in arcs:
        b.arc_to((rx, ry), rot, bool(l), bool(s), (x, y))
    p = b.close().build()
    pts, kinds, sub, closed = O.OraclePath.parse(svg).export()
    assert list(p.kinds) == list(kinds) and list(p.closed) == list(closed) and list(p.subpath_offsets) == list(sub)
    got, want = np.asarray(p.points), np.asarray(pts).reshape(-1, 2)
    assert got.shape == want.shape
    assert np.array_equal(got, want) or np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), np.abs(got - want).max()


@pytest.mark.parametrize("case", range(len(DEGENERATE)))
def test_degenerate_arcs_are_lines(case):
    start, (rx, ry, rot, l, s, x, y), n_lines = DEGENERATE[case]
    # a lead-in line that ends where the arc starts
    p = rb.Path.builder().move_to((start[0] + 3.0, start[1] - 2.0)).line_to(start).arc_to((rx, ry), rot, bool(l), bool(s), (x, y)).build()
    # the line that replaces the arc; `line_to` drops it when it has no length (src/path.rs:895-903)
    assert list(p.kinds) == [2] * (1 + n_lines)
    assert np.isfinite(np.asarray(p.points)).all()
    if n_lines:
        assert np.array_equal(np.asarray(p.points)[-1], [x, y])


def test_circle_from_two_arcs_has_the_right_area():
    """`PathBuilder::circle` of the reference is two half-turn arcs (src/path.rs:974-987): area of the flattened cubics = pi r^2"""
    r, c = 40.0, (50.0, 50.0)
    b = rb.Path.builder().move_to((c[0] + r, c[1]))
    b.arc_to((r, r), 0.0, False, False, (c[0] - r, c[1])).arc_to((r, r), 0.0, False, False, (c[0] + r, c[1])).close()
    p = b.build()
    assert list(p.kinds) == [4, 4, 4, 4]
    lines = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed).flatten()
    area = 0.5 * abs(np.sum(lines[:, 0] * lines[:, 3] - lines[:, 2] * lines[:, 1]))
    assert abs(area - np.pi * r * r) / (np.pi * r * r) < 5e-3  # chords at flatness 0.05 on r = 40
