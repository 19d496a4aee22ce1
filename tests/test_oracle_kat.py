"""Pins the CPU oracle (oracle/) to every known-answer test the reference holds for the fill path.

Each test cites the reference test it restates (paths relative to /root/reference).  Runs on CPU.
"""
import math

import numpy as np
import pytest

import oracle as O

EPS = np.finfo(np.float64).eps

SQUIRREL = """
M12 1C9.79 1 8 2.31 8 3.92c0 1.94.5 3.03 0 6.08 0-4.5-2.77-6.34-4-6.34.05-.5-.48
-.66-.48-.66s-.22.11-.3.34c-.27-.31-.56-.27-.56-.27l-.13.58S.7 4.29 .68 6.87c.2.33
1.53.6 2.47.43.89.05.67.79.47.99C2.78 9.13 2 8 1 8S0 9 1 9s1 1 3 1c-3.09 1.2 0 4 0 4
H3c-1 0-1 1-1 1h6c3 0 5-1 5-3.47 0-.85-.43-1.79 -1-2.53-1.11-1.46.23-2.68 1-2
.77.68 3 1 3-2 0-2.21-1.79-4-4-4zM2.5 6 c-.28 0-.5-.22-.5-.5s.22-.5.5-.5.5.22.5.5
-.22.5-.5.5z
"""


def approx(a, b, tol=EPS):
    assert abs(a - b) <= tol, (a, b)


def test_signed_difference_line():
    """src/rasterize.rs:946-1003 test_signed_difference_line"""
    img = np.zeros((2, 5))
    O.signed_difference_line(img, (0.5, 1.0, 3.5, 0.0))
    a0 = (0.5 * (1.0 / 6.0)) / 2.0
    a1 = ((1.0 / 6.0) + (3.0 / 6.0)) / 2.0
    a2 = ((3.0 / 6.0) + (5.0 / 6.0)) / 2.0
    approx(img[0, 0], -a0)
    approx(img[0, 1], a0 - a1)
    approx(img[0, 2], a1 - a2)
    approx(img[0, 3], a0 - a1)
    approx(img[0, 4], -a0)
    approx(img.sum(), -1.0)

    img[:] = 0
    O.signed_difference_line(img, (-1.0, 0.0, 1.0, 1.0))
    approx(img[0, 0], 3.0 / 4.0)
    approx(img[0, 1], 1.0 / 4.0)

    img[:] = 0
    O.signed_difference_line(img, (0.0, -0.5, 2.0, 1.5))
    approx(img[0, 0], 1.0 / 8.0)
    approx(img[0, 1], 1.0 - 2.0 / 8.0)
    approx(img[0, 2], 1.0 / 8.0)
    approx(img[1, 1], 1.0 / 8.0)
    approx(img[1, 2], 0.5 - 1.0 / 8.0)

    img[:] = 0
    O.signed_difference_line(img, (0.1, 0.1, 1.9, 0.9))
    approx(img[0, 0], 0.18)
    approx(img[0, 1], 0.44)
    approx(img[0, 2], 0.18)

    img[:] = 0
    O.signed_difference_line(img, (0.1, 0.1, 0.9, 0.9))
    approx(img[0, 0], 0.4)
    approx(img[0, 1], 0.8 - 0.4)

    img[:] = 0
    O.signed_difference_line(img, (0.5, 0.5, 0.5, 1.75))
    approx(img[0, 0], 1.0 / 4.0)
    approx(img[0, 1], 1.0 / 4.0)
    approx(img[1, 0], 3.0 / 8.0)
    approx(img[1, 1], 3.0 / 8.0)


def test_rasterizer_octagon():
    """src/rasterize.rs:1065-1120 test_rasterizer (SignedDifference arm)"""
    expected = np.array([
        0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0,
        0.0, 0.5, 1.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.0,
        0.0, 0.0, 0.25, 0.75, 1.0, 0.75, 0.25, 0.0, 0.0,
        0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
    ]).reshape(4, 9)
    path = O.OraclePath.parse("M1,0 L1,1 L2,2 L4,3 L5,3 L7,2 L8,1 L8,0 Z")
    img = np.zeros((4, 9))
    path.mask(O.IDENTITY, O.EVENODD, img)
    assert np.abs(img - expected).max() <= EPS
    # square that goes exactly through the sides of the image: must not crash
    path = O.OraclePath.parse("M1,1 v2 h1 v-2 Z")
    img = np.zeros((3, 3))
    path.mask(O.IDENTITY, O.EVENODD, img)
    assert np.isfinite(img).all()


def test_fill_rule():
    """src/rasterize.rs:1123-1161 test_fill_rule"""
    path = O.OraclePath.parse("""
        M50,0 21,90 98,35 2,35 79,90z
        M110,0 h90 v90 h-90z
        M130,20 h50 v50 h-50 z
        M210,0  h90 v90 h-90 z
        M230,20 v50 h50 v-50 z
    """)
    (w, h), tr, _ = path.size()
    assert (w, h) == (300, 92)
    img = np.zeros((h, w))
    path.mask(tr, O.EVENODD, img)
    y, x0, x1, x2 = 50, 50, 150, 250
    approx(img[y, x0], 0.0, 1e-6)
    approx(img[y, x1], 0.0, 1e-6)
    approx(img[y, x2], 0.0, 1e-6)
    approx(img.sum(), 13130.0, 1.0)
    img[:] = 0
    path.mask(tr, O.NONZERO, img)
    approx(img[y, x0], 1.0, 1e-6)
    approx(img[y, x1], 1.0, 1e-6)
    approx(img[y, x2], 0.0, 1e-6)
    approx(img.sum(), 16492.5, 1.0)


def test_bbox():
    """src/path.rs:1098-1105 test_bbox"""
    bb = O.OraclePath.parse(SQUIRREL).bbox()
    approx(bb[0], 0.25)
    approx(bb[1], 1.0)
    approx(bb[2] - bb[0], 15.75)
    approx(bb[3] - bb[1], 14.0)


def _builder_path(cmds):
    # builds through the SVG front-end with absolute commands only (== PathBuilder calls)
    return O.OraclePath.parse(cmds)


def test_path_parse():
    """src/path.rs:1117-1179 test_path_parse"""
    path = O.OraclePath.parse(SQUIRREL)
    ns, npnt, nsub = path.counts()
    assert (ns, nsub) == (26, 2)
    pts, kinds, sub, closed = path.export()
    assert list(closed) == [1, 1]
    assert list(sub) == [0, 22, 26]
    # reference builder coordinates (end points of each segment), 4 significant digits like `{:#}`
    ends = [(8.0, 3.92), (8.0, 10.0), (4.0, 3.66), (3.52, 3.0), (3.22, 3.34), (2.66, 3.07), (2.53, 3.65), (0.68, 6.87),
            (3.15, 7.3), (3.62, 8.29), (1.0, 8.0), (1.0, 9.0), (4.0, 10.0), (4.0, 14.0), (3.0, 14.0), (2.0, 15.0),
            (8.0, 15.0), (13.0, 11.53), (12.0, 9.0), (13.0, 7.0), (16.0, 5.0), (12.0, 1.0), (2.0, 5.5), (2.5, 5.0),
            (3.0, 5.5), (2.5, 6.0)]
    offs = np.concatenate([[0], np.cumsum(kinds.astype(np.int64))]).astype(np.int64)
    for i, e in enumerate(ends):
        p = pts[offs[i + 1] - 1]
        assert abs(p[0] - e[0]) < 1e-3 and abs(p[1] - e[1]) < 1e-3, (i, p, e)
    # second control point of the first smooth curve `s-.22.11-.3.34` reflects (3.52,3.0)
    assert np.allclose(pts[offs[4] + 1], (3.52, 3.0), atol=1e-12)

    p = O.OraclePath.parse(" M0,0L1-1L1,0ZL0,1 L1,1Z ")
    pts, kinds, sub, closed = p.export()
    assert list(sub) == [0, 2, 4] and list(closed) == [1, 1]
    assert np.array_equal(pts, np.array([[0, 0], [1, -1], [1, -1], [1, 0], [0, 0], [0, 1], [0, 1], [1, 1]], dtype=float))

    ref = np.array([[0.5, -3.0], [-11.0, -0.11]])
    p1 = O.OraclePath.parse("M.5-3-11-.11").export()[0]
    p2 = O.OraclePath.parse(" m.5,-3 -11.5\n2.89 ").export()[0]
    assert np.allclose(p1, ref, atol=1e-12) and np.allclose(p2, ref, atol=1e-12)


def test_parse_scalar():
    """src/svg.rs:668-677 test_parse_scalar"""
    text = "1 .22e0.32 3.21e-3-1.24 1e4"
    expect = [1.0, 0.22, 0.32, 3.21e-3, -1.24, 1e4]
    pos = 0
    for e in expect:
        v, n = O.parse_scalar(text[pos:])
        approx(v, e, 4 * EPS * max(1.0, abs(e)))
        pos += n


def test_parse_transform():
    """src/svg.rs:680-695 test_parse_transform (expected matrix printed at 6 significant digits)"""
    tr = O.transform_parse("""
            translate(1 2)
            skewX(30deg)
            matrix(1  2 3 4 -3-7)
            scale(2,1)
            rotate(10 1 2)
            rotate(1rad)
    """)
    m00, m01, m02, m10, m11, m12 = tr
    got = [m00, m10, m01, m11, m02, m12]
    exp = [6.56129, 5.23393, -1.92617, -2.14614, -5.23999, -4.1231]
    for g, e in zip(got, exp):
        assert abs(g - e) < 6e-6 * max(1, abs(e)), (got, exp)


def test_flatten_connected():
    """src/path.rs:1192-1211 test_flatten: consecutive lines of a subpath are connected and the
    whole-path flatten equals the per-subpath flatten"""
    path = O.OraclePath.parse(SQUIRREL)
    tr = O.transform_mul(O.rotate(math.pi / 3.0), O.translate(-10.0, -20.0))
    lines = path.flatten(tr, 0.05, True)
    pts, kinds, sub, closed = path.export()
    offs = np.concatenate([[0], np.cumsum(kinds.astype(np.int64))]).astype(np.int64)
    k = 0
    for sp in range(len(closed)):
        s0, s1 = sub[sp], sub[sp + 1]
        spp = O.OraclePath.from_flat(pts[offs[s0]:offs[s1]], kinds[s0:s1], [0, s1 - s0], [closed[sp]])
        sl = spp.flatten(tr, 0.05, True)
        for a, b in zip(sl[:-1], sl[1:]):
            assert abs(a[2] - b[0]) < EPS and abs(a[3] - b[1]) < EPS
        assert np.array_equal(lines[k:k + len(sl)], sl)
        k += len(sl)
    assert k == len(lines)


def test_split_halves():
    """src/curve.rs:1570-1582 test_split: split() halves meet at at(0.5) — checked through flatten of a
    cubic with a threshold that forces exactly one split"""
    p = O.OraclePath.parse("M0,0 C0,10 10,10 10,0")
    lines = p.flatten(O.IDENTITY, 2.0, False)
    # f=16*flat^2: u=(0,30)-(10,0)=(-10,30), v=(30,30)-(0,0)-(20,0)=(10,30): 100+900=1000 >= 64 -> split
    assert len(lines) >= 2
    mid = 0.125 * np.array([0, 0]) + 0.375 * np.array([0, 10]) + 0.375 * np.array([10, 10]) + 0.125 * np.array([10, 0])
    assert any(np.allclose(l[2:], mid, atol=1e-12) for l in lines)


def test_flatten_length():
    """src/curve.rs:1622-1663 test_length: flatten at tol 1e-5 reproduces the arc length (2e-4 relative).
    Reference lengths: quad/cubic of the test, values from the assertions there are recomputed here by
    dense sampling."""
    p = O.OraclePath.parse("M158,70 C210,250 25,190 219,89")
    lines = p.flatten(O.IDENTITY, 1e-5, False)
    length = np.hypot(lines[:, 2] - lines[:, 0], lines[:, 3] - lines[:, 1]).sum()
    t = np.linspace(0, 1, 2_000_001)
    P = np.array([[158, 70], [210, 250], [25, 190], [219, 89]], dtype=float)
    c = ((1 - t) ** 3)[:, None] * P[0] + (3 * (1 - t) ** 2 * t)[:, None] * P[1] + (3 * (1 - t) * t ** 2)[:, None] * P[2] + (t ** 3)[:, None] * P[3]
    ref = np.hypot(*np.diff(c, axis=0).T).sum()
    assert abs(length - ref) / ref < 2e-4


def test_stroke():
    """src/path.rs:1225-1269 test_stroke (expected strings at 4-5 significant digits)"""
    def pts_of(svg):
        return O.OraclePath.parse(svg).export()

    def check(p, ref_svg):
        rp, rk, rs, rc = pts_of(ref_svg)
        pp, pk, psub, pc = p.export()
        assert list(pk) == list(rk), (list(pk), list(rk))
        assert list(psub) == list(rs) and list(pc) == list(rc)
        assert np.abs(pp - rp).max() < 1e-4, np.abs(pp - rp).max()

    path = O.OraclePath.parse("M2,2L8,2C11,2 11,8 8,8L5,4")
    check(path.stroke(1.0), """
        M2,1.5 L8,1.5 C9.80902,1.5 10.75,3.38197 10.75,5 10.75,6.61803 9.80902,8.5 8,8.5 L7.75,8.5 7.6,8.3 4.6,4.3
        5.4,3.7 8.4,7.7 8,7.5 C9.19098,7.5 9.75,6.38197 9.75,5 9.75,3.61803 9.19098,2.5 8,2.5 L2,2.5 2,1.5 Z
    """)
    check(path.stroke(1.0, "round", 4.0, "round"), """
        M2,1.5 L8,1.5 C9.80902,1.5 10.75,3.38197 10.75,5 10.75,6.61803 9.80902,8.5 8,8.5 7.84274,8.5 7.69436,8.42581
        7.6,8.3 L4.6,4.3 C4.43542,4.08057 4.48057,3.76458 4.7,3.6 4.91943,3.43542 5.23542,3.48057 5.4,3.7 L8.4,7.7 8,7.5
        C9.19098,7.5 9.75,6.38197 9.75,5 9.75,3.61803 9.19098,2.5 8,2.5 L2,2.5 C1.72571,2.5 1.5,2.27429 1.5,2
        1.5,1.72571 1.72571,1.5 2,1.5 Z
    """)
    path = O.OraclePath.parse("M2,2L8,2C11,2 11,8 8,8L5,4Z")
    check(path.stroke(1.0, "round", 4.0, "round"), """
        M2,1.5 L8,1.5 C9.80902,1.5 10.75,3.38197 10.75,5 10.75,6.61803 9.80902,8.5 8,8.5 7.84274,8.5 7.69436,8.42581
        7.6,8.3 L4.6,4.3 4.72265,4.41603 1.72265,2.41603 C1.53984,2.29415 1.45778,2.06539 1.52145,1.85511 1.58512,1.64482
        1.78029,1.5 2,1.5 ZM5.4,3.7 L8.4,7.7 8,7.5 C9.19098,7.5 9.75,6.38197 9.75,5 9.75,3.61803 9.19098,2.5 8,2.5
        L2,2.5 2.27735,1.58397 5.27735,3.58397 C5.32451,3.61542 5.36599,3.65465 5.4,3.7 Z
    """)


def test_spread():
    """src/grad.rs:515-524 test_spread"""
    REFLECT, REPEAT = 2, 1
    approx(O.lib().orc_spread_at(REFLECT, 0.3), 0.3, 1e-6)
    approx(O.lib().orc_spread_at(REFLECT, -0.3), 0.3, 1e-6)
    approx(O.lib().orc_spread_at(REFLECT, 1.3), 0.7, 1e-6)
    approx(O.lib().orc_spread_at(REFLECT, -1.3), 0.7, 1e-6)
    approx(O.lib().orc_spread_at(REPEAT, 0.3), 0.3)
    approx(O.lib().orc_spread_at(REPEAT, -0.3), 0.7)


def test_grad_stops():
    """src/grad.rs:527-539 test_grad_stops — via a linear-colour linear gradient along x in [0,1]"""
    stops = [(0.0, (1, 0, 0, 1)), (0.5, (0, 1, 0, 1)), (1.0, (0, 0, 1, 1))]
    g = O.OraclePaint.linear(stops, (0, 0), (1, 0), linear_colors=True)
    assert np.array_equal(g.at(-1.0, 0), np.float32([1, 0, 0, 1]))
    assert np.array_equal(g.at(0.25, 0), np.float32([0.5, 0.5, 0, 1]))
    assert np.array_equal(g.at(0.75, 0), np.float32([0, 0.5, 0.5, 1]))
    assert np.array_equal(g.at(2.0, 0), np.float32([0, 0, 1, 1]))


def test_radial_grad():
    """src/grad.rs:542-567 test_radial_grad"""
    g = O.OraclePaint.radial([], (0.5, 0.0), 0.5, fcenter=(0.25, 0.0), fradius=0.1, units=1, linear_colors=True)
    assert g.radial_offset(0.25, 0.0) < 0.0
    approx(g.radial_offset(0.675, 0.0), 0.5)
    approx(g.radial_offset(1.0, 0.0), 1.0)
    # empty stop list -> single opaque black stop (src/grad.rs:92-97)
    assert np.array_equal(g.at(0.6, 0.0), np.float32([0, 0, 0, 1]))


def test_lin_grad():
    """src/grad.rs:570-602 test_ling_grad"""
    c0, c1, c2 = O.parse_color("#89155180"), O.parse_color("#ff272d80"), O.parse_color("#ff272d00")
    g = O.OraclePaint.linear([(0.0, c0), (0.5, c1), (1.0, c2)], (0, 0), (1, 1), linear_colors=True)
    assert np.array_equal(g.at(-0.5, -0.5), c0)
    assert np.array_equal(g.at(0.0, 0.0), c0)
    assert np.array_equal(g.at(1.0, 0.0), c1)
    assert np.array_equal(g.at(0.0, 1.0), c1)
    assert np.array_equal(g.at(1.0, 1.0), c2)
    assert np.array_equal(g.at(1.5, 1.5), c2)


def test_lin_and_srgb():
    """src/color.rs:547-554 test_lin_and_srgb (scalar functions)"""
    L = O.lib()
    for i in range(255):
        v = np.float32(i / 255.0)
        approx(float(v), L.orc_linear_to_srgb(L.orc_srgb_to_linear(v)), 1e-4)
        approx(float(v), L.orc_srgb_to_linear(L.orc_linear_to_srgb(v)), 1e-4)


def test_conversion():
    """src/color.rs:539-545 test_conversion: RGBA -> LinColor -> RGBA identity for #ff804010"""
    c = np.array([0xff, 0x80, 0x40, 0x10], dtype=np.uint8)
    for simd in (True, False):
        O.set_simd_x86(simd)
        try:
            assert np.array_equal(O.lin_to_rgba(O.rgba_to_lin(c)), c)
        finally:
            O.set_simd_x86(True)


def test_color_parse():
    """src/color.rs:524-536 test_color_parse (hex forms and /alpha suffix)"""
    assert np.array_equal(O.lin_to_rgba(O.parse_color("#01020304")), [1, 2, 3, 4])
    assert np.array_equal(O.lin_to_rgba(O.parse_color("#aabbcc")), [170, 187, 204, 255])
    assert np.array_equal(O.lin_to_rgba(O.parse_color("#000000")), [0, 0, 0, 255])
    assert O.parse_color("#010203/.25")[3] == np.float32(63 / 255.0)


def test_simd_constants():
    """x86 s2l(1.0) = 1.0008736, l2s(1.0) = 0.99996865 on all four lanes"""
    one = np.ones(4, dtype=np.float32)
    assert abs(float(O.s2l(one)[3]) - 1.0008736) < 2e-7
    assert abs(float(O.l2s(one)[3]) - 0.99996865) < 2e-7
    O.set_simd_x86(False)
    try:
        assert O.s2l(one)[3] == 1.0 and O.l2s(one)[3] == 1.0
    finally:
        O.set_simd_x86(True)


SCENE = r"""
{
    "type": "transform",
    "tr": "translate(7, 7) rotate(45, 7, 7) scale(10)",
    "child": { "type": "fill", "paint": "#ff8040", "path": "M0,0 h1 v1 h-1 z" }
}
"""


def test_scene_view():
    """src/scene.rs:669-693 test_scene_view"""
    sc = O.OracleScene.load_json(SCENE)
    x, y, img = sc.render(O.IDENTITY, view=(1.0, 2.0, 23.0, 23.0))
    assert (x, y, img.shape[1], img.shape[0]) == (1, 2, 22, 21)
    x, y, img = sc.render(O.IDENTITY)
    assert (x, y, img.shape[1], img.shape[0]) == (6, 4, 16, 15)
    assert img[..., 3].max() > 0.99


def test_lcg():
    """LCG of benches/scene_bench.rs:53-88, first three uniforms from seed 0"""
    s, vals = 0, []
    for _ in range(3):
        v, s = O.lcg_uniform(s)
        vals.append(v)
    assert vals == [0.0011632625591320167, 0.2702386477625217, 0.31892294627116813]


def test_glyph_shape():
    g = O.OraclePath.glyph(1)
    assert g.counts() == (18, 72, 3)
    pts = g.export()[0]
    assert pts.min() >= 4.0 and pts.max() <= 60.0


def test_mask_iter_matches_mask():
    """mask_iter (src/rasterize.rs:313-355) agrees with mask (:299-311) away from the overflow column"""
    path = O.OraclePath.parse(SQUIRREL)
    (w, h), tr, _ = path.size(O.scale(6.0, 6.0))
    img = np.zeros((h, w))
    path.mask(tr, O.NONZERO, img)
    px = path.mask_iter(tr, w, h, O.NONZERO)
    dense = np.zeros((h, w))
    for x, y, a in px:
        dense[y, x] = a
    assert np.abs(dense - img).max() < 1e-12


def test_mask_threads_matches_mask():
    path = O.OraclePath.parse(SQUIRREL)
    (w, h), tr, _ = path.size(O.scale(10.0, 10.0))
    a = np.zeros((h, w))
    b = np.zeros((h, w))
    path.mask(tr, O.EVENODD, a)
    path.mask_threads(tr, O.EVENODD, b, threads=5)
    assert np.abs(a - b).max() < 1e-12


def test_nan_panics():
    """src/path.rs:765-767: flattening a segment with NaN panics -> error"""
    p = O.OraclePath.from_flat([[0, 0], [float("nan"), 1]], [2], [0, 1], [0])
    with pytest.raises(ValueError):
        p.flatten()


def test_right_edge_wrap_quirk():
    """A reference DEFECT the oracle reproduces and the GPU path deliberately does not.

    `signed_difference_line` cuts a line that crosses x == width at that column and keeps the inside half, choosing it with
    `p0.x() < width` (src/rasterize.rs:377-383).  When p0.x == width exactly and the line continues to the right, the test
    is false and the OUTSIDE half is kept; its cells land at `row_offset + x0i` with x0i > width, i.e. in the first columns
    of the NEXT row (src/rasterize.rs:437-444), or out of bounds (a panic) on the last row.  Expected values traced by hand
    from the reference for M0,0 L5,0 L7,4 L0,4 Z on a 6 x 6 image (width - 1 == 5): the geometric answer is 1.0 in rows 0..3.
    """
    p = O.OraclePath.parse(b"M0,0 L5,0 L7,4 L0,4 Z")
    img = np.zeros((6, 6))
    p.mask(O.IDENTITY, O.NONZERO, img)
    expected = np.array([
        [1.0, 1.0, 1.0, 1.0, 1.0, 0.25],     # col 0: -1 (left edge), col 5: +0.75
        [0.75, 0.75, 0.75, 0.75, 0.75, 0.5],  # col 0: -1 + 0.25 wrapped from row 0
        [0.25, 0.25, 0.25, 0.25, 0.25, 0.25],  # col 0: -1 + 0.75 wrapped from row 1
        [0.25, 0.0, 0.0, 0.0, 0.0, 0.0],      # row 2's cells (x0i = 6) land here: col 0 +0.75, col 1 +0.25
        [0.25, 1.0, 1.0, 1.0, 1.0, 1.0],      # row 3's cells land in row 4
        [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    ])
    np.testing.assert_allclose(img, expected, atol=1e-12)


def test_batch_threads_helper_runs():
    """bench.py's multi-threaded CPU arm for glyph batches (timing helper: it renders into private scratch images)."""
    glyphs = [O.OraclePath.glyph(i + 1) for i in range(5)]
    O.batch_threads(glyphs, O.IDENTITY, O.NONZERO, None, 64, 64, 2)
    O.batch_threads(glyphs, O.IDENTITY, O.EVENODD, O.OraclePaint.solid([0.0, 0.0, 0.0, 1.0]), 64, 64, 3)
