"""GPU: randomised parity sweep.  Seeded random paths (lines, quads, cubics; open and closed subpaths; coordinates that
overhang every edge of the canvas), random affine transforms and canvas sizes on both sides of the tile boundaries
(64 / 128 / 1024 / 2048 columns, 8 / 64 rows), both fill rules:

  * `Path::flatten` lines bit-identical to the oracle's (count and end points),
  * `Rasterizer::mask` within 1e-4 per pixel,
  * `mask_iter` coverage (dense form) within 1e-4.

Coordinates are quantised to 1/64 + an irrational offset so that no flattened end point sits exactly on the last column
(the reference's right-edge wrap defect, tests/test_oracle_kat.py::test_right_edge_wrap_quirk)."""
import numpy as np
import pytest

import oracle as O
import rasterize_b200 as rb
from helpers import opath

pytestmark = pytest.mark.gpu

COV_TOL = 1e-4
SIZES = [(7, 5), (63, 64), (65, 33), (129, 70), (500, 8), (1023, 17), (1025, 9), (2050, 24), (3000, 40), (96, 1000)]


@pytest.fixture(scope="module")
def rast():
    r = rb.GpuRasterizer()
    yield r
    r.close()


def random_path(rng, w, h, n_sub, n_seg):
    b = rb.Path.builder()

    def pt():
        # up to 30 % beyond every edge
        return (float(rng.uniform(-0.3 * w, 1.3 * w)) + 0.0137, float(rng.uniform(-0.3 * h, 1.3 * h)) + 0.0071)

    for _ in range(n_sub):
        b.move_to(pt())
        for _ in range(n_seg):
            k = rng.integers(0, 3)
            if k == 0:
                b.line_to(pt())
            elif k == 1:
                b.quad_to(pt(), pt())
            else:
                b.cubic_to(pt(), pt(), pt())
        if rng.random() < 0.7:
            b.close()
    return b.build()


@pytest.mark.parametrize("case", range(len(SIZES)))
def test_random_paths_match_oracle(rast, case):
    w, h = SIZES[case]
    rng = np.random.default_rng(1000 + case)
    for rep in range(8):
        p = random_path(rng, w, h, n_sub=int(rng.integers(1, 5)), n_seg=int(rng.integers(1, 9)))
        a = float(rng.uniform(-0.4, 0.4))
        s = float(rng.uniform(0.6, 1.4))
        tr = np.array([s * np.cos(a), -s * np.sin(a), float(rng.uniform(-0.1, 0.1)) * w,
                       s * np.sin(a), s * np.cos(a), float(rng.uniform(-0.1, 0.1)) * h])
        op = opath(p)
        lines = rast.flatten(p, tr, True)
        olines = op.flatten(tr)
        assert lines.shape == olines.shape and np.array_equal(lines, olines), (case, rep)
        for rule, orule in ((rb.FillRule.NonZero, O.NONZERO), (rb.FillRule.EvenOdd, O.EVENODD)):
            img = np.zeros((h, w))
            rast.mask(p, tr, img, rule)
            ref = np.zeros((h, w))
            op.mask(tr, orule, ref)
            assert np.abs(img - ref).max() <= COV_TOL, (case, rep, rule, float(np.abs(img - ref).max()))
        cov = rast.coverage(p, tr, rb.Size(w, h), rb.FillRule.NonZero)
        ocov = np.zeros((h, w))
        for px in op.mask_iter(tr, w, h, O.NONZERO):
            ocov[px[1], px[0]] = px[2]
        assert np.abs(cov - ocov).max() <= COV_TOL, (case, rep)
