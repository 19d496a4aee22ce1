"""CPU checks of the committed fixtures (tests/golden/) against the oracle: the flat assets reproduce the
known answers in expected.json, and the Fill-job form of each scene composites to exactly what the oracle's own
`Scene::render` produced when the fixture was generated."""
import hashlib

import numpy as np
import pytest

import oracle as O
from helpers import opath, render_scene_oracle
import assets


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["squirrel", "tv", "rust", "ava", "huyak", "tv_stroked", "squirrel_stroked", "material"])
def test_path_fixture_known_answers(name):
    p = assets.load_path(name)
    e = assets.expected()["paths"][name]
    op = opath(p)
    assert list(op.counts()) == e["counts"]
    assert np.allclose(op.bbox(), e["bbox"], rtol=0, atol=0)
    (w, h), tr, _ = op.size()
    assert [w, h] == e["size"]
    lines = op.flatten(tr)
    assert len(lines) == e["lines_at_size"] and digest(lines) == e["lines_digest"]
    if "mask_sum_nonzero" in e and name != "material":
        img = np.zeros((h, w))
        op.mask(tr, O.NONZERO, img)
        assert abs(img.sum() - e["mask_sum_nonzero"]) < 1e-9 * max(1.0, e["mask_sum_nonzero"])


def test_survey_counts():
    """line counts of the BASELINE configs"""
    ex = assets.expected()["paths"]
    assert ex["squirrel"]["c1"]["lines"] == 666 and ex["squirrel"]["c1"]["size"] == [512, 453]
    assert ex["material"]["c2"]["lines"] == 109691
    assert ex["tv_stroked"]["c5"]["lines"] == 14118 and ex["tv_stroked"]["counts"][0] == 208
    assert ex["squirrel"]["lines_at_size"] == 300 and ex["material"]["lines_at_size"] == 70621
    assert assets.expected()["scenes"]["firefox_2048"]["lines"] == 7899


@pytest.mark.parametrize("name", ["squirrel_cli_512", "linear_colors", "firefox_256", "many_circles_64"])
def test_scene_fixture_matches_scene_render(name):
    sc = assets.load_scene(name)
    img = render_scene_oracle(sc)
    e = assets.expected()["scenes"][name]
    assert len(sc.fills) == e["n_jobs"]
    assert digest(O.lin_to_rgba(img)) == e["rgba_digest"]


def test_glyph_generator_matches_oracle():
    """bench.py's Python glyph generator == the oracle's"""
    import rasterize_b200 as rb
    import bench
    for seed, e in assets.expected()["glyphs"].items():
        g = bench.glyph_path(rb, int(seed))
        og = O.OraclePath.glyph(int(seed))
        assert np.array_equal(g.points, og.export()[0])
        assert digest(g.points) == e["points_digest"]
        assert len(og.flatten()) == e["lines"]
