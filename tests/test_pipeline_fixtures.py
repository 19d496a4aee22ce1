"""CPU: the pipeline fixtures (tests/golden/pipelines) and the fixture-level oracle composition used by the GPU parity tests
reproduce the oracle's own `Scene::render` of the original scene text (digests recorded by make_golden.py)."""
import hashlib

import numpy as np
import pytest

import oracle as O
from helpers import render_pipeline_oracle
import assets

NAMES = ["grad", "grad_1024", "nested", "nested_900", "firefox_512"]


@pytest.mark.parametrize("name", NAMES)
def test_pipeline_fixture_matches_scene_render(name):
    pl = assets.load_pipeline(name)
    e = assets.expected()["pipelines"][name]
    assert len(pl.nodes) == e["n_nodes"]
    assert [n.kind for n in pl.nodes] == e["kinds"]
    x, y, img = render_pipeline_oracle(pl)
    assert [x, y, img.shape[1], img.shape[0]] == e["layer"]
    rgba = O.lin_to_rgba(img)
    assert hashlib.sha256(np.ascontiguousarray(rgba).tobytes()).hexdigest() == e["rgba_digest"]
    np.testing.assert_allclose(img.reshape(-1, 4).sum(0, dtype=np.float64), e["lin_sum"], rtol=1e-9)


def test_node_kinds_cover_every_arm():
    kinds = set(k for name in NAMES for k in assets.expected()["pipelines"][name]["kinds"])
    assert kinds == {0, 1, 2, 3}  # Fill, Group, Opacity, Clip
