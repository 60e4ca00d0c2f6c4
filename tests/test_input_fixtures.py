"""The input fixtures (configs/*.yaml, adaptive-sph_b200/data/split-patterns.yaml) carry the reference's parameter values in
this repository's own layout (tools/make_input_fixtures.py).  In the build container, where the reference is mounted, every
value is compared with the reference's file; everywhere, both loaders must read the fixtures."""
import os

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAIRS = [("configs/default-config.yaml", "default-config.yaml"), ("configs/default-config-web.yaml", "default-config-web.yaml"),
         ("configs/default-scene.yaml", "default-scene.yaml"), ("configs/default-scene-web.yaml", "default-scene-web.yaml"),
         ("configs/motivation-scene2.yaml", "media/motivation-scene2.yaml"),
         ("configs/ratio-stress-test-scene.yaml", "media/ratio-stress-test-scene.yaml"),
         ("adaptive-sph_b200/data/split-patterns.yaml", "split-patterns.yaml")]


def _load(path):
    with open(path) as f:
        return yaml.load(f, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))


def _equal(a, b):
    if isinstance(a, dict):
        return isinstance(b, dict) and set(a) == set(b) and all(_equal(a[k], b[k]) for k in a)
    if isinstance(a, list):
        return isinstance(b, list) and len(a) == len(b) and all(_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, bool) or isinstance(b, bool) or a is None or b is None or isinstance(a, str) or isinstance(b, str):
        return a == b
    return float(a) == float(b)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("ours,theirs", PAIRS)
def test_fixture_values_equal_the_reference(ours, theirs):
    assert _equal(_load(os.path.join(ROOT, ours)), _load(os.path.join(REF, theirs)))


def test_fixtures_load(asph):
    for name in ("default-config.yaml", "default-config-web.yaml"):
        p = asph.SimulationParams.from_yaml(os.path.join(ROOT, "configs", name))
        assert p["pressure_solver_method"] == "HybridDFSPH" and float(p["sdf_gradient_eps"]) == 1e-5
    assert asph.scene_particle_count(asph.SceneConfig.from_yaml(os.path.join(ROOT, "configs", "default-scene.yaml"))) == 1035
    sp = asph.load_split_patterns_from_file()
    assert sp.max_children == 59 and sp.pos.shape == (sum(range(2, 60)), 2) and sp.pos.dtype == np.float32
