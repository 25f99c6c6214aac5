"""CPU: scene generators (SURVEY.md 8d inputs)."""
import hashlib
import json
import os

import numpy as np
import pytest

from gatling_b200 import scenes
from oracle import orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _hashes(desc):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.scene_hashes(desc)


def test_cornell_matches_reference_asset_fixture():
    """scenes.cornell() == the reference's cornell.usda as frozen in tests/golden/scene_hashes.json."""
    with open(os.path.join(GOLDEN, "scene_hashes.json")) as f:
        want = json.load(f)["hashes"]
    got = _hashes(scenes.cornell())
    assert got == want


@pytest.mark.skipif(not os.path.exists("/root/reference/cornell.usda"), reason="reference checkout not present (GPU box)")
def test_cornell_matches_reference_asset_live():
    from gatling_b200 import usda
    a = usda.load("/root/reference/cornell.usda")
    assert _hashes(a) == _hashes(scenes.cornell())
    assert a.camera == scenes.cornell().camera


def test_cornell_counts():
    d = scenes.cornell()
    assert len(d.meshes) == 8 and d.triangle_count == 46
    assert sum(len(m.vertices) for m in d.meshes) == 138                 # faceVarying normals force de-indexing
    assert abs(d.camera.vfov - 0.3995) < 1e-3
    assert d.settings.nextEventEstimation is False and d.settings.maxBounces == 4 and d.settings.spp == 64


def test_pcg32_stream_matches_sequential_hash():
    import ctypes as C
    L = orc.lib()
    seed = 0x2545F491
    v = scenes.pcg32_stream(seed, 64)
    s = C.c_uint32(seed)
    for i in range(64):
        want = L.orc_hash_pcg32(C.byref(s))      # advances the LCG state in place, returns the permuted word
        assert int(v[i]) == want
    u = scenes.pcg32_uniform(seed, 1000)
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.05


def test_random_triangles_definition():
    d = scenes.random_triangles(1000, 64, 36, spp=2)
    m = d.meshes[0]
    assert len(m.faces) == 1000 and len(m.vertices) == 3000 and d.triangle_count == 1000
    p = m.vertices["pos"].reshape(1000, 3, 3)
    assert np.abs(p[:, 0]).max() <= 1.0 and np.abs(p[:, 1] - p[:, 0]).max() <= 0.02 + 1e-6
    n = m.vertices["norm"].reshape(1000, 3, 3)
    assert np.allclose(np.linalg.norm(n[:, 0], axis=1), 1.0, atol=1e-5)
    assert np.array_equal(n[:, 0], n[:, 1])
    t = m.vertices["tangent"]
    assert np.abs(np.einsum("ij,ij->i", t, m.vertices["norm"])).max() < 1e-5
    # deterministic
    d2 = scenes.random_triangles(1000, 64, 36, spp=2)
    assert d2.meshes[0].vertices.tobytes() == m.vertices.tobytes()
