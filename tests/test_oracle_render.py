"""CPU: the oracle integrator -- regression anchor, closure consistency (sample vs evaluate), energy
conservation (white furnace), and the reference's accumulation arithmetic."""
import ctypes as C
import dataclasses
import json
import os
import sys

import numpy as np
import pytest

from gatling_b200 import scenes
from oracle import orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_cornell_regression_anchor():
    d = scenes.cornell(32, 32, spp=4, max_bounces=4)
    img, _, rays = orc.OracleScene(d).render()
    want = np.load(os.path.join(GOLDEN, "cornell_oracle_32.npy"))
    with open(os.path.join(GOLDEN, "cornell_oracle_32.json")) as f:
        meta = json.load(f)
    assert rays[0] == meta["closest_rays"] and rays[1] == meta["shadow_rays"]
    assert np.allclose(img, want, rtol=1e-5, atol=1e-6)


def test_feature_regression_anchors():
    """Textures, primvars, participating media, NEE with every light type, OpenPBR: today's oracle against the committed images."""
    sys.path.insert(0, GOLDEN)
    import make_golden
    want = np.load(os.path.join(GOLDEN, "feature_anchors.npz"))
    for name, d in make_golden.feature_scenes().items():
        img, _, rays = orc.OracleScene(d).render()
        assert list(rays) == list(want[name + "_rays"]), name
        assert np.allclose(img, want[name], rtol=1e-5, atol=1e-6), name


def test_empty_stage_1x1():
    """The reference's Render.Empty1x1 case (src/hdGatling/main.cpp:584): an empty stage at 1x1 is the colour AOV's clear
    value after its trip through the RGBA8 fallback dome texture (Gi.cpp:2183-2199)."""
    e = scenes.cornell(1, 1, spp=4)
    e.meshes = []
    e.clear_color = (0.25, 0.5, 1.0, 1.0)
    img, _, rays = orc.OracleScene(e).render()
    assert img.shape == (1, 1, 4)
    assert np.allclose(img[0, 0], [63 / 255, 127 / 255, 1.0, 1.0], atol=1e-6)
    assert rays[0] == 4 and rays[1] == 0            # one primary ray per sample, all of them escape


def test_progressive_accumulation_matches_single_pass():
    d = scenes.cornell(24, 24, spp=8, max_bounces=3)
    o = orc.OracleScene(d)
    full, _, _ = o.render()
    a, _, _ = o.render(spp=4)
    b, _, _ = o.render(sample_offset=4, prev=a, spp=4)         # rp_main.rgen:506-515
    assert np.allclose(b, full, rtol=1e-5, atol=1e-6)


def test_shards_partition_the_image():
    d = scenes.cornell(40, 24, spp=2, max_bounces=3)
    o = orc.OracleScene(d)
    full, _, _ = o.render()
    acc = np.zeros_like(full)
    for r in range(4):
        part, _, _ = o.render(shard=(r, 4, 8))
        assert ((part[..., 3] == 0) | (part[..., 3] == 1)).all()
        acc += part
    assert np.array_equal(acc, full)


def _mat(m):
    md = orc.MaterialDesc()
    md.closure = m.closure
    md.params = (C.c_float * 48)(*[float(x) for x in m.params])
    return md


MATERIALS = [
    ("diffuse", scenes.material_diffuse((0.8, 0.6, 0.4))),
    ("ups_default", scenes.material_usd_preview_surface(diffuse=(0.8, 0.8, 0.8))),
    ("ups_metal_coat", scenes.material_usd_preview_surface(diffuse=(0.9, 0.6, 0.2), metallic=1.0, roughness=0.3, clearcoat=1.0, clearcoatRoughness=0.1)),
    ("openpbr_default", scenes.material_openpbr()),
    ("openpbr_half_metal_tinted_coat", scenes.material_openpbr(base_color=(0.7, 0.3, 0.2), base_metalness=0.5, specular_roughness=0.5, specular_color=(0.9, 0.8, 0.7), coat_weight=0.6, coat_roughness=0.35, coat_color=(0.8, 0.9, 1.0), subsurface_weight=0.3)),
    ("openpbr_metal_coat", scenes.material_openpbr(base_color=(0.9, 0.6, 0.3), base_metalness=1.0, specular_roughness=0.4, coat_weight=1.0, coat_roughness=0.3)),
]


@pytest.mark.parametrize("name,mat", MATERIALS)
def test_closure_sample_vs_evaluate(name, mat):
    """E[bsdf_over_pdf] over xi must equal the integral of evaluate()'s (diffuse+glossy) over the hemisphere,
    and stay <= 1 per channel (energy conservation); tested for reflection-only parameter sets."""
    L = orc.lib()
    md = _mat(mat)
    state = (C.c_float * 9)(0, 0, 1, 0, 0, 1, 1, 0, 0)
    rng = np.random.default_rng(1)
    for k1 in [(0.0, 0.0, 1.0), (0.6, 0.0, 0.8), (0.0, -0.8, 0.6)]:
        n = 60000
        acc = np.zeros(3)
        k2 = (C.c_float * 3)(); w = (C.c_float * 3)(); ev = C.c_int32()
        xis = rng.random((n, 4)).astype(np.float32)
        for i in range(n):
            L.orc_bsdf_sample(C.byref(md), state, (C.c_float * 3)(*k1), (C.c_float * 4)(*xis[i]), k2, w, C.byref(ev))
            if ev.value != 0:
                acc += (w[0], w[1], w[2])
        mc_sample = acc / n
        # uniform-hemisphere quadrature of evaluate()
        m = 60000
        u = rng.random((m, 2))
        z = u[:, 0]; r = np.sqrt(1 - z * z); phi = 2 * np.pi * u[:, 1]
        dirs = np.stack([r * np.cos(phi), r * np.sin(phi), z], 1).astype(np.float32)
        tot = np.zeros(3)
        dif = (C.c_float * 3)(); glo = (C.c_float * 3)(); pdf = C.c_float()
        for i in range(m):
            L.orc_bsdf_evaluate(C.byref(md), state, (C.c_float * 3)(*k1), (C.c_float * 3)(*dirs[i]), dif, glo, C.byref(pdf))
            tot += (dif[0] + glo[0], dif[1] + glo[1], dif[2] + glo[2])
        mc_eval = tot / m * 2 * np.pi
        assert np.all(mc_sample <= 1.0 + 0.02), (name, k1, mc_sample)
        assert np.allclose(mc_sample, mc_eval, rtol=0.06, atol=0.01), (name, k1, mc_sample, mc_eval)


def test_white_furnace():
    """A closed diffuse-white (tint 1) room lit only by the white background seen through nothing: every
    path that escapes returns exactly 1, paths cut at max-bounces return 0 -> pixel = P(escape) <= 1; an
    open box under a white background must therefore never exceed 1 (no energy gain anywhere)."""
    d = scenes.cornell(24, 24, spp=16, max_bounces=12)
    d.materials = [scenes.material_diffuse((1.0, 1.0, 1.0)) for _ in d.materials]
    d.settings = dataclasses.replace(d.settings, rrBounceOffset=100)   # Russian roulette reweights by 1/p > 1 per sample
    img, _, _ = orc.OracleScene(d).render()
    assert img[..., :3].max() <= 1.0 + 1e-4
    assert img[..., :3].mean() > 0.5
