"""CPU: the oracle's restatement of the reference's texture lookup (mdl_interface.glsl:8-38,127-145; sampler linear /
Repeat) against closed-form expectations, and a textured render through the oracle integrator."""
import dataclasses

import numpy as np

from gatling_b200 import scenes
from oracle import orc


def _grid(w=4, h=3):
    t = np.zeros((h, w, 4), np.float32)
    for y in range(h):
        for x in range(w):
            t[y, x] = (x, y, 10 * y + x, 1.0)
    return scenes.Texture(t)


def test_texel_centres_and_bilinear_weights():
    t = _grid()
    for y in range(3):
        for x in range(4):
            c, clipped = orc.tex_lookup(t, (x + 0.5) / 4, (y + 0.5) / 3)
            assert not clipped and np.array_equal(c, t.texels[y, x])
    # half-way between texels (0,0) and (1,0) / (0,1): exact averages
    c, _ = orc.tex_lookup(t, 1.0 / 4, 0.5 / 3)
    assert np.allclose(c, (t.texels[0, 0] + t.texels[0, 1]) / 2, atol=1e-6)
    c, _ = orc.tex_lookup(t, 0.5 / 4, 1.0 / 3)
    assert np.allclose(c, (t.texels[0, 0] + t.texels[1, 0]) / 2, atol=1e-6)


def test_wrap_modes():
    t = _grid()
    R, CL, MI, CP = scenes.WRAP_REPEAT, scenes.WRAP_CLAMP, scenes.WRAP_MIRRORED_REPEAT, scenes.WRAP_CLIP
    centre = lambda x, y: ((x + 0.5) / 4, (y + 0.5) / 3)
    u, v = centre(1, 2)
    assert np.array_equal(orc.tex_lookup(t, u + 3.0, v - 2.0, R, R)[0], t.texels[2, 1])            # repeat: integer shifts are invisible
    assert np.array_equal(orc.tex_lookup(t, 7.3, v, CL, R)[0], t.texels[2, 3])                     # clamp: to the last texel centre
    assert np.array_equal(orc.tex_lookup(t, -0.4, v, CL, R)[0], t.texels[2, 0])
    assert np.array_equal(orc.tex_lookup(t, 1.0 + (1.0 - u), v, MI, R)[0], t.texels[2, 1])         # mirrored: 1..2 runs backwards
    assert np.array_equal(orc.tex_lookup(t, 2.0 + u, v, MI, R)[0], t.texels[2, 1])
    c, clipped = orc.tex_lookup(t, 1.2, v, CP, R)
    assert clipped and not c.any()                                                                # clip: zero + flag
    c, clipped = orc.tex_lookup(t, u, -0.01, R, CP)
    assert clipped
    assert not orc.tex_lookup(t, 1.0, 1.0, CP, CP)[1]                                              # the interval is closed (glsl :130-131)
    # repeat addressing of the bilinear footprint: u = 0 blends the last and the first column
    c, _ = orc.tex_lookup(t, 0.0, v, R, R)
    assert np.allclose(c, (t.texels[2, 3] + t.texels[2, 0]) / 2, atol=1e-6)


def test_formats_and_srgb():
    rgba8 = scenes.Texture(np.array([[[0, 128, 255, 64]]], np.uint8))
    assert np.allclose(orc.tex_lookup(rgba8, 0.5, 0.5)[0], np.array([0, 128, 255, 64], np.float32) / 255)
    half = scenes.Texture(np.array([[[0.5, -2.0, 6.1e-5, 1000.0]]], np.float16))
    assert np.array_equal(orc.tex_lookup(half, 0.5, 0.5)[0], half.texels[0, 0].astype(np.float32))
    r32 = scenes.Texture(np.array([[[0.25]]], np.float32))
    assert np.array_equal(orc.tex_lookup(r32, 0.5, 0.5)[0], np.array([0.25, 0, 0, 1], np.float32))
    # MaterialX srgb_texture: linear segment below the break point, 2.4 power above; continuous, fixed points 0 and 1
    assert orc.srgb_texture_to_linear(0.0) == 0.0
    assert abs(orc.srgb_texture_to_linear(1.0) - 1.0) < 1e-6
    assert abs(orc.srgb_texture_to_linear(0.02) - 0.02 * 0.07738015800714493) < 1e-9
    assert abs(orc.srgb_texture_to_linear(0.5) - ((0.5 + 0.055) / 1.055) ** 2.4) < 1e-5
    b = 0.03928571566939354
    assert abs(orc.srgb_texture_to_linear(b) - orc.srgb_texture_to_linear(b + 1e-6)) < 1e-5


def test_textured_render_changes_the_albedo_and_the_image():
    d = scenes.textured_cornell(48, 48, spp=4)
    plain = scenes.cornell(48, 48, spp=4)
    plain.settings = d.settings
    plain.rect_lights = d.rect_lights
    for m in plain.meshes:
        scenes.assign_planar_uv(m)
    o, p = orc.OracleScene(d), orc.OracleScene(plain)
    img = o.render()[0]
    ref = p.render()[0]
    assert np.isfinite(img).all() and np.abs(img - ref).mean() > 1e-2
    # the floor's checker shows up in the albedo AOV: two clearly separated populations of red values
    alb = o.render_aovs([16])[16]
    floor = alb[2:14, 8:40, 0]
    assert floor.std() > 0.1
    # the unmapped Normal AOV is untouched by the normal map (rp_main.chit:147,252 use the copy taken before *_init)
    assert np.array_equal(o.render_aovs([1])[1], p.render_aovs([1])[1])
    # determinism
    assert np.array_equal(img, orc.OracleScene(d).render()[0])
