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


# ---- the other MDL runtime lookups: tex_lookup_float4_3d (:45-64), tex_texel_float4_2d / _3d (:167-185, :91-110) -------------
def _volume(w=4, h=3, d=2):
    t = np.zeros((d, h, w, 4), np.float32)
    for z in range(d):
        for y in range(h):
            for x in range(w):
                t[z, y, x] = (x, y, z, 100 * z + 10 * y + x)
    return scenes.Texture(t, depth=d)


def test_3d_lookup_texel_centres_and_trilinear_weights():
    t = _volume()
    R = scenes.WRAP_REPEAT
    for z in range(2):
        for y in range(3):
            for x in range(4):
                c, clipped = orc.tex_lookup_3d(t, (x + 0.5) / 4, (y + 0.5) / 3, (z + 0.5) / 2, R, R, R)
                assert not clipped and np.array_equal(c, t.texels[z, y, x])
    c, _ = orc.tex_lookup_3d(t, 0.5 / 4, 0.5 / 3, 1.0 / 2, R, R, R)              # half-way between the two slices
    assert np.allclose(c, (t.texels[0, 0, 0] + t.texels[1, 0, 0]) / 2, atol=1e-6)
    c, _ = orc.tex_lookup_3d(t, 1.0 / 4, 1.0 / 3, 1.0 / 2, R, R, R)              # the centre of a 2x2x2 cell: the mean of its corners
    assert np.allclose(c, t.texels[0:2, 0:2, 0:2].reshape(-1, 4).mean(0), atol=1e-5)
    # independent numpy restatement at random coordinates (repeat addressing on all three axes)
    rng = np.random.default_rng(4)
    tex = t.texels.astype(np.float64)
    for u, v, w in rng.uniform(-1.5, 2.5, (64, 3)):
        c, _ = orc.tex_lookup_3d(t, u, v, w, R, R, R)
        fu, fv, fw = (np.float32(u) - np.floor(np.float32(u))), (np.float32(v) - np.floor(np.float32(v))), (np.float32(w) - np.floor(np.float32(w)))
        x, y, z = fu * 4 - 0.5, fv * 3 - 0.5, fw * 2 - 0.5
        x0, y0, z0 = int(np.floor(x)), int(np.floor(y)), int(np.floor(z))
        ax, ay, az = x - x0, y - y0, z - z0
        want = np.zeros(4)
        for dz, wz in ((0, 1 - az), (1, az)):
            for dy, wy in ((0, 1 - ay), (1, ay)):
                for dx, wx in ((0, 1 - ax), (1, ax)):
                    want += wz * wy * wx * tex[(z0 + dz) % 2, (y0 + dy) % 3, (x0 + dx) % 4]
        assert np.allclose(c, want, atol=2e-4), (u, v, w)


def test_3d_wrap_and_clip_along_w():
    t = _volume()
    R, CL, CP = scenes.WRAP_REPEAT, scenes.WRAP_CLAMP, scenes.WRAP_CLIP
    u, v = (1 + 0.5) / 4, (2 + 0.5) / 3
    assert np.array_equal(orc.tex_lookup_3d(t, u, v, 5.0, R, R, CL)[0], t.texels[1, 2, 1])          # clamp: to the last slice centre
    assert np.array_equal(orc.tex_lookup_3d(t, u, v, -3.0, R, R, CL)[0], t.texels[0, 2, 1])
    c, clipped = orc.tex_lookup_3d(t, u, v, 1.01, R, R, CP)
    assert clipped and not c.any()
    assert not orc.tex_lookup_3d(t, u, v, 1.0, R, R, CP)[1]                                        # closed interval (:49)


def test_texel_fetches_return_zero_outside():
    t2 = _grid()
    assert np.array_equal(orc.tex_texel(t2, 3, 2), t2.texels[2, 3])
    for x, y in ((-1, 0), (4, 0), (0, -1), (0, 3)):
        assert not orc.tex_texel(t2, x, y).any()                                                  # :177-180
    t3 = _volume()
    assert np.array_equal(orc.tex_texel(t3, 2, 1, 1), t3.texels[1, 1, 2])
    for x, y, z in ((0, 0, -1), (0, 0, 2), (4, 0, 0), (0, 3, 1)):
        assert not orc.tex_texel(t3, x, y, z).any()                                               # :101-104
    r32 = scenes.Texture(np.array([[[0.25]]], np.float32))
    assert np.array_equal(orc.tex_texel(r32, 0, 0), np.array([0.25, 0, 0, 1], np.float32))


def test_textured_cutout_follows_the_mask():
    """rp_main.ahit:35-61 with an opacity that is a texture: a quad whose cutout input reads a two-valued mask lets exactly
    the rays through that land on the zero texels (opacity 0 -> always ignored, 1 -> never), for closest AND any-hit rays."""
    d = scenes.cornell(48, 48, spp=1)
    quad = [m for m in d.meshes if m.name == "BackPlane"][0]
    scenes.assign_planar_uv(quad)
    v = quad.vertices.copy(); v["u"] = (v["u"] + 0.25) / 1.5; v["v"] = (v["v"] + 0.25) / 1.5; quad.vertices = v      # uv = [0, 1]^2 over the quad
    mask = np.zeros((4, 4, 1), np.float32); mask[::2, ::2] = 1.0; mask[1::2, 1::2] = 1.0                           # checker of texels
    d.textures = [scenes.Texture(mask)]
    mat = scenes.material_usd_preview_surface("mask", diffuse=(0.5, 0.5, 0.5))
    mat.textures = [scenes.TextureBinding(scenes.UPS["opacity"], 0, (0,), scenes.WRAP_CLAMP, scenes.WRAP_CLAMP, lookup=scenes.LOOKUP_TEXEL_2D)]
    d.materials = [mat]
    d.meshes = [dataclasses.replace(quad, material=0)]
    d.settings = dataclasses.replace(d.settings, jitteredSampling=False)
    o = orc.OracleScene(d)
    rays = orc.camera_rays(d, 0); rays[:, 3] = 0.0; rays[:, 7] = 3.0e38
    hits = o.trace(rays)
    anyh = o.trace(rays, any_hit=True)
    assert np.array_equal(hits["t"] >= 0, anyh["t"] >= 0)
    # what an opaque quad would return, and the uv of those hits
    mat.textures = []
    d.materials = [mat]
    solid = orc.OracleScene(d).trace(rays)
    on = solid["t"] >= 0
    assert on.mean() > 0.3
    tex = orc.OracleScene(d).render_aovs([4])[4].reshape(-1, 4)[:, :2]          # Texcoords AOV of the opaque quad = the uv of each pixel's hit
    tx = np.clip(np.floor(tex[:, 0] * 4).astype(int), 0, 3); ty = np.clip(np.floor(tex[:, 1] * 4).astype(int), 0, 3)
    want = on & (mask[ty, tx, 0] > 0.5)
    inner = on & (np.abs(tex[:, 0] * 4 - np.round(tex[:, 0] * 4)) > 0.05) & (np.abs(tex[:, 1] * 4 - np.round(tex[:, 1] * 4)) > 0.05)   # away from texel borders
    assert np.array_equal((hits["t"] >= 0)[inner], want[inner])
    assert 0.3 < (hits["t"] >= 0)[on].mean() < 0.7


def test_lookup_cornell_renders_and_shows_holes():
    d = scenes.lookup_cornell(48, 48, spp=4)
    o = orc.OracleScene(d)
    img, _, rays = o.render()
    assert np.isfinite(img).all() and rays[0] > 0 and rays[1] > 0
    aov = o.render_aovs([7, 16])
    op = aov[7][..., :3].reshape(-1, 3)
    assert len(np.unique(np.round(op, 3), axis=0)) > 20              # viridis(opacity) over a soft-edged alpha: many values, not 2
    # the same scene with opaque boxes traces fewer closest rays per path on average (nothing passes through)
    solid = scenes.lookup_cornell(48, 48, spp=4)
    for m in solid.materials:
        m.textures = [b for b in m.textures if b.param not in (scenes.UPS["opacity"], scenes.OPBR["geometry_opacity"])]
    img2 = orc.OracleScene(solid).render()[0]
    assert np.abs(img - img2).mean() > 1e-3


# ---- inputs built from several lookups (gtlc.h GtlcCombineOp) ----------------------------------------------------------------
def test_layered_bindings_equal_the_baked_images():
    """multiply / add / mix of lookups, a mask carried in a scratch float, constant operands: with texel fetches of float images the
    combination can be done ahead of time in numpy with the same float32 operations -- both forms of the scene render bit-identically,
    and differently from the untextured box."""
    a, _, ra = orc.OracleScene(scenes.layered_cornell(48, 48, spp=4)).render()
    b, _, rb = orc.OracleScene(scenes.layered_cornell(48, 48, spp=4, baked=True)).render()
    assert ra == rb and np.array_equal(a, b)
    plain = scenes.cornell(48, 48, spp=4)
    layered = scenes.layered_cornell(48, 48, spp=4)
    plain.settings, plain.rect_lights = layered.settings, layered.rect_lights
    c = orc.OracleScene(plain).render()[0]
    assert np.abs(a - c).mean() > 1e-2
    # the order of the bindings is the order of evaluation: the mask must be in its scratch float BEFORE the mix reads it
    swapped = scenes.layered_cornell(48, 48, spp=4)
    back = [m for m in swapped.materials if m.name == "BackMixConstant"][0]
    back.textures = [back.textures[1], back.textures[2], back.textures[0]]          # mix first (factor still 0), mask last
    e = orc.OracleScene(swapped).render()[0]
    assert not np.array_equal(e, a)


def test_rgba8_texels_decode_to_c_over_255():
    """Every byte value through the texel fetch: (float) c / 255, correctly rounded."""
    t = scenes.Texture(np.arange(256, dtype=np.uint8).reshape(8, 8, 4))
    for i in range(64):
        got = orc.tex_texel(t, i % 8, i // 8)
        assert np.array_equal(got, (np.arange(4 * i, 4 * i + 4).astype(np.float32) / np.float32(255.0)))
