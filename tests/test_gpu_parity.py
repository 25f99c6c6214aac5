"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Bars: traversal hit records bit-exact (t, u, v, primitive,
instance) and visit counts identical; images |a-b| <= 1e-3 + 1e-3|b| on >= 99.9 % of channels and
RMSE < 1e-3 at identical spp/settings."""
import dataclasses

import numpy as np
import pytest

from gatling_b200 import scenes
from oracle import orc
from helpers import image_metrics, random_rays

pytestmark = pytest.mark.gpu


def _trace_parity(gi, desc, rays, walker=True):
    """`walker=False`: scenes with TEXTURED cutouts -- the oracle's stand-alone BVH8 walker sees node and triangle bytes only and
    cannot evaluate a material input, so visit counts are not compared there (hits still are, against the oracle's scene query)."""
    b = gi.BuiltScene(desc)
    try:
        o = orc.OracleScene(desc)
        nodes, tris = gi.get_bvh(b.scene)
        h_gpu, v_gpu = gi.trace_rays(b.scene, rays, want_visits=True)
        h_ref = o.trace(rays)
        for k in ("prim", "inst", "t", "u", "v"):
            assert np.array_equal(h_gpu[k], h_ref[k]), "closest hit field %s differs from the oracle" % k
        if walker:
            h_walk, v_walk = orc.bvh8_trace(nodes, tris, rays)
            assert np.array_equal(v_gpu, v_walk), "node/triangle visit counts differ from the oracle's BVH8 walk"
        a_gpu = gi.trace_rays(b.scene, rays, any_hit=True)
        a_ref = o.trace(rays, any_hit=True)
        assert np.array_equal(a_gpu["t"] >= 0, a_ref["t"] >= 0), "any-hit occlusion differs"
        return h_gpu
    finally:
        b.close()


def test_trace_parity_cornell(gpu):
    d = scenes.cornell(96, 96, spp=1)
    cam = orc.camera_rays(d, 0)
    cam[:, 3] = 0.0
    cam[:, 7] = 3.0e38
    rays = np.concatenate([cam, random_rays(50000, 7)])
    h = _trace_parity(gpu, d, rays)
    assert (h["inst"] != 0xFFFFFFFF).mean() > 0.5


def test_trace_parity_random_triangles(gpu):
    d = scenes.random_triangles(100000, 64, 64, spp=1)
    cam = orc.camera_rays(d, 0)
    cam[:, 3] = 0.0
    cam[:, 7] = 3.0e38
    short = random_rays(20000, 11, tmax=0.05)       # shadow-like short segments
    rays = np.concatenate([cam, random_rays(200000, 3), short])
    _trace_parity(gpu, d, rays)


def test_trace_edge_cases(gpu):
    d = scenes.cornell(8, 8, spp=1)
    rays = random_rays(64, 5)
    rays[:16, 7] = 0.0            # empty interval (disabled shadow ray, rgen:409-410)
    rays[16:32, 3] = 5.0          # tMin beyond the scene
    rays[32:40, 4:7] = [0.0, 1.0, 0.0]   # axis-aligned directions (zero components)
    rays[40:48, 4:7] = [0.0, 0.0, -1.0]
    _trace_parity(gpu, d, rays)
    b = gpu.BuiltScene(d)
    try:
        h = gpu.trace_rays(b.scene, np.zeros((0, 8), np.float32))   # empty queue
        assert len(h) == 0
    finally:
        b.close()


def _render_parity(gi, desc, rmse_bar=1e-3, within_bar=0.999, flipped_samples=0):
    """`flipped_samples` > 0: the scene makes path decisions (lobe choice, Russian roulette, shadow tests) from values that
    pass through libm on the CPU and CUDA's math library on the GPU; a last-ulp difference can flip one decision and with
    it one whole sample (up to maxSampleValue / spp in one pixel).  Such a test tolerates that many outlier pixels in the
    RMSE and the same number of rays in the counts; the 99.9 % criterion is unchanged."""
    b = gi.BuiltScene(desc)
    try:
        o = orc.OracleScene(desc)
        img = b.render().copy()
        ref, _, rays = o.render()
        st = gi.get_stats()
        rmse, within = image_metrics(img, ref)
        if flipped_samples:
            err = ((img[..., :3].astype(np.float64) - ref[..., :3]) ** 2).sum(-1).reshape(-1)
            rmse = float(np.sqrt(np.sort(err)[:-flipped_samples].sum() / (3 * err.size)))
        assert rmse < rmse_bar, "RMSE %.3e" % rmse
        assert within >= within_bar, "only %.5f of channels within tolerance" % within
        slack = 8 * flipped_samples
        assert abs(st["closestRays"] - rays[0]) <= slack and abs(st["shadowRays"] - rays[1]) <= slack, (st, rays)
        assert np.all(img[..., 3] == 1.0)
        return img, ref
    finally:
        b.close()


def test_render_parity_cornell(gpu):
    _render_parity(gpu, scenes.cornell(128, 128, spp=16, max_bounces=4))


def test_render_parity_random_triangles(gpu):
    _render_parity(gpu, scenes.random_triangles(50000, 160, 90, spp=4, max_bounces=8))


def _lit_cornell(w=96, h=96, spp=8):
    d = scenes.cornell(w, h, spp=spp, max_bounces=5)
    d.settings = dataclasses.replace(d.settings, nextEventEstimation=True)
    d.rect_lights.append(scenes.RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(12, 12, 12), width=0.6, height=0.6))
    d.sphere_lights.append(scenes.SphereLight(position=(0.5, -0.5, 0.2), base_emission=(4, 3, 2), radius=(0.1, 0.1, 0.1)))
    d.distant_lights.append(scenes.DistantLight(direction=(0.2, 0.9, -0.4), base_emission=(1.0, 1.0, 1.2), angle=0.05))
    d.disk_lights.append(scenes.DiskLight(origin=(-0.6, 0.0, 0.5), t0=(0, 1, 0), t1=(0, 0, 1), base_emission=(3, 5, 3), radius=(0.2, 0.15)))
    return d


def test_render_parity_nee_all_light_types(gpu):
    _render_parity(gpu, _lit_cornell())


def test_render_parity_settings_variants(gpu):
    d = scenes.cornell(64, 64, spp=4, max_bounces=6)
    d.settings = dataclasses.replace(d.settings, filterImportanceSampling=False, rrBounceOffset=1, maxSampleValue=2.0)
    _render_parity(gpu, d)
    d.settings = dataclasses.replace(d.settings, jitteredSampling=False, clippingPlanes=True)
    d.camera = dataclasses.replace(d.camera, clipStart=6.5, clipEnd=7.6)
    _render_parity(gpu, d)
    d.settings = dataclasses.replace(d.settings, depthOfField=True, jitteredSampling=True, clippingPlanes=False)
    d.camera = dataclasses.replace(d.camera, fStop=8.0, focusDistance=7.0, focalLength=0.5)
    _render_parity(gpu, d)


def test_render_parity_dome_light(gpu):
    """R11 (rp_main.miss:39-87) with data that can fail: a gradient lat-long dome with one hot texel, a rotation about all
    three axes, and domeLightCameraVisible on / off, over an empty stage (pure mapping) and over the Cornell geometry (bounce
    rays see the dome, primary rays the fallback).  The CPU suite pins the oracle's mapping to numpy (tests/test_dome_cpu.py)."""
    from test_dome_cpu import empty_scene, gradient_dome
    q = np.array([0.18, -0.37, 0.11, 0.0]); q[3] = np.sqrt(1.0 - (q[:3] ** 2).sum())
    dome = scenes.DomeLight(gradient_dome(32, 16, hot=(9, 21)), rotation=tuple(float(x) for x in q), base_emission=(0.5, 2.0, 1.5))
    for visible in (True, False):
        e = empty_scene(96, 64, dome, domeLightCameraVisible=visible)
        e.clear_color = (0.2, 0.4, 0.6, 1.0)
        img, ref = _render_parity(gpu, e, rmse_bar=1e-4)
        assert (img[..., :3].std(axis=(0, 1)).max() > 0.05) == visible   # a picture of the dome, or the flat fallback colour
        c = scenes.cornell(64, 64, spp=8, max_bounces=4)
        c.meshes = c.meshes[:3]                                   # open box: most bounce rays escape to the dome
        c.dome_light = dome
        c.clear_color = (0.2, 0.4, 0.6, 1.0)
        c.settings = dataclasses.replace(c.settings, domeLightCameraVisible=visible)
        _render_parity(gpu, c)
    identity = scenes.DomeLight(dome.texels)                      # the rotation is not ignored
    b = gpu.BuiltScene(empty_scene(96, 64, identity))
    try:
        plain = b.render().copy()
    finally:
        b.close()
    b = gpu.BuiltScene(empty_scene(96, 64, dome))
    try:
        rotated = b.render().copy()
    finally:
        b.close()
    assert np.abs(plain[..., :3] - rotated[..., :3] / np.array(dome.base_emission, np.float32)).max() > 0.3


def test_render_empty_1x1(gpu):
    """Render.Empty1x1 of the reference's test-suite (src/hdGatling/main.cpp:584): no geometry, one pixel."""
    e = scenes.cornell(1, 1, spp=4)
    e.meshes = []
    e.clear_color = (0.25, 0.5, 1.0, 1.0)
    img, ref = _render_parity(gpu, e)
    assert img.shape == (1, 1, 4) and np.allclose(img[0, 0], [63 / 255, 127 / 255, 1.0, 1.0], atol=1e-6)


def test_progressive_accumulation(gpu):
    d = scenes.cornell(64, 64, spp=4, max_bounces=4)
    b = gpu.BuiltScene(d)
    try:
        o = orc.OracleScene(d)
        img1 = b.render().copy()
        img2 = b.render().copy()          # second giRender call refines: sampleOffset = 4
        ref1, _, _ = o.render(sample_offset=0)
        ref2, _, _ = o.render(sample_offset=4, prev=ref1)
        for a, r in ((img1, ref1), (img2, ref2)):
            rmse, within = image_metrics(a, r)
            assert rmse < 1e-3 and within >= 0.999
        # changing the camera resets the accumulation (Gi.cpp:1859-1987)
        d2 = dataclasses.replace(d, camera=dataclasses.replace(d.camera, vfov=0.5))
        b.desc = d2
        img3 = b.render().copy()
        ref3, _, _ = orc.OracleScene(d2).render()
        rmse, within = image_metrics(img3, ref3)
        assert rmse < 1e-3 and within >= 0.999
    finally:
        b.close()


def test_wave_splitting_is_invisible(gpu):
    """Small path pools force many waves / pixel chunks; pixels must not change."""
    d = scenes.cornell(64, 48, spp=6, max_bounces=4)
    b = gpu.BuiltScene(d)
    try:
        full = b.render().copy()
    finally:
        b.close()
    for pool in (64 * 48 * 2, 1500):
        gpu.set_tuning(maxPathsPerWave=pool)
        b = gpu.BuiltScene(d)
        try:
            img = b.render().copy()
        finally:
            b.close()
            gpu.set_tuning(maxPathsPerWave=1 << 24)
        assert np.array_equal(img, full), "pool %d changed the image" % pool


def test_shard_sum_equals_full(gpu):
    d = _lit_cornell(80, 64, spp=4)
    b = gpu.BuiltScene(d)
    try:
        full = b.render().copy()
    finally:
        b.close()
    acc = np.zeros_like(full)
    for rank in range(3):
        b = gpu.BuiltScene(d)
        try:
            gpu.set_shard(b.scene, rank, 3, 16)
            part = b.render().copy()
            o, _, _ = orc.OracleScene(d).render(shard=(rank, 3, 16))
            rmse, within = image_metrics(part, o)
            assert rmse < 1e-3 and within >= 0.999
            acc += part
        finally:
            b.close()
    assert np.array_equal(acc, full)


def test_sample_range_shards_sum_to_the_image(gpu):
    """SURVEY.md 8e, the alternative to tiles: every rank renders all pixels and a share of the samples; the partial images
    (each equal to the oracle's partial) sum to the single-GPU image up to the order of the float additions, also across
    progressive calls and with an empty sample range (5 ranks, 4 samples)."""
    d = _lit_cornell(72, 56, spp=4)
    b = gpu.BuiltScene(d)
    try:
        full = [b.render().copy(), b.render().copy()]
        o = orc.OracleScene(d)
        for world in (3, 5):
            acc = [np.zeros_like(full[0]), np.zeros_like(full[0])]
            for rank in range(world):
                gpu.set_shard(b.scene, rank, world, gpu.SHARD_SAMPLES)
                one = b.render().copy()
                two = b.render().copy()
                o1, _, _ = o.render(shard=(rank, world, o.SHARD_SAMPLES))
                o2, _, _ = o.render(sample_offset=4, prev=o1, shard=(rank, world, o.SHARD_SAMPLES))
                for part, want in ((one, o1), (two, o2)):
                    rmse, within = image_metrics(part, want)
                    assert rmse < 1e-3 and within >= 0.999
                    assert np.all(part[..., 3] == (1.0 if rank == 0 else 0.0))
                acc[0] += one; acc[1] += two
            for a, f in zip(acc, full):
                assert np.all(a[..., 3] == 1.0)
                assert np.allclose(a[..., :3], f[..., :3], rtol=2e-6, atol=1e-7)
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(b.render(), full[0])
    finally:
        b.close()


@pytest.mark.logs_errors
def test_errors_are_loud(gpu):
    d = scenes.cornell(16, 16, spp=1)
    b = gpu.BuiltScene(d)
    try:
        st = dataclasses.replace(d.settings, mediumStackSize=9)
        with pytest.raises(RuntimeError):
            b.render(settings=st)
        assert "mediumStackSize" in gpu.giLastError()
    finally:
        b.close()


def test_render_parity_openpbr_material_stress(gpu):
    """C5 stand-in at test size: 16 OpenPBR parameter sets (diffuse / coat / metal / glass), dome + distant light, NEE on."""
    _render_parity(gpu, scenes.material_stress(tris_per_mesh=400, meshes=16, width=96, height=72, spp=4))


def test_render_parity_kitchen_standin(gpu):
    """C3 stand-in at test size: instanced boxes / tubes / tori in a room, rect + distant lights, NEE on."""
    _render_parity(gpu, scenes.kitchen_standin(tris_per_instance=300, instances=60, meshes=12, materials=8, width=96, height=72, spp=4))


def test_material_sort_is_invisible(gpu):
    d = scenes.material_stress(tris_per_mesh=300, meshes=25, width=80, height=60, spp=3)
    imgs = []
    for mode in (0, 1):
        gpu.set_tuning(sortByMaterial=mode)
        b = gpu.BuiltScene(d)
        try:
            imgs.append(b.render().copy())
            st = gpu.get_stats()
            assert (st["msSort"] > 0) == (mode == 1)
        finally:
            b.close()
            gpu.set_tuning(sortByMaterial=-1)
    assert np.array_equal(imgs[0], imgs[1])


AOVS = [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]   # everything but the ClockCycles debug AOV (a per-pixel timer of the megakernel)
INT_AOVS = (11, 13, 14)


def _aov_parity(gi, desc, second_call=False):
    clear = {11: (-1,), 13: (-1,), 14: (-1,), 1: (0.5, 0.5, 1.0, 0.0), 12: (1.0,)}
    b = gi.BuiltScene(desc)
    try:
        o = orc.OracleScene(desc)
        got = b.render_aovs(AOVS, clear)
        ref = o.render_aovs(AOVS, clear)
        if second_call:            # progressive refinement of the accumulating AOVs (Color, Normal, Albedo)
            got = b.render_aovs(AOVS, clear)
            ref = o.render_aovs(AOVS, clear, sample_offset=desc.settings.spp, prev=ref)
        for aov in AOVS:
            g, r = got[gi.GiAovId(aov)], ref[aov]
            if aov == 2:               # NEE debug AOV: red / green per pixel; a shadow test may flip with the last ulp of the light sample
                assert (np.abs(g[..., :3] - r[..., :3]).max(-1) > 0).mean() < 2e-3, "NEE AOV differs on %.4f of the pixels" % (np.abs(g[..., :3] - r[..., :3]).max(-1) > 0).mean()
                assert len(np.unique(g[..., :3].reshape(-1, 3), axis=0)) >= (2 if desc.settings.nextEventEstimation else 1)
            elif aov in INT_AOVS:
                assert (g != r).mean() < 1e-3, "integer AOV %d differs on %.4f of the pixels" % (aov, (g != r).mean())
            else:
                diff = np.abs(g[..., :3 if g.shape[-1] == 4 else 1].astype(np.float64) - r[..., :3 if r.shape[-1] == 4 else 1])
                assert np.sqrt(np.mean(diff ** 2)) < 2e-3, "AOV %d RMSE %.3e" % (aov, np.sqrt(np.mean(diff ** 2)))
                assert (diff <= 1e-3 + 1e-3 * np.abs(r[..., :diff.shape[-1]])).mean() >= 0.998, "AOV %d" % aov
        return got
    finally:
        b.close()


def test_aov_parity_cornell(gpu):
    got = _aov_parity(gpu, scenes.cornell(64, 64, spp=3, max_bounces=4))
    ids = got[gpu.GiAovId.ObjectId]
    assert set(np.unique(ids)) <= set(range(-1, 9)) and (ids > 0).mean() > 0.5     # 8 meshes, object ids 1..8, clear value -1


def test_aov_parity_instances_and_progressive(gpu):
    d = scenes.kitchen_standin(tris_per_instance=200, instances=30, meshes=6, materials=6, width=64, height=48, spp=2)
    got = _aov_parity(gpu, d, second_call=True)
    assert got[gpu.GiAovId.InstanceId].max() >= 3      # instanced props carry their instance index


def test_clock_cycles_aov(gpu):
    """The reference's in-tree profiler (rp_main.rgen:187-189,501-504): per pixel, the cycles its samples' paths spent in the
    kernels.  Not a parity quantity (it is a timer); what must hold: every pixel counts something, pixels whose paths bounce
    around the box cost more than pixels that leave through the open front, and binding the AOV leaves the image alone."""
    d = scenes.cornell(64, 64, spp=4, max_bounces=6)
    d.meshes = d.meshes[:4]                              # light + floor + ceiling + back wall: half the camera rays escape at once
    b = gpu.BuiltScene(d)
    try:
        plain = b.render().copy()
    finally:
        b.close()
    b = gpu.BuiltScene(d)
    try:
        got = b.render_aovs([gpu.GiAovId.Color, gpu.GiAovId.ClockCycles, gpu.GiAovId.Bounces])
    finally:
        b.close()
    assert np.array_equal(got[gpu.GiAovId.Color], plain)
    clk = got[gpu.GiAovId.ClockCycles][..., 0]
    assert np.isfinite(clk).all() and clk.min() > 0.0
    assert not got[gpu.GiAovId.ClockCycles][..., 1:3].any()                 # uvec3(cycles, 0, 0)
    ids = b_ids = None
    bb = gpu.BuiltScene(d)
    try:
        hit = bb.render_aovs([gpu.GiAovId.Barycentrics])[gpu.GiAovId.Barycentrics][..., :3].sum(-1) > 0.5
    finally:
        bb.close()
    assert hit.any() and (~hit).any()
    assert clk[hit].mean() > 1.5 * clk[~hit].mean()


def test_color_only_render_is_unchanged_by_aov_support(gpu):
    d = scenes.cornell(48, 48, spp=2, max_bounces=3)
    b = gpu.BuiltScene(d)
    try:
        plain = b.render().copy()
    finally:
        b.close()
    b = gpu.BuiltScene(d)
    try:
        with_aovs = b.render_aovs([0, 1, 16, 12])[gpu.GiAovId.Color]
    finally:
        b.close()
    assert np.array_equal(plain, with_aovs)


def _cutout_cornell(w=96, h=96, spp=6):
    """Cornell box whose tall box is a 40 % cutout and whose short box is 85 % opaque; NEE on so shadow rays cross cutouts."""
    d = _lit_cornell(w, h, spp)
    d.materials = list(d.materials) + [scenes.material_usd_preview_surface("Cutout40", diffuse=(0.2, 0.4, 0.9), opacity=0.4),
                                       scenes.material_usd_preview_surface("Cutout85", diffuse=(0.9, 0.8, 0.2), opacity=0.85)]
    # the boxes of cornell.usda sit exactly ON the floor: once they are see-through, their bottom faces z-fight with the
    # floor and the winner flips with the last ulp of the ray (libm vs CUDA sin/cos) -- lift them to keep the test sharp
    for k, mat in ((6, len(d.materials) - 2), (7, len(d.materials) - 1)):
        xf = d.meshes[k].transform.copy()
        xf[3, 2] += 0.03
        d.meshes[k] = dataclasses.replace(d.meshes[k], material=mat, transform=xf)
    return d


def test_cutout_trace_and_render_parity(gpu):
    d = _cutout_cornell()
    cam = orc.camera_rays(d, 0)
    cam[:, 3] = 0.0
    cam[:, 7] = 3.0e38
    _trace_parity(gpu, d, np.concatenate([cam, random_rays(30000, 21)]))     # seed 0 on both sides
    img, ref = _render_parity(gpu, d)
    opaque = _lit_cornell(96, 96, 6)
    b = gpu.BuiltScene(opaque)
    try:
        img_opaque = b.render().copy()
    finally:
        b.close()
    assert np.abs(img - img_opaque).mean() > 1e-3, "the cutout materials changed nothing"
    got = _aov_parity(gpu, dataclasses.replace(d, width=48, height=48))
    op = got[gpu.GiAovId.Opacity][..., :3]
    assert len(np.unique(op.reshape(-1, 3), axis=0)) >= 3      # clear, opaque red (chit:202), viridis(0.4) / viridis(0.85)


def test_textured_cutout_and_lookup_modes_parity(gpu):
    """R13 with an opacity that is a material INPUT (rp_main.ahit:35-61: mdl_cutout_opacity on a shading state) and the rest
    of R15: filtered / texel-fetched 2-D and 3-D lookups (mdl_interface.glsl:45-64, 91-110, 167-185), integer primvars
    (:417-479).  Hit records bit-exact for closest and any-hit rays, image + AOVs against the oracle; every traversal kernel
    that can meet such a triangle (batched default, per-lane, counting) gives the same hits."""
    d = scenes.lookup_cornell(96, 96, spp=8)
    cam = orc.camera_rays(d, 0)
    cam[:, 3] = 0.0
    cam[:, 7] = 3.0e38
    rays = np.concatenate([cam, random_rays(40000, 23)])
    h = _trace_parity(gpu, d, rays, walker=False)                  # per-lane counting kernel + any-hit, seed 0 on both sides
    b = gpu.BuiltScene(d)
    try:
        for variant in (0, 1):
            gpu.set_tuning(traceVariant=variant)
            hv = gpu.trace_rays(b.scene, rays)
            for k in ("prim", "inst", "t", "u", "v"):
                assert np.array_equal(hv[k], h[k]), (variant, k)
    finally:
        gpu.set_tuning(traceVariant=0)
        b.close()
    img, ref = _render_parity(gpu, d, flipped_samples=4)
    got = _aov_parity(gpu, dataclasses.replace(d, width=48, height=48))
    op = got[gpu.GiAovId.Opacity][..., :3]
    assert len(np.unique(np.round(op.reshape(-1, 3), 3), axis=0)) > 20         # viridis over a soft-edged alpha mask
    solid = scenes.lookup_cornell(96, 96, spp=8)
    for m in solid.materials:
        m.textures = [t for t in m.textures if t.param not in (scenes.UPS["opacity"], scenes.OPBR["geometry_opacity"])]
    bs = gpu.BuiltScene(solid)
    try:
        img_solid = bs.render().copy()
    finally:
        bs.close()
    assert np.abs(img - img_solid).mean() > 1e-3, "the opacity textures changed nothing"


def test_primvar_driven_cutout_parity(gpu):
    """UsdPreviewSurface.opacity <- displayOpacity: the cutout input read from a primvar (vertex ramp on the tall box, per-face
    integer flags on the short one, a missing primvar with fallback 0.3 on the back wall) inside the any-hit test."""
    d = _lit_cornell(96, 96, 6)
    mats = list(d.materials)

    def cut(name, fallback):
        m = scenes.material_usd_preview_surface(name, diffuse=(0.3, 0.6, 0.8))
        m.primvars = [scenes.PrimvarBinding(scenes.UPS["opacity"], 1, "displayOpacity", (fallback, 0, 0, 0))]
        return m
    mats += [cut("ramp", 1.0), cut("flags", 1.0), cut("missing", 0.3)]
    for i, m in enumerate(d.meshes):
        if m.name in ("Box1", "Box2"):
            xf = np.asarray(m.transform, np.float32).copy(); xf[3, 2] += 0.03          # no z-fighting with the floor once see-through
            if m.name == "Box1":
                pv = scenes.Primvar("displayOpacity", scenes.INTERP_VERTEX, np.linspace(0.05, 1.0, len(m.vertices), dtype=np.float32)[:, None])
                d.meshes[i] = dataclasses.replace(m, material=len(mats) - 3, transform=xf, primvars=[pv])
            else:
                pv = scenes.Primvar("displayOpacity", scenes.INTERP_UNIFORM, (np.arange(len(m.faces)) % 2).astype(np.int32)[:, None])
                d.meshes[i] = dataclasses.replace(m, material=len(mats) - 2, transform=xf, primvars=[pv])
        elif m.name == "BackPlane":
            d.meshes[i] = dataclasses.replace(m, material=len(mats) - 1)
    d.materials = mats
    cam = orc.camera_rays(d, 0)
    cam[:, 3] = 0.0
    cam[:, 7] = 3.0e38
    h = _trace_parity(gpu, d, np.concatenate([cam, random_rays(30000, 29)]), walker=False)
    assert (h["t"] > 0).any()
    _render_parity(gpu, d)
    got = _aov_parity(gpu, dataclasses.replace(d, width=48, height=48))
    assert len(np.unique(np.round(got[gpu.GiAovId.Opacity][..., :3].reshape(-1, 3), 3), axis=0)) > 10


def test_traversal_variants_are_invisible(gpu):
    """The traversal kernels (GtlcTuning.traceVariant 0 = 7 batched default, 1 per-lane loop, 6 TMA-staged ray tiles) walk the BVH in a different order; the
    closest hit is order-independent (smallest t, ties by (instance, primitive)), so hit records and images must be
    bit-identical to each other -- closest, any-hit and shadow mode, with cutouts."""
    d = scenes.random_triangles(60000, 64, 64, spp=1)
    rays = np.concatenate([random_rays(150000, 21), random_rays(20000, 22, tmax=0.05)])
    lit = dataclasses.replace(_cutout_cornell(), width=64, height=64)
    ref = None
    try:
        for variant in (1, 0, 6, 7):
            gpu.set_tuning(traceVariant=variant)
            b = gpu.BuiltScene(d)
            try:
                h = gpu.trace_rays(b.scene, rays)
                a = gpu.trace_rays(b.scene, rays, any_hit=True)
            finally:
                b.close()
            b = gpu.BuiltScene(lit)
            try:
                img = b.render().copy()
            finally:
                b.close()
            got = ([h[k].copy() for k in ("prim", "inst", "t", "u", "v")], a["t"] >= 0, img)
            if ref is None:
                ref = got
                continue
            for x, y in zip(ref[0], got[0]):
                assert np.array_equal(x, y), "variant %d: closest hits differ" % variant
            assert np.array_equal(ref[1], got[1]), "variant %d: any-hit occlusion differs" % variant
            assert np.array_equal(ref[2], got[2]), "variant %d: image differs" % variant
    finally:
        gpu.set_tuning(traceVariant=0)


def test_device_bvh_build_parity(gpu):
    """GtlcTuning.bvhBuilder = 1 / 2: flatten + Morton sort + Karras tree or PLOC + SAH table + collapse on the GPU.  The node and
    triangle arrays must equal the oracle's restatement (oracle/orc_lbvh8.h) byte for byte; the flattened records must
    equal the host flatten; traversal and images must not change (closest hits do not depend on the tree)."""
    from test_lbvh8_cpu import rank_order, with_non_finite_triangles
    inst_scene = scenes.kitchen_standin(tris_per_instance=300, instances=60, meshes=12, materials=8, width=64, height=48, spp=2)
    cases = [scenes.cornell(48, 48, spp=2), scenes.random_triangles(200000, 64, 64, spp=1), inst_scene,
             scenes.random_triangles(1, 8, 8, spp=1), scenes.random_triangles(3, 8, 8, spp=1), scenes.random_triangles(7, 8, 8, spp=1),
             with_non_finite_triangles()[0]]       # inactive triangles: a point at the origin for the builders, never hit
    try:
        for d, builder in [(d, k) for d in cases for k in (1, 2)]:
            gpu.set_tuning(bvhBuilder=0)
            b = gpu.BuiltScene(d)
            try:
                _, host_tris = gpu.get_bvh(b.scene)
                img_host = b.render().copy()
                ids_host = b.render_aovs([0, 11, 14]) if d is inst_scene else None     # Color, ObjectId, InstanceId
            finally:
                b.close()
            gpu.set_tuning(bvhBuilder=builder)
            b = gpu.BuiltScene(d)
            try:
                nodes, tris = gpu.get_bvh(b.scene)
                st = gpu.get_stats()
                recs = rank_order(tris)
                assert np.array_equal(recs, rank_order(host_tris), equal_nan=True), "device flatten differs from the host flatten"
                o_nodes, o_tris, o_depth = orc.lbvh8_build(recs, builder)
                assert len(nodes) == len(o_nodes) and np.array_equal(nodes, o_nodes), "BVH8 nodes differ from the oracle's build"
                assert np.array_equal(tris, o_tris, equal_nan=True), "leaf-order triangles differ from the oracle's build"
                assert st["bvhMaxDepth"] == o_depth and st["bvhNodeCount"] == len(o_nodes) // 80
                rays = random_rays(40000, 31)
                h_gpu, v_gpu = gpu.trace_rays(b.scene, rays, want_visits=True)
                h_walk, v_walk = orc.bvh8_trace(nodes, tris, rays)
                for k in ("prim", "inst", "t", "u", "v"):
                    assert np.array_equal(h_gpu[k], h_walk[k]), k
                assert np.array_equal(v_gpu, v_walk)
                assert np.array_equal(b.render(), img_host), "image changed with the device-built BVH"
                if ids_host is not None:       # the per-instance id table does not come from the builder: it must survive a device build
                    ids = b.render_aovs([0, 11, 14])
                    for aov in (gpu.GiAovId.ObjectId, gpu.GiAovId.InstanceId):
                        assert np.array_equal(ids[aov], ids_host[aov]), aov
                    assert len(np.unique(ids[gpu.GiAovId.InstanceId])) > 1
            finally:
                b.close()
        # an empty scene
        e = scenes.cornell(8, 8, spp=1)
        e.meshes = []
        b = gpu.BuiltScene(e)
        try:
            nodes, tris = gpu.get_bvh(b.scene)
            assert len(nodes) == 80 and len(tris) == 0 and not nodes.any()
            assert np.all(gpu.trace_rays(b.scene, random_rays(64, 2))["t"] == -1)
        finally:
            b.close()
    finally:
        gpu.set_tuning(bvhBuilder=0)


def test_render_parity_textured_materials(gpu):
    """R15: texture lookups (all four formats, every wrap mode, UsdTransform2d, sRGB, scale/bias, clip fallback), a
    tangent-space normal map and a textured emission, image and AOVs against the oracle."""
    d = scenes.textured_cornell(96, 96, spp=8)
    img, ref = _render_parity(gpu, d)
    plain = scenes.cornell(96, 96, spp=8)
    plain.settings, plain.rect_lights = d.settings, d.rect_lights
    b = gpu.BuiltScene(plain)
    try:
        assert np.abs(b.render() - img).mean() > 1e-2, "the textures changed nothing"
    finally:
        b.close()
    _aov_parity(gpu, dataclasses.replace(d, width=48, height=48))
    _render_parity(gpu, scenes.textured_cornell(64, 64, spp=4, nee=False))
    # material sort and wave splitting stay invisible with textured materials
    imgs = []
    for sort, pool in ((0, 1 << 24), (1, 4096)):
        gpu.set_tuning(sortByMaterial=sort, maxPathsPerWave=pool)
        b = gpu.BuiltScene(dataclasses.replace(d, width=48, height=48))
        try:
            imgs.append(b.render().copy())
        finally:
            b.close()
            gpu.set_tuning(sortByMaterial=-1, maxPathsPerWave=1 << 24)
    assert np.array_equal(imgs[0], imgs[1])


def test_unorm8_texels_decode_exactly(gpu):
    """A texel-fetched RGBA8 image holding all 256 byte values renders bit-identically to the float image of the same values c / 255
    (correctly rounded division in the decode, textures.cuh tex_fetch)."""
    bytes_ = np.arange(256, dtype=np.uint8).reshape(8, 8, 4)
    imgs = []
    for tex in (scenes.Texture(bytes_), scenes.Texture((bytes_.astype(np.float32) / np.float32(255.0)).astype(np.float32))):
        d = scenes.cornell(64, 64, spp=4)
        for m in d.meshes:
            scenes.assign_planar_uv(m)
        d.textures = [tex]
        for i, name in enumerate(("bottomplane", "backplane", "leftplane")):
            mat = scenes.material_usd_preview_surface("Bytes%d" % i, diffuse=(0.5, 0.5, 0.5))
            mat.textures = [scenes.TextureBinding(scenes.UPS["diffuse"], 0, ((0, 1, 2), (3, 0, 1), (2, 3, 0))[i], lookup=scenes.LOOKUP_TEXEL_2D, uv_scale=(0.9, 0.9), uv_translation=(0.05, 0.05))]
            d.materials = list(d.materials) + [mat]
            for k, mesh in enumerate(d.meshes):
                if name in mesh.name.lower():
                    d.meshes[k] = dataclasses.replace(mesh, material=len(d.materials) - 1)
        b = gpu.BuiltScene(d)
        try:
            imgs.append(b.render().copy())
        finally:
            b.close()
    assert imgs[0][..., :3].std() > 0.05
    assert np.array_equal(imgs[0], imgs[1])


def test_layered_bindings_parity_and_baked_equivalence(gpu):
    """Inputs built from several lookups (gtlc.h GtlcCombineOp: image x image, image + image, mix through a scratch float, constant
    operands) -- what a MaterialX graph with multiply / add / mix nodes in front of a shader input becomes: against the oracle, and
    bit for bit against the same scene with ONE pre-combined float image per input (texel fetches: the same float32 operations)."""
    d = scenes.layered_cornell(96, 96, spp=8)
    img, ref = _render_parity(gpu, d)
    b = gpu.BuiltScene(scenes.layered_cornell(96, 96, spp=8, baked=True))
    try:
        assert np.array_equal(b.render(), img)
    finally:
        b.close()
    plain = scenes.cornell(96, 96, spp=8)
    plain.settings, plain.rect_lights = d.settings, d.rect_lights
    b = gpu.BuiltScene(plain)
    try:
        assert np.abs(b.render() - img).mean() > 1e-2, "the bindings changed nothing"
    finally:
        b.close()
    _aov_parity(gpu, dataclasses.replace(d, width=48, height=48))


@pytest.mark.logs_errors
def test_texture_errors_are_loud(gpu):
    assert gpu.giCreateTexture(np.zeros((0, 4, 4), np.uint8)) is None and "empty texture" in gpu.giLastError()
    s = gpu.giCreateScene()
    tex = gpu.giCreateTexture(np.zeros((2, 2, 4), np.uint8))
    assert tex
    bad = scenes.TextureBinding(47, 0, (0, 1, 2))                                  # writes floats 47..49 of a 48-float block
    assert gpu.giCreateMaterial(s, scenes.CLOSURE_DIFFUSE, [0.5] * 3, "m", [bad], [tex]) is None
    assert "outside the parameter block" in gpu.giLastError()
    ok = scenes.TextureBinding(0, 0, (0, 1, 2))
    assert gpu.giCreateMaterial(s, scenes.CLOSURE_DIFFUSE, [0.5] * 3, "m", [ok], [s]) is None         # a handle, but not a texture
    m = gpu.giCreateMaterial(s, scenes.CLOSURE_DIFFUSE, [0.5] * 3, "m", [ok], [tex])
    assert m
    # the lookup modes: an unknown mode, a 3-D texture read with a 2-D lookup, a bad wrap mode along w, and a cutout input
    # that is bound as part of a colour are refused; rays with tMin < 0 are refused by the ray query
    vol = gpu.giCreateTexture(scenes.procedural_volume(4, 3))
    assert vol
    assert gpu.giCreateMaterial(s, scenes.CLOSURE_DIFFUSE, [0.5] * 3, "m", [scenes.TextureBinding(0, 0, (0, 1, 2), lookup=7)], [tex]) is None
    assert "lookup mode" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, scenes.CLOSURE_DIFFUSE, [0.5] * 3, "m", [ok], [vol]) is None
    assert "3-D texture with a 2-D lookup" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, scenes.CLOSURE_DIFFUSE, [0.5] * 3, "m", [scenes.TextureBinding(0, 0, (0, 1, 2), lookup=scenes.LOOKUP_FILTERED_3D, wrap_w=9)], [vol]) is None
    assert "wrap mode" in gpu.giLastError()
    ups = scenes.material_usd_preview_surface("x")
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [scenes.TextureBinding(scenes.UPS["opacity"] - 1, 0, (0, 1, 2))], [tex]) is None
    assert "cutout input" in gpu.giLastError()
    # the combine word: unknown operation, a factor outside the block + scratch floats, anything but a plain lookup on the normal or
    # the cutout input, a write past the scratch floats; a constant operand needs no texture, a lookup does
    T, cmb = scenes.TextureBinding, scenes.combine
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.UPS["diffuse"], 0, (0, 1, 2), combine=7)], [tex]) is None
    assert "combine word" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.UPS["diffuse"], 0, (0, 1, 2), combine=cmb(scenes.COMBINE_MIX, 60))], [tex]) is None
    assert "combine word" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.BIND_NORMAL, 0, (0, 1, 2), combine=cmb(scenes.COMBINE_ADD))], [tex]) is None
    assert "normal" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.UPS["opacity"], 0, (0,), combine=cmb(scenes.COMBINE_MULTIPLY))], [tex]) is None
    assert "ONE plain lookup" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.UPS["opacity"], 0, (0,)), T(scenes.UPS["opacity"], 0, (1,))], [tex]) is None
    assert "ONE plain lookup" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.BIND_SCRATCH + 2, 0, (0, 1, 2))], [tex]) is None
    assert "outside the parameter block" in gpu.giLastError()
    assert gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.UPS["diffuse"], -1, (0, 1, 2))], [tex]) is None
    assert "no valid texture" in gpu.giLastError()
    m2 = gpu.giCreateMaterial(s, ups.closure, ups.params, "m", [T(scenes.BIND_SCRATCH + 1, 0, (0, 1, 2)), T(scenes.UPS["diffuse"], -1, (0, 1, 2), bias=(0.2, 0.3, 0.4, 0.0),
                              combine=cmb(scenes.COMBINE_MIX, scenes.BIND_SCRATCH + 3, constant=True))], [tex])
    assert m2
    gpu.giDestroyMaterial(m2)
    d = scenes.cornell(8, 8, spp=1)
    b = gpu.BuiltScene(d)
    try:
        rays = random_rays(4, 1); rays[2, 3] = -0.5
        with pytest.raises(RuntimeError, match="tMin"):
            gpu.trace_rays(b.scene, rays)
    finally:
        b.close()
    gpu.giDestroyTexture(vol)
    gpu.giDestroyTexture(tex)            # deferred: the material still references it
    gpu.giDestroyMaterial(m)
    gpu.giDestroyScene(s)


@pytest.mark.logs_errors      # ends with a refused mediumStackSize
def test_render_parity_volumes(gpu):
    """R14: medium stack (push / pop, nested media), free-flight sampling, Henyey-Greenstein scattering, NEE from inside
    nothing special (shadow rays are plain any-hit rays), against the oracle; plus the Beer-only mediumStackSize = 0 path."""
    for n in (2, 1, 0):
        _render_parity(gpu, scenes.volume_cornell(96, 96, spp=8, medium_stack_size=n))
    _render_parity(gpu, scenes.volume_cornell(64, 64, spp=6, medium_stack_size=2, nee=False))
    _aov_parity(gpu, scenes.volume_cornell(48, 48, spp=4, medium_stack_size=2))
    d = scenes.volume_cornell(32, 32, spp=2, medium_stack_size=5)
    b = gpu.BuiltScene(d)
    try:
        with pytest.raises(RuntimeError, match="exceeds the 4 nested media"):
            b.render()
    finally:
        b.close()


def test_full_size_properties_c2(gpu):
    """Config C2 at BASELINE.json's full size (1 M triangles, 1920x1080), where the oracle is too slow to compare
    pixel by pixel: size-independent properties of the path."""
    d = scenes.random_triangles(1_000_000, 1920, 1080, spp=2, max_bounces=8)
    b = gpu.BuiltScene(d)
    try:
        img = b.render().copy()
        st = gpu.get_stats()
        # furnace bound: Lambert 0.8 everywhere, no emitters, white background => every sample is in [0, 1]
        assert np.isfinite(img).all() and img[..., :3].min() >= 0.0 and img[..., :3].max() <= 1.0 + 1e-6
        assert np.all(img[..., 3] == 1.0)
        assert st["paths"] == 1920 * 1080 * 2 and st["paths"] <= st["closestRays"] <= 8 * st["paths"]
        # the image does not depend on how the work is cut into waves or shards (bitwise)
        gpu.set_tuning(maxPathsPerWave=1 << 20)
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(b.render(), img)
        gpu.set_tuning(maxPathsPerWave=1 << 24)
        acc = np.zeros_like(img)
        for r in range(3):
            gpu.set_shard(b.scene, r, 3, 64)
            part = b.render()
            acc[..., :3] += part[..., :3]
            acc[..., 3] = np.maximum(acc[..., 3], part[..., 3])
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(acc, img)
        # closest-hit queries: shrinking the interval to just beyond / just before the hit
        rays = random_rays(400000, 77)
        h = gpu.trace_rays(b.scene, rays)
        hit = h["t"] > 0
        assert 0.2 < hit.mean() < 1.0
        beyond = rays[hit].copy(); beyond[:, 7] = np.nextafter(h["t"][hit], np.float32(np.inf))
        h2 = gpu.trace_rays(b.scene, beyond)
        for k in ("t", "u", "v", "prim", "inst"):
            assert np.array_equal(h2[k], h[k][hit]), k                                        # idempotent
        before = rays[hit].copy(); before[:, 7] = h["t"][hit]
        h3 = gpu.trace_rays(b.scene, before)                                                  # open interval: the hit itself is excluded
        assert np.all((h3["t"] < h["t"][hit]))
        occl = gpu.trace_rays(b.scene, rays, any_hit=True)
        assert np.array_equal(occl["t"] >= 0, hit)                                            # any-hit agrees with closest-hit on occlusion
        # the hit point lies on the reported triangle: barycentric reconstruction == origin + t * direction (to rounding)
        _, tris = gpu.get_bvh(b.scene)
        t = tris.reshape(-1, 12)
        prim_of = np.empty(len(t), np.int64); prim_of[t[:, 3].view(np.uint32)] = np.arange(len(t))
        rec = t[prim_of[h["prim"][hit]]]
        v0, v1, v2 = rec[:, 0:3], rec[:, 4:7], rec[:, 8:11]
        u, v = h["u"][hit, None], h["v"][hit, None]
        p_bary = v0 * (1 - u - v) + v1 * u + v2 * v
        p_ray = rays[hit, 0:3] + rays[hit, 4:7] * h["t"][hit, None]
        assert np.abs(p_bary - p_ray).max() < 2e-5
    finally:
        gpu.set_tuning(maxPathsPerWave=1 << 24)
        b.close()


def test_full_size_c1_parity(gpu):
    """Config C1 exactly as BASELINE.json states it (cornell.usda, 512x512, spp 64, max-bounces 4, UsdPreviewSurface): the whole
    image against the oracle, traced-ray counts equal."""
    _render_parity(gpu, scenes.cornell(512, 512, spp=64, max_bounces=4))


def test_full_size_c3_band_parity_and_properties(gpu):
    """The headline config C3 at BASELINE.json's full geometry and resolution (4.97 M triangles in 401 instances, 24 OpenPBR
    materials, NEE, 1920x1080).  The oracle is too slow for the whole frame: it renders a band of rows in the middle of the
    image, which must agree with the same rows of the GPU frame; hit records of random rays are bit-exact against the oracle;
    and the frame does not depend on wave size, material sort or traversal kernel (bitwise)."""
    d = scenes.kitchen_standin(width=1920, height=1080, spp=2, max_bounces=8)
    assert d.triangle_count > 4_900_000
    b = gpu.BuiltScene(d)
    try:
        img = b.render().copy()
        st = gpu.get_stats()
        assert np.isfinite(img).all() and np.all(img[..., 3] == 1.0) and img[..., :3].min() >= 0.0
        assert st["paths"] == 1920 * 1080 * 2 and st["shadowRays"] > 0
        o = orc.OracleScene(d)
        r0, r1 = 536, 544
        band, _, rays = o.render(rows=(r0, r1), spp=2)
        got, ref = img[r0:r1, :, :3].astype(np.float64), band[r0:r1, :, :3].astype(np.float64)
        err = ((got - ref) ** 2).sum(-1).reshape(-1)
        rmse = float(np.sqrt(np.sort(err)[:-8].sum() / (3 * err.size)))          # up to 8 samples may flip on a last-ulp decision (see _render_parity)
        assert rmse < 1e-3, "band RMSE %.3e" % rmse
        within = float((np.abs(got - ref) <= 1e-3 + 1e-3 * np.abs(ref)).mean())
        assert within >= 0.999, within
        assert rays[0] > (r1 - r0) * 1920 * 2
        probe = random_rays(60000, 41)
        probe[:, 0:3] *= np.array([3.5, 2.5, 1.2], np.float32); probe[:, 2] += 1.5          # inside the room
        h = gpu.trace_rays(b.scene, probe)
        h_ref = o.trace(probe)
        for k in ("prim", "inst", "t", "u", "v"):
            assert np.array_equal(h[k], h_ref[k]), k
        assert (h["t"] > 0).mean() > 0.9
        o.close()
        gpu.set_tuning(maxPathsPerWave=1 << 21)
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(b.render(), img), "wave size changed the image"
        gpu.set_tuning(maxPathsPerWave=1 << 24, sortByMaterial=0)
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(b.render(), img), "the material sort changed the image"
        gpu.set_tuning(sortByMaterial=-1, traceVariant=1)
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(b.render(), img), "the traversal kernel changed the image"
    finally:
        gpu.set_tuning(maxPathsPerWave=1 << 24, sortByMaterial=-1, traceVariant=0)
        b.close()


def test_full_size_c5_band_parity_and_properties(gpu):
    """Config C5 at BASELINE.json's full size (64 displaced spheres of 31 250 triangles, one OpenPBR parameter set each, dome +
    distant light, NEE, 1920x1080): a band of rows against the oracle, bit-exact hit records, and the properties the domain
    offers at any size -- the frame does not depend on the material sort (64 bins in use), the wave size or the shard count."""
    d = scenes.material_stress(width=1920, height=1080, spp=2, max_bounces=8)
    assert d.triangle_count >= 1_900_000 and len(d.materials) == 64
    b = gpu.BuiltScene(d)
    try:
        img = b.render().copy()
        st = gpu.get_stats()
        assert np.isfinite(img).all() and np.all(img[..., 3] == 1.0) and img[..., :3].min() >= 0.0
        assert st["paths"] == 1920 * 1080 * 2 and st["shadowRays"] > 0
        o = orc.OracleScene(d)
        r0, r1 = 500, 506
        band, _, rays = o.render(rows=(r0, r1), spp=2)
        got, ref = img[r0:r1, :, :3].astype(np.float64), band[r0:r1, :, :3].astype(np.float64)
        err = ((got - ref) ** 2).sum(-1).reshape(-1)
        rmse = float(np.sqrt(np.sort(err)[:-8].sum() / (3 * err.size)))
        assert rmse < 1e-3, "band RMSE %.3e" % rmse
        assert float((np.abs(got - ref) <= 1e-3 + 1e-3 * np.abs(ref)).mean()) >= 0.999
        probe = random_rays(40000, 43)
        probe[:, 0:3] *= np.array([8.0, 1.0, 8.0], np.float32); probe[:, 1] -= 6.0             # in front of the wall of spheres
        probe[:, 5] = np.abs(probe[:, 5]) + 0.5                                                # towards it
        h = gpu.trace_rays(b.scene, probe)
        h_ref = o.trace(probe)
        for k in ("prim", "inst", "t", "u", "v"):
            assert np.array_equal(h[k], h_ref[k]), k
        assert (h["t"] > 0).mean() > 0.3
        o.close()
        gpu.set_tuning(sortByMaterial=0, maxPathsPerWave=1 << 21)
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(b.render(), img), "material sort / wave size changed the image"
        gpu.set_tuning(sortByMaterial=-1, maxPathsPerWave=1 << 24)
        acc = np.zeros_like(img)
        for r in range(2):
            gpu.set_shard(b.scene, r, 2, 32)
            part = b.render()
            acc[..., :3] += part[..., :3]
            acc[..., 3] = np.maximum(acc[..., 3], part[..., 3])
        gpu.set_shard(b.scene, 0, 1, 64)
        assert np.array_equal(acc, img), "tile shards do not sum to the frame"
    finally:
        gpu.set_tuning(maxPathsPerWave=1 << 24, sortByMaterial=-1)
        b.close()


@pytest.mark.logs_errors      # a too-short primvar is ignored, loudly
def test_render_parity_primvars(gpu):
    """R15 scene data: material inputs read from mesh / instancer primvars (constant, instance, uniform, vertex
    interpolation; overrides, fallbacks) -- image and AOVs against the oracle."""
    d = scenes.primvar_cornell(96, 96, spp=8)
    img, _ = _render_parity(gpu, d, flipped_samples=4)
    got = _aov_parity(gpu, dataclasses.replace(d, width=48, height=48))
    alb = got[gpu.GiAovId.Albedo][..., :3]
    assert len(np.unique(np.round(alb.reshape(-1, 3), 3), axis=0)) > 50
    # a primvar whose array is too short for its interpolation is ignored, loudly
    assert "ignored" in gpu.giLastError() and "LeftPlane" in gpu.giLastError()
    # instancer primvars set after the fact dirty the scene
    b = gpu.BuiltScene(d)
    try:
        first = b.render().copy()
        box2 = [i for i, m in enumerate(d.meshes) if m.name == "Box2"][0]
        gpu.giSetMeshInstancerPrimvars(b.meshes[box2], [scenes.Primvar("displayColor", scenes.INTERP_INSTANCE, np.full((4, 3), 0.05, np.float32))])
        assert np.abs(b.render() - first).mean() > 1e-4
    finally:
        b.close()


@pytest.mark.logs_errors      # the primvar scene carries a too-short primvar, ignored loudly
def test_committed_golden_images(gpu):
    """The CUDA path against the oracle images committed under tests/golden/ (made by tests/golden/make_golden.py in the
    build container), not against an oracle run of the moment."""
    import os
    import sys
    golden = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, golden)
    import make_golden
    want = np.load(os.path.join(golden, "feature_anchors.npz"))
    for name, d in make_golden.feature_scenes().items():
        b = gpu.BuiltScene(d)
        try:
            img = b.render().copy()
            st = gpu.get_stats()
        finally:
            b.close()
        rmse, within = image_metrics(img, want[name])
        assert rmse < 1e-3 and within >= 0.999, "%s: RMSE %.3e, %.5f within tolerance" % (name, rmse, within)
        assert abs(st["closestRays"] - int(want[name + "_rays"][0])) <= 8 and abs(st["shadowRays"] - int(want[name + "_rays"][1])) <= 8, name
    c1 = np.load(os.path.join(golden, "cornell_oracle_32.npy"))
    b = gpu.BuiltScene(scenes.cornell(32, 32, spp=4, max_bounces=4))
    try:
        rmse, within = image_metrics(b.render(), c1)
    finally:
        b.close()
    assert rmse < 1e-3 and within >= 0.999
