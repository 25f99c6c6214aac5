"""CPU: primvar ("scene data") lookups of the oracle, scene_data_lookup_float{,2,3} of mdl_interface.glsl:281-417: the
albedo AOV of a Lambert-like surface shows the interpolated primvar directly."""
import dataclasses

import numpy as np

from gatling_b200 import scenes
from oracle import orc


def _one_quad(primvars, binding, instancer=(), instance_ids=None, n_inst=1):
    d = scenes.cornell(32, 32, spp=1)
    floor = [m for m in d.meshes if m.name == "BackPlane"][0]
    mat = scenes.material_diffuse((0.5, 0.5, 0.5), "pv")
    mat.primvars = [binding]
    d.materials = [mat]
    m = dataclasses.replace(floor, material=0, primvars=list(primvars), instancer_primvars=list(instancer))
    if instance_ids is not None:
        m.instance_transforms = np.stack([np.eye(4, dtype=np.float32)] * n_inst)
        m.instance_ids = instance_ids
    d.meshes = [m]
    return d


def test_vertex_uniform_constant_and_fallback():
    base = scenes.cornell(8, 8).meshes[3]
    nv, nf = len(base.vertices), len(base.faces)
    cols = np.linspace(0.1, 0.9, nv * 3, dtype=np.float32).reshape(nv, 3)
    d = _one_quad([scenes.Primvar("displayColor", scenes.INTERP_VERTEX, cols)], scenes.PrimvarBinding(0, 3, "displayColor", (1, 0, 0, 0)))
    alb = orc.OracleScene(d).render_aovs([16, 3])
    a, bary = alb[16][..., :3], alb[3][..., :3]
    hit = bary.sum(-1) > 0.5
    assert hit.mean() > 0.3
    assert a[hit].min() >= 0.1 - 1e-6 and a[hit].max() <= 0.9 + 1e-6 and a[hit].std() > 0.05     # a convex combination of the vertex values
    # uniform: exactly one colour per face (up to the three-term blend's rounding)
    fcols = np.array([[0.2, 0.4, 0.6], [0.7, 0.1, 0.3]], np.float32)[:nf]
    d = _one_quad([scenes.Primvar("displayColor", scenes.INTERP_UNIFORM, fcols)], scenes.PrimvarBinding(0, 3, "displayColor"))
    a = orc.OracleScene(d).render_aovs([16])[16][..., :3]
    seen = np.unique(np.round(a[hit], 4), axis=0)
    assert len(seen) <= nf and all(np.abs(fcols - s).sum(1).min() < 1e-3 for s in seen)
    # constant + scalar primvar broadcast into one component; missing primvar -> fallback
    d = _one_quad([scenes.Primvar("k", scenes.INTERP_CONSTANT, np.array([[0.25]], np.float32))], scenes.PrimvarBinding(1, 1, "k"))
    a = orc.OracleScene(d).render_aovs([16])[16][..., :3]
    assert np.allclose(a[hit], (0.5, 0.25, 0.5), atol=1e-6)
    d = _one_quad([], scenes.PrimvarBinding(0, 3, "nope", (0.3, 0.6, 0.9, 0)))
    a = orc.OracleScene(d).render_aovs([16])[16][..., :3]
    assert np.allclose(a[hit], (0.3, 0.6, 0.9), atol=1e-6)


def test_instance_interpolation_uses_instance_ids_and_mesh_overrides_instancer():
    inst_cols = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    d = _one_quad([], scenes.PrimvarBinding(0, 3, "c"), instancer=[scenes.Primvar("c", scenes.INTERP_INSTANCE, inst_cols)],
                  instance_ids=np.array([2], np.int32), n_inst=1)
    a = orc.OracleScene(d).render_aovs([16, 3])
    hit = a[3][..., :3].sum(-1) > 0.5
    assert np.allclose(a[16][..., :3][hit], (0, 0, 1), atol=1e-6)                  # InstanceIds[gl_InstanceID] = 2
    d.meshes[0].primvars = [scenes.Primvar("c", scenes.INTERP_CONSTANT, np.array([[0.4, 0.4, 0.1]], np.float32))]
    a = orc.OracleScene(d).render_aovs([16])[16][..., :3]
    assert np.allclose(a[hit], (0.4, 0.4, 0.1), atol=1e-6)                        # Gi.cpp:921-928


def test_primvar_cornell_renders():
    d = scenes.primvar_cornell(40, 40, spp=2)
    img, _, rays = orc.OracleScene(d).render()
    assert np.isfinite(img).all() and rays[0] > 0
    alb = orc.OracleScene(d).render_aovs([16])[16][..., :3]
    assert len(np.unique(np.round(alb.reshape(-1, 3), 3), axis=0)) > 50


def test_integer_primvars_take_the_nearest_vertex():
    """scene_data_lookup_int / _int3 (mdl_interface.glsl:417-479): no interpolation -- the element of the vertex with the largest
    barycentric weight, per face for uniform data; the value enters the parameter block as a float."""
    base = scenes.cornell(8, 8).meshes[3]
    nv, nf = len(base.vertices), len(base.faces)
    ids = (np.arange(nv * 3).reshape(nv, 3) % 5).astype(np.int32)                 # small integers, distinct per vertex
    d = _one_quad([scenes.Primvar("tag", scenes.INTERP_VERTEX, ids)], scenes.PrimvarBinding(0, 3, "tag", (9, 9, 9, 0)))
    a = orc.OracleScene(d).render_aovs([16, 3])
    alb, bary = a[16][..., :3], a[3][..., :3]
    hit = bary.sum(-1) > 0.5
    seen = np.unique(alb[hit], axis=0)
    assert len(seen) <= nv and all((ids.astype(np.float32) == s).all(1).any() for s in seen)      # only whole vertex records, never blends
    assert len(seen) >= 3
    # the record is the one of the dominant barycentric weight
    f = base.faces
    face_aov = orc.OracleScene(d).render_aovs([13])[13][..., 0]
    ys, xs = np.nonzero(hit)
    for y, x in list(zip(ys, xs))[::7]:
        b = bary[y, x]
        k = (0 if b[0] > b[2] else 2) if b[0] > b[1] else (1 if b[1] > b[2] else 2)          # mdl_interface.glsl:436-443
        assert np.array_equal(alb[y, x], ids[f[face_aov[y, x]][k]].astype(np.float32))
    # uniform int, scalar, and the fallback for a missing name
    per_face = np.array([[3], [7]], np.int32)[:nf]
    d = _one_quad([scenes.Primvar("k", scenes.INTERP_UNIFORM, per_face)], scenes.PrimvarBinding(2, 1, "k"))
    alb = orc.OracleScene(d).render_aovs([16])[16][..., :3]
    assert set(np.unique(alb[hit][:, 2]).tolist()) <= {3.0, 7.0}


def test_primvar_driven_cutout_opacity():
    """UsdPreviewSurface.opacity <- UsdPrimvarReader(displayOpacity): the any-hit test evaluates the reader at the candidate hit
    (rp_main.ahit:35-61 runs mdl_cutout_opacity on a shading state, which can read scene data).  Per-face opacities 0 / 1 let
    exactly one triangle of the quad stop rays; a vertex ramp lets more rays through where it is low."""
    base = scenes.cornell(8, 8).meshes[3]
    nv, nf = len(base.vertices), len(base.faces)
    ups = scenes.material_usd_preview_surface("op", diffuse=(0.5, 0.5, 0.5))

    def scene(primvars, fallback=1.0):
        d = scenes.cornell(48, 48, spp=1)
        quad = [m for m in d.meshes if m.name == "BackPlane"][0]
        mat = scenes.material_usd_preview_surface("op", diffuse=(0.5, 0.5, 0.5))
        mat.primvars = [scenes.PrimvarBinding(scenes.UPS["opacity"], 1, "displayOpacity", (fallback, 0, 0, 0))]
        d.materials = [mat]
        d.meshes = [dataclasses.replace(quad, material=0, primvars=list(primvars))]
        d.settings = dataclasses.replace(d.settings, jitteredSampling=False)
        return d
    rays = orc.camera_rays(scene([]), 0); rays[:, 3] = 0.0; rays[:, 7] = 3.0e38
    solid = orc.OracleScene(scene([])).trace(rays)                      # missing primvar -> fallback 1 -> opaque
    on = solid["t"] >= 0
    assert on.mean() > 0.3
    assert not (orc.OracleScene(scene([], fallback=0.0)).trace(rays)["t"] >= 0).any()      # fallback 0: fully transparent
    per_face = np.array([[0.0], [1.0]], np.float32)[:nf]
    h = orc.OracleScene(scene([scenes.Primvar("displayOpacity", scenes.INTERP_UNIFORM, per_face)])).trace(rays)
    assert np.array_equal(h["t"] >= 0, on & (solid["prim"] == 1))
    a = orc.OracleScene(scene([scenes.Primvar("displayOpacity", scenes.INTERP_UNIFORM, per_face)])).trace(rays, any_hit=True)
    assert np.array_equal(a["t"] >= 0, h["t"] >= 0)
    flags = np.array([[1], [0]], np.int32)[:nf]                          # the same through an integer primvar
    hi = orc.OracleScene(scene([scenes.Primvar("displayOpacity", scenes.INTERP_UNIFORM, flags)])).trace(rays)
    assert np.array_equal(hi["t"] >= 0, on & (solid["prim"] == 0))
    ramp = np.linspace(0.0, 1.0, nv, dtype=np.float32)[:, None]
    hr = orc.OracleScene(scene([scenes.Primvar("displayOpacity", scenes.INTERP_VERTEX, ramp)])).trace(rays)
    # with the ray-query seed (0) the acceptance draw is one number per triangle: a hit survives iff the opacity interpolated at
    # it reaches that threshold -- per triangle, every surviving hit has a higher opacity than every ignored one
    f = base.faces
    op = ((1.0 - solid["u"] - solid["v"]) * ramp[f[np.minimum(solid["prim"], nf - 1), 0], 0] + solid["u"] * ramp[f[np.minimum(solid["prim"], nf - 1), 1], 0]
          + solid["v"] * ramp[f[np.minimum(solid["prim"], nf - 1), 2], 0])
    kept = hr["t"] >= 0
    assert np.array_equal(kept & ~on, np.zeros_like(on)) and (on & ~kept).any() and kept.any()
    for tri in range(nf):
        m = on & (solid["prim"] == tri)
        if (m & kept).any() and (m & ~kept).any():
            assert op[m & kept].min() >= op[m & ~kept].max() - 1e-6
