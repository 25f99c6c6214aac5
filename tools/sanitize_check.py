#!/usr/bin/env python
"""One small pass over every kernel family of libgtlc.so, meant to run under compute-sanitizer on the GPU box:

  compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck --error-exitcode 9 python tools/sanitize_check.py

Covers: raygen, the warp-batched traversal kernel (closest / any-hit / shadow, with and without evaluated cutouts), the
per-lane counting kernel, material sort, shade (diffuse, UsdPreviewSurface, OpenPBR + NEE, textures, primvars, media),
resolve / AOV resolve, the LBVH and PLOC device builders, gtlc_trace_rays.  Sizes are tiny: the sanitizer slows kernels 20-100x.
"""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatling_b200 import gi, scenes  # noqa: E402

W, H = 48, 36
assert gi.giInitialize(0) == gi.GiStatus.Ok, gi.giLastError()


def render(d, aovs=None, **tuning):
    if tuning:
        gi.set_tuning(**tuning)
    b = gi.BuiltScene(d)
    try:
        img = b.render().copy() if not aovs else b.render_aovs(aovs)
        rays = np.zeros((256, 8), np.float32)
        rng = np.random.default_rng(5)
        rays[:, 0:3] = rng.uniform(-1, 1, (256, 3)); rays[:, 3] = 0.0
        rays[:, 4:7] = rng.normal(size=(256, 3)); rays[:, 7] = 1e30
        gi.trace_rays(b.scene, rays)
        gi.trace_rays(b.scene, rays, any_hit=True)
    finally:
        b.close()
    return img


lit = scenes.cornell(W, H, spp=2, max_bounces=4)
lit.settings = dataclasses.replace(lit.settings, nextEventEstimation=True)
lit.rect_lights.append(scenes.RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
done = []
for name, d, kw in (
        ("cornell + NEE", lit, {}),
        ("per-lane counting kernel", lit, {"countVisits": 1}),
        ("LBVH build", lit, {"countVisits": 0, "bvhBuilder": 1}),
        ("PLOC build", lit, {"bvhBuilder": 2}),
        ("textures", scenes.textured_cornell(W, H, spp=2, max_bounces=4), {"bvhBuilder": 0}),
        ("3-D / texel lookups + evaluated cutout", scenes.lookup_cornell(W, H, spp=2, max_bounces=4), {}),
        ("primvars", scenes.primvar_cornell(W, H, spp=2, max_bounces=4), {}),
        ("inputs combined from several lookups", scenes.layered_cornell(W, H, spp=2, max_bounces=4), {}),
        ("media", scenes.volume_cornell(W, H, spp=2, max_bounces=6), {}),
        ("OpenPBR instances + sort", scenes.kitchen_standin(tris_per_instance=200, instances=12, meshes=6, materials=5, width=W, height=H, spp=2, max_bounces=4), {"sortByMaterial": 1}),
        ("triangle soup", scenes.random_triangles(3000, W, H, spp=2, max_bounces=4), {"sortByMaterial": -1})):
    img = render(d, **kw)
    assert np.isfinite(img).all(), name
    done.append(name)
img = render(lit, aovs=[gi.GiAovId.Color, gi.GiAovId.Normal, gi.GiAovId.Depth, gi.GiAovId.FaceId, gi.GiAovId.ClockCycles])
done.append("AOVs")
gi.giTerminate()
print("SANITIZE PASS:", "; ".join(done), flush=True)
