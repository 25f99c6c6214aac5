#!/usr/bin/env python
"""Runs the non-headline configs of BASELINE.json (C1 cornell, C3 kitchen stand-in, C5 material stress) on one GPU at a
reduced sample count and prints one JSON line each: Mray/s, kernel times, BVH build time, sort on/off."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatling_b200 import gi, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="c1,c3,c5")
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--scale", type=float, default=1.0, help="triangle-count scale of c3/c5")
ap.add_argument("--bvh-builder", type=int, default=0, help="0 host binned SAH, 1 device LBVH, 2 device PLOC")
ap.add_argument("--trace-variant", type=int, default=0)
ap.add_argument("--ups", action="store_true", help="replace the OpenPBR materials of c3/c5 by UsdPreviewSurface ones (same colours / roughness / metalness)")
ap.add_argument("--tri-threshold", type=int, default=0)
ap.add_argument("--textured", action="store_true", help="c3/c5: planar uvs on every mesh and, on every OpenPBR material, a 1024x1024 sRGB RGBA8 image on base_color + a 512x512 R32F image on specular_roughness")
ap.add_argument("--refill-sweep", default="", help="comma list of refill thresholds (idle lanes at which a warp fetches new rays) measured after the normal lines, sort as in the last of them")
a = ap.parse_args()
assert gi.giInitialize(0) == gi.GiStatus.Ok, gi.giLastError()


def make(name):
    if name == "c1":
        return scenes.cornell(512, 512, spp=64, max_bounces=4)
    if name == "c3":
        return scenes.kitchen_standin(tris_per_instance=int(12500 * a.scale), spp=a.spp)
    if name == "c5":
        return scenes.material_stress(tris_per_mesh=int(31250 * a.scale), spp=a.spp)
    raise SystemExit("unknown config " + name)


for name in a.configs.split(","):
    t0 = time.perf_counter()
    d = make(name)
    if a.ups:
        O, U = scenes.OPBR, scenes.UPS
        d.materials = [m if m.closure != scenes.CLOSURE_OPENPBR else scenes.material_usd_preview_surface(m.name, diffuse=tuple(m.params[O["base_color"]:O["base_color"] + 3]),
                       roughness=float(m.params[O["specular_roughness"]]), metallic=float(m.params[O["base_metalness"]]), clearcoat=float(m.params[O["coat_weight"]])) for m in d.materials]
    if a.textured:
        for m in d.meshes:
            scenes.assign_planar_uv(m)
        d.textures = [scenes.procedural_texture("checker", 1024), scenes.procedural_texture("roughness", 512)]
        for m in d.materials:
            if m.closure == scenes.CLOSURE_OPENPBR:
                m.textures = [scenes.TextureBinding(scenes.OPBR["base_color"], 0, (0, 1, 2), srgb=True, uv_scale=(4.0, 4.0)),
                              scenes.TextureBinding(scenes.OPBR["specular_roughness"], 1, (0,), scale=(0.6, 1, 1, 1), bias=(0.05, 0, 0, 0), uv_scale=(3.0, 3.0))]
    t_gen = time.perf_counter() - t0
    import ctypes as C
    gi.set_tuning(bvhBuilder=a.bvh_builder, traceVariant=a.trace_variant, reserved=(C.c_int32 * 3)(0, 0, a.tri_threshold))
    b = gi.BuiltScene(d)
    gi.set_tuning(countVisits=1)                     # untimed counting pass: exact visit counts of the identical rays (algorithmic bytes)
    b.render()
    cs = gi.get_stats()
    gi.set_tuning(countVisits=0)
    visits = {"closest_nodes_per_ray": cs["nodeVisits"] / max(cs["closestRays"], 1), "closest_tris_per_ray": cs["triVisits"] / max(cs["closestRays"], 1),
              "shadow_nodes_per_ray": cs["shadowNodeVisits"] / max(cs["shadowRays"], 1), "shadow_tris_per_ray": cs["shadowTriVisits"] / max(cs["shadowRays"], 1),
              "closest_hits": cs["closestHits"], "node_visits": cs["nodeVisits"], "tri_visits": cs["triVisits"],
              "shadow_node_visits": cs["shadowNodeVisits"], "shadow_tri_visits": cs["shadowTriVisits"]}
    for sort in ((-1,) if len(d.materials) == 1 else (0, 1)):
        gi.set_tuning(sortByMaterial=sort)
        best = None
        for rep in range(3):
            gi.set_shard(b.scene, 0, 1, 64)
            b.render()
            s = gi.get_stats()
            if best is None or s["msTotal"] < best["msTotal"]:
                best = s
        rays = best["closestRays"] + best["shadowRays"]
        print(json.dumps({"config": name, "scene": d.name, "triangles": d.triangle_count, "materials": len(d.materials), "spp": d.settings.spp,
                          "sort": sort, "textured": bool(a.textured), "Mray_s": rays / best["msTotal"] / 1e3, "ms_total": best["msTotal"], "closest": best["closestRays"],
                          "shadow": best["shadowRays"], "ms_trace": best["msTraceClosest"], "ms_shadow": best["msTraceAny"], "ms_shade": best["msShade"],
                          "ms_sort": best["msSort"], "bvh_builder": a.bvh_builder, "trace_variant": a.trace_variant, "bvh_build_ms": best["msBvhBuild"], "bvh_device_build_ms": best["msBvhDeviceBuild"], "bvh_nodes": best["bvhNodeCount"], "depth": best["bvhMaxDepth"],
                          "scene_gen_s": round(t_gen, 2), "launches": best["kernelLaunches"], **visits}), flush=True)
    for rf in [int(x) for x in a.refill_sweep.split(",") if x]:
        gi.set_tuning(reserved=(C.c_int32 * 3)(rf, 0, a.tri_threshold))
        best = None
        for rep in range(3):
            gi.set_shard(b.scene, 0, 1, 64)
            b.render()
            s = gi.get_stats()
            if best is None or s["msTotal"] < best["msTotal"]:
                best = s
        print(json.dumps({"config": name, "refill": rf, "ms_total": best["msTotal"], "ms_trace": best["msTraceClosest"], "ms_shadow": best["msTraceAny"], "ms_shade": best["msShade"]}), flush=True)
    b.close()
gi.giTerminate()
