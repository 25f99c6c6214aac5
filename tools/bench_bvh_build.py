"""Host SAH vs device LBVH / PLOC BVH8 build: build time and the trace speed of the resulting tree (random triangle soup)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from gatling_b200 import gi, scenes
assert gi.giInitialize(0) == gi.GiStatus.Ok
for tris in (1_000_000, 5_000_000):
    d = scenes.random_triangles(tris, 1920, 1080, 8, 8)
    for builder in (0, 1, 1, 2, 2):
        gi.set_tuning(bvhBuilder=builder)
        t0 = time.time(); b = gi.BuiltScene(d); t1 = time.time()
        img = b.render(); s = gi.get_stats()
        rays = s["closestRays"] + s["shadowRays"]
        print("builder %d: scene create %.2f s, bvh build+upload %.1f ms (device build %.1f ms), nodes %d depth %d, %.1f Mray/s (trace %.2f ms)" % (builder, t1 - t0, s["msBvhBuild"], s["msBvhDeviceBuild"], s["bvhNodeCount"], s["bvhMaxDepth"], rays / s["msTotal"] / 1e3, s["msTraceClosest"]), flush=True)
        b.close()
