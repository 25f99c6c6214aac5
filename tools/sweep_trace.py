#!/usr/bin/env python
"""GPU tuning sweep: traversal kernel variants / knobs on config C2 at reduced spp.  Prints one line per
setting; all settings must produce the bit-identical image."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatling_b200 import gi, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=1_000_000)
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--variants", default="0,1,6")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--refill", default="0")
ap.add_argument("--ctas", default="0")
ap.add_argument("--trithr", default="0")
a = ap.parse_args()

assert gi.giInitialize(0) == gi.GiStatus.Ok, gi.giLastError()
d = scenes.random_triangles(a.tris, 1920, 1080, a.spp, 8)
b = gi.BuiltScene(d)
ref = None
gi.set_tuning(countVisits=1)
b.render()
s = gi.get_stats()
print("counting pass: rays %d  nodes/ray %.2f  tris/ray %.2f  max nodes of one ray %d  non-finite rays %d"
      % (s["closestRays"], s["nodeVisits"] / s["closestRays"], s["triVisits"] / s["closestRays"], s["maxNodeVisitsPerRay"], s["nonFiniteRays"]), flush=True)
gi.set_tuning(countVisits=0)
import ctypes as C
combos = [(v, rf, ct) for v in [int(x) for x in a.variants.split(",")] for rf in [int(x) for x in a.refill.split(",")] for ct in [int(x) for x in a.ctas.split(",")]]
combos = [(v, rf, ct, tt) for (v, rf, ct) in combos for tt in ([int(x) for x in a.trithr.split(",")] if v == 5 else [0])]
for v, rf, ct, tt in combos:
    gi.set_tuning(traceVariant=v, skipReadback=0, reserved=(C.c_int32 * 3)(rf, ct, tt))
    best = None
    for _ in range(a.reps):
        gi.set_shard(b.scene, 0, 1, 64)
        img = b.render().copy()
        s = gi.get_stats()
        if best is None or s["msTotal"] < best["msTotal"]:
            best = s
    if ref is None:
        ref = img
    same = bool(np.array_equal(ref, img))
    rays = best["closestRays"] + best["shadowRays"]
    print("variant %d refill %d ctas/SM %d triThr %d: %.1f Mray/s total  trace %.2f ms  shade %.2f ms  total %.2f ms  identical=%s depth=%d"
          % (v, rf, ct, tt, rays / best["msTotal"] / 1e3, best["msTraceClosest"], best["msShade"], best["msTotal"], same, best["bvhMaxDepth"]), flush=True)
b.close()
gi.giTerminate()
