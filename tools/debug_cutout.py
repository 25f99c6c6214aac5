import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from gatling_b200 import gi, scenes
from oracle import orc
from helpers import random_rays
assert gi.giInitialize(0) == gi.GiStatus.Ok
d = scenes.lookup_cornell(96, 96, spp=8)
cam = orc.camera_rays(d, 0); cam[:, 3] = 0.0; cam[:, 7] = 3.0e38
rays = np.concatenate([cam, random_rays(40000, 23)])
b = gi.BuiltScene(d)
o = orc.OracleScene(d)
for variant, visits in ((0, True), (0, False), (1, False)):
    gi.set_tuning(traceVariant=variant)
    h = gi.trace_rays(b.scene, rays, want_visits=visits)
    h = h[0] if visits else h
    r = o.trace(rays)
    bad = (h["prim"] != r["prim"]) | (h["inst"] != r["inst"])
    print("variant", variant, "visits", visits, "mismatch", int(bad.sum()), "of", len(rays))
    if bad.any():
        idx = np.nonzero(bad)[0][:8]
        for i in idx:
            print("  ray", i, "gpu", h["inst"][i], h["prim"][i], h["t"][i], "ref", r["inst"][i], r["prim"][i], r["t"][i])
        print("  gpu inst histogram of mismatches", np.unique(h["inst"][bad], return_counts=True))
        print("  ref inst histogram of mismatches", np.unique(r["inst"][bad], return_counts=True))
names = [m.name for m in d.meshes]
print(names, [m.material for m in d.meshes])
img = b.render().copy()
ref = o.render()[0]
print("image rmse", float(np.sqrt(np.mean((img[..., :3] - ref[..., :3]) ** 2))))
b.close()
