import sys, os, dataclasses
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gatling_b200 import gi, scenes
from oracle import orc
import test_gpu_parity as T
assert gi.giInitialize(0) == gi.GiStatus.Ok
for nee in (False, True):
    for bounces in (1, 2, 5):
        d = T._cutout_cornell(64, 64, 4)
        d.settings = dataclasses.replace(d.settings, nextEventEstimation=nee, maxBounces=bounces)
        b = gi.BuiltScene(d); img = b.render().copy(); st = gi.get_stats(); b.close()
        ref, _, rays = orc.OracleScene(d).render()
        print("nee", nee, "bounces", bounces, "rmse %.3e" % np.sqrt(np.mean((img[..., :3] - ref[..., :3]) ** 2)), "rays gpu", st["closestRays"], st["shadowRays"], "oracle", rays, flush=True)
