#!/usr/bin/env python
"""How much does ray order matter to the traversal kernel?  Traces the same rays of config C2 in three orders:
camera order (coherent), randomly permuted, and sorted by (direction octant, Morton code of the origin)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gatling_b200 import gi, scenes  # noqa: E402
from oracle import orc  # noqa: E402  (camera_rays helper only)

assert gi.giInitialize(0) == gi.GiStatus.Ok
d = scenes.random_triangles(1_000_000, 1920, 1080, 1, 8)
b = gi.BuiltScene(d)
rng = np.random.default_rng(1)


def run(name, rays):
    best = None
    for _ in range(3):
        hits = gi.trace_rays(b.scene, rays)
        s = gi.get_stats()
        best = s["msTraceClosest"] if best is None else min(best, s["msTraceClosest"])
    print("%-34s %8d rays  %7.2f ms  %7.1f Mray/s  hit %.2f" % (name, len(rays), best, len(rays) / best / 1e3, (hits["t"] >= 0).mean()), flush=True)
    return hits


def morton_sort(rays, bits=7):
    o = rays[:, 0:3]
    lo, hi = o.min(0), o.max(0)
    q = np.clip(((o - lo) / np.maximum(hi - lo, 1e-9) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    code = np.zeros(len(rays), np.int64)
    for k in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> k) & 1) << (3 * k + a)
    octant = (rays[:, 4] < 0).astype(np.int64) | ((rays[:, 5] < 0).astype(np.int64) << 1) | ((rays[:, 6] < 0).astype(np.int64) << 2)
    return np.argsort((octant << (3 * bits)) | code, kind="stable")


cam = orc.camera_rays(d, 0)
cam[:, 3] = 0
cam[:, 7] = 3e38
h = run("camera rays, pixel order", cam)
run("camera rays, shuffled", cam[rng.permutation(len(cam))])
# secondary rays: diffuse bounce off the primary hit points
hit = h["t"] >= 0
p = cam[hit, 0:3] + cam[hit, 4:7] * h["t"][hit, None]
dirs = rng.normal(size=(hit.sum(), 3)).astype(np.float32)
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
sec = np.concatenate([p - cam[hit, 4:7] * 1e-3, np.full((hit.sum(), 1), 1e-4, np.float32), dirs, np.full((hit.sum(), 1), 3e38, np.float32)], 1).astype(np.float32)
run("bounce rays, path order", sec)
run("bounce rays, shuffled", sec[rng.permutation(len(sec))])
run("bounce rays, octant+Morton(7b) sorted", sec[morton_sort(sec, 7)])
run("bounce rays, octant+Morton(5b) sorted", sec[morton_sort(sec, 5)])
b.close()
