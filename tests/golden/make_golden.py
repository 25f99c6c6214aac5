#!/usr/bin/env python
"""Generates the committed golden fixtures.  Run HERE (the build container), never on the GPU box:

  python tests/golden/make_golden.py

* scene_hashes.json -- SHA-256 of the Cornell-box scene arrays as parsed from the reference's own
  /root/reference/cornell.usda by gatling_b200/usda.py (pins scenes.cornell() to the reference asset).
* kat.json          -- known-answer vectors of the pure functions of common.glsl, produced by the CPU
  oracle.  The reference has no numeric KATs (parity unpinned), so these freeze the
  oracle's outputs; integer functions are additionally cross-checked against an independent numpy
  restatement in tests/test_oracle_kat.py.
* cornell_oracle_32.npy -- float32 32x32 spp 4 oracle render of C1 (regression anchor for the oracle).
* reference_lfs.json -- the Git-LFS pointers (sha256, size) of the reference's golden test images.
* feature_anchors.npz  -- 32x32 spp 4 oracle renders of the texture / primvar / volume / cutout test scenes plus
  their ray counts: the -m gpu suite compares the CUDA path against these committed images as well as against a
  fresh oracle run, so an accidental change of the oracle itself cannot go unnoticed.
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gatling_b200 import scenes, usda  # noqa: E402
from oracle import orc  # noqa: E402


def scene_hashes(desc):
    out = {}
    for m in desc.meshes:
        h = hashlib.sha256()
        h.update(m.vertices.tobytes()); h.update(np.ascontiguousarray(m.faces, np.uint32).tobytes())
        h.update(np.ascontiguousarray(m.transform, np.float32).tobytes()); h.update(bytes([m.double_sided, m.left_handed, m.material]))
        out[m.name] = h.hexdigest()
    for m in desc.materials:
        out["material:" + m.name] = hashlib.sha256(np.ascontiguousarray(m.params, np.float32).tobytes() + bytes([m.closure])).hexdigest()
    c = desc.camera
    out["camera"] = hashlib.sha256(np.array(list(c.position) + list(c.forward) + list(c.up) + [c.vfov, c.focalLength, c.clipStart, c.clipEnd], np.float32).tobytes()).hexdigest()
    return out


def f3(v):
    return (C.c_float * 3)(*v)


def kat():
    L = orc.lib()
    k = {}
    xs = [0, 1, 2, 3, 0xFFFFFFFF, 0x12345678, 0xdeadbeef, 262144, 262145 * 65]
    k["hash_theironborn"] = [[x, L.orc_hash_theironborn(x)] for x in xs]
    pcg = []
    for x in xs:
        s = C.c_uint32(x)
        r = L.orc_hash_pcg32(C.byref(s))
        pcg.append([x, s.value, r])
    k["hash_pcg32"] = pcg
    k["uint_as_float"] = [[x, float(L.orc_uint_as_float(x))] for x in xs]
    k["rng_init"] = [[p, s, L.orc_rng_init(p, s)] for p, s in [(0, 0), (1, 0), (1, 1), (2, 0), (262143, 63), (2073599, 255), (12345, 678)]]
    seq = []
    for seed in (0, 1, 0xabcdef01):
        s = C.c_uint32(seed)
        vals = [float(L.orc_rng_next1f(C.byref(s))) for _ in range(8)]
        seq.append([seed, vals, s.value])
    k["rng_next1f_sequence"] = seq
    dirs = [(0, 0, 1), (0, 0, -1), (1, 0, 0), (0.6, -0.48, 0.64), (-0.267261237, 0.534522474, -0.801783681), (1e-3, 1e-3, -0.999999)]
    ob, enc, henc, dec = [], [], [], []
    for d in dirs:
        b1, b2 = (C.c_float * 3)(), (C.c_float * 3)()
        L.orc_orthonormal_basis(f3(d), b1, b2)
        ob.append([list(d), list(b1), list(b2)])
        e = L.orc_encode_direction(f3(d)); enc.append([list(d), e])
        h = L.orc_host_encode_direction(f3(d)); henc.append([list(d), h])
        o = (C.c_float * 3)(); L.orc_decode_direction(e, o); dec.append([e, list(o)])
    k["orthonormal_basis"] = ob
    k["encode_direction"] = enc
    k["host_encode_direction"] = henc
    k["decode_direction"] = dec
    off = []
    for p, n in [((0.5, -0.25, 2.0), (0, 0, 1)), ((0.001, 0.01, -0.02), (0.6, 0.0, -0.8)), ((-100.0, 50.0, 1e-4), (-0.57735, 0.57735, 0.57735)), ((0.03125, -0.03125, 0.03124), (1, 0, 0))]:
        o = (C.c_float * 3)(); L.orc_offset_ray_origin(f3(p), f3(n), o); off.append([list(p), list(n), list(o)])
    k["offset_ray_origin"] = off
    xis = [(0.0, 0.0), (0.5, 0.5), (0.25, 0.75), (0.999, 0.001), (0.1, 0.9)]
    sh, ss, sd, fg = [], [], [], []
    for xi in xis:
        x2 = (C.c_float * 2)(*xi)
        o = (C.c_float * 3)(); L.orc_sample_hemisphere(x2, o); sh.append([list(xi), list(o)])
        o = (C.c_float * 3)(); L.orc_sample_sphere(x2, f3((0.5, 0.25, 2.0)), o); ss.append([list(xi), list(o)])
        o2 = (C.c_float * 2)(); L.orc_sample_disk(x2, (C.c_float * 2)(0.5, 0.25), o2); sd.append([list(xi), list(o2)])
        o2 = (C.c_float * 2)(); L.orc_fis_gauss(x2, o2); fg.append([list(xi), list(o2)])
    k["sample_hemisphere"] = sh; k["sample_sphere_r(0.5,0.25,2)"] = ss; k["sample_disk_r(0.5,0.25)"] = sd; k["fis_gauss"] = fg
    k["luminance"] = [[list(c), float(L.orc_luminance(f3(c)))] for c in [(1, 1, 1), (0.2, 0.5, 0.3), (8.5, 6, 4)]]
    k["pack_half2x16"] = [[a, b, L.orc_pack_half2x16(a, b)] for a, b in [(1.0, 1.0), (0.1, 100.0), (0.0, 65504.0), (1e-5, -2.5), (0.333333, 3.14159)]]
    tri = []
    for o_, d_, p0, p1, p2 in [((0, 0, -1), (0, 0, 1), (-1, -1, 0), (1, -1, 0), (0, 1, 0)),
                               ((0.1, 0.2, -3), (0.05, -0.02, 1), (-1, -1, 0.5), (2, -1, 0.25), (0, 3, -0.25)),
                               ((0, 0, -1), (0, 0, 1), (0, 0, 0), (1, 0, 0), (0, 1, 0)),      # exactly through a vertex
                               ((0.5, 0, -1), (0, 0, 1), (0, 0, 0), (1, 0, 0), (0, 1, 0)),    # exactly on an edge
                               ((5, 5, -1), (0, 0, 1), (0, 0, 0), (1, 0, 0), (0, 1, 0))]:     # miss
        tuv = (C.c_float * 3)()
        hit = L.orc_tri_intersect(f3(o_), f3(d_), f3(p0), f3(p1), f3(p2), tuv)
        tri.append([list(o_), list(d_), list(p0), list(p1), list(p2), hit, list(tuv) if hit else None])
    k["tri_intersect"] = tri
    return k


def feature_scenes():
    import dataclasses
    lit = scenes.cornell(32, 32, spp=4, max_bounces=5)
    lit.settings = dataclasses.replace(lit.settings, nextEventEstimation=True)
    lit.rect_lights.append(scenes.RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(12, 12, 12), width=0.6, height=0.6))
    lit.sphere_lights.append(scenes.SphereLight(position=(0.5, -0.5, 0.2), base_emission=(4, 3, 2), radius=(0.1, 0.1, 0.1)))
    return {"textured": scenes.textured_cornell(32, 32, spp=4), "primvars": scenes.primvar_cornell(32, 32, spp=4),
            "volumes": scenes.volume_cornell(32, 32, spp=4, medium_stack_size=2), "nee": lit,
            "stress": scenes.material_stress(tris_per_mesh=200, meshes=12, width=32, height=32, spp=4),
            "lookups": scenes.lookup_cornell(32, 32, spp=4), "layered": scenes.layered_cornell(32, 32, spp=4)}


def feature_anchors():
    out = {}
    for name, d in feature_scenes().items():
        img, _, rays = orc.OracleScene(d).render()
        out[name] = img
        out[name + "_rays"] = np.array(rays, np.int64)
    return out


def reference_lfs_pointers():
    """sha256 + size of every golden image of the reference's graphical tests (src/hdGatling/testenv/*/ref*.png).  The repository
    snapshot holds Git-LFS pointers only; a pointer still pins the file bit for bit (tests/test_reference_pin.py reproduces the
    one small enough to be reconstructed, Render.Empty1x1/ref.png)."""
    base = "/root/reference/src/hdGatling/testenv"
    out = {}
    for case in sorted(os.listdir(base)):
        d = os.path.join(base, case)
        if not os.path.isdir(d):
            continue
        for f in sorted(os.listdir(d)):
            if f.startswith("ref") and f.endswith(".png"):
                with open(os.path.join(d, f), "rb") as fh:
                    head = fh.read(200).decode("ascii", "replace")
                if head.startswith("version https://git-lfs"):
                    oid = [ln.split(":", 1)[1].strip() for ln in head.splitlines() if ln.startswith("oid sha256:")][0]
                    size = [int(ln.split()[1]) for ln in head.splitlines() if ln.startswith("size ")][0]
                    out["%s/%s" % (case, f)] = {"sha256": oid, "size": size}
    return out


def main():
    with open(os.path.join(HERE, "reference_lfs.json"), "w") as f:
        json.dump(reference_lfs_pointers(), f, indent=1, sort_keys=True)
    if "--lfs-only" in sys.argv:
        return
    if "--add-missing-anchors" in sys.argv:        # a new feature scene: the committed anchors stay byte for byte, only the missing ones are rendered
        path = os.path.join(HERE, "feature_anchors.npz")
        have = dict(np.load(path))
        for name, d_ in feature_scenes().items():
            if name not in have:
                img, _, rays = orc.OracleScene(d_).render()
                have[name] = img
                have[name + "_rays"] = np.array(rays, np.int64)
                print("added anchor", name)
        np.savez_compressed(path, **have)
        return
    ref = "/root/reference/cornell.usda"
    desc = usda.load(ref)
    with open(os.path.join(HERE, "scene_hashes.json"), "w") as f:
        json.dump({"source": ref, "hashes": scene_hashes(desc)}, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat(), f, indent=1)
    d = scenes.cornell(32, 32, spp=4, max_bounces=4)
    img, _, rays = orc.OracleScene(d).render()
    np.save(os.path.join(HERE, "cornell_oracle_32.npy"), img)
    with open(os.path.join(HERE, "cornell_oracle_32.json"), "w") as f:
        json.dump({"closest_rays": rays[0], "shadow_rays": rays[1]}, f)
    np.savez_compressed(os.path.join(HERE, "feature_anchors.npz"), **feature_anchors())
    print("golden fixtures written")


if __name__ == "__main__":
    main()
