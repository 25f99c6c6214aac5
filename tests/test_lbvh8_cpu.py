"""CPU: the oracle's restatement of the DEVICE BVH8 builder (oracle/orc_lbvh8.h) produces a valid tree -- walked by
the oracle's traversal it gives bit-identical closest hits to the oracle's own binary BVH -- on the same inputs
the GPU parity test later feeds to the CUDA builder (tests/test_gpu_parity.py::test_device_bvh_build_parity)."""
import numpy as np

from gatling_b200 import gi, scenes
from oracle import orc
from helpers import random_rays


def rank_order(tris):
    """Triangle records of any BVH8 of a scene -> the builder's input order: ascending (instance, primitive)."""
    t = np.asarray(tris, np.float32).reshape(-1, 12)
    prim = t[:, 3].view(np.uint32).astype(np.uint64)
    inst = t[:, 7].view(np.uint32).astype(np.uint64)
    return t[np.argsort((inst << np.uint64(32)) | prim, kind="stable")]


def _check(desc, rays, mode=1):
    b = gi.BuiltScene(desc)
    try:
        _, host_tris = gi.get_bvh(b.scene)          # host builder: only the flattened triangle records are used
    finally:
        b.close()
    recs = rank_order(host_tris)
    nodes, tris, depth = orc.lbvh8_build(recs, mode)
    n = len(recs)
    assert len(tris) == n * 12 and len(nodes) % 80 == 0
    # every input triangle is referenced exactly once, records unchanged
    assert np.array_equal(rank_order(tris), recs)
    o = orc.OracleScene(desc)
    ref = o.trace(rays)
    walk, visits = orc.bvh8_trace(nodes, tris, rays)
    for k in ("prim", "inst", "t", "u", "v"):
        assert np.array_equal(ref[k], walk[k]), k
    a_ref = o.trace(rays, any_hit=True)
    a_walk, _ = orc.bvh8_trace(nodes, tris, rays, any_hit=True)
    assert np.array_equal(a_ref["t"] >= 0, a_walk["t"] >= 0)
    return nodes, tris, depth, visits


def test_cornell():
    d = scenes.cornell(32, 32, spp=1)
    cam = orc.camera_rays(d, 0); cam[:, 3] = 0; cam[:, 7] = 3e38
    nodes, tris, depth, _ = _check(d, np.concatenate([cam, random_rays(5000, 1)]))
    assert 1 <= depth <= 8


def test_random_triangles_and_determinism():
    d = scenes.random_triangles(30000, 32, 32, spp=1)
    nodes, tris, depth, visits = _check(d, np.concatenate([random_rays(20000, 2), random_rays(5000, 4, tmax=0.1)]))
    assert visits[:, 0].mean() > 5
    assert len(nodes) // 80 < 30000 // 2        # the collapse really produces wide nodes
    assert depth <= 24                          # far below the traversal stack (48)
    b = gi.BuiltScene(d)
    try:
        _, host_tris = gi.get_bvh(b.scene)
    finally:
        b.close()
    nodes2, tris2, _ = orc.lbvh8_build(rank_order(host_tris))
    assert np.array_equal(nodes, nodes2) and np.array_equal(tris, tris2)


def test_instances_duplicates_and_tiny_scenes():
    d = scenes.cornell(16, 16, spp=1)
    box = d.meshes[6]
    inst = []
    for k in range(5):
        m = np.eye(4, dtype=np.float32); m[3, :3] = (0.3 * k - 0.6, -0.5, 0.1 * k)
        inst.append(m)
    inst.append(inst[2].copy())                 # two coincident instances: identical Morton keys, tie on (inst, prim)
    box.instance_transforms = np.stack(inst)
    _check(d, random_rays(20000, 9))
    for count in (1, 2, 3, 4, 9):
        t = scenes.random_triangles(count, 8, 8, spp=1)
        nodes, tris, depth, _ = _check(t, random_rays(4000, 20 + count))
        assert depth == (1 if count <= 8 else 2) or count > 3
    nodes, tris, depth = orc.lbvh8_build(np.zeros((0, 12), np.float32))
    assert len(nodes) == 80 and len(tris) == 0 and depth == 0


def test_ploc_trees_are_valid_and_better():
    """bvhBuilder = 2 (PLOC): same validity checks; on structured geometry (tessellated surfaces) the locally-ordered
    clustering needs fewer node visits than the Karras radix tree over the same Morton order."""
    d = scenes.material_stress(tris_per_mesh=1500, meshes=12, width=32, height=32, spp=1)
    d.dome_light = None
    cam = orc.camera_rays(d, 0); cam[:, 3] = 0; cam[:, 7] = 3e38
    rays = np.concatenate([cam, random_rays(8000, 2, scale=3.0)])
    _, _, depth_l, visits_l = _check(d, rays, mode=1)
    nodes, tris, depth_p, visits_p = _check(d, rays, mode=2)
    assert depth_p <= 24
    assert visits_p[:, 0].mean() < visits_l[:, 0].mean()
    _check(scenes.random_triangles(20000, 32, 32, spp=1), np.concatenate([random_rays(10000, 2), random_rays(3000, 4, tmax=0.1)]), mode=2)
    c = scenes.cornell(16, 16, spp=1)
    cam = orc.camera_rays(c, 0); cam[:, 3] = 0; cam[:, 7] = 3e38
    _check(c, np.concatenate([cam, random_rays(3000, 5)]), mode=2)
    for count in (1, 2, 3, 4, 9, 40):
        _check(scenes.random_triangles(count, 8, 8, spp=1), random_rays(3000, 30 + count), mode=2)


def with_non_finite_triangles(n=3000, bad=200, seed=5):
    """random_triangles(n) with `bad` triangles given a NaN / infinite / absurdly large coordinate; returns (scene, victims)."""
    d = scenes.random_triangles(n, 8, 8, spp=1)
    m = d.meshes[0]
    v = m.vertices.copy()
    rng = np.random.default_rng(seed)
    victims = rng.choice(len(m.faces), bad, replace=False)
    specials = np.array([np.nan, np.inf, -np.inf, 3e38, -3e38, 1e30], np.float32)
    for k, f in enumerate(victims):
        v["pos"][m.faces[f][k % 3], k % 3] = specials[k % len(specials)]
    m.vertices = v
    return d, victims


def test_non_finite_triangles_are_inactive_in_the_device_builders():
    """The restated device builders (LBVH and PLOC) with inactive triangles in the input: a valid tree of sane depth whose
    walk never reports a victim and finds every other triangle exactly as the scene without the victims does."""
    d, victims = with_non_finite_triangles()
    b = gi.BuiltScene(d)
    try:
        _, host_tris = gi.get_bvh(b.scene)
    finally:
        b.close()
    recs = rank_order(host_tris)
    rays = np.ascontiguousarray(random_rays(20000, 9), np.float32)
    m = d.meshes[0]
    keep = np.setdiff1d(np.arange(len(m.faces)), victims)
    clean = scenes.random_triangles(3000, 8, 8, spp=1)
    clean.meshes[0].faces = clean.meshes[0].faces[keep]
    ref = orc.OracleScene(clean).trace(rays)
    for mode in (1, 2):
        nodes, tris, depth = orc.lbvh8_build(recs, mode)
        assert np.array_equal(rank_order(tris), recs, equal_nan=True) and depth <= 24
        assert np.isfinite(nodes.view(np.float32).reshape(-1, 20)[:, :3]).all()       # grid origins: no NaN reached a node
        hits, _ = orc.bvh8_trace(nodes, tris, rays)
        hit = hits["t"] >= 0
        assert not np.isin(hits["prim"][hit], victims).any()
        assert np.array_equal(hits["t"], ref["t"]) and np.array_equal(hits["prim"][hit], keep[ref["prim"][hit]])
