"""CPU: the product's host-built BVH8 (through the C-ABI, no GPU needed) walked by the oracle's
restatement of the traversal gives bit-identical closest hits to the oracle's own binary BVH."""
import numpy as np
import pytest

from gatling_b200 import gi, scenes
from oracle import orc
from helpers import random_rays


def _check(desc, rays):
    b = gi.BuiltScene(desc)
    try:
        nodes, tris = gi.get_bvh(b.scene)
        assert len(nodes) % 80 == 0 and len(tris) % 12 == 0
        assert len(tris) // 12 == desc.triangle_count
        o = orc.OracleScene(desc)
        ref = o.trace(rays)
        walk, visits = orc.bvh8_trace(nodes, tris, rays)
        for k in ("prim", "inst", "t", "u", "v"):
            assert np.array_equal(ref[k], walk[k]), k
        a_ref = o.trace(rays, any_hit=True)
        a_walk, _ = orc.bvh8_trace(nodes, tris, rays, any_hit=True)
        assert np.array_equal(a_ref["t"] >= 0, a_walk["t"] >= 0)
        return nodes, tris, visits
    finally:
        b.close()


def test_cornell_bvh8():
    d = scenes.cornell(32, 32, spp=1)
    cam = orc.camera_rays(d, 0); cam[:, 3] = 0; cam[:, 7] = 3e38
    nodes, tris, visits = _check(d, np.concatenate([cam, random_rays(5000, 1)]))
    assert len(tris) // 12 == 46


def test_random_triangles_bvh8():
    d = scenes.random_triangles(30000, 32, 32, spp=1)
    nodes, tris, visits = _check(d, np.concatenate([random_rays(20000, 2), random_rays(5000, 4, tmax=0.1)]))
    assert visits[:, 0].mean() > 5       # the walk really descends
    # every triangle is referenced exactly once
    prim = tris.reshape(-1, 12)[:, 3].view(np.uint32)
    assert np.array_equal(np.sort(prim), np.arange(30000, dtype=np.uint32))


def test_instanced_and_transformed_meshes():
    d = scenes.cornell(16, 16, spp=1)
    box = d.meshes[6]
    xf = np.eye(4, dtype=np.float32)
    inst = []
    for k in range(5):
        m = xf.copy(); m[3, :3] = (0.3 * k - 0.6, -0.5, 0.1 * k)
        inst.append(m)
    box.instance_transforms = np.stack(inst)
    _check(d, random_rays(20000, 9))


def test_degenerate_inputs():
    d = scenes.random_triangles(64, 8, 8, spp=1)
    m = d.meshes[0]
    v = m.vertices.copy()
    v["pos"][0:3] = v["pos"][0]          # zero-area triangle
    v["pos"][3:6] = [[0, 0, 0], [1, 1, 1], [2, 2, 2]]   # collinear
    v["pos"][6:12] = v["pos"][12:18]     # duplicated (coincident) triangles -> tie break on primitive id
    m.vertices = v
    _check(d, random_rays(20000, 13))
    # a scene without geometry
    e = scenes.cornell(8, 8, spp=1)
    e.meshes = []
    b = gi.BuiltScene(e)
    try:
        nodes, tris = gi.get_bvh(b.scene)
        assert len(nodes) == 80 and len(tris) == 0
        h, _ = orc.bvh8_trace(nodes, tris, random_rays(100, 1))
        assert np.all(h["t"] == -1)
    finally:
        b.close()


def test_host_build_is_independent_of_the_thread_count(tmp_path):
    """The builder runs as OpenMP tasks with atomically allocated node slots; the emitted BVH8 (breadth-first order from the
    root) must not depend on how many threads took part."""
    import hashlib
    import os
    import subprocess
    import sys
    script = tmp_path / "h.py"
    script.write_text(
        "import sys, hashlib\n"
        "sys.path.insert(0, %r)\n"
        "from gatling_b200 import gi, scenes\n"
        "for d in (scenes.random_triangles(300000, 8, 8, 1), scenes.kitchen_standin(tris_per_instance=400, instances=60, meshes=12, materials=4, width=8, height=8, spp=1)):\n"
        "    d.dome_light = None\n"
        "    b = gi.BuiltScene(d); n, t = gi.get_bvh(b.scene); b.close()\n"
        "    print(hashlib.sha256(n.tobytes() + t.tobytes()).hexdigest())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for threads in ("1", "5"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        outs.append(subprocess.check_output([sys.executable, str(script)], env=env, text=True))
    assert outs[0] == outs[1] and len(outs[0].split()) == 2


def test_host_builder_output_is_pinned():
    """The host builder was restructured for speed several times under the rule "same bytes out"; these two trees pin it.
    (An intended change of the tree -- a new SAH heuristic -- updates the hashes together with DESIGN.md.)"""
    import hashlib
    want = {1000: "9e6f71296d08a507a07e5836e05fec4a47e48620ef2868939eafd47abf056885",
            20000: "adfabed4cd37a402d1da5d8de01d3805b7980685e450c058fa613388c08373bd",
            300000: "bfcc74ef60a9e5325835051bda411eddc07d229130825babb84c7ff6635714fb"}     # large enough for the team-wide binning / partition passes
    for n, digest in want.items():
        b = gi.BuiltScene(scenes.random_triangles(n, 8, 8, 1))
        try:
            nodes, tris = gi.get_bvh(b.scene)
        finally:
            b.close()
        assert hashlib.sha256(nodes.tobytes() + tris.tobytes()).hexdigest() == digest


def test_analysis_walkers_bound_the_production_walk():
    """The oracle's analysis walkers (ideal near-to-far order, culling deferred node groups at pop) only count visits; they
    must never need MORE node visits than the production order they are compared with in DESIGN.md section 8."""
    import ctypes as C
    d = scenes.random_triangles(40000, 16, 16, spp=1)
    b = gi.BuiltScene(d)
    try:
        nodes, tris = gi.get_bvh(b.scene)
    finally:
        b.close()
    rays = np.ascontiguousarray(random_rays(20000, 3), np.float32)
    _, visits = orc.bvh8_trace(nodes, tris, rays)
    L = orc.lib()
    L.orc_bvh8_trace_ordered.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
    L.orc_bvh8_trace_groupcull.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
    production = int(visits[:, 0].sum())
    counts = {}
    for name, fn, mode in (("queue", L.orc_bvh8_trace_ordered, 0), ("sorted", L.orc_bvh8_trace_ordered, 1), ("group", L.orc_bvh8_trace_groupcull, 0),
                           ("group1", L.orc_bvh8_trace_groupcull, 2), ("all", L.orc_bvh8_trace_groupcull, 1)):
        out = (C.c_uint64 * 2)()
        fn(nodes.ctypes.data, tris.ctypes.data, rays.ctypes.data, len(rays), out, mode)
        counts[name] = int(out[0])
    assert counts["queue"] <= counts["sorted"] <= production
    assert counts["queue"] <= counts["group"] <= counts["group1"] <= counts["all"] <= production
    assert counts["group"] < production                  # culling at pop saves node fetches whenever rays hit something
    # where the wasted node fetches come from: grid-box miss <= union-of-children miss <= no child hit, and the purely
    # geometric part (the parent's quantised box is looser than the node's own extent) is small
    L.orc_bvh8_grid_earlyout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    out = (C.c_uint64 * 5)()
    L.orc_bvh8_grid_earlyout(nodes.ctypes.data, tris.ctypes.data, rays.ctypes.data, len(rays), out)
    visited, grid, empty, union, union_geom = (int(x) for x in out)
    assert visited == production
    assert grid <= union <= empty < visited
    assert union_geom <= union and union_geom < 0.05 * visited


def test_non_finite_triangles_are_inactive():
    """Triangles with NaN / infinite / absurdly large coordinates must neither crash the builder nor disturb the tree
    around them (Vulkan calls them inactive, and the watertight test can never report them): the finite triangles are
    found exactly as in the scene without the bad ones."""
    d = scenes.random_triangles(3000, 8, 8, spp=1)
    m = d.meshes[0]
    good = m.vertices.copy()
    bad = good.copy()
    rng = np.random.default_rng(5)
    victims = rng.choice(len(m.faces), 200, replace=False)
    specials = np.array([np.nan, np.inf, -np.inf, 3e38, -3e38, 1e30], np.float32)
    for k, f in enumerate(victims):
        v = m.faces[f][k % 3]
        bad["pos"][v, k % 3] = specials[k % len(specials)]
    m.vertices = bad
    b = gi.BuiltScene(d)
    try:
        nodes, tris = gi.get_bvh(b.scene)
    finally:
        b.close()
    assert len(tris) == 3000 * 12
    rays = np.ascontiguousarray(random_rays(20000, 9), np.float32)
    hits, _ = orc.bvh8_trace(nodes, tris, rays)
    assert not np.isin(hits["prim"][hits["t"] >= 0], victims).any()
    # reference: the same scene with the bad triangles removed (primitive ids keep their meaning through faceIds order)
    keep = np.setdiff1d(np.arange(len(m.faces)), victims)
    m.vertices = good
    m.faces = m.faces[keep]
    b = gi.BuiltScene(d)
    try:
        nodes2, tris2 = gi.get_bvh(b.scene)
    finally:
        b.close()
    hits2, _ = orc.bvh8_trace(nodes2, tris2, rays)
    assert np.array_equal(hits["t"], hits2["t"])
    hit = hits["t"] >= 0
    assert np.array_equal(hits["prim"][hit], keep[hits2["prim"][hit]])


def _scene_with_extra_triangles(n_random, extra_pos):
    """random_triangles(n_random) plus the triangles given as (k, 3, 3) positions, appended to the same mesh."""
    d = scenes.random_triangles(n_random, 8, 8, spp=1)
    m = d.meshes[0]
    nv = len(m.vertices)
    extra = np.zeros(len(extra_pos) * 3, m.vertices.dtype)
    extra[:] = m.vertices[0]
    extra["pos"] = np.asarray(extra_pos, np.float32).reshape(-1, 3)
    m.vertices = np.concatenate([m.vertices, extra])
    m.faces = np.concatenate([m.faces, (nv + np.arange(len(extra_pos) * 3, dtype=m.faces.dtype)).reshape(-1, 3)])
    return d


@pytest.mark.parametrize("case", ["collinear_in_soup", "points_on_a_line", "geometric_progression"])
def test_degenerate_clusters_stay_shallow(case):
    """Zero-area clusters make every SAH candidate cost 0 (the strict comparison alone peels 1/16 of the range per level),
    and a geometric progression of sizes makes SAH peel ONE triangle per level: both used to exceed the traversal stack
    and gtlc_scene_get_bvh refused the scene.  The reference's driver BVH renders such meshes."""
    if case == "collinear_in_soup":
        x = np.linspace(-0.9, 0.9, 200, dtype=np.float32)
        pos = np.zeros((200, 3, 3), np.float32)
        pos[:, 0, 0] = x; pos[:, 1, 0] = x + 0.001; pos[:, 2, 0] = x + 0.002     # segments on the x axis: zero-area boxes
        d = _scene_with_extra_triangles(50000, pos)
    elif case == "points_on_a_line":
        x = np.linspace(-1, 1, 100000, dtype=np.float32)
        pos = np.zeros((100000, 3, 3), np.float32)
        pos[:, :, 1] = x[:, None]
        d = _scene_with_extra_triangles(16, pos)
    else:
        k = np.arange(400)
        s = (2.0 ** (-k * 0.25)).astype(np.float32)                              # each triangle 16 % smaller, nested at one corner
        pos = np.zeros((400, 3, 3), np.float32)
        pos[:, 1, 0] = s; pos[:, 2, 1] = s; pos[:, :, 2] = (k * 1e-4)[:, None]
        d = _scene_with_extra_triangles(16, pos)
    b = gi.BuiltScene(d)
    try:
        nodes, tris = gi.get_bvh(b.scene)          # raises when the builder refuses the scene
    finally:
        b.close()
    assert len(tris) // 12 == d.triangle_count
    rays = np.ascontiguousarray(random_rays(4000, 21), np.float32)
    o = orc.OracleScene(d)
    ref = o.trace(rays)
    walk, visits = orc.bvh8_trace(nodes, tris, rays)
    for key in ("prim", "inst", "t", "u", "v"):
        assert np.array_equal(ref[key], walk[key]), key
    o.close()


def test_depth_capped_build_is_a_valid_tree(monkeypatch):
    """The safety net behind the builder (a second build with a binary-depth cap when the first wide tree is deeper than the
    traversal stack) cannot be reached with float coordinates once zero-cost ties are balanced, so it is exercised directly."""
    d = scenes.random_triangles(30000, 8, 8, spp=1)
    rays = np.ascontiguousarray(random_rays(8000, 5), np.float32)
    _, _, plain = _check(d, rays)
    monkeypatch.setenv("GTLC_BVH_DEPTH_CAP", "16")       # ceil(log2 30000) = 15: SAH only while its splits stay balanced, median splits otherwise
    _, _, capped = _check(d, rays)
    assert capped[:, 0].mean() > plain[:, 0].mean()     # a worse tree, but the same hits (checked by _check)
