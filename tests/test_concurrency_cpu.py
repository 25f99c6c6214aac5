"""Hydra syncs prims from many threads (renderDelegate.cpp:375-382): every giCreate*/giSet*/giDestroy* of meshes and
materials may run concurrently on one scene (Gi.cpp takes scene->mutex).  ctypes releases the GIL during the calls, so
Python threads really collide inside the library."""
import threading

import numpy as np

from gatling_b200 import gi, scenes


def _quad(z):
    pts = np.array([[-1, -1, z], [1, -1, z], [1, 1, z], [-1, 1, z]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    return scenes.bake_vertices(pts, nrm), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)


def test_concurrent_prim_sync_on_one_scene():
    scene = gi.giCreateScene()
    assert scene
    threads, per_thread = 8, 40
    kept = [[] for _ in range(threads)]
    errors = []

    def worker(t):
        try:
            rng = np.random.default_rng(t)
            for i in range(per_thread):
                mat = gi.giCreateMaterial(scene, 0, [0.8, 0.8, 0.8] + [0.0] * 45, name="m%d_%d" % (t, i))
                verts, faces = _quad(float(t * per_thread + i))
                mesh = gi.giCreateMesh(scene, verts, faces, id=t * 1000 + i, name="q%d_%d" % (t, i))
                assert mat and mesh
                gi.giSetMeshMaterial(mesh, mat)
                gi.giSetMeshTransform(mesh, np.eye(4, dtype=np.float32))
                inst = np.tile(np.eye(4, dtype=np.float32), (1 + i % 3, 1, 1))
                inst[:, 3, 0] = np.arange(len(inst)) * 3.0
                gi.giSetMeshInstanceTransforms(mesh, inst)
                gi.giSetMeshVisibility(mesh, True)
                if rng.random() < 0.25:              # some prims go away again while other threads keep adding
                    gi.giDestroyMesh(mesh)
                    gi.giDestroyMaterial(mat)
                else:
                    kept[t].append((mesh, mat, len(inst)))
        except Exception as e:                        # pragma: no cover
            errors.append(e)

    pool = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    for th in pool:
        th.start()
    for th in pool:
        th.join()
    assert not errors, errors
    expected_tris = sum(2 * n for per in kept for (_, _, n) in per)
    nodes, tris = gi.get_bvh(scene)
    assert len(tris) == expected_tris * 12 and len(nodes) % 80 == 0
    # every surviving instance is in the tree exactly once: (instance index, primitive) pairs are unique
    recs = tris.reshape(-1, 12).view(np.uint32)
    pairs = set(zip(recs[:, 7].tolist(), recs[:, 3].tolist()))
    assert len(pairs) == expected_tris
    for per in kept:
        for mesh, mat, _ in per:
            gi.giDestroyMesh(mesh)
            gi.giDestroyMaterial(mat)
    nodes, tris = gi.get_bvh(scene)
    assert len(tris) == 0 and len(nodes) == 80
    gi.giDestroyScene(scene)
