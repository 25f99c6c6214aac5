"""CPU: the oracle's restatements against the reference's OWN shader functions.

oracle/ref_glsl/build_ref.py compiles the pure functions of /root/reference/src/gi/shaders (common.glsl, colormap.glsl,
rp_main_payload.glsl, the helper functions of rp_main.rgen / .chit / .miss, apply_wrap_and_crop and mdl_adapt_normal of
mdl_interface.glsl) as C++ -- from where they lie, through a GLSL shim whose built-ins are the oracle's -- into
oracle/_ref/libglslref.so.  Every function must return the same BITS as the oracle's hand-written restatement on seeded random
inputs and on the edge cases of its domain: what is pinned is the reference's code (constants, operation order, branches)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libglslref.so")


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(SO):
        if not os.path.isdir("/root/reference/src/gi/shaders"):
            pytest.skip("oracle/_ref/libglslref.so is built from /root/reference, which only exists in the build container")
        sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_glsl"))
        import build_ref
        build_ref.build()
    return C.CDLL(SO), orc.lib()


F3 = C.c_float * 3


def f3(v):
    return F3(*[float(x) for x in v])


def bits(a):
    return np.asarray(a, np.float32).tobytes()


def call3(fn, *args, n=3):
    out = (C.c_float * n)()
    fn(*args, out)
    return bits(out[:])


def unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def test_rng_and_hashes(libs):
    R, O = libs
    for L in (R, O):
        for name in ("hash_theironborn", "rng_init"):
            getattr(L, ("ref_" if L is R else "orc_") + name).restype = C.c_uint32
    R.ref_uint_as_float.restype = O.orc_uint_as_float.restype = C.c_float
    R.ref_rng_next1f.restype = O.orc_rng_next1f.restype = C.c_float
    rng = np.random.default_rng(1)
    for x in list(rng.integers(0, 2 ** 32, 2000, dtype=np.uint64)) + [0, 1, 0xFFFFFFFF, 0x80000000]:
        x = int(x)
        assert R.ref_hash_theironborn(C.c_uint32(x)) == O.orc_hash_theironborn(C.c_uint32(x))
        assert R.ref_rng_init(C.c_uint32(x), C.c_uint32(x >> 7)) == O.orc_rng_init(C.c_uint32(x), C.c_uint32(x >> 7))
        assert bits([R.ref_uint_as_float(C.c_uint32(x))]) == bits([O.orc_uint_as_float(C.c_uint32(x))])
        a, b = C.c_uint32(x), C.c_uint32(x)
        fa, fb = (C.c_float * 4)(), (C.c_float * 4)()
        R.ref_rng_next4f(C.byref(a), fa); O.orc_rng_next4f(C.byref(b), fb)
        assert a.value == b.value and bits(fa[:]) == bits(fb[:])
        assert bits([R.ref_rng_next1f(C.byref(a))]) == bits([O.orc_rng_next1f(C.byref(b))]) and a.value == b.value


def test_geometry_helpers(libs):
    R, O = libs
    R.ref_encode_direction.restype = O.orc_encode_direction.restype = C.c_uint32
    R.ref_luminance.restype = O.orc_luminance.restype = C.c_float
    R.ref_safe_div.restype = O.orc_safe_div.restype = C.c_float
    R.ref_safe_div.argtypes = O.orc_safe_div.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(2)
    dirs = [unit(rng.normal(size=3)) for _ in range(3000)] + [np.array(v, np.float32) for v in
            ((0, 0, 1), (0, 0, -1), (1, 0, 0), (0, -1, 0), (0.6, 0, -0.8), (-0.0, 0.0, 1.0), (0, 1, -0.0))]
    for d in dirs:
        b1r, b2r, b1o, b2o = F3(), F3(), F3(), F3()
        R.ref_orthonormal_basis(f3(d), b1r, b2r); O.orc_orthonormal_basis(f3(d), b1o, b2o)
        assert bits(b1r[:]) == bits(b1o[:]) and bits(b2r[:]) == bits(b2o[:])
        e = R.ref_encode_direction(f3(d))
        assert e == O.orc_encode_direction(f3(d))
        assert call3(R.ref_decode_direction, C.c_uint32(e)) == call3(O.orc_decode_direction, C.c_uint32(e))
        p = (rng.normal(size=3) * 10.0 ** rng.integers(-4, 4)).astype(np.float32)
        assert call3(R.ref_offset_ray_origin, f3(p), f3(d)) == call3(O.orc_offset_ray_origin, f3(p), f3(d))
        assert bits([R.ref_luminance(f3(p))]) == bits([O.orc_luminance(f3(p))])
    for p in ((0, 0, 0), (1 / 32, -1 / 32, 0.03124), (-0.0, 5e-3, -7.0)):        # the |p| < 1/32 branch of offset_ray_origin, per component
        for d in dirs[:50]:
            assert call3(R.ref_offset_ray_origin, f3(p), f3(d)) == call3(O.orc_offset_ray_origin, f3(p), f3(d))
    for a, b in ((1.0, 0.0), (0.0, 0.0), (3.0, -0.0), (1.0, 3.0), (-2.5, 1e-30)):
        assert bits([R.ref_safe_div(a, b)]) == bits([O.orc_safe_div(a, b)])


def test_sampling_transforms(libs):
    R, O = libs
    rng = np.random.default_rng(3)
    F2 = C.c_float * 2
    xis = [rng.random(2).astype(np.float32) for _ in range(3000)] + [np.array(v, np.float32) for v in ((0, 0), (1, 1), (0.5, 0.5), (0, 1), (0.25, 0.75), (0.5, 0.25))]
    for xi in xis:
        x = F2(*xi)
        assert call3(R.ref_sample_hemisphere, x) == call3(O.orc_sample_hemisphere, x)
        rad = (rng.random(3) * 2).astype(np.float32)
        assert call3(R.ref_sample_sphere, x, f3(rad)) == call3(O.orc_sample_sphere, x, f3(rad))
        r2 = F2(*rad[:2])
        assert call3(R.ref_sample_disk, x, r2, n=2) == call3(O.orc_sample_disk, x, r2, n=2)
        assert call3(R.ref_fis_gauss, x, n=2) == call3(O.orc_fis_gauss, x, n=2)


def test_colormaps(libs):
    R, O = libs
    for fn in (R.ref_colormap, O.orc_colormap):
        fn.argtypes = [C.c_int, C.c_float, C.c_void_p]
    for t in list(np.linspace(0, 1, 513, dtype=np.float32)) + [np.float32(-0.5), np.float32(1.5)]:
        for ref_id, orc_id in ((0, 0), (3, 3)):                      # viridis (Opacity AOV), inferno (Bounces AOV): the two the path uses
            a, b = F3(), F3()
            R.ref_colormap(ref_id, float(t), a); O.orc_colormap(orc_id, float(t), b)
            assert bits(a[:]) == bits(b[:]), (ref_id, t)


def test_volume_and_roulette_helpers(libs):
    R, O = libs
    for L, pre in ((R, "ref_"), (O, "orc_")):
        getattr(L, pre + "sample_distance").restype = C.c_float
        getattr(L, pre + "sample_distance").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        getattr(L, pre + "sample_hg_cos").restype = C.c_float
        getattr(L, pre + "sample_hg_cos").argtypes = [C.c_float, C.c_float]
        getattr(L, pre + "sample_volume_scattering_direction").argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        getattr(L, pre + "russian_roulette").argtypes = [C.c_float, C.c_void_p, C.c_float]
    rng = np.random.default_rng(4)
    for k in range(3000):
        albedo, thr, sig = (rng.random(3).astype(np.float32) for _ in range(3))
        if k % 50 == 0:
            thr = np.zeros(3, np.float32)                                           # the `sum <= 1e-9` branch
        xi = float(np.float32(rng.random()))
        pr, po = F3(), F3()
        a = R.ref_sample_distance(f3(albedo), f3(thr), f3(sig), xi, pr)
        b = O.orc_sample_distance(f3(albedo), f3(thr), f3(sig), xi, po)
        assert bits([a]) == bits([b]) and bits(pr[:]) == bits(po[:])
        g = float(np.float32(rng.uniform(-0.95, 0.95))) if k % 7 else float(np.float32(rng.uniform(-1e-3, 1e-3)))
        assert bits([R.ref_sample_hg_cos(xi, g)]) == bits([O.orc_sample_hg_cos(xi, g)])
        d = unit(rng.normal(size=3))
        dr, do = f3(d), f3(d)
        x2 = (C.c_float * 2)(*rng.random(2).astype(np.float32))
        R.ref_sample_volume_scattering_direction(x2, g, dr); O.orc_sample_volume_scattering_direction(x2, g, do)
        assert bits(dr[:]) == bits(do[:])
        tr, to = f3(thr * 2), f3(thr * 2)
        inv = float(np.float32(rng.uniform(0.2, 1.0)))
        assert R.ref_russian_roulette(xi, tr, inv) == O.orc_russian_roulette(xi, to, inv) and bits(tr[:]) == bits(to[:])


def test_payload_bitfield(libs):
    R, O = libs
    for L, pre in ((R, "ref_"), (O, "orc_")):
        for n in ("payload_get_medium_idx", "payload_set_medium_idx", "payload_increment_walk", "payload_get_walk"):
            getattr(L, pre + n).restype = C.c_uint32
    rng = np.random.default_rng(5)
    vals = [int(x) for x in rng.integers(0, 2 ** 32, 3000, dtype=np.uint64)] + [0, 0xFFFFFFFF, 0x00FFF000, 0x00FFE000, 0x0F000000, 0x00000FFF]
    for v in vals:
        v32 = C.c_uint32(v)
        for mss in (0, 2):
            assert R.ref_payload_get_medium_idx(v32, mss) == O.orc_payload_get_medium_idx(v32, mss)
        for idx in (0, 1, 2, 15, 16, 255, 300):
            assert R.ref_payload_set_medium_idx(v32, C.c_uint32(idx)) == O.orc_payload_set_medium_idx(v32, C.c_uint32(idx))
        assert R.ref_payload_increment_walk(v32) == O.orc_payload_increment_walk(v32)
        assert R.ref_payload_get_walk(v32) == O.orc_payload_get_walk(v32)


def test_dome_and_texture_helpers(libs):
    R, O = libs
    R.ref_apply_wrap_and_crop.restype = O.orc_apply_wrap_and_crop.restype = C.c_float
    R.ref_apply_wrap_and_crop.argtypes = O.orc_apply_wrap_and_crop.argtypes = [C.c_float, C.c_int, C.c_int]
    rng = np.random.default_rng(6)
    for k in range(3000):
        q = unit(rng.normal(size=4)); d = unit(rng.normal(size=3))
        q4 = (C.c_float * 4)(*q)
        assert call3(R.ref_quat_rotate_dir, q4, f3(d)) == call3(O.orc_quat_rotate_dir, q4, f3(d))
        assert call3(R.ref_dome_light_uv, f3(d), n=2) == call3(O.orc_dome_light_uv, f3(d), n=2)
        coord = float(np.float32(rng.uniform(-3, 3)))
        for wrap in (0, 1, 2):                                                  # clamp, repeat, mirrored repeat (clip is decided by the caller)
            for res in (1, 7, 64):
                assert bits([R.ref_apply_wrap_and_crop(coord, wrap, res)]) == bits([O.orc_apply_wrap_and_crop(coord, wrap, res)])
        gn = unit(rng.normal(size=3)); n = unit(gn + 0.4 * rng.normal(size=3)); w = unit(rng.normal(size=3))
        assert call3(R.ref_mdl_adapt_normal, f3(w), f3(gn), f3(n)) == call3(O.orc_mdl_adapt_normal, f3(w), f3(gn), f3(n))
    for d in ((0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1)):
        assert call3(R.ref_dome_light_uv, f3(d), n=2) == call3(O.orc_dome_light_uv, f3(d), n=2)


LIGHT_DT = {
    "sphere": np.dtype([("pos", "3f4"), ("dsp", "u4"), ("emission", "3f4"), ("area", "f4"), ("radius", "3f4"), ("pad", "f4")]),
    "distant": np.dtype([("dir", "3f4"), ("angle", "f4"), ("emission", "3f4"), ("dsp", "u4"), ("pad", "3f4"), ("invPdf", "f4")]),
    "rect": np.dtype([("origin", "3f4"), ("width", "f4"), ("emission", "3f4"), ("height", "f4"), ("frame", "2u4"), ("dsp", "u4"), ("pad", "f4")]),
    "disk": np.dtype([("origin", "3f4"), ("rx", "f4"), ("emission", "3f4"), ("ry", "f4"), ("frame", "2u4"), ("dsp", "u4"), ("pad", "f4")]),
}


def test_sample_light(libs):
    """rp_main.chit:30-129 with every light type, light selection by k4.x / k4.y, zero-area lights, the distant light's cone."""
    R, O = libs
    assert [R.ref_struct_sizes(i) for i in range(4)] == [48, 48, 48, 48] == [LIGHT_DT[k].itemsize for k in ("sphere", "distant", "rect", "disk")]
    O.orc_encode_direction.restype = C.c_uint32
    rng = np.random.default_rng(7)
    for trial in range(40):
        counts = rng.integers(0, 4, 4).astype(np.uint32)
        if counts.sum() == 0:
            counts[trial % 4] = 1
        sp = np.zeros(max(counts[0], 1), LIGHT_DT["sphere"]); di = np.zeros(max(counts[1], 1), LIGHT_DT["distant"])
        re = np.zeros(max(counts[2], 1), LIGHT_DT["rect"]); dk = np.zeros(max(counts[3], 1), LIGHT_DT["disk"])
        for a in (sp, di, re, dk):
            a["emission"] = rng.random((len(a), 3)) * 10
            a["dsp"] = rng.integers(0, 2 ** 32, len(a), dtype=np.uint64).astype(np.uint32)
        sp["pos"] = rng.normal(size=(len(sp), 3)); sp["radius"] = rng.random((len(sp), 3)) * 0.3; sp["area"] = rng.random(len(sp)) * (rng.random(len(sp)) > 0.2)
        di["dir"] = [unit(rng.normal(size=3)) for _ in di]; di["angle"] = rng.random(len(di)) * 0.2 * (rng.random(len(di)) > 0.3); di["invPdf"] = rng.random(len(di)) + 0.5
        for a in (re, dk):
            a["origin"] = rng.normal(size=(len(a), 3))
            a["frame"] = [[O.orc_encode_direction(f3(unit(rng.normal(size=3)))) for _ in range(2)] for _ in a]
        re["width"] = rng.random(len(re)) * (rng.random(len(re)) > 0.2); re["height"] = rng.random(len(re))
        dk["rx"] = rng.random(len(dk)) * (rng.random(len(dk)) > 0.2); dk["ry"] = rng.random(len(dk))
        cnt = (C.c_uint32 * 4)(*[int(c) for c in counts])
        mult, expo = float(np.float32(rng.uniform(0.5, 2))), float(np.float32(rng.uniform(-1, 1)))
        for k in range(200):
            k4 = (C.c_float * 4)(*rng.random(4).astype(np.float32))
            if k < 8:
                k4[0] = (0.0, 1.0)[k & 1]; k4[1] = (0.0, 1.0)[(k >> 1) & 1]; k4[2] = (0.0, 0.5)[(k >> 2) & 1]
            pos = f3(rng.normal(size=3).astype(np.float32))
            outs = []
            for fn in (R.ref_sample_light, O.orc_sample_light):
                fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
                o8 = (C.c_float * 8)(); dsp = C.c_uint32(0)
                fn(cnt, sp.ctypes.data, di.ctypes.data, re.ctypes.data, dk.ctypes.data, mult, expo, k4, pos, o8, C.byref(dsp))
                outs.append((bits(o8[:]), dsp.value))
            assert outs[0] == outs[1], (trial, k, list(counts), list(k4))


def test_camera_rays(libs):
    """rp_main.rgen:191-288, the head of main(): pixel -> camera ray with the RNG draws of jittering / Gauss filter importance sampling /
    depth of field, the zero-component guard and the clip range along the ray -- compiled once per set of pipeline switches -- against
    the oracle's camera_ray + clip_range_along, fed with the uniform values the oracle derives from the camera (Gi.cpp:2373-2423)."""
    import dataclasses
    from gatling_b200 import scenes
    R, O = libs
    d = scenes.cornell(37, 23, spp=1)
    cams = [d.camera,
            dataclasses.replace(d.camera, fStop=1.4, focalLength=50.0, focusDistance=2.5, clipStart=0.3, clipEnd=40.0),
            dataclasses.replace(d.camera, position=(0.3, -2.0, 0.5), forward=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0), vfov=1.2, fStop=0.0)]   # axis-aligned: zero direction components
    for cam in cams:
        for flags, (jit, fis, dof) in ((0, (0, 0, 0)), (1, (1, 0, 0)), (3, (1, 1, 0)), (7, (1, 1, 1)), (4, (0, 0, 1))):
            dd = dataclasses.replace(d, camera=cam, settings=dataclasses.replace(d.settings, jitteredSampling=bool(jit), filterImportanceSampling=bool(fis), depthOfField=bool(dof)))
            r = orc.OracleScene.render_desc(type("S", (), {"desc": dd})())
            cam12 = (C.c_float * 12)(); bits2 = (C.c_uint32 * 2)()
            O.orc_camera_ubo(C.byref(r), cam12, bits2)
            for sample in (0, 5):
                for py in range(0, dd.height, 3):
                    for px in range(0, dd.width, 2):
                        a, b = (C.c_float * 8)(), (C.c_float * 8)()
                        ra, rb = C.c_uint32(0), C.c_uint32(0)
                        assert R.ref_camera_ray(flags, cam12, bits2, sample, px, py, a, C.byref(ra)) == 1
                        O.orc_camera_ray_full(C.byref(r), sample, px, py, b, C.byref(rb))
                        assert bits(a[:]) == bits(b[:]) and ra.value == rb.value, (flags, sample, px, py, list(a), list(b))


def test_shading_state(libs):
    """mdl_shading_state.glsl (whole file): interpolation of the packed vertices, normal / tangent decoding, transforms through the instance
    matrices, the front-face flip, re-orthonormalised tangent frame and uv -- at random hits of every instance of scenes with rotated,
    scaled and mirrored instances -- against the oracle's setup_shading_state on the same packed buffers."""
    from gatling_b200 import scenes
    R, O = libs
    O.orc_scene_views.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    O.orc_setup_shading_state.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    R.ref_setup_shading_state.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(9)
    mirrored = scenes.cornell(8, 8, spp=1)
    m0 = np.asarray(mirrored.meshes[-1].transform, np.float32).copy(); m0[0, :3] *= -1.0          # a left-handed instance
    mirrored.meshes[-1].transform = m0
    checked = 0
    for d in (scenes.cornell(8, 8, spp=1), mirrored, scenes.kitchen_standin(tris_per_instance=60, instances=9, meshes=5, materials=4, width=8, height=8, spp=1),
              scenes.textured_cornell(8, 8, spp=1)):
        o = orc.OracleScene(d)
        h = o.handle
        faces, fverts = C.c_void_p(), C.c_void_p()
        bases = (C.c_uint32 * 3)(); o2w, w2o = (C.c_float * 12)(), (C.c_float * 12)(); count = C.c_uint32(0)
        O.orc_scene_views(h, 0, C.byref(faces), C.byref(fverts), bases, o2w, w2o, C.byref(count))
        for inst in range(count.value):
            O.orc_scene_views(h, inst, C.byref(faces), C.byref(fverts), bases, o2w, w2o, C.byref(count))
            mesh_faces = C.c_void_p(faces.value + 12 * bases[1])
            for _ in range(40):
                prim = int(rng.integers(0, bases[2]))
                u = float(np.float32(rng.random())); v = float(np.float32(rng.random() * (1.0 - u)))
                if _ < 3:
                    u, v = ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))[_]
                ray = f3(unit(rng.normal(size=3)))
                a, b = (C.c_float * 18)(), (C.c_float * 18)()
                fa, fb = C.c_int(0), C.c_int32(0)
                R.ref_setup_shading_state(mesh_faces, fverts, bases[0], o2w, w2o, prim, u, v, ray, a, C.byref(fa))
                O.orc_setup_shading_state(h, inst, prim, u, v, ray, b, C.byref(fb))
                assert fa.value == fb.value and bits(a[:]) == bits(b[:]), (d.name, inst, prim, u, v, list(a), list(b))
                checked += 1
        o.close()
    assert checked > 1000


# ---- the whole integrator: rp_main.rgen / .chit / .miss / _shadow.miss compiled from the reference's files -------------------------------
def reference_integrator_render(libs, desc, sample_offset=0, prev=None, aov_ids=None, clear_values=None):
    """The scene rendered by the reference's own shader code (orc.ReferenceIntegrator) and by the oracle: (image or {aov: image}, ray counts,
    the oracle's image or {aov: image}, the oracle's ray counts)."""
    o = orc.OracleScene(desc)
    ri = orc.ReferenceIntegrator()
    got, counts, _ = ri.render(o, sample_offset=sample_offset, prev=prev, aov_ids=aov_ids, clear_values=clear_values)
    if aov_ids is None:
        want, _, rays = o.render(sample_offset=sample_offset, prev=prev)
    else:
        want, rays = o.render_aovs(aov_ids, clear_values, sample_offset=sample_offset), None
    o.close()
    return got, counts, want, rays


def _lit(d):
    import dataclasses
    from gatling_b200 import scenes
    d.settings = dataclasses.replace(d.settings, nextEventEstimation=True)
    if not d.rect_lights:
        d.rect_lights.append(scenes.RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
    return d


def test_reference_integrator_cornell(libs):
    """The reference's ray-generation, closest-hit, miss and shadow-miss shaders, compiled from its own files, render the Cornell box
    (NEE on a rect light + sphere / distant / disk lights, progressive second call): the oracle's image must be the same bits, and so
    must the ray counts."""
    from gatling_b200 import scenes
    d = _lit(scenes.cornell(40, 30, spp=3, max_bounces=6))
    d.sphere_lights.append(scenes.SphereLight(position=(0.5, -0.5, 0.2), base_emission=(4, 3, 2), radius=(0.1, 0.1, 0.1)))
    d.distant_lights.append(scenes.DistantLight(direction=(0.3, 0.5, -0.8), base_emission=(1.0, 1.0, 1.2), angle=0.05))
    img, counts, want, rays = reference_integrator_render(libs, d)
    assert counts == tuple(rays) and rays[1] > 0
    assert want[..., :3].mean() > 0.05
    assert np.array_equal(img, want)
    img2, counts2, want2, rays2 = reference_integrator_render(libs, d, sample_offset=3, prev=want)
    assert counts2 == tuple(rays2) and np.array_equal(img2, want2) and not np.array_equal(want2, want)


def test_reference_integrator_materials_and_media(libs):
    """Same comparison on scenes that walk the rest of the closest-hit and miss shaders: layered OpenPBR / UsdPreviewSurface closures with
    transmission (medium toggle), thin walls and emission under a dome light; textured and primvar-driven inputs with a normal map
    (what *_init() leaves in the state); nested scattering media with the two-deep medium stack (stepVolume, walk counters, push / pop)."""
    import dataclasses
    from gatling_b200 import scenes
    stress = scenes.material_stress(tris_per_mesh=60, meshes=16, width=36, height=28, spp=2, max_bounces=7)
    kitchen = scenes.kitchen_standin(tris_per_instance=40, instances=14, meshes=7, materials=6, width=36, height=28, spp=2, max_bounces=6)
    textured = scenes.textured_cornell(32, 32, spp=2)
    primvars = scenes.primvar_cornell(32, 32, spp=2)
    layered = scenes.layered_cornell(32, 32, spp=2)
    media = scenes.volume_cornell(32, 32, spp=3, max_bounces=12, medium_stack_size=2)
    beer = scenes.volume_cornell(32, 32, spp=2, max_bounces=8, medium_stack_size=0)
    for d in (stress, kitchen, textured, primvars, layered, media, beer):
        d = _lit(d) if not d.settings.nextEventEstimation else d
        img, counts, want, rays = reference_integrator_render(libs, d)
        assert counts == tuple(rays), d.name
        assert np.isfinite(want).all() and want[..., :3].mean() > 0.01, d.name
        assert np.array_equal(img, want), (d.name, float(np.abs(img - want).max()))


def test_reference_integrator_other_switches(libs):
    """The pipeline built with the other compile-time switches: depth of field and clipping planes on, no jitter, no progressive
    accumulation, the dome light hidden from the camera (fallback texel for primary rays)."""
    import dataclasses
    from gatling_b200 import scenes
    d = _lit(scenes.cornell(36, 28, spp=3, max_bounces=5))
    d.meshes = d.meshes[:-1]                                         # a gap for rays to leave through
    dome = np.zeros((8, 16, 4), np.float32); dome[..., :3] = np.linspace(0.2, 2.0, 16)[None, :, None]; dome[2, 5, :3] = 30.0
    d.dome_light = scenes.DomeLight(dome, rotation=(0.0, 0.3826834, 0.0, 0.9238795), base_emission=(1.0, 0.9, 0.8))
    d.camera = dataclasses.replace(d.camera, fStop=2.0, focalLength=0.5, focusDistance=6.0, clipStart=6.6, clipEnd=7.7)
    d.settings = dataclasses.replace(d.settings, depthOfField=True, clippingPlanes=True, jitteredSampling=False, filterImportanceSampling=False,
                                     progressiveAccumulation=False, domeLightCameraVisible=False)
    img, counts, want, rays = reference_integrator_render(libs, d)
    assert counts == tuple(rays) and np.array_equal(img, want)
    plain = dataclasses.replace(d, settings=dataclasses.replace(d.settings, clippingPlanes=False))
    assert not np.array_equal(orc.OracleScene(plain).render()[0], want)          # the clip range really cut something


def test_reference_integrator_auxiliary_aovs(libs):
    """The build with every auxiliary AOV bound (rp_main.rgen:132-183 clearAovs, :431-436 NEE, :483-486 Bounces, :517-520 normal
    re-normalisation; rp_main.chit:191-290): Normal, NEE, Barycentrics, Texcoords, Bounces, Opacity, Tangents, Bitangents, ThinWalled,
    ObjectId, Depth, InstanceId, DoubleSided, Albedo and Color against the oracle's replay of the megakernel's per-pixel write order,
    with non-trivial clear values.  (ClockCycles is a timer; FaceId reads a byte-packed table the oracle does not model.)"""
    from gatling_b200 import scenes
    ids = [a for a in range(17) if a not in (6, 13)]
    clear = {a: ((7, 0, 0, 0) if a in (11, 14) else (0.25 + 0.01 * a, 0.5, 0.125 * (a % 3), 0.75)) for a in ids if a != 0}
    for d in (_lit(scenes.cornell(30, 22, spp=3, max_bounces=5)),
              _lit(scenes.kitchen_standin(tris_per_instance=40, instances=14, meshes=7, materials=6, width=30, height=22, spp=2, max_bounces=5)),
              scenes.textured_cornell(28, 28, spp=2)):
        if d.name.startswith("cornell"):
            d.meshes = d.meshes[:-2]                                  # pixels that hit nothing keep the clear values
        got, counts, want, _ = reference_integrator_render(libs, d, aov_ids=ids, clear_values=clear)
        for a in ids:
            g, w = got[a], want[a]
            if a == 0:
                assert np.array_equal(g, w), (d.name, a)
            else:
                assert np.array_equal(g, w[..., :g.shape[-1]]), (d.name, a, float(np.abs(g.astype(np.float64) - w[..., :g.shape[-1]]).max()))
        assert len(np.unique(got[11])) > 1 and got[12].std() > 0.0 and got[16][..., :3].std() > 0.0


def test_host_light_encoders(libs):
    """The reference's HOST code for lights -- every gi{Create,Set}{Sphere,Distant,Rect,Disk}Light* of src/gi/impl/Gi.cpp:2573-2941 and
    `_EncodeDirection` (:287-300), compiled from the reference's file -- against the light records the oracle builds for the same user
    values (orc.light_arrays): Knud-Thomsen area, the cone's inverse pdf (in double, as written), tangent-frame packing, layout."""
    from gatling_b200 import scenes
    R, O = libs
    R.refhost_set_pack_half(C.cast(O.orc_pack_half2x16, C.c_void_p))
    R.refhost_encode_direction.restype = O.orc_host_encode_direction.restype = C.c_uint32
    rng = np.random.default_rng(11)
    fl = lambda v: float(np.float32(v))
    d = scenes.cornell(4, 4, spp=1)
    for k in range(300):
        v3 = lambda s=1.0: tuple(fl(x) for x in rng.normal(size=3) * s)
        ds = (fl(rng.random()), fl(rng.random() * 2))
        sph = scenes.SphereLight(position=v3(3), base_emission=v3(5), radius=tuple(fl(x) for x in rng.random(3) * 2 * (1 if k % 9 else 0)), diffuse=ds[0], specular=ds[1])
        dis = scenes.DistantLight(direction=tuple(unit(rng.normal(size=3)).tolist()), base_emission=v3(2), angle=fl(rng.random() * (1 if k % 5 else 0)), diffuse=ds[0], specular=ds[1])
        rec = scenes.RectLight(origin=v3(2), t0=v3(), t1=v3(), base_emission=v3(9), width=fl(rng.random() * 3), height=fl(rng.random()), diffuse=ds[0], specular=ds[1])
        dsk = scenes.DiskLight(origin=v3(2), t0=v3(), t1=v3(), base_emission=v3(9), radius=(fl(rng.random()), fl(rng.random() * 2)), diffuse=ds[0], specular=ds[1])
        d.sphere_lights, d.distant_lights, d.rect_lights, d.disk_lights = [sph], [dis], [rec], [dsk]
        sp, di, re, dk = orc.light_arrays(d)
        out = (C.c_uint8 * 48)()
        R.refhost_sphere_light.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        R.refhost_sphere_light(1, f3(sph.position), f3(sph.base_emission), f3(sph.radius), sph.diffuse, sph.specular, out)
        assert bytes(out)[:44] == sp.tobytes()[:44], (k, sph)                                   # (the last 4 bytes are padding)
        R.refhost_distant_light.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        R.refhost_distant_light(1, f3(dis.direction), f3(dis.base_emission), dis.angle, dis.diffuse, dis.specular, out)
        b, w = bytes(out), di.tobytes()
        assert b[:32] == w[:32] and b[44:] == w[44:], (k, dis)                                  # (bytes 32..43 are padding)
        R.refhost_rect_light.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_float] * 4 + [C.c_void_p]
        R.refhost_rect_light(1, f3(rec.origin), f3(rec.t0), f3(rec.t1), f3(rec.base_emission), rec.width, rec.height, rec.diffuse, rec.specular, out)
        assert bytes(out)[:44] == re.tobytes()[:44], (k, rec)
        R.refhost_disk_light.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_float] * 4 + [C.c_void_p]
        R.refhost_disk_light(1, f3(dsk.origin), f3(dsk.t0), f3(dsk.t1), f3(dsk.base_emission), dsk.radius[0], dsk.radius[1], dsk.diffuse, dsk.specular, out)
        assert bytes(out)[:44] == dk.tobytes()[:44], (k, dsk)
        v = f3(rng.normal(size=3) * 10.0 ** rng.integers(-3, 3))
        assert R.refhost_encode_direction(v) == O.orc_host_encode_direction(v)


def test_host_uniform_block(libs):
    """giRender's uniform block (src/gi/impl/Gi.cpp:2375-2425, compiled from the reference's file against the reference's own Gi.h): camera
    vectors normalised, lens radius, packed image size / bounce counts / clip range, 1 / total samples, dome-light values and light counts
    -- against what the oracle derives from the same camera and settings (fill_ctx -> orc_ubo_values)."""
    import dataclasses
    from gatling_b200 import scenes
    R, O = libs
    R.refhost_set_pack_half(C.cast(O.orc_pack_half2x16, C.c_void_p))
    O.orc_ubo_values.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.refhost_uniform_data.argtypes = [C.c_void_p] * 5 + [C.c_uint32] * 3 + [C.c_void_p] * 2
    rng = np.random.default_rng(13)
    fl = lambda v: float(np.float32(v))
    for k in range(200):
        d = scenes.cornell(int(rng.integers(1, 4000)), int(rng.integers(1, 3000)), spp=int(rng.integers(1, 5000)))
        cam = scenes.Camera(position=tuple(fl(x) for x in rng.normal(size=3) * 5), forward=tuple(fl(x) for x in rng.normal(size=3) * 3), up=tuple(fl(x) for x in rng.normal(size=3)),
                            vfov=fl(rng.uniform(0.1, 2.5)), fStop=fl(rng.uniform(0.5, 22.0) * (k % 3 != 0)), focusDistance=fl(rng.uniform(0.1, 50)), focalLength=fl(rng.uniform(0.01, 0.2)),
                            clipStart=fl(rng.uniform(0.001, 2.0)), clipEnd=fl(rng.uniform(5.0, 70000.0)), exposure=fl(rng.normal()))
        st = dataclasses.replace(d.settings, lightIntensityMultiplier=fl(rng.uniform(0.1, 4)), maxSampleValue=fl(rng.uniform(1, 100)), metersPerSceneUnit=fl(rng.uniform(0.01, 10)),
                                 rrInvMinTermProb=fl(rng.uniform(0.5, 1)), maxBounces=int(rng.integers(0, 4096)), maxVolumeWalkLength=int(rng.integers(0, 64)), rrBounceOffset=int(rng.integers(0, 64)))
        d = dataclasses.replace(d, camera=cam, settings=st)
        dome7 = None
        if k % 2:
            q = unit(rng.normal(size=4)); em = tuple(fl(x) for x in rng.random(3) * 3)
            d.dome_light = scenes.DomeLight(np.ones((2, 4, 4), np.float32), rotation=tuple(float(x) for x in q), base_emission=em)
            dome7 = (C.c_float * 7)(*q, *em)
        d.sphere_lights = [scenes.SphereLight(position=(0, 0, 0), base_emission=(1, 1, 1), radius=(1, 1, 1))] * int(rng.integers(0, 3))
        d.distant_lights = [scenes.DistantLight(direction=(0, 0, -1), base_emission=(1, 1, 1), angle=0.0)] * int(rng.integers(0, 3))
        sample_offset = int(rng.integers(0, 100000))
        o = orc.OracleScene(d)
        r = o.render_desc(sample_offset)
        fo, uo = (C.c_float * 28)(), (C.c_uint32 * 16)()
        O.orc_ubo_values(o.handle, C.byref(r), fo, uo)
        o.close()
        cam16 = (C.c_float * 16)(*cam.position, *cam.forward, *cam.up, cam.vfov, cam.fStop, cam.focusDistance, cam.focalLength, cam.clipStart, cam.clipEnd, cam.exposure)
        set_f = (C.c_float * 4)(st.lightIntensityMultiplier, st.maxSampleValue, st.metersPerSceneUnit, st.rrInvMinTermProb)
        set_u = (C.c_uint32 * 4)(st.maxBounces, st.maxVolumeWalkLength, st.rrBounceOffset, st.spp)
        counts = (C.c_uint32 * 4)(len(d.sphere_lights), len(d.distant_lights), len(d.rect_lights), len(d.disk_lights))
        fr, ur = (C.c_float * 28)(), (C.c_uint32 * 16)()
        R.refhost_uniform_data(cam16, set_f, set_u, counts, dome7, sample_offset, d.width, d.height, fr, ur)
        assert bits(fr[:25]) == bits(fo[:25]), (k, list(fr), list(fo))
        assert list(ur[:11]) == list(uo[:11]), (k, list(ur), list(uo))


def test_mdl_texture_runtime(libs):
    """mdl_interface.glsl:40-205 (tex_lookup_float4_2d / _3d, tex_texel_float4_2d / _3d) compiled from the reference's file; the hardware's
    part of a lookup (size query, linear filter with Repeat addressing, texel fetch) is supplied by the oracle's samplers.  Against the
    oracle's restatements: clip test, wrap + crop order, out-of-range texel fetches, for every wrap mode and all four texel formats."""
    from gatling_b200 import scenes
    R, O = libs
    R.ref_tex_function.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    sample, texel = C.cast(O.orc_tex_hw_sample, C.c_void_p), C.cast(O.orc_tex_hw_texel, C.c_void_p)
    rng = np.random.default_rng(17)
    flat = [scenes.procedural_texture(k, 16) for k in ("checker", "roughness", "emission", "noise")]
    vols = [scenes.procedural_volume(5, 3), scenes.procedural_volume(4, 8, channels=1)]
    def ref(mode, tex, coord=(0, 0, 0), icoord=(0, 0, 0), wraps=(1, 1, 1)):
        d, keep = orc._texture_desc(tex)
        rows, cols = tex.rows_cols
        out = (C.c_float * 4)()
        R.ref_tex_function(mode, C.byref(d), sample, texel, cols, rows, max(tex.depth, 1), (C.c_float * 3)(*coord), (C.c_int * 3)(*icoord), wraps[0], wraps[1], wraps[2], out)
        return np.array(out[:], np.float32)
    for tex in flat:
        for _ in range(400):
            u, v = (float(np.float32(x)) for x in rng.uniform(-2.5, 2.5, 2))
            wu, wv = (int(x) for x in rng.integers(0, 4, 2))
            got = ref(0, tex, (u, v, 0), wraps=(wu, wv, 1))
            want, clipped = orc.tex_lookup(tex, u, v, wu, wv)
            assert np.array_equal(got, np.zeros(4, np.float32) if clipped else want), (u, v, wu, wv)
            x, y = (int(t) for t in rng.integers(-3, 20, 2))
            assert np.array_equal(ref(2, tex, icoord=(x, y, 0)), orc.tex_texel(tex, x, y))
    for tex in vols:
        for _ in range(400):
            c = tuple(float(np.float32(x)) for x in rng.uniform(-2.0, 2.0, 3))
            w3 = tuple(int(x) for x in rng.integers(0, 4, 3))
            got = ref(1, tex, c, wraps=w3)
            want, clipped = orc.tex_lookup_3d(tex, *c, *w3)
            assert np.array_equal(got, np.zeros(4, np.float32) if clipped else want), (c, w3)
            ic = tuple(int(t) for t in rng.integers(-2, 7, 3))
            assert np.array_equal(ref(3, tex, icoord=ic), orc.tex_texel(tex, *ic))


def test_scene_data_lookups(libs):
    """mdl_interface.glsl:266-452 ("scene data" = primvars: scene_data_isvalid, get_scene_data_indices, scene_data_lookup_float / _float2 /
    _float3 / _color / _int) compiled from the reference's file, buffer references as pointers: every primvar reader of the primvar test
    scenes at random hits -- constant, per-instance, per-face and per-vertex interpolation, float and integer data, missing primvars --
    against the values the oracle writes into the parameter block."""
    from gatling_b200 import scenes
    R, O = libs
    O.orc_scene_views.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    O.orc_hit_begin.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
    O.orc_primvar_entry.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    O.orc_instance_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    R.ref_scene_data_lookup.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(21)
    kinds = set()
    for d in (scenes.primvar_cornell(8, 8, spp=1), scenes.lookup_cornell(8, 8, spp=1)):
        o = orc.OracleScene(d)
        h = o.handle
        faces, fverts = C.c_void_p(), C.c_void_p()
        bases = (C.c_uint32 * 3)(); o2w, w2o = (C.c_float * 12)(), (C.c_float * 12)(); count = C.c_uint32(0)
        O.orc_scene_views(h, 0, C.byref(faces), C.byref(fverts), bases, o2w, w2o, C.byref(count))
        inst_ids, obj_ids = np.zeros(count.value, np.int32), np.zeros(count.value, np.int32)
        O.orc_instance_ids(h, inst_ids.ctypes.data, obj_ids.ctypes.data)
        flat_meshes = [m for m in d.meshes for _ in range(1 if m.instance_transforms is None else len(m.instance_transforms))]
        assert len(flat_meshes) == count.value
        for inst, mesh in enumerate(flat_meshes):
            readers = getattr(d.materials[mesh.material], "primvars", None) or []
            if not readers:
                continue
            O.orc_scene_views(h, inst, C.byref(faces), C.byref(fverts), bases, o2w, w2o, C.byref(count))
            all_faces = np.ctypeslib.as_array(C.cast(faces, C.POINTER(C.c_uint32)), shape=((bases[1] + bases[2]) * 3,))
            for _ in range(60):
                prim = int(rng.integers(0, bases[2]))
                u = float(np.float32(rng.random())); v = float(np.float32(rng.random() * (1.0 - u)))
                O.orc_hit_begin(h, inst, prim, u, v, f3((0.0, 1.0, 0.0)))
                params = np.zeros(48, np.float32)
                O.orc_hit_params(params.ctypes.data)
                idx = (C.c_uint32 * 3)(*[int(x) for x in all_faces[(bases[1] + prim) * 3:(bases[1] + prim) * 3 + 3]])
                for k, rd in enumerate(readers):
                    info = (C.c_uint32 * 2)(); data = C.c_void_p()
                    O.orc_primvar_entry(h, inst, k, info, C.byref(data))
                    present, stride, interp, is_int = bool(info[1] & 0x100), info[1] & 15, (info[1] >> 4) & 15, bool(info[1] & 0x200)
                    if present and stride != rd.count:
                        continue                                        # a reader narrower / wider than its data: undefined in the reference (reads the neighbour)
                    word = ((interp << 30) | ((stride - 1) << 28)) if present else 0xFFFFFFFF
                    out = (C.c_float * 3)()
                    fb = (C.c_float * 3)(*[float(x) for x in rd.fallback[:3]])
                    R.ref_scene_data_lookup(word, data, idx, u, v, prim, max(int(inst_ids[inst]), 0), rd.count, int(is_int), fb, out)
                    want = params[rd.param:rd.param + rd.count]
                    assert bits(out[:rd.count]) == bits(want), (d.name, inst, k, rd, list(out), list(want))
                    kinds.add((present, interp if present else -1, is_int))
        o.close()
    assert {(True, i, False) for i in range(4)} <= kinds and (False, -1, False) in kinds and any(k[2] for k in kinds), kinds


def test_test_harness_conversion(libs):
    """How the reference's graphical tests turn a float AOV into the 8-bit image they byte-compare (src/hdGatling/main.cpp:136-143
    `_AccurateLinearToSrgb`, :465-487 the conversion loop), compiled from the reference's file, against the restatement
    tools/compare_with_reference.py uses to reproduce the one reconstructible golden image (tests/test_reference_pin.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import compare_with_reference as cw
    R, _ = libs
    R.refhost_accurate_linear_to_srgb.restype = C.c_float
    R.refhost_accurate_linear_to_srgb.argtypes = [C.c_float]
    R.refhost_harness_bytes.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(23)
    xs = np.concatenate([rng.random(4000), rng.random(500) * 0.01, [0.0, 0.0031308, 0.00313081, 1.0, 0.5, 0.999999, 1e-8]]).astype(np.float32)
    got = np.array([R.refhost_accurate_linear_to_srgb(float(x)) for x in xs], np.float32)
    assert np.array_equal(got, cw.accurate_linear_to_srgb(xs))
    for comp in (4, 3, 1):
        img = (rng.random((9, 13, comp)) * 1.002).astype(np.float32)                       # [0, 1.002): the range a resolved colour / AOV buffer holds
        img[0, 0] = 1.0; img[0, 1] = 0.0
        for gamma in (1, 0):
            out = np.zeros((9, 13, 4), np.uint8)
            R.refhost_harness_bytes(img.ctypes.data, 13, 9, comp, 0, gamma, out.ctypes.data)
            assert np.array_equal(out, cw.harness_bytes(img, gamma_encode=bool(gamma))), (comp, gamma)
    ids = rng.integers(-1, 1 << 24, (9, 13), dtype=np.int64).astype(np.int32)
    out = np.zeros((9, 13, 4), np.uint8)
    R.refhost_harness_bytes(ids.ctypes.data, 13, 9, 1, 1, 0, out.ctypes.data)
    assert np.array_equal(out, cw.harness_bytes(ids))


def test_reference_integrator_row_bands_and_no_nee(libs):
    """What bench.py's reference arm runs: a band of image rows at an overridden sample count, on the OpenMP team, with and without next-event
    estimation (configs C3 / C5 and C2 of BASELINE.json): the band equals the oracle's band, ray counts included."""
    from gatling_b200 import scenes
    for d in (_lit(scenes.kitchen_standin(tris_per_instance=40, instances=14, meshes=7, materials=6, width=48, height=40, spp=8, max_bounces=6)),
              scenes.random_triangles(600, 48, 40, spp=8, max_bounces=5)):
        o = orc.OracleScene(d)
        got, counts, secs = orc.ReferenceIntegrator().render(o, rows=(13, 29), spp=3)
        want, _, rays = o.render(rows=(13, 29), spp=3)
        o.close()
        assert counts == tuple(rays) and secs > 0.0
        assert np.array_equal(got[13:29], want[13:29]) and want[13:29, :, :3].mean() > 0.01
        assert not got[:13].any() and not got[29:].any()


def test_host_vertex_packing(libs):
    """Gi.cpp:845-859 (the loop of _giBuildGeometryStructures that turns a mesh's GiVertex array into the rp::FVertex records
    mdl_shading_state.glsl unpacks), compiled from the reference's file against its own Gi.h: normals and tangents through `_EncodeDirection`,
    position + bitangent sign, uv -- against the oracle's packed vertex buffer for the same meshes."""
    from gatling_b200 import scenes
    R, O = libs
    O.orc_scene_views.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.refhost_pack_vertices.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    for d in (scenes.textured_cornell(8, 8, spp=1), scenes.kitchen_standin(tris_per_instance=60, instances=9, meshes=5, materials=4, width=8, height=8, spp=1),
              scenes.material_stress(tris_per_mesh=50, meshes=4, width=8, height=8, spp=1)):
        o = orc.OracleScene(d)
        verts = np.ascontiguousarray(np.concatenate([m.vertices for m in d.meshes]))
        assert verts.dtype.itemsize == 48
        out = np.zeros((len(verts), 8), np.float32)
        R.refhost_pack_vertices(verts.ctypes.data, len(verts), out.ctypes.data)
        faces, fverts = C.c_void_p(), C.c_void_p()
        bases = (C.c_uint32 * 3)(); o2w, w2o = (C.c_float * 12)(), (C.c_float * 12)(); count = C.c_uint32(0)
        O.orc_scene_views(o.handle, 0, C.byref(faces), C.byref(fverts), bases, o2w, w2o, C.byref(count))
        want = np.ctypeslib.as_array(C.cast(fverts, C.POINTER(C.c_float)), shape=(len(verts), 8))
        assert out.tobytes() == want.tobytes(), d.name
        o.close()


def test_caller_side_fallback_tangents(libs):
    """src/hdGatling/mesh.cpp:83-86, 231-261 (`_CalculateBitangentSign`, `_DuffOrthonormalBasis`, `_CalculateFallbackTangents`: the tangent
    frame hdGatling attaches to meshes without authored tangents or uvs), compiled from the reference's file -- against scenes.bake_vertices,
    which prepares the geometry of this repository's test and bench scenes the same way."""
    from gatling_b200 import scenes
    R, _ = libs
    R.refhost_fallback_tangents.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(29)
    normals = np.array([unit(v) for v in rng.normal(size=(4000, 3))] + [(0, 0, 1), (0, 0, -1), (1, 0, 0), (0, -1, 0), (0.6, 0.0, -0.8)], np.float32)
    t, s = np.zeros_like(normals), np.zeros(len(normals), np.float32)
    R.refhost_fallback_tangents(np.ascontiguousarray(normals).ctypes.data, len(normals), t.ctypes.data, s.ctypes.data)
    v = scenes.bake_vertices(np.zeros_like(normals), normals)
    assert np.array_equal(v["tangent"], t) and np.array_equal(v["bitangentSign"], s)
    assert set(np.unique(s)) <= {-1.0, 1.0}
