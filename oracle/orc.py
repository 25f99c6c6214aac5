"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
gatling_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """`make` decides: the library is rebuilt whenever a source or include/gtlc.h is newer (a stale liboracle.so read the
    C-ABI structs with an old layout once).  Without make / a compiler the existing library is used as it is."""
    import shutil
    if shutil.which("make") is None and os.path.exists(_LIB_PATH) and not force:
        return
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


class TextureDesc(C.Structure):  # GtlcTextureDesc
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_int32), ("depth", C.c_uint32), ("texels", C.c_void_p)]


class TextureBinding(C.Structure):  # GtlcTextureBinding
    _fields_ = [("paramIndex", C.c_int32), ("componentCount", C.c_int32), ("channels", C.c_int32 * 3), ("wrapU", C.c_int32),
                ("wrapV", C.c_int32), ("srgb", C.c_int32), ("texture", C.c_uint64), ("fallback", C.c_float * 4),
                ("scale", C.c_float * 4), ("bias", C.c_float * 4), ("uvRotation", C.c_float), ("uvScale", C.c_float * 2),
                ("uvTranslation", C.c_float * 2), ("lookup", C.c_int32), ("wrapW", C.c_int32), ("xyzScale", C.c_float * 3),
                ("xyzOffset", C.c_float * 3), ("combine", C.c_int32)]


class PrimvarBinding(C.Structure):  # GtlcPrimvarBinding
    _fields_ = [("paramIndex", C.c_int32), ("componentCount", C.c_int32), ("name", C.c_char_p), ("fallback", C.c_float * 4)]


class MaterialDesc(C.Structure):  # GtlcMaterialDesc
    _fields_ = [("closure", C.c_int32), ("name", C.c_char_p), ("params", C.c_float * 48),
                ("textures", C.POINTER(TextureBinding)), ("textureCount", C.c_uint32), ("primvarCount", C.c_uint32),
                ("primvars", C.POINTER(PrimvarBinding))]


MAX_PRIMVAR_BINDINGS = 8


class CameraDesc(C.Structure):  # GtlcCameraDesc
    _fields_ = [("position", C.c_float * 3), ("forward", C.c_float * 3), ("up", C.c_float * 3), ("vfov", C.c_float),
                ("fStop", C.c_float), ("focusDistance", C.c_float), ("focalLength", C.c_float), ("clipStart", C.c_float),
                ("clipEnd", C.c_float), ("exposure", C.c_float)]


class RenderSettings(C.Structure):  # GtlcRenderSettings
    _fields_ = [("clippingPlanes", C.c_int32), ("depthOfField", C.c_int32), ("domeLightCameraVisible", C.c_int32),
                ("filterImportanceSampling", C.c_int32), ("frame", C.c_float), ("jitteredSampling", C.c_int32),
                ("lightIntensityMultiplier", C.c_float), ("maxBounces", C.c_uint32), ("maxSampleValue", C.c_float),
                ("maxVolumeWalkLength", C.c_uint32), ("mediumStackSize", C.c_uint32), ("metersPerSceneUnit", C.c_float),
                ("nextEventEstimation", C.c_int32), ("progressiveAccumulation", C.c_int32), ("rrBounceOffset", C.c_uint32),
                ("rrInvMinTermProb", C.c_float), ("spp", C.c_uint32), ("time", C.c_float)]


INSTANCE_DTYPE = np.dtype([("vertexBase", "<u4"), ("faceBase", "<u4"), ("faceCount", "<u4"), ("flags", "<u4"),
                           ("material", "<u4"), ("instanceId", "<i4"), ("objectId", "<i4"), ("pad", "<u4"),
                           ("objectToWorld", "<f4", 12)])
assert INSTANCE_DTYPE.itemsize == 80

SPHERE_LIGHT_DTYPE = np.dtype([("pos", "<f4", 3), ("diffuseSpecularPacked", "<u4"), ("baseEmission", "<f4", 3), ("area", "<f4"),
                               ("radiusXYZ", "<f4", 3), ("padding", "<f4")])
DISTANT_LIGHT_DTYPE = np.dtype([("direction", "<f4", 3), ("angle", "<f4"), ("baseEmission", "<f4", 3),
                                ("diffuseSpecularPacked", "<u4"), ("padding", "<f4", 3), ("invPdf", "<f4")])
RECT_LIGHT_DTYPE = np.dtype([("origin", "<f4", 3), ("width", "<f4"), ("baseEmission", "<f4", 3), ("height", "<f4"),
                             ("tangentFramePacked", "<u4", 2), ("diffuseSpecularPacked", "<u4"), ("padding", "<f4")])
DISK_LIGHT_DTYPE = np.dtype([("origin", "<f4", 3), ("radiusX", "<f4"), ("baseEmission", "<f4", 3), ("radiusY", "<f4"),
                             ("tangentFramePacked", "<u4", 2), ("diffuseSpecularPacked", "<u4"), ("padding", "<f4")])
for _d in (SPHERE_LIGHT_DTYPE, DISTANT_LIGHT_DTYPE, RECT_LIGHT_DTYPE, DISK_LIGHT_DTYPE):
    assert _d.itemsize == 48

HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("inst", "<u4")])


class SceneDescC(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertexCount", C.c_uint32),
                ("faces", C.c_void_p), ("faceCount", C.c_uint32),
                ("instances", C.c_void_p), ("instanceCount", C.c_uint32),
                ("materials", C.POINTER(MaterialDesc)), ("materialCount", C.c_uint32),
                ("sphereLights", C.c_void_p), ("sphereLightCount", C.c_uint32),
                ("distantLights", C.c_void_p), ("distantLightCount", C.c_uint32),
                ("rectLights", C.c_void_p), ("rectLightCount", C.c_uint32),
                ("diskLights", C.c_void_p), ("diskLightCount", C.c_uint32),
                ("domeTexels", C.c_void_p), ("domeW", C.c_uint32), ("domeH", C.c_uint32),
                ("domeRotation", C.c_float * 4), ("domeEmission", C.c_float * 3), ("pad", C.c_uint32), ("faceIds", C.c_void_p),
                ("textures", C.POINTER(TextureDesc)), ("textureCount", C.c_uint32), ("pad2", C.c_uint32),
                ("primvarInfos", C.c_void_p), ("primvarData", C.c_void_p), ("primvarDataCount", C.c_uint64)]


class AovDescC(C.Structure):
    _fields_ = [("mask", C.c_uint32), ("pad", C.c_uint32), ("buffers", C.c_void_p * 17), ("clear", (C.c_uint8 * 16) * 17)]


class RenderDescC(C.Structure):
    _fields_ = [("camera", CameraDesc), ("settings", RenderSettings), ("clearColor", C.c_float * 4),
                ("width", C.c_uint32), ("height", C.c_uint32), ("sampleOffset", C.c_uint32),
                ("rowBegin", C.c_uint32), ("rowEnd", C.c_uint32),
                ("shardRank", C.c_uint32), ("shardCount", C.c_uint32), ("shardTile", C.c_uint32)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(SceneDescC)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(RenderDescC), C.c_void_p, C.c_void_p]
        L.orc_render_aovs.restype = C.c_double
        L.orc_render_aovs.argtypes = [C.c_void_p, C.POINTER(RenderDescC), C.c_void_p, C.c_void_p, C.POINTER(AovDescC)]
        L.orc_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_camera_rays.argtypes = [C.POINTER(RenderDescC), C.c_uint32, C.c_void_p]
        L.orc_hash_theironborn.restype = C.c_uint32
        L.orc_hash_theironborn.argtypes = [C.c_uint32]
        L.orc_hash_pcg32.restype = C.c_uint32
        L.orc_hash_pcg32.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_uint_as_float.restype = C.c_float
        L.orc_uint_as_float.argtypes = [C.c_uint32]
        L.orc_rng_init.restype = C.c_uint32
        L.orc_rng_init.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_rng_next1f.restype = C.c_float
        L.orc_rng_next1f.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_luminance.restype = C.c_float
        L.orc_pack_half2x16.restype = C.c_uint32
        L.orc_pack_half2x16.argtypes = [C.c_float, C.c_float]
        L.orc_encode_direction.restype = C.c_uint32
        L.orc_host_encode_direction.restype = C.c_uint32
        L.orc_num_threads.restype = C.c_int
        L.orc_tri_intersect.restype = C.c_int
        if hasattr(L, "orc_bvh8_trace"):
            L.orc_bvh8_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int,
                                         C.c_void_p, C.c_void_p]
        if hasattr(L, "orc_lbvh8_build"):
            L.orc_lbvh8_build.restype = C.c_uint32
            L.orc_lbvh8_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return _lib


def _f3(x):
    return (C.c_float * 3)(*[float(v) for v in x])


def _fa(x, n):
    a = np.ascontiguousarray(np.asarray(x, np.float32).reshape(n))
    return a, a.ctypes.data_as(C.c_void_p)


def compose_transform(mesh_xf, inst_xf):
    """Row-major 3x4 object->world of (mesh.transform, instanceTransform) as Gi.cpp:1191 composes it:
    world = inst o mesh, i.e. USD row-vector product M_mesh . M_inst.  Evaluated in float64 in a fixed
    order and rounded once (the product host code does the same)."""
    a = np.asarray(mesh_xf, np.float32).astype(np.float64)
    b = np.asarray(inst_xf, np.float32).astype(np.float64)
    r = a[:, [0]] * b[[0], :]
    for k in (1, 2, 3):
        r = r + a[:, [k]] * b[[k], :]
    g = r.T  # column-vector convention
    return g[:3, :].astype(np.float32).reshape(12)


def pack_half(a):
    return int(np.float16(a).view(np.uint16))


def camera_c(cam) -> CameraDesc:
    c = CameraDesc()
    c.position = _f3(cam.position)
    c.forward = _f3(cam.forward)
    c.up = _f3(cam.up)
    for k in ("vfov", "fStop", "focusDistance", "focalLength", "clipStart", "clipEnd", "exposure"):
        setattr(c, k, float(getattr(cam, k)))
    return c


def settings_c(st) -> RenderSettings:
    s = RenderSettings()
    for name, ctype in RenderSettings._fields_:
        v = getattr(st, name)
        setattr(s, name, float(v) if ctype is C.c_float else int(v))
    return s


def light_arrays(desc):
    """Host-side light encoding exactly as Gi.cpp:2573-2941 does it (restated through the oracle's own
    encoders)."""
    L = lib()
    sp = np.zeros(len(desc.sphere_lights), SPHERE_LIGHT_DTYPE)
    for i, l in enumerate(desc.sphere_lights):
        L.orc_sphere_light_area.restype = C.c_float
        L.orc_sphere_light_area.argtypes = [C.c_float] * 3
        area = np.float32(L.orc_sphere_light_area(*[float(np.float32(v)) for v in l.radius]))
        sp[i] = (l.position, L.orc_pack_half2x16(l.diffuse, l.specular), l.base_emission, area, l.radius, 0.0)
    di = np.zeros(len(desc.distant_lights), DISTANT_LIGHT_DTYPE)
    for i, l in enumerate(desc.distant_lights):
        L.orc_distant_light_inv_pdf.restype = C.c_float
        L.orc_distant_light_inv_pdf.argtypes = [C.c_float]
        inv_pdf = np.float32(L.orc_distant_light_inv_pdf(float(np.float32(l.angle))))
        di[i] = (l.direction, l.angle, l.base_emission, L.orc_pack_half2x16(l.diffuse, l.specular), (0, 0, 0), inv_pdf)
    re = np.zeros(len(desc.rect_lights), RECT_LIGHT_DTYPE)
    for i, l in enumerate(desc.rect_lights):
        t0 = L.orc_host_encode_direction(_f3(l.t0))
        t1 = L.orc_host_encode_direction(_f3(l.t1))
        re[i] = (l.origin, l.width, l.base_emission, l.height, (t0, t1), L.orc_pack_half2x16(l.diffuse, l.specular), 0.0)
    dk = np.zeros(len(desc.disk_lights), DISK_LIGHT_DTYPE)
    for i, l in enumerate(desc.disk_lights):
        t0 = L.orc_host_encode_direction(_f3(l.t0))
        t1 = L.orc_host_encode_direction(_f3(l.t1))
        dk[i] = (l.origin, l.radius[0], l.base_emission, l.radius[1], (t0, t1), L.orc_pack_half2x16(l.diffuse, l.specular), 0.0)
    return sp, di, re, dk


class OracleScene:
    """Flattens a gatling_b200.scenes.SceneDesc and builds the oracle's scene."""

    def __init__(self, desc):
        self.desc = desc
        L = lib()
        verts, faces, inst = [], [], []
        vbase = fbase = 0
        for mi, m in enumerate(desc.meshes):
            verts.append(m.vertices)
            faces.append(np.ascontiguousarray(m.faces, np.uint32))
            xfs = m.instance_transforms if m.instance_transforms is not None else np.eye(4, dtype=np.float32)[None]
            flags = (1 if m.left_handed else 0) | (2 if m.double_sided else 0)
            for k, ixf in enumerate(xfs):
                rec = np.zeros(1, INSTANCE_DTYPE)
                rec["vertexBase"] = vbase
                rec["faceBase"] = fbase
                rec["faceCount"] = len(m.faces)
                rec["flags"] = flags
                rec["material"] = m.material
                rec["instanceId"] = k if getattr(m, "instance_ids", None) is None else int(m.instance_ids[k])
                rec["objectId"] = m.object_id
                rec["objectToWorld"] = compose_transform(m.transform, ixf)
                inst.append(rec)
            vbase += len(m.vertices)
            fbase += len(m.faces)
        if verts:
            self.vertices = np.ascontiguousarray(np.concatenate(verts))
            self.faces = np.ascontiguousarray(np.concatenate(faces).reshape(-1, 3))
            self.instances = np.ascontiguousarray(np.concatenate(inst))
        else:                                             # an empty stage (the reference's Render.Empty1x1 test)
            from gatling_b200.scenes import VERTEX_DTYPE
            self.vertices = np.zeros(0, VERTEX_DTYPE)
            self.faces = np.zeros((0, 3), np.uint32)
            self.instances = np.zeros(0, INSTANCE_DTYPE)
        self._keep = []
        self.mats = (MaterialDesc * len(desc.materials))()
        for i, m in enumerate(desc.materials):
            self.mats[i].closure = m.closure
            self.mats[i].name = None
            self.mats[i].params = (C.c_float * 48)(*[float(x) for x in m.params])
            if getattr(m, "textures", None):
                arr = (TextureBinding * len(m.textures))()
                for k, b in enumerate(m.textures):
                    arr[k].paramIndex = b.param
                    arr[k].componentCount = len(b.channels)
                    arr[k].channels = (C.c_int32 * 3)(*(list(b.channels) + [0, 0, 0])[:3])
                    arr[k].wrapU, arr[k].wrapV, arr[k].srgb = b.wrap_u, b.wrap_v, int(b.srgb)
                    arr[k].texture = b.texture + 1          # 1-based index into SceneDesc.textures
                    arr[k].fallback = (C.c_float * 4)(*b.fallback)
                    arr[k].scale = (C.c_float * 4)(*b.scale)
                    arr[k].bias = (C.c_float * 4)(*b.bias)
                    arr[k].uvRotation = b.uv_rotation
                    arr[k].uvScale = (C.c_float * 2)(*b.uv_scale)
                    arr[k].uvTranslation = (C.c_float * 2)(*b.uv_translation)
                    arr[k].lookup, arr[k].wrapW = b.lookup, b.wrap_w
                    arr[k].xyzScale = (C.c_float * 3)(*b.xyz_scale)
                    arr[k].xyzOffset = (C.c_float * 3)(*b.xyz_offset)
                    arr[k].combine = b.combine
                self._keep.append(arr)
                self.mats[i].textures = arr
                self.mats[i].textureCount = len(m.textures)
            if getattr(m, "primvars", None):
                parr = (PrimvarBinding * len(m.primvars))()
                for k, b in enumerate(m.primvars):
                    parr[k].paramIndex, parr[k].componentCount, parr[k].name = b.param, b.count, b.name.encode()
                    parr[k].fallback = (C.c_float * 4)(*b.fallback)
                self._keep.append(parr)
                self.mats[i].primvars = parr
                self.mats[i].primvarCount = len(m.primvars)
        self.sp, self.di, self.re, self.dk = light_arrays(desc)
        d = SceneDescC()
        d.vertices = self.vertices.ctypes.data
        d.vertexCount = len(self.vertices)
        d.faces = self.faces.ctypes.data
        d.faceCount = len(self.faces)
        d.instances = self.instances.ctypes.data
        d.instanceCount = len(self.instances)
        d.materials = self.mats
        d.materialCount = len(desc.materials)
        d.sphereLights = self.sp.ctypes.data if len(self.sp) else None
        d.sphereLightCount = len(self.sp)
        d.distantLights = self.di.ctypes.data if len(self.di) else None
        d.distantLightCount = len(self.di)
        d.rectLights = self.re.ctypes.data if len(self.re) else None
        d.rectLightCount = len(self.re)
        d.diskLights = self.dk.ctypes.data if len(self.dk) else None
        d.diskLightCount = len(self.dk)
        if desc.dome_light is not None:
            self.dome = np.ascontiguousarray(desc.dome_light.texels, np.float32)
            d.domeTexels = self.dome.ctypes.data
            d.domeH, d.domeW = self.dome.shape[0], self.dome.shape[1]
            d.domeRotation = (C.c_float * 4)(*desc.dome_light.rotation)
            d.domeEmission = (C.c_float * 3)(*desc.dome_light.base_emission)
        # scene data (Gi.cpp:905-1018): per instance, the primvar each binding of its material resolves to --
        # mesh primvars override instancer primvars; too-short arrays count as "not found"
        if any(getattr(m, "primvars", None) for m in desc.materials):
            infos = np.zeros((len(self.instances), MAX_PRIMVAR_BINDINGS, 2), np.uint32)
            chunks, offset, row = [], 0, 0
            for m in desc.meshes:
                n_inst = 1 if m.instance_transforms is None else len(m.instance_transforms)
                ids = np.arange(n_inst) if getattr(m, "instance_ids", None) is None else np.asarray(m.instance_ids)
                resolved = []
                for k, b in enumerate(desc.materials[m.material].primvars[:MAX_PRIMVAR_BINDINGS]):
                    found = None
                    for p in list(getattr(m, "instancer_primvars", [])) + list(getattr(m, "primvars", [])):
                        if p.name == b.name and np.size(p.data):
                            found = p
                    if found is None:
                        resolved.append((0, 0))
                        continue
                    is_int = np.issubdtype(np.asarray(found.data).dtype, np.integer)       # Int..Int4: scene_data_lookup_int, stored as bit patterns
                    a = np.ascontiguousarray(found.data, np.int32).view(np.float32) if is_int else np.ascontiguousarray(found.data, np.float32)
                    a = a.reshape(len(a), -1)
                    need = {0: 1, 1: int(max(ids.max(), 0)) + 1, 2: len(m.faces), 3: len(m.vertices)}[found.interpolation]
                    if len(a) < need:
                        resolved.append((0, 0))
                        continue
                    resolved.append((offset, a.shape[1] | (found.interpolation << 4) | 0x100 | (0x200 if is_int else 0)))
                    chunks.append(a.reshape(-1))
                    offset += a.size
                for _ in range(n_inst):
                    for k, r in enumerate(resolved):
                        infos[row, k] = r
                    row += 1
            self.pv_infos = np.ascontiguousarray(infos)
            self.pv_data = np.ascontiguousarray(np.concatenate(chunks)) if chunks else np.zeros(1, np.float32)
            d.primvarInfos = self.pv_infos.ctypes.data
            d.primvarData = self.pv_data.ctypes.data
            d.primvarDataCount = len(self.pv_data)
        if getattr(desc, "textures", None):
            self.texs = (TextureDesc * len(desc.textures))()
            for i, t in enumerate(desc.textures):
                a = np.ascontiguousarray(t.texels)
                self._keep.append(a)
                self.texs[i].height, self.texs[i].width = t.rows_cols
                self.texs[i].depth = t.depth
                self.texs[i].format = t.format
                self.texs[i].texels = a.ctypes.data
            d.textures = self.texs
            d.textureCount = len(desc.textures)
        self.handle = L.orc_scene_create(C.byref(d))

    def close(self):
        if self.handle:
            lib().orc_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    SHARD_SAMPLES = 0xFFFFFFFF   # as the tile size of `shard=(rank, world, tile)`: sample-range sharding (gtlc.h GTLC_SHARD_SAMPLES)

    def render_desc(self, sample_offset=0, rows=None, shard=None, spp=None) -> RenderDescC:
        d = self.desc
        r = RenderDescC()
        r.camera = camera_c(d.camera)
        r.settings = settings_c(d.settings)
        if spp is not None:
            r.settings.spp = spp
        r.clearColor = (C.c_float * 4)(*d.clear_color)
        r.width, r.height = d.width, d.height
        r.sampleOffset = sample_offset
        r.rowBegin, r.rowEnd = rows if rows else (0, d.height)
        r.shardRank, r.shardCount, r.shardTile = shard if shard else (0, 1, 64)
        return r

    def render(self, sample_offset=0, prev=None, rows=None, shard=None, spp=None):
        """Returns (rgba float32[h, w, 4] with row 0 = bottom, seconds, (closestRays, shadowRays))."""
        d = self.desc
        img = np.zeros((d.height, d.width, 4), np.float32) if prev is None else np.ascontiguousarray(prev, np.float32).copy()
        counts = np.zeros(2, np.uint64)
        r = self.render_desc(sample_offset, rows, shard, spp)
        secs = lib().orc_render(self.handle, C.byref(r), img.ctypes.data, counts.ctypes.data)
        return img, secs, (int(counts[0]), int(counts[1]))

    def render_aovs(self, aov_ids, clear_values=None, sample_offset=0, prev=None, spp=None):
        """Colour + auxiliary AOVs; `prev` = dict of the previous call's outputs for progressive refinement."""
        d = self.desc
        INT = (11, 13, 14)
        out = {}
        a = AovDescC()
        for aov in aov_ids:
            aov = int(aov)
            if aov == 0:
                continue
            if aov in INT:
                buf = np.zeros((d.height, d.width, 1), np.int32)
            elif aov == 12:
                buf = np.zeros((d.height, d.width, 1), np.float32)
            else:
                buf = np.zeros((d.height, d.width, 4), np.float32)
            if prev is not None and aov in prev:
                buf[...] = prev[aov]
            out[aov] = buf
            a.mask |= (1 << aov)
            a.buffers[aov] = buf.ctypes.data
            clear = (clear_values or {}).get(aov, (0.0, 0.0, 0.0, 0.0))
            raw = bytes(clear) if isinstance(clear, (bytes, bytearray)) else (np.asarray(clear, np.int32).tobytes() if aov in INT else np.asarray(clear, np.float32).tobytes())
            raw = (raw + b"\0" * 16)[:16]
            a.clear[aov] = (C.c_uint8 * 16)(*raw)
        img = np.zeros((d.height, d.width, 4), np.float32) if prev is None or 0 not in prev else np.ascontiguousarray(prev[0], np.float32).copy()
        counts = np.zeros(2, np.uint64)
        r = self.render_desc(sample_offset, None, None, spp)
        lib().orc_render_aovs(self.handle, C.byref(r), img.ctypes.data, counts.ctypes.data, C.byref(a))
        out[0] = img
        return out

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(len(rays), HIT_DTYPE)
        lib().orc_trace(self.handle, rays.ctypes.data, len(rays), 1 if any_hit else 0, hits.ctypes.data)
        return hits


def camera_rays(desc, sample_index=0):
    """(n, 8) float32: ox oy oz rngbits dx dy dz 0 for every pixel (rp_main.rgen:191-271)."""
    tmp = type("T", (), {})()
    r = RenderDescC()
    r.camera = camera_c(desc.camera)
    r.settings = settings_c(desc.settings)
    r.clearColor = (C.c_float * 4)(*desc.clear_color)
    r.width, r.height = desc.width, desc.height
    out = np.zeros((desc.width * desc.height, 8), np.float32)
    lib().orc_camera_rays(C.byref(r), sample_index, out.ctypes.data)
    return out


def bvh8_trace(nodes, tris, rays, any_hit=False):
    """Walks the product's BVH8 bytes with the oracle's restatement of the traversal order; returns
    (hits, visits[n, 2])."""
    L = lib()
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    hits = np.zeros(len(rays), HIT_DTYPE)
    visits = np.zeros((len(rays), 2), np.uint32)
    L.orc_bvh8_trace(nodes.ctypes.data, len(nodes) // 80, tris.ctypes.data, len(tris) // 12, rays.ctypes.data, len(rays),
                     1 if any_hit else 0, hits.ctypes.data, visits.ctypes.data)
    return hits, visits


def lbvh8_build(tris_rank_order, mode=1):
    """CPU restatement of the device BVH8 builder (orc_lbvh8.h): triangle records (n x 12 float32, in
    (instance, primitive) order) -> (nodes bytes [nodeCount*80], tris float32 [n*12], depth).
    mode 1 = Karras radix tree (bvhBuilder 1), mode 2 = PLOC (bvhBuilder 2)."""
    L = lib()
    t = np.ascontiguousarray(tris_rank_order, np.float32).reshape(-1, 12)
    n = len(t)
    nodes = np.zeros(max(n, 1) * 20, np.uint32)
    out = np.zeros(max(n, 1) * 12, np.float32)
    depth = C.c_uint32(0)
    count = L.orc_lbvh8_build(t.ctypes.data, n, nodes.ctypes.data, out.ctypes.data, C.byref(depth), int(mode))
    return nodes[:count * 20].view(np.uint8).copy(), out[:n * 12].copy(), int(depth.value)


def tex_lookup(texture, u, v, wrap_u=1, wrap_v=1):
    """One tex_lookup_float4_2d of the oracle's restatement; returns (rgba float32[4], clipped)."""
    L = lib()
    d, keep = _texture_desc(texture)
    out = (C.c_float * 4)()
    L.orc_tex_lookup.argtypes = [C.POINTER(TextureDesc), C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    clipped = L.orc_tex_lookup(C.byref(d), float(u), float(v), int(wrap_u), int(wrap_v), out)
    return np.array(out[:], np.float32), bool(clipped)


def _texture_desc(texture):
    a = np.ascontiguousarray(texture.texels)
    d = TextureDesc()
    d.height, d.width = texture.rows_cols
    d.depth = texture.depth
    d.format = texture.format
    d.texels = a.ctypes.data
    return d, a


def tex_lookup_3d(texture, u, v, w, wrap_u, wrap_v, wrap_w):
    """tex_lookup_float4_3d (mdl_interface.glsl:45-64) of the oracle: (rgba, clipped)."""
    L = lib()
    d, keep = _texture_desc(texture)
    out = (C.c_float * 4)()
    L.orc_tex_lookup_3d.argtypes = [C.POINTER(TextureDesc), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]
    clipped = L.orc_tex_lookup_3d(C.byref(d), float(u), float(v), float(w), int(wrap_u), int(wrap_v), int(wrap_w), out)
    return np.array(out[:], np.float32), bool(clipped)


def tex_texel(texture, x, y, z=0):
    """tex_texel_float4_2d / _3d (mdl_interface.glsl:167-185, 91-110) of the oracle."""
    L = lib()
    d, keep = _texture_desc(texture)
    out = (C.c_float * 4)()
    L.orc_tex_texel.argtypes = [C.POINTER(TextureDesc), C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orc_tex_texel.restype = None
    L.orc_tex_texel(C.byref(d), int(x), int(y), int(z), out)
    return np.array(out[:], np.float32)


def srgb_texture_to_linear(c):
    L = lib()
    L.orc_srgb_texture_to_linear.restype = C.c_float
    L.orc_srgb_texture_to_linear.argtypes = [C.c_float]
    return float(L.orc_srgb_texture_to_linear(float(c)))


# ---- the reference's integrator, compiled from its own shader files (oracle/ref_glsl, oracle/_ref/libglslref.so) ------------------------------
class _RefCallbacks(C.Structure):
    _fields_ = [("scene", C.c_void_p), ("renderDesc", C.c_void_p)] + [(n, C.c_void_p) for n in
                ("trace", "views", "hitBegin", "mappedNormal", "volume", "bsdfSample", "bsdfEvaluate", "bsdfAlbedo", "edf", "dome")]


class ReferenceIntegrator:
    """rp_main.rgen / .chit / .miss / _shadow.miss of the reference, compiled as C++ from the reference's files (oracle/ref_glsl/build_ref.py),
    rendering an OracleScene: ray queries, the materials' mdl_* entry points and the dome texture are served by the oracle through a callback
    table, everything else -- camera rays, the bounce loop, NEE, roulette, media, accumulation, AOVs -- is the reference's own code.
    TEST INFRASTRUCTURE (and bench.py's reference arm)."""
    SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libglslref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.SO)

    def __init__(self):
        if not self.available():
            raise RuntimeError("oracle/_ref/libglslref.so is missing: it is built from /root/reference by oracle/ref_glsl/build_ref.py in the build container")
        self.R = C.CDLL(self.SO)
        self.R.ref_render.argtypes = [C.c_void_p] * 15
        self.R.ref_render.restype = C.c_int

    def render(self, scene: "OracleScene", sample_offset=0, prev=None, rows=None, spp=None, aov_ids=None, clear_values=None):
        """Returns (rgba float32[h, w, 4] -- or {aov id: array} with `aov_ids` --, (closest, shadow) ray counts, seconds)."""
        import time
        O, desc = lib(), scene.desc
        r = scene.render_desc(sample_offset, None, None, spp)
        cb = _RefCallbacks()
        cb.scene, cb.renderDesc = scene.handle, C.cast(C.pointer(r), C.c_void_p)
        for field, name in (("trace", "orc_trace_one"), ("views", "orc_scene_views"), ("hitBegin", "orc_hit_begin"), ("mappedNormal", "orc_hit_mapped_normal"),
                            ("volume", "orc_hit_volume"), ("bsdfSample", "orc_hit_bsdf_sample"), ("bsdfEvaluate", "orc_hit_bsdf_evaluate"),
                            ("bsdfAlbedo", "orc_hit_bsdf_albedo"), ("edf", "orc_hit_edf"), ("dome", "orc_dome_fetch")):
            setattr(cb, field, C.cast(getattr(O, name), C.c_void_p))
        f28, u16 = (C.c_float * 28)(), (C.c_uint32 * 16)()
        O.orc_ubo_values.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        O.orc_ubo_values(scene.handle, C.byref(r), f28, u16)
        lights = [C.c_void_p() for _ in range(4)]
        O.orc_light_buffers.argtypes = [C.c_void_p] * 5
        O.orc_light_buffers(scene.handle, *[C.byref(p) for p in lights])
        img = np.zeros((desc.height, desc.width, 4), np.float32) if prev is None else np.ascontiguousarray(prev if aov_ids is None else prev[0], np.float32).copy()
        counts = (C.c_uint64 * 2)()
        row_arg = (C.c_uint32 * 2)(*rows) if rows else None
        if aov_ids is None:
            t0 = time.perf_counter()
            ok = self.R.ref_render(C.byref(cb), f28, u16, lights[0], lights[1], lights[2], lights[3], img.ctypes.data, counts, None, None, None, None, None, row_arg)
            secs = time.perf_counter() - t0
            if ok != 1:
                raise RuntimeError("no reference pipeline was compiled for these render settings (oracle/ref_glsl/ref_exports.cpp)")
            return img, (int(counts[0]), int(counts[1])), secs
        INT, n = (11, 14), desc.width * desc.height
        bufs = {a: (np.zeros((n, 1), np.int32) if a in INT else np.zeros((n, 1), np.float32) if a == 12 else np.zeros((n, 3), np.float32)) for a in range(17)}
        ptrs = (C.c_void_p * 17)(*[bufs[a].ctypes.data for a in range(17)])
        clear_f, clear_i = np.zeros((17, 4), np.float32), np.zeros((17, 4), np.int32)
        for a, v in (clear_values or {}).items():
            (clear_i if a in INT else clear_f)[a, :len(v)] = v
        ninst = sum(1 if m.instance_transforms is None else len(m.instance_transforms) for m in desc.meshes) + 1
        inst_ids, obj_ids = np.zeros(ninst, np.int32), np.zeros(ninst, np.int32)
        O.orc_instance_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        O.orc_instance_ids(scene.handle, inst_ids.ctypes.data, obj_ids.ctypes.data)
        t0 = time.perf_counter()
        ok = self.R.ref_render(C.byref(cb), f28, u16, lights[0], lights[1], lights[2], lights[3], img.ctypes.data, counts, ptrs, clear_f.ctypes.data, clear_i.ctypes.data,
                               inst_ids.ctypes.data, obj_ids.ctypes.data, row_arg)
        secs = time.perf_counter() - t0
        if ok != 1:
            raise RuntimeError("no reference pipeline with auxiliary AOVs was compiled for these render settings")
        got = {0: img}
        for a in aov_ids:
            if a != 0:
                got[a] = bufs[a].reshape(desc.height, desc.width, -1)
        return got, (int(counts[0]), int(counts[1])), secs
