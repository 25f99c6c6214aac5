"""Python mirror of the reference's `gi` renderer API (/root/reference/src/gi/gtl/gi/Gi.h:199-261).

Same function names, argument meaning and error behaviour (GiStatus / None on failure, no exceptions
for renderer errors), implemented over the C-ABI in include/gtlc.h.  The C++ facade with the exact
reference signatures is include/gtl/gi/Gi.h + gatling_b200/gi_cpp/Gi.cpp; this module is what the tests
and bench.py drive.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from . import scenes as S


class GiStatus(enum.IntEnum):  # Gi.h:32
    Ok = 0
    Error = 1


class GiAovId(enum.IntEnum):  # Gi.h:34-54
    Color = 0
    Normal = 1
    NEE = 2
    Barycentrics = 3
    Texcoords = 4
    Bounces = 5
    ClockCycles = 6
    Opacity = 7
    Tangents = 8
    Bitangents = 9
    ThinWalled = 10
    ObjectId = 11
    Depth = 12
    FaceId = 13
    InstanceId = 14
    DoubleSided = 15
    Albedo = 16


class GiRenderBufferFormat(enum.IntEnum):  # Gi.h:68-73
    Int32 = 0
    Float32 = 1
    Float32Vec4 = 2


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _lib():
    return L.load()


def giLastError() -> str:
    return (_lib().gtlc_last_error() or b"").decode()


def giErrorCount() -> int:
    """Error lines logged so far (gtlc_error_count): lets a test fail on any error line like the reference's fixture."""
    return int(_lib().gtlc_error_count())


# ---- lifecycle (Gi.h:199-200) ---------------------------------------------------------------------
def giInitialize(device_index: int = -1) -> GiStatus:
    p = L.InitParams(device_index, 0)
    return GiStatus(_lib().gtlc_initialize(C.byref(p)))


def giTerminate():
    _lib().gtlc_terminate()


# ---- scene (Gi.h:220-221) ---------------------------------------------------------------------------
def giCreateScene():
    return _lib().gtlc_scene_create() or None


def giDestroyScene(scene):
    _lib().gtlc_scene_destroy(scene)


# ---- materials (Gi.h:204-207).  The reference takes MaterialX / MDL; the binding takes the lowered
# closure + parameter block (see gtlc.h) -- gi_cpp/Gi.cpp does the MaterialX-XML lowering. ------------
def giCreateTexture(texels: np.ndarray):
    """Decoded image -> texture handle.  texels[row, col, channels], row 0 = bottom; dtype uint8 (RGBA8), float16 (RGBA16F),
    float32 with 1 channel (R32F) or 4 channels (RGBA32F).  (The reference decodes files through imgio, out of scope.)"""
    t = texels if isinstance(texels, S.Texture) else S.Texture(np.ascontiguousarray(texels))
    t = S.Texture(np.ascontiguousarray(t.texels), t.depth)
    d = L.TextureDesc()
    d.height, d.width = t.rows_cols
    d.depth = t.depth
    d.format = t.format
    d.texels = t.texels.ctypes.data
    return _lib().gtlc_texture_create(C.byref(d)) or None


def giDestroyTexture(tex):
    _lib().gtlc_texture_destroy(tex)


def texture_binding_c(b: "S.TextureBinding", handle) -> "L.TextureBinding":
    c = L.TextureBinding()
    c.paramIndex = b.param
    c.componentCount = len(b.channels)
    c.channels = (C.c_int32 * 3)(*(list(b.channels) + [0, 0, 0])[:3])
    c.wrapU, c.wrapV, c.srgb = b.wrap_u, b.wrap_v, int(b.srgb)
    c.texture = handle
    c.fallback = (C.c_float * 4)(*b.fallback)
    c.scale = (C.c_float * 4)(*b.scale)
    c.bias = (C.c_float * 4)(*b.bias)
    c.uvRotation = b.uv_rotation
    c.uvScale = (C.c_float * 2)(*b.uv_scale)
    c.uvTranslation = (C.c_float * 2)(*b.uv_translation)
    c.lookup, c.wrapW = b.lookup, b.wrap_w
    c.xyzScale = (C.c_float * 3)(*b.xyz_scale)
    c.xyzOffset = (C.c_float * 3)(*b.xyz_offset)
    c.combine = b.combine
    return c


def giCreateMaterial(scene, closure: int, params: Sequence[float], name: str = "", textures=(), texture_handles=(), primvars=()):
    """`textures`: scenes.TextureBinding list whose .texture indexes `texture_handles`; `primvars`: scenes.PrimvarBinding list."""
    d = L.MaterialDesc()
    d.closure = closure
    d.name = name.encode() if name else None
    arr = np.zeros(48, np.float32)
    arr[:len(params)] = params
    d.params = (C.c_float * 48)(*[float(x) for x in arr])
    if textures:
        arr_b = (L.TextureBinding * len(textures))(*[texture_binding_c(b, texture_handles[b.texture] if b.texture >= 0 else 0) for b in textures])
        d.textures = arr_b
        d.textureCount = len(textures)
    if primvars:
        arr_p = (L.PrimvarBinding * len(primvars))()
        for i, b in enumerate(primvars):
            arr_p[i].paramIndex, arr_p[i].componentCount, arr_p[i].name = b.param, b.count, b.name.encode()
            arr_p[i].fallback = (C.c_float * 4)(*b.fallback)
        d.primvars = arr_p
        d.primvarCount = len(primvars)
    return _lib().gtlc_material_create(scene, C.byref(d)) or None


def giDestroyMaterial(mat):
    _lib().gtlc_material_destroy(mat)


# ---- meshes (Gi.h:209-216) ----------------------------------------------------------------------------
def _primvar_descs(primvars):
    """scenes.Primvar list -> (ctypes array, keep-alive list)."""
    arr = (L.PrimvarDesc * max(len(primvars), 1))()
    keep = []
    for i, p in enumerate(primvars):
        is_int = np.issubdtype(np.asarray(p.data).dtype, np.integer)
        a = np.ascontiguousarray(p.data, np.int32 if is_int else np.float32)
        a = a.reshape(len(a), -1)
        keep.append(a)
        arr[i].name = p.name.encode()
        arr[i].type = a.shape[1] - 1 + (4 if is_int else 0)    # GTLC_PRIMVAR_FLOAT .. VEC4, GTLC_PRIMVAR_INT .. INT4
        arr[i].interpolation = p.interpolation
        arr[i].data = a.ctypes.data
        arr[i].dataSize = a.nbytes
    return arr, keep


def giCreateMesh(scene, vertices: np.ndarray, faces: np.ndarray, *, id: int = 0, isDoubleSided=True, isLeftHanded=False,
                 name: str = "", faceIds: Optional[np.ndarray] = None, primvars=()):
    v = np.ascontiguousarray(vertices)
    assert v.dtype.itemsize == 48, "vertices must be GiVertex records (48 bytes)"
    f = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    d = L.MeshDesc()
    d.faceCount = len(f)
    d.faces = f.ctypes.data
    fid = None
    if faceIds is not None:
        fid = np.ascontiguousarray(faceIds, np.int32)
        d.faceIds = fid.ctypes.data
    d.id = id
    d.isDoubleSided = int(isDoubleSided)
    d.isLeftHanded = int(isLeftHanded)
    d.name = name.encode() if name else None
    d.maxFaceId = int(fid.max()) if fid is not None and len(fid) else max(len(f) - 1, 0)
    d.vertexCount = len(v)
    d.vertices = v.ctypes.data
    if primvars:
        arr, keep = _primvar_descs(primvars)
        d.primvars = arr
        d.primvarCount = len(primvars)
    return _lib().gtlc_mesh_create(scene, C.byref(d)) or None


def giSetMeshTransform(mesh, mat4x4):
    m = np.ascontiguousarray(mat4x4, np.float32).reshape(16)
    _lib().gtlc_mesh_set_transform(mesh, m.ctypes.data)


def giSetMeshInstanceTransforms(mesh, transforms):
    t = np.ascontiguousarray(transforms, np.float32).reshape(-1, 16)
    _lib().gtlc_mesh_set_instance_transforms(mesh, len(t), t.ctypes.data)


def giSetMeshInstancerPrimvars(mesh, primvars):
    arr, keep = _primvar_descs(primvars)
    _lib().gtlc_mesh_set_instancer_primvars(mesh, len(primvars), arr)


def giSetMeshInstanceIds(mesh, ids):
    a = np.ascontiguousarray(ids, np.int32)
    _lib().gtlc_mesh_set_instance_ids(mesh, len(a), a.ctypes.data)


def giSetMeshMaterial(mesh, mat):
    _lib().gtlc_mesh_set_material(mesh, mat or 0)


def giSetMeshVisibility(mesh, visible: bool):
    _lib().gtlc_mesh_set_visibility(mesh, int(visible))


def giDestroyMesh(mesh):
    _lib().gtlc_mesh_destroy(mesh)


# ---- lights (Gi.h:223-257) -----------------------------------------------------------------------------
def giCreateSphereLight(scene): return _lib().gtlc_sphere_light_create(scene) or None
def giDestroySphereLight(scene, l): _lib().gtlc_sphere_light_destroy(scene, l)
def giSetSphereLightPosition(l, p): _lib().gtlc_sphere_light_set_position(l, _f3(p))
def giSetSphereLightBaseEmission(l, rgb): _lib().gtlc_sphere_light_set_base_emission(l, _f3(rgb))
def giSetSphereLightRadius(l, rx, ry, rz): _lib().gtlc_sphere_light_set_radius(l, rx, ry, rz)
def giSetSphereLightDiffuseSpecular(l, d, s): _lib().gtlc_sphere_light_set_diffuse_specular(l, d, s)
def giCreateDistantLight(scene): return _lib().gtlc_distant_light_create(scene) or None
def giDestroyDistantLight(scene, l): _lib().gtlc_distant_light_destroy(scene, l)
def giSetDistantLightDirection(l, d): _lib().gtlc_distant_light_set_direction(l, _f3(d))
def giSetDistantLightBaseEmission(l, rgb): _lib().gtlc_distant_light_set_base_emission(l, _f3(rgb))
def giSetDistantLightAngle(l, a): _lib().gtlc_distant_light_set_angle(l, a)
def giSetDistantLightDiffuseSpecular(l, d, s): _lib().gtlc_distant_light_set_diffuse_specular(l, d, s)
def giCreateRectLight(scene): return _lib().gtlc_rect_light_create(scene) or None
def giDestroyRectLight(scene, l): _lib().gtlc_rect_light_destroy(scene, l)
def giSetRectLightOrigin(l, o): _lib().gtlc_rect_light_set_origin(l, _f3(o))
def giSetRectLightTangents(l, t0, t1): _lib().gtlc_rect_light_set_tangents(l, _f3(t0), _f3(t1))
def giSetRectLightBaseEmission(l, rgb): _lib().gtlc_rect_light_set_base_emission(l, _f3(rgb))
def giSetRectLightDimensions(l, w, h): _lib().gtlc_rect_light_set_dimensions(l, w, h)
def giSetRectLightDiffuseSpecular(l, d, s): _lib().gtlc_rect_light_set_diffuse_specular(l, d, s)
def giCreateDiskLight(scene): return _lib().gtlc_disk_light_create(scene) or None
def giDestroyDiskLight(scene, l): _lib().gtlc_disk_light_destroy(scene, l)
def giSetDiskLightOrigin(l, o): _lib().gtlc_disk_light_set_origin(l, _f3(o))
def giSetDiskLightTangents(l, t0, t1): _lib().gtlc_disk_light_set_tangents(l, _f3(t0), _f3(t1))
def giSetDiskLightBaseEmission(l, rgb): _lib().gtlc_disk_light_set_base_emission(l, _f3(rgb))
def giSetDiskLightRadius(l, rx, ry): _lib().gtlc_disk_light_set_radius(l, rx, ry)
def giSetDiskLightDiffuseSpecular(l, d, s): _lib().gtlc_disk_light_set_diffuse_specular(l, d, s)


def giCreateDomeLight(scene, texels: np.ndarray):
    """The reference takes a file path and decodes through imgio (out of scope); here: lat-long float32[h, w, 4]."""
    t = np.ascontiguousarray(texels, np.float32)
    return _lib().gtlc_dome_light_create(scene, t.shape[1], t.shape[0], t.ctypes.data) or None


def giDestroyDomeLight(l): _lib().gtlc_dome_light_destroy(l)
def giSetDomeLightRotation(l, quat): _lib().gtlc_dome_light_set_rotation(l, (C.c_float * 4)(*quat))
def giSetDomeLightBaseEmission(l, rgb): _lib().gtlc_dome_light_set_base_emission(l, _f3(rgb))
def giSetDomeLightDiffuseSpecular(l, d, s): _lib().gtlc_dome_light_set_diffuse_specular(l, d, s)


# ---- render buffers (Gi.h:259-261) ---------------------------------------------------------------------
class RenderBuffer:
    def __init__(self, handle, width, height, fmt):
        self.handle, self.width, self.height, self.format = handle, width, height, fmt

    def mem(self) -> np.ndarray:
        """giGetRenderBufferMem: the persistently mapped (pinned) host memory, row 0 = bottom."""
        ptr = _lib().gtlc_render_buffer_mem(self.handle)
        comps = 4 if self.format == GiRenderBufferFormat.Float32Vec4 else 1
        ctype = C.c_int32 if self.format == GiRenderBufferFormat.Int32 else C.c_float
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(self.height, self.width, comps))
        return arr

    def device_ptr(self) -> int:
        return _lib().gtlc_render_buffer_device_mem(self.handle)


def giCreateRenderBuffer(width, height, fmt=GiRenderBufferFormat.Float32Vec4) -> Optional[RenderBuffer]:
    h = _lib().gtlc_render_buffer_create(width, height, int(fmt))
    return RenderBuffer(h, width, height, fmt) if h else None


def giDestroyRenderBuffer(rb: RenderBuffer):
    _lib().gtlc_render_buffer_destroy(rb.handle)


def giGetRenderBufferMem(rb: RenderBuffer) -> np.ndarray:
    return rb.mem()


# ---- render (Gi.h:218) ----------------------------------------------------------------------------------
def _camera_c(cam: S.Camera) -> L.CameraDesc:
    c = L.CameraDesc()
    c.position = _f3(cam.position)
    c.forward = _f3(cam.forward)
    c.up = _f3(cam.up)
    for k in ("vfov", "fStop", "focusDistance", "focalLength", "clipStart", "clipEnd", "exposure"):
        setattr(c, k, float(getattr(cam, k)))
    return c


def _settings_c(st: S.RenderSettings) -> L.RenderSettings:
    s = L.RenderSettings()
    for name, ctype in L.RenderSettings._fields_:
        v = getattr(st, name)
        setattr(s, name, float(v) if ctype is C.c_float else int(v))
    return s


def giRender(scene, camera: S.Camera, settings: S.RenderSettings, aov_bindings, dome_light=None) -> GiStatus:
    """aov_bindings: list of (GiAovId, clearValue bytes-or-floats, RenderBuffer)."""
    arr = (L.AovBinding * len(aov_bindings))()
    for i, (aov, clear, rb) in enumerate(aov_bindings):
        arr[i].aovId = int(aov)
        raw = bytes(clear) if isinstance(clear, (bytes, bytearray)) else np.asarray(clear, np.float32).tobytes()
        raw = (raw + b"\0" * 16)[:16]
        arr[i].clearValue = (C.c_uint8 * 16)(*raw)
        arr[i].renderBuffer = rb.handle
    p = L.RenderParams()
    p.aovBindingCount = len(aov_bindings)
    p.aovBindings = arr
    p.camera = _camera_c(camera)
    p.domeLight = dome_light or 0
    p.renderSettings = _settings_c(settings)
    p.scene = scene
    return GiStatus(_lib().gtlc_render(C.byref(p)))


# ---- extensions (no Gi.h counterpart) ---------------------------------------------------
SHARD_SAMPLES = 0xFFFFFFFF   # GTLC_SHARD_SAMPLES: as `tile`, selects sample-range sharding (every pixel, a share of the samples)


def set_shard(scene, rank, world_size, tile=64):
    _lib().gtlc_scene_set_shard(scene, rank, world_size, tile)


COMM_ID_BYTES = 128


def comm_get_unique_id() -> bytes:
    """Rank 0: the NCCL unique id (128 bytes) the other ranks need for comm_init; carry it with any host-side channel."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    if _lib().gtlc_comm_get_unique_id(buf) != 0:
        raise RuntimeError("gtlc_comm_get_unique_id failed: " + giLastError())
    return buf.raw


def comm_init(unique_id: bytes, rank: int, world_size: int):
    """Joins the library's NCCL communicator: giRender of a sharded scene then ends with the HDR sum-reduce to rank 0."""
    assert len(unique_id) == COMM_ID_BYTES
    buf = C.create_string_buffer(unique_id, COMM_ID_BYTES)
    if _lib().gtlc_comm_init(buf, rank, world_size) != 0:
        raise RuntimeError("gtlc_comm_init failed: " + giLastError())


def comm_destroy():
    _lib().gtlc_comm_destroy()


def set_tuning(**kw):
    t = L.Tuning()
    _lib().gtlc_get_tuning(C.byref(t))
    for k, v in kw.items():
        setattr(t, k, v)
    _lib().gtlc_set_tuning(C.byref(t))


def get_stats() -> dict:
    s = L.Stats()
    _lib().gtlc_get_stats(C.byref(s))
    return s.as_dict()


HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("inst", "<u4")])


def trace_rays(scene, rays: np.ndarray, any_hit=False, want_visits=False):
    r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    hits = np.zeros(len(r), HIT_DTYPE)
    visits = np.zeros((len(r), 2), np.uint32) if want_visits else None
    st = _lib().gtlc_trace_rays(scene, r.ctypes.data, len(r), int(any_hit), hits.ctypes.data,
                                visits.ctypes.data if want_visits else None)
    if st != 0:
        raise RuntimeError("gtlc_trace_rays failed: " + giLastError())
    return (hits, visits) if want_visits else hits


def get_bvh(scene):
    """(nodes uint8[n*80], tris float32[m*12]) copies of the host BVH8."""
    v = L.BvhView()
    if _lib().gtlc_scene_get_bvh(scene, C.byref(v)) != 0:
        raise RuntimeError("gtlc_scene_get_bvh failed: " + giLastError())
    nodes = np.ctypeslib.as_array(C.cast(v.nodes, C.POINTER(C.c_uint8)), shape=(v.nodeCount * 80,)).copy()
    tris = np.ctypeslib.as_array(C.cast(v.triangles, C.POINTER(C.c_float)), shape=(v.triCount * 12,)).copy() if v.triCount else np.zeros(0, np.float32)
    return nodes, tris


# ---- SceneDesc -> gi calls (what hdGatling's prim sync does, src/hdGatling/mesh.cpp:470-610) --------------
class BuiltScene:
    def __init__(self, desc: S.SceneDesc):
        self.desc = desc
        self.scene = giCreateScene()
        self.textures = [giCreateTexture(t) for t in desc.textures]
        if any(t is None for t in self.textures):
            raise RuntimeError("giCreateTexture failed: " + giLastError())
        self.materials = [giCreateMaterial(self.scene, m.closure, m.params, m.name, m.textures, self.textures, m.primvars) for m in desc.materials]
        if any(h is None for h in self.materials):
            raise RuntimeError("giCreateMaterial failed: " + giLastError())
        self.meshes = []
        for m in desc.meshes:
            h = giCreateMesh(self.scene, m.vertices, m.faces, id=m.object_id, isDoubleSided=m.double_sided,
                             isLeftHanded=m.left_handed, name=m.name, primvars=m.primvars)
            if h is None:
                raise RuntimeError("giCreateMesh failed: " + giLastError())
            giSetMeshTransform(h, m.transform)
            if m.instance_transforms is not None:
                giSetMeshInstanceTransforms(h, m.instance_transforms)
            if m.instance_ids is not None:
                giSetMeshInstanceIds(h, m.instance_ids)
            if m.instancer_primvars:
                giSetMeshInstancerPrimvars(h, m.instancer_primvars)
            giSetMeshMaterial(h, self.materials[m.material])
            self.meshes.append(h)
        self.lights = []
        for l in desc.sphere_lights:
            h = giCreateSphereLight(self.scene)
            giSetSphereLightPosition(h, l.position); giSetSphereLightBaseEmission(h, l.base_emission)
            giSetSphereLightRadius(h, *l.radius); giSetSphereLightDiffuseSpecular(h, l.diffuse, l.specular)
            self.lights.append(("sphere", h))
        for l in desc.distant_lights:
            h = giCreateDistantLight(self.scene)
            giSetDistantLightDirection(h, l.direction); giSetDistantLightBaseEmission(h, l.base_emission)
            giSetDistantLightAngle(h, l.angle); giSetDistantLightDiffuseSpecular(h, l.diffuse, l.specular)
            self.lights.append(("distant", h))
        for l in desc.rect_lights:
            h = giCreateRectLight(self.scene)
            giSetRectLightOrigin(h, l.origin); giSetRectLightTangents(h, l.t0, l.t1); giSetRectLightBaseEmission(h, l.base_emission)
            giSetRectLightDimensions(h, l.width, l.height); giSetRectLightDiffuseSpecular(h, l.diffuse, l.specular)
            self.lights.append(("rect", h))
        for l in desc.disk_lights:
            h = giCreateDiskLight(self.scene)
            giSetDiskLightOrigin(h, l.origin); giSetDiskLightTangents(h, l.t0, l.t1); giSetDiskLightBaseEmission(h, l.base_emission)
            giSetDiskLightRadius(h, *l.radius); giSetDiskLightDiffuseSpecular(h, l.diffuse, l.specular)
            self.lights.append(("disk", h))
        self.dome = None
        if desc.dome_light is not None:
            self.dome = giCreateDomeLight(self.scene, desc.dome_light.texels)
            giSetDomeLightRotation(self.dome, desc.dome_light.rotation)
            giSetDomeLightBaseEmission(self.dome, desc.dome_light.base_emission)
        self.color = None

    def render(self, spp=None, settings=None) -> np.ndarray:
        """One giRender call; returns the colour AOV (float32[h, w, 4], row 0 = bottom).  Successive calls
        accumulate (sampleOffset) exactly like the reference."""
        d = self.desc
        if self.color is None:
            self.color = giCreateRenderBuffer(d.width, d.height, GiRenderBufferFormat.Float32Vec4)
            if self.color is None:
                raise RuntimeError("giCreateRenderBuffer failed: " + giLastError())
        st = settings or d.settings
        if spp is not None:
            import dataclasses
            st = dataclasses.replace(st, spp=spp)
        status = giRender(self.scene, d.camera, st, [(GiAovId.Color, d.clear_color, self.color)], self.dome)
        if status != GiStatus.Ok:
            raise RuntimeError("giRender failed: " + giLastError())
        return self.color.mem()

    def render_aovs(self, aov_ids, clear_values=None, spp=None, settings=None):
        """giRender with several AOV bindings.  Returns {GiAovId: ndarray} (Color: float32[h,w,4]; vec3 AOVs: float32[h,w,4];
        Depth: float32[h,w,1]; ObjectId/FaceId/InstanceId: int32[h,w,1]).  Buffers persist across calls (progressive refinement)."""
        d = self.desc
        if not hasattr(self, "aov_buffers"):
            self.aov_buffers = {}
        bindings = []
        for aov in aov_ids:
            aov = GiAovId(aov)
            fmt = (GiRenderBufferFormat.Int32 if aov in (GiAovId.ObjectId, GiAovId.FaceId, GiAovId.InstanceId)
                   else GiRenderBufferFormat.Float32 if aov == GiAovId.Depth else GiRenderBufferFormat.Float32Vec4)
            if aov not in self.aov_buffers:
                rb = giCreateRenderBuffer(d.width, d.height, fmt)
                if rb is None:
                    raise RuntimeError("giCreateRenderBuffer failed: " + giLastError())
                self.aov_buffers[aov] = rb
            clear = (clear_values or {}).get(aov, d.clear_color if aov == GiAovId.Color else (0.0, 0.0, 0.0, 0.0))
            if fmt == GiRenderBufferFormat.Int32 and not isinstance(clear, (bytes, bytearray)):
                clear = np.asarray(clear, np.int32).tobytes()
            bindings.append((aov, clear, self.aov_buffers[aov]))
        st = settings or d.settings
        if spp is not None:
            import dataclasses
            st = dataclasses.replace(st, spp=spp)
        status = giRender(self.scene, d.camera, st, bindings, self.dome)
        if status != GiStatus.Ok:
            raise RuntimeError("giRender failed: " + giLastError())
        return {aov: self.aov_buffers[aov].mem().copy() for aov, _, _ in bindings}

    def close(self):
        if self.scene is None:
            return
        if self.color is not None:
            giDestroyRenderBuffer(self.color)
        for rb in getattr(self, "aov_buffers", {}).values():
            giDestroyRenderBuffer(rb)
        for h in self.meshes:
            giDestroyMesh(h)
        for kind, h in self.lights:
            {"sphere": giDestroySphereLight, "distant": giDestroyDistantLight, "rect": giDestroyRectLight, "disk": giDestroyDiskLight}[kind](self.scene, h)
        if self.dome:
            giDestroyDomeLight(self.dome)
        for h in self.materials:
            giDestroyMaterial(h)
        for h in getattr(self, "textures", []):
            giDestroyTexture(h)
        giDestroyScene(self.scene)
        self.scene = None
