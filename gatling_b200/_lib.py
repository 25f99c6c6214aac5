"""ctypes loader of the C-ABI library (gatling_b200/libgtlc.so).  There is no CPU fallback: if the
library is missing the import of the binding fails loudly, and gtlc_initialize() fails without a
CUDA device."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTLC_LIB_PATH") or os.path.join(HERE, "libgtlc.so")   # the override is for tuning experiments (tools/)
HEADER = os.path.join(HERE, "..", "include", "gtlc.h")

H = C.c_uint64


class PrimvarDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", C.c_int32), ("interpolation", C.c_int32), ("data", C.c_void_p), ("dataSize", C.c_uint64)]


class PrimvarBinding(C.Structure):
    _fields_ = [("paramIndex", C.c_int32), ("componentCount", C.c_int32), ("name", C.c_char_p), ("fallback", C.c_float * 4)]


class TextureDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_int32), ("depth", C.c_uint32), ("texels", C.c_void_p)]


class TextureBinding(C.Structure):
    _fields_ = [("paramIndex", C.c_int32), ("componentCount", C.c_int32), ("channels", C.c_int32 * 3), ("wrapU", C.c_int32),
                ("wrapV", C.c_int32), ("srgb", C.c_int32), ("texture", C.c_uint64), ("fallback", C.c_float * 4),
                ("scale", C.c_float * 4), ("bias", C.c_float * 4), ("uvRotation", C.c_float), ("uvScale", C.c_float * 2),
                ("uvTranslation", C.c_float * 2), ("lookup", C.c_int32), ("wrapW", C.c_int32), ("xyzScale", C.c_float * 3),
                ("xyzOffset", C.c_float * 3), ("combine", C.c_int32)]


class MaterialDesc(C.Structure):
    _fields_ = [("closure", C.c_int32), ("name", C.c_char_p), ("params", C.c_float * 48),
                ("textures", C.POINTER(TextureBinding)), ("textureCount", C.c_uint32), ("primvarCount", C.c_uint32),
                ("primvars", C.POINTER(PrimvarBinding))]


class MeshDesc(C.Structure):
    _fields_ = [("faceCount", C.c_uint32), ("faces", C.c_void_p), ("faceIds", C.c_void_p), ("id", C.c_int32),
                ("isDoubleSided", C.c_int32), ("isLeftHanded", C.c_int32), ("name", C.c_char_p), ("maxFaceId", C.c_uint32),
                ("vertexCount", C.c_uint32), ("vertices", C.c_void_p), ("primvars", C.POINTER(PrimvarDesc)),
                ("primvarCount", C.c_uint32), ("reserved", C.c_uint32)]


class CameraDesc(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("forward", C.c_float * 3), ("up", C.c_float * 3), ("vfov", C.c_float),
                ("fStop", C.c_float), ("focusDistance", C.c_float), ("focalLength", C.c_float), ("clipStart", C.c_float),
                ("clipEnd", C.c_float), ("exposure", C.c_float)]


class RenderSettings(C.Structure):
    _fields_ = [("clippingPlanes", C.c_int32), ("depthOfField", C.c_int32), ("domeLightCameraVisible", C.c_int32),
                ("filterImportanceSampling", C.c_int32), ("frame", C.c_float), ("jitteredSampling", C.c_int32),
                ("lightIntensityMultiplier", C.c_float), ("maxBounces", C.c_uint32), ("maxSampleValue", C.c_float),
                ("maxVolumeWalkLength", C.c_uint32), ("mediumStackSize", C.c_uint32), ("metersPerSceneUnit", C.c_float),
                ("nextEventEstimation", C.c_int32), ("progressiveAccumulation", C.c_int32), ("rrBounceOffset", C.c_uint32),
                ("rrInvMinTermProb", C.c_float), ("spp", C.c_uint32), ("time", C.c_float)]


class AovBinding(C.Structure):
    _fields_ = [("aovId", C.c_int32), ("clearValue", C.c_uint8 * 16), ("renderBuffer", H)]


class RenderParams(C.Structure):
    _fields_ = [("aovBindingCount", C.c_uint32), ("aovBindings", C.POINTER(AovBinding)), ("camera", CameraDesc),
                ("domeLight", H), ("renderSettings", RenderSettings), ("scene", H)]


class InitParams(C.Structure):
    _fields_ = [("deviceIndex", C.c_int32), ("reserved", C.c_int32)]


class Tuning(C.Structure):
    _fields_ = [("maxPathsPerWave", C.c_uint32), ("sortByMaterial", C.c_int32), ("skipReadback", C.c_int32),
                ("countVisits", C.c_int32), ("traceVariant", C.c_int32), ("reserved", C.c_int32 * 3), ("bvhBuilder", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("closestRays", C.c_uint64), ("shadowRays", C.c_uint64), ("nodeVisits", C.c_uint64), ("triVisits", C.c_uint64),
                ("kernelLaunches", C.c_uint64), ("paths", C.c_uint64), ("msTotal", C.c_float), ("msTraceClosest", C.c_float),
                ("msTraceAny", C.c_float), ("msShade", C.c_float), ("msRaygen", C.c_float), ("msSort", C.c_float),
                ("msResolve", C.c_float), ("msBvhBuild", C.c_float), ("bvhNodeCount", C.c_uint32), ("bvhTriCount", C.c_uint32),
                ("bvhMaxDepth", C.c_uint32), ("maxNodeVisitsPerRay", C.c_uint32), ("nonFiniteRays", C.c_uint64),
                ("msBvhDeviceBuild", C.c_float), ("reservedStats", C.c_float), ("closestHits", C.c_uint64),
                ("shadowNodeVisits", C.c_uint64), ("shadowTriVisits", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class BvhView(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("nodeCount", C.c_uint32), ("triangles", C.c_void_p), ("triCount", C.c_uint32)]


_F3 = C.POINTER(C.c_float)
_SIGNATURES = {
    "gtlc_material_desc_defaults": (None, [C.c_int32, C.POINTER(MaterialDesc)]),
    "gtlc_initialize": (C.c_int, [C.POINTER(InitParams)]),
    "gtlc_terminate": (None, []),
    "gtlc_last_error": (C.c_char_p, []),
    "gtlc_error_count": (C.c_uint64, []),
    "gtlc_scene_create": (H, []),
    "gtlc_scene_destroy": (None, [H]),
    "gtlc_material_create": (H, [H, C.POINTER(MaterialDesc)]),
    "gtlc_material_destroy": (None, [H]),
    "gtlc_mesh_create": (H, [H, C.POINTER(MeshDesc)]),
    "gtlc_mesh_set_transform": (None, [H, C.c_void_p]),
    "gtlc_mesh_set_instance_transforms": (None, [H, C.c_uint32, C.c_void_p]),
    "gtlc_mesh_set_instancer_primvars": (None, [H, C.c_uint32, C.POINTER(PrimvarDesc)]),
    "gtlc_mesh_set_instance_ids": (None, [H, C.c_uint32, C.c_void_p]),
    "gtlc_mesh_set_material": (None, [H, H]),
    "gtlc_mesh_set_visibility": (None, [H, C.c_int32]),
    "gtlc_mesh_destroy": (None, [H]),
    "gtlc_sphere_light_create": (H, [H]), "gtlc_sphere_light_destroy": (None, [H, H]),
    "gtlc_sphere_light_set_position": (None, [H, _F3]), "gtlc_sphere_light_set_base_emission": (None, [H, _F3]),
    "gtlc_sphere_light_set_radius": (None, [H, C.c_float, C.c_float, C.c_float]),
    "gtlc_sphere_light_set_diffuse_specular": (None, [H, C.c_float, C.c_float]),
    "gtlc_distant_light_create": (H, [H]), "gtlc_distant_light_destroy": (None, [H, H]),
    "gtlc_distant_light_set_direction": (None, [H, _F3]), "gtlc_distant_light_set_base_emission": (None, [H, _F3]),
    "gtlc_distant_light_set_angle": (None, [H, C.c_float]),
    "gtlc_distant_light_set_diffuse_specular": (None, [H, C.c_float, C.c_float]),
    "gtlc_rect_light_create": (H, [H]), "gtlc_rect_light_destroy": (None, [H, H]),
    "gtlc_rect_light_set_origin": (None, [H, _F3]), "gtlc_rect_light_set_tangents": (None, [H, _F3, _F3]),
    "gtlc_rect_light_set_base_emission": (None, [H, _F3]), "gtlc_rect_light_set_dimensions": (None, [H, C.c_float, C.c_float]),
    "gtlc_rect_light_set_diffuse_specular": (None, [H, C.c_float, C.c_float]),
    "gtlc_disk_light_create": (H, [H]), "gtlc_disk_light_destroy": (None, [H, H]),
    "gtlc_disk_light_set_origin": (None, [H, _F3]), "gtlc_disk_light_set_tangents": (None, [H, _F3, _F3]),
    "gtlc_disk_light_set_base_emission": (None, [H, _F3]), "gtlc_disk_light_set_radius": (None, [H, C.c_float, C.c_float]),
    "gtlc_disk_light_set_diffuse_specular": (None, [H, C.c_float, C.c_float]),
    "gtlc_dome_light_create": (H, [H, C.c_uint32, C.c_uint32, C.c_void_p]), "gtlc_dome_light_destroy": (None, [H]),
    "gtlc_dome_light_set_rotation": (None, [H, _F3]), "gtlc_dome_light_set_base_emission": (None, [H, _F3]),
    "gtlc_dome_light_set_diffuse_specular": (None, [H, C.c_float, C.c_float]),
    "gtlc_texture_create": (H, [C.POINTER(TextureDesc)]), "gtlc_texture_destroy": (None, [H]),
    "gtlc_render_buffer_create": (H, [C.c_uint32, C.c_uint32, C.c_int32]), "gtlc_render_buffer_destroy": (None, [H]),
    "gtlc_render_buffer_mem": (C.c_void_p, [H]), "gtlc_render_buffer_device_mem": (C.c_void_p, [H]),
    "gtlc_render": (C.c_int, [C.POINTER(RenderParams)]),
    "gtlc_scene_set_shard": (None, [H, C.c_uint32, C.c_uint32, C.c_uint32]),
    "gtlc_comm_get_unique_id": (C.c_int, [C.c_void_p]), "gtlc_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "gtlc_comm_destroy": (None, []), "gtlc_comm_world_size": (C.c_int, []),
    "gtlc_set_tuning": (None, [C.POINTER(Tuning)]), "gtlc_get_tuning": (None, [C.POINTER(Tuning)]),
    "gtlc_get_stats": (None, [C.POINTER(Stats)]),
    "gtlc_trace_rays": (C.c_int, [H, C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p]),
    "gtlc_scene_get_bvh": (C.c_int, [H, C.POINTER(BvhView)]),
}

_lib = None


def declared_symbols():
    """Names of every GTLC_API function declared in include/gtlc.h (for the export test)."""
    import re
    with open(HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"GTLC_API[^;(]*?\b(gtlc_\w+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("gatling_b200/libgtlc.so is missing -- run `python -m gatling_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback for the product path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
