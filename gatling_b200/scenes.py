"""Synthetic scene descriptions for the configs of BASELINE.json.

This is the USD-free stand-in for what `hdGatling` hands to the `gi` API: meshes arrive already
triangulated and de-indexed the way /root/reference/src/hdGatling/mesh.cpp:837-1138 does it
(fan triangulation, faceVarying normals force de-indexing, Duff-basis fallback tangents when no
texcoords are authored, mesh.cpp:236-268), cameras like renderPass.cpp:191-228, render settings
with the defaults of renderDelegate.cpp:93-110.

Everything is plain numpy; both the product binding (gatling_b200.gi) and the test oracle consume
the same `SceneDesc`.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional

import numpy as np

# ---- closure ids / parameter slots (include/gtlc.h) -------------------------------------------
CLOSURE_DIFFUSE = 0
CLOSURE_USD_PREVIEW_SURFACE = 1
CLOSURE_OPENPBR = 2
MATERIAL_PARAM_COUNT = 48

UPS = dict(diffuse=0, emissive=3, useSpecularWorkflow=6, specular=7, metallic=10, roughness=11,
           clearcoat=12, clearcoatRoughness=13, opacity=14, ior=15)
OPBR = dict(base_weight=0, base_color=1, base_diffuse_roughness=4, base_metalness=5, specular_weight=6,
            specular_color=7, specular_roughness=10, specular_ior=11, specular_roughness_anisotropy=12,
            transmission_weight=13, transmission_color=14, transmission_depth=17, coat_weight=18,
            coat_color=19, coat_roughness=22, coat_ior=23, coat_darkening=24, fuzz_weight=25,
            fuzz_color=26, fuzz_roughness=29, emission_luminance=30, emission_color=31,
            thin_film_weight=34, thin_film_thickness=35, thin_film_ior=36, geometry_opacity=37,
            geometry_thin_walled=38, subsurface_weight=39, subsurface_color=40, transmission_scatter=43,
            transmission_scatter_anisotropy=46)

VERTEX_DTYPE = np.dtype([("pos", "<f4", 3), ("u", "<f4"), ("norm", "<f4", 3), ("v", "<f4"),
                         ("tangent", "<f4", 3), ("bitangentSign", "<f4")])  # GiVertex, Gi.h:109-117
assert VERTEX_DTYPE.itemsize == 48


@dataclasses.dataclass
class Texture:
    """Decoded image; texels[row, col, channel] with row 0 = bottom (v = 0).  uint8 -> RGBA8 unorm, float16 -> RGBA16F,
    float32 with 1 channel -> R32F, float32 with 4 channels -> RGBA32F.  depth > 1: a 3-D texture, texels[slice, row, col, channel]."""
    texels: np.ndarray
    depth: int = 1

    @property
    def format(self) -> int:
        t = self.texels
        if t.dtype == np.uint8:
            return 0
        if t.dtype == np.float16:
            return 1
        return 2 if (t.ndim == 2 or t.shape[-1] == 1) else 3

    @property
    def rows_cols(self):
        t = self.texels
        return (t.shape[1], t.shape[2]) if self.depth > 1 else (t.shape[0], t.shape[1])


WRAP_CLAMP, WRAP_REPEAT, WRAP_MIRRORED_REPEAT, WRAP_CLIP = 0, 1, 2, 3
LOOKUP_FILTERED_2D, LOOKUP_TEXEL_2D, LOOKUP_FILTERED_3D, LOOKUP_TEXEL_3D = 0, 1, 2, 3    # GtlcTextureLookup
BIND_NORMAL = -1
BIND_SCRATCH, BIND_SCRATCH_COUNT = MATERIAL_PARAM_COUNT, 4                               # scratch floats behind the parameter block
COMBINE_SET, COMBINE_MULTIPLY, COMBINE_ADD, COMBINE_MIX, COMBINE_CONSTANT = 0, 1, 2, 3, 0x10   # GtlcCombineOp


def combine(op, factor_param=0, constant=False) -> int:
    """GTLC_COMBINE(op, factorParam) [| GTLC_COMBINE_CONSTANT]"""
    return int(op) | (int(factor_param) << 8) | (COMBINE_CONSTANT if constant else 0)


@dataclasses.dataclass
class TextureBinding:   # GtlcTextureBinding: UsdUVTexture (+ UsdTransform2d) feeding one material input
    param: int                      # first index of the parameter block, or BIND_NORMAL
    texture: int                    # index into SceneDesc.textures (-1 with a constant operand)
    channels: tuple = (0, 1, 2)     # one entry per written float
    wrap_u: int = WRAP_REPEAT
    wrap_v: int = WRAP_REPEAT
    srgb: bool = False
    fallback: tuple = (0.0, 0.0, 0.0, 1.0)
    scale: tuple = (1.0, 1.0, 1.0, 1.0)
    bias: tuple = (0.0, 0.0, 0.0, 0.0)
    uv_rotation: float = 0.0
    uv_scale: tuple = (1.0, 1.0)
    uv_translation: tuple = (0.0, 0.0)
    lookup: int = LOOKUP_FILTERED_2D
    wrap_w: int = WRAP_REPEAT
    xyz_scale: tuple = (1.0, 1.0, 1.0)      # 3-D lookups: coordinate = object-space position * scale + offset
    xyz_offset: tuple = (0.0, 0.0, 0.0)
    combine: int = 0                # combine(op, factor_param, constant): how the value meets what the input already holds (gtlc.h)


INTERP_CONSTANT, INTERP_INSTANCE, INTERP_UNIFORM, INTERP_VERTEX = 0, 1, 2, 3


@dataclasses.dataclass
class Primvar:           # GiPrimvarData: data float32[elements, 1..4] (Float..Vec4) or int32[elements, 1..4] (Int..Int4)
    name: str
    interpolation: int
    data: np.ndarray


@dataclasses.dataclass
class PrimvarBinding:    # GtlcPrimvarBinding: a material input read from a primvar of the hit mesh
    param: int
    count: int
    name: str
    fallback: tuple = (0.0, 0.0, 0.0, 0.0)


@dataclasses.dataclass
class Material:
    closure: int
    params: np.ndarray  # float32[48]
    name: str = ""
    textures: list = dataclasses.field(default_factory=list)   # TextureBinding
    primvars: list = dataclasses.field(default_factory=list)   # PrimvarBinding


def material_diffuse(tint, name="") -> Material:
    p = np.zeros(MATERIAL_PARAM_COUNT, np.float32)
    p[0:3] = tint
    return Material(CLOSURE_DIFFUSE, p, name)


def material_usd_preview_surface(name="", **kw) -> Material:
    """UsdPreviewSurface with spec defaults (diffuse .18, roughness .5, ior 1.5, ...)."""
    p = np.zeros(MATERIAL_PARAM_COUNT, np.float32)
    p[UPS["diffuse"]:UPS["diffuse"] + 3] = 0.18
    p[UPS["roughness"]] = 0.5
    p[UPS["clearcoatRoughness"]] = 0.01
    p[UPS["opacity"]] = 1.0
    p[UPS["ior"]] = 1.5
    for k, v in kw.items():
        i = UPS[k]
        if np.ndim(v) == 0:
            p[i] = v
        else:
            p[i:i + 3] = v
    return Material(CLOSURE_USD_PREVIEW_SURFACE, p, name)


def material_openpbr(name="", **kw) -> Material:
    """open_pbr_surface with the defaults of src/gi/mtlx/open_pbr_surface.mtlx:11-84."""
    p = np.zeros(MATERIAL_PARAM_COUNT, np.float32)
    d = dict(base_weight=1.0, base_color=(0.8, 0.8, 0.8), specular_weight=1.0, specular_color=(1, 1, 1),
             specular_roughness=0.3, specular_ior=1.5, transmission_color=(1, 1, 1), coat_color=(1, 1, 1),
             coat_ior=1.6, coat_darkening=1.0, fuzz_color=(1, 1, 1), fuzz_roughness=0.5,
             emission_color=(1, 1, 1), thin_film_thickness=0.5, thin_film_ior=1.4, geometry_opacity=1.0,
             subsurface_color=(0.8, 0.8, 0.8))
    d.update(kw)
    for k, v in d.items():
        i = OPBR[k]
        if np.ndim(v) == 0:
            p[i] = v
        else:
            p[i:i + 3] = v
    return Material(CLOSURE_OPENPBR, p, name)


@dataclasses.dataclass
class Mesh:
    vertices: np.ndarray            # VERTEX_DTYPE[n]
    faces: np.ndarray               # uint32[m, 3]
    material: int                   # index into SceneDesc.materials
    transform: np.ndarray = dataclasses.field(default_factory=lambda: np.eye(4, dtype=np.float32))  # USD row-major, row-vector convention
    instance_transforms: Optional[np.ndarray] = None   # float32[k, 4, 4]; None = one identity instance
    double_sided: bool = True
    left_handed: bool = False
    object_id: int = 0
    name: str = ""
    primvars: list = dataclasses.field(default_factory=list)             # Primvar
    instancer_primvars: list = dataclasses.field(default_factory=list)   # Primvar (giSetMeshInstancerPrimvars)
    instance_ids: Optional[np.ndarray] = None                            # giSetMeshInstanceIds; None = 0..n-1


@dataclasses.dataclass
class Camera:  # GiCameraDesc, Gi.h:95-107
    position: tuple
    forward: tuple
    up: tuple
    vfov: float
    fStop: float = 0.0
    focusDistance: float = 0.0
    focalLength: float = 0.05
    clipStart: float = 0.1
    clipEnd: float = 100.0
    exposure: float = 0.0


@dataclasses.dataclass
class RenderSettings:  # GiRenderSettings, Gi.h:139-159; defaults renderDelegate.cpp:93-110
    clippingPlanes: bool = False
    depthOfField: bool = False
    domeLightCameraVisible: bool = True
    filterImportanceSampling: bool = True
    frame: float = 0.0
    jitteredSampling: bool = True
    lightIntensityMultiplier: float = 1.0
    maxBounces: int = 13
    maxSampleValue: float = 10.0
    maxVolumeWalkLength: int = 7
    mediumStackSize: int = 0
    metersPerSceneUnit: float = 1.0
    nextEventEstimation: bool = False
    progressiveAccumulation: bool = True
    rrBounceOffset: int = 3
    rrInvMinTermProb: float = 0.95
    spp: int = 1
    time: float = 0.0


@dataclasses.dataclass
class SphereLight:
    position: tuple
    base_emission: tuple
    radius: tuple = (0.5, 0.5, 0.5)
    diffuse: float = 1.0
    specular: float = 1.0


@dataclasses.dataclass
class DistantLight:
    direction: tuple
    base_emission: tuple
    angle: float = 0.0
    diffuse: float = 1.0
    specular: float = 1.0


@dataclasses.dataclass
class RectLight:
    origin: tuple
    t0: tuple
    t1: tuple
    base_emission: tuple
    width: float = 1.0
    height: float = 1.0
    diffuse: float = 1.0
    specular: float = 1.0


@dataclasses.dataclass
class DiskLight:
    origin: tuple
    t0: tuple
    t1: tuple
    base_emission: tuple
    radius: tuple = (0.5, 0.5)
    diffuse: float = 1.0
    specular: float = 1.0


@dataclasses.dataclass
class DomeLight:
    texels: np.ndarray  # float32[h, w, 4] lat-long
    rotation: tuple = (0.0, 0.0, 0.0, 1.0)
    base_emission: tuple = (1.0, 1.0, 1.0)


@dataclasses.dataclass
class SceneDesc:
    name: str
    meshes: List[Mesh]
    materials: List[Material]
    camera: Camera
    settings: RenderSettings
    width: int
    height: int
    clear_color: tuple = (1.0, 1.0, 1.0, 1.0)   # renderDelegate.cpp:227-230 (test harness background)
    sphere_lights: List[SphereLight] = dataclasses.field(default_factory=list)
    distant_lights: List[DistantLight] = dataclasses.field(default_factory=list)
    rect_lights: List[RectLight] = dataclasses.field(default_factory=list)
    disk_lights: List[DiskLight] = dataclasses.field(default_factory=list)
    dome_light: Optional[DomeLight] = None
    textures: List[Texture] = dataclasses.field(default_factory=list)

    @property
    def triangle_count(self) -> int:
        n = 0
        for m in self.meshes:
            k = 1 if m.instance_transforms is None else len(m.instance_transforms)
            n += len(m.faces) * k
        return n


# ---- reproducible random stream: the reference's own hash_pcg32 (common.glsl:93-98) -------------
def pcg32_stream(seed: int, count: int) -> np.ndarray:
    """count outputs of hash_pcg32 applied to an LCG state starting at `seed` (vectorised)."""
    a = np.uint32(747796405)
    c = np.uint32(2891336453)
    with np.errstate(over="ignore"):
        pw = np.empty(count + 1, np.uint32)
        pw[0] = 1
        pw[1:] = a
        pw = np.cumprod(pw, dtype=np.uint32)            # a^i mod 2^32
        geo = np.cumsum(pw[:-1], dtype=np.uint32)       # sum_{k<i+1} a^k
        state = pw[1:] * np.uint32(seed) + c * geo      # state after i+1 steps
        word = ((state >> ((state >> np.uint32(28)) + np.uint32(4))) ^ state) * np.uint32(277803737)
        return (word >> np.uint32(22)) ^ word


def pcg32_uniform(seed: int, count: int) -> np.ndarray:
    """floats in [0,1) via uintAsFloat (common.glsl:44-47)."""
    w = pcg32_stream(seed, count)
    return ((w >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)


# ---- vertex baking helpers (hdGatling/mesh.cpp conventions) ------------------------------------
def duff_basis(n: np.ndarray):
    """_DuffOrthonormalBasis, mesh.cpp:225-233 (float32 arithmetic)."""
    n = n.astype(np.float32)
    sign = np.where(n[:, 2] >= 0, np.float32(1), np.float32(-1))
    a = np.float32(-1.0) / (sign + n[:, 2])
    b = n[:, 0] * n[:, 1] * a
    t = np.stack([np.float32(1) + sign * n[:, 0] * n[:, 0] * a, sign * b, -sign * n[:, 0]], 1)
    bt = np.stack([b, sign + n[:, 1] * n[:, 1] * a, -n[:, 1]], 1)
    return t.astype(np.float32), bt.astype(np.float32)


def bake_vertices(points: np.ndarray, normals: np.ndarray, uvs: Optional[np.ndarray] = None) -> np.ndarray:
    """_BakeMeshGeometry (mesh.cpp:271-330) for the no-authored-tangent, no-texcoord case:
    fallback tangents from the Duff basis, bitangentSign = sign(dot(cross(t, b), n)) (mesh.cpp:83-86)."""
    n = len(points)
    t, b = duff_basis(normals)
    sign = np.where(np.einsum("ij,ij->i", np.cross(t, b), normals) > 0, 1.0, -1.0).astype(np.float32)
    v = np.zeros(n, VERTEX_DTYPE)
    v["pos"] = points
    v["norm"] = normals
    v["tangent"] = t
    v["bitangentSign"] = sign
    if uvs is not None:
        v["u"] = uvs[:, 0]
        v["v"] = uvs[:, 1]
    return v


def polygon_mesh(points, face_vertex_counts, face_vertex_indices, facevarying_normals):
    """Fan-triangulate (HdMeshUtil::ComputeTriangleIndices) and de-index a polygon mesh with
    faceVarying normals (mesh.cpp:916-920,1011-1017)."""
    points = np.asarray(points, np.float32)
    normals = np.asarray(facevarying_normals, np.float32)
    tri_pts, tri_nrm = [], []
    off = 0
    for cnt in face_vertex_counts:
        for i in range(1, cnt - 1):
            for corner in (0, i, i + 1):
                tri_pts.append(points[face_vertex_indices[off + corner]])
                tri_nrm.append(normals[off + corner])
        off += cnt
    tri_pts = np.asarray(tri_pts, np.float32)
    tri_nrm = np.asarray(tri_nrm, np.float32)
    faces = np.arange(len(tri_pts), dtype=np.uint32).reshape(-1, 3)
    return bake_vertices(tri_pts, tri_nrm), faces


# ---- C1: cornell box --------------------------------------------------------------------------
# Geometry tables restate /root/reference/cornell.usda:1-247 (8 meshes, 46 triangles, Z-up).
_BOX_POINTS = [(-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1), (1, -1, 1), (1, 1, -1), (1, 1, 1)]
_BOX_COUNTS = [4] * 6
_BOX_INDICES = [0, 1, 3, 2, 2, 3, 7, 6, 6, 7, 5, 4, 4, 5, 1, 0, 2, 6, 4, 0, 7, 3, 1, 5]
_BOX_NORMALS = [(-1, -0.0, 0)] * 4 + [(0, 1, 0)] * 4 + [(1, -0.0, 0)] * 4 + [(0, -1, 0)] * 4 + [(0, 0, -1)] * 4 + [(0, -0.0, 1)] * 4
_F = np.float32


def _cornell_tables():
    light_pts = [(-0.5, -0.5, 0.98), (-0.5, -0.5, 1), (-0.5, 0.5, 0.98), (-0.5, 0.5, 1),
                 (0.5, -0.5, 0.98), (0.5, -0.5, 1), (0.5, 0.5, 0.98), (0.5, 0.5, 1)]
    quad = ([4], [0, 1, 3, 2])
    eye = np.eye(4, dtype=np.float64)
    box1 = np.array([(0.28844988346099854, 0.19823384284973145, 0, 0), (-0.19823384284973145, 0.28844988346099854, 0, 0),
                     (0, 0, 0.699999988079071, 0), (-0.3499999940395355, 0.3499999940395355, -0.30000001192092896, 1)])
    box2 = np.array([(0.2831651270389557, -0.09908341616392136, 0, 0), (0.09908341616392136, 0.2831651270389557, 0, 0),
                     (0, 0, 0.3501630127429962, 0), (0.45550230145454407, -0.41113391518592834, -0.6499999761581421, 1)])
    # name, material, xform, points, counts, indices, normals
    return [
        ("Light", "Light", eye, light_pts, _BOX_COUNTS, _BOX_INDICES, _BOX_NORMALS),
        ("BottomPlane", "White", eye, [(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (1, 1, -1)], *quad, [(0, 0, 1)] * 4),
        ("TopPlane", "White", eye, [(1, -1, 0.99999994), (-1, -1, 1.0000001), (1, 1, 0.99999994), (-1, 1, 1.0000001)], *quad,
         [(-8.940697e-8, 0, -1)] * 4),
        ("BackPlane", "White", eye, [(-1, 1, -1), (1, 1, -1), (-1, 0.99999994, 1), (1, 0.99999994, 1)], *quad,
         [(0, -1, -2.9802322e-8)] * 4),
        ("LeftPlane", "Red", eye, [(-0.99999994, -1, 1), (-1, -1, -1), (-0.99999994, 1, 1), (-1, 1, -1)], *quad,
         [(1, 0, -2.9802322e-8)] * 4),
        ("RightPlane", "Green", eye, [(1, -1, -1), (1, -1, 1), (1, 1, -1), (1, 1, 1)], *quad, [(-1, -0.0, 0)] * 4),
        ("Box1", "White", box1, _BOX_POINTS, _BOX_COUNTS, _BOX_INDICES, _BOX_NORMALS),
        ("Box2", "White", box2, _BOX_POINTS, _BOX_COUNTS, _BOX_INDICES, _BOX_NORMALS),
    ]


def cornell_camera() -> Camera:
    # cornell.usda:13 camera xform; renderPass.cpp:191-228: forward = dir(0,0,-1), up = dir(0,1,0)
    m = np.array([(1, 0, 0, 0), (0, -4.371138828673793e-8, 1, 0), (0, -1, -4.371138828673793e-8, 0), (0, -7, 0, 1)], np.float64)
    pos = (np.array([0, 0, 0, 1.0]) @ m)[:3]
    fwd = (np.array([0, 0, -1.0, 0]) @ m)[:3]
    up = (np.array([0, 1.0, 0, 0]) @ m)[:3]
    fwd /= np.linalg.norm(fwd)
    up /= np.linalg.norm(up)
    aperture = np.float32(20.25) * np.float32(0.1)    # GfCamera::APERTURE_UNIT
    focal = np.float32(50.0) * np.float32(0.1)        # GfCamera::FOCAL_LENGTH_UNIT
    vfov = np.float32(2.0) * np.float32(math.atan(float(aperture / (np.float32(2.0) * focal))))
    return Camera(position=tuple(float(_F(x)) for x in pos), forward=tuple(float(_F(x)) for x in fwd),
                  up=tuple(float(_F(x)) for x in up), vfov=float(vfov), focalLength=float(focal),
                  clipStart=0.1, clipEnd=100.0)


def cornell(width=512, height=512, spp=64, max_bounces=4) -> SceneDesc:
    """Config C1: cornell.usda 512x512 spp 64 max-bounces 4, UsdPreviewSurface, NEE off, background (1,1,1)."""
    mats = {
        "Light": material_usd_preview_surface("Light", diffuse=(0.8, 0.8, 0.8), emissive=(8.5, 6, 4)),
        "White": material_usd_preview_surface("White", diffuse=(0.8, 0.8, 0.8)),
        "Red": material_usd_preview_surface("Red", diffuse=(1, 0, 0)),
        "Green": material_usd_preview_surface("Green", diffuse=(0, 1, 0)),
    }
    names = list(mats)
    meshes = []
    for i, (name, mat, xf, pts, counts, idx, nrm) in enumerate(_cornell_tables()):
        v, f = polygon_mesh(pts, counts, idx, nrm)
        meshes.append(Mesh(v, f, names.index(mat), np.asarray(xf, np.float32), None, True, False, i + 1, name))
    st = RenderSettings(spp=spp, maxBounces=max_bounces)
    return SceneDesc("cornell", meshes, [mats[n] for n in names], cornell_camera(), st, width, height)


# ---- textured Cornell box (R15 test scene) ----------------------------------------------------
def procedural_texture(kind: str, size: int = 64, seed: int = 7) -> Texture:
    """Deterministic test images (the reference reads files through imgio; none ship with it)."""
    y, x = np.mgrid[0:size, 0:size]
    r = pcg32_uniform(seed, size * size * 4).reshape(size, size, 4)
    if kind == "checker":                       # RGBA8, sRGB-encoded colours
        c = ((x // (size // 8) + y // (size // 8)) & 1).astype(np.float32)
        img = np.stack([0.9 * c + 0.1, 0.25 + 0.5 * (1 - c), 0.2 + 0.6 * r[..., 2], np.ones_like(c)], -1)
        return Texture((img * 255 + 0.5).astype(np.uint8))
    if kind == "roughness":                     # R32F scalar map
        return Texture((0.05 + 0.9 * (0.5 + 0.5 * np.sin(x * 0.35) * np.cos(y * 0.23))).astype(np.float32)[..., None])
    if kind == "normal":                        # RGBA8 tangent-space normals, encoded n * 0.5 + 0.5
        nx = 0.45 * np.sin(x * 2 * np.pi * 3 / size); ny = 0.45 * np.cos(y * 2 * np.pi * 2 / size)
        nz = np.sqrt(np.maximum(0.0, 1 - nx * nx - ny * ny))
        img = np.stack([nx * 0.5 + 0.5, ny * 0.5 + 0.5, nz * 0.5 + 0.5, np.ones_like(nx)], -1)
        return Texture((img * 255 + 0.5).astype(np.uint8))
    if kind == "emission":                      # RGBA16F HDR blobs
        d = np.exp(-(((x - size / 2) ** 2 + (y - size / 2) ** 2) / (0.08 * size * size)))
        return Texture(np.stack([6.0 * d, 2.5 * d + 0.2 * r[..., 1], 0.5 * d, np.ones_like(d)], -1).astype(np.float16))
    if kind == "noise":                         # RGBA32F
        return Texture(r.astype(np.float32))
    raise ValueError(kind)


def assign_planar_uv(mesh: Mesh) -> None:
    """u, v = the two object-space axes along which the mesh is largest, mapped to [-0.25, 1.25] (so wrap modes matter)."""
    p = mesh.vertices["pos"]
    lo, hi = p.min(0), p.max(0)
    ax = np.argsort(-(hi - lo))[:2]
    ax.sort()
    ext = np.maximum(hi - lo, 1e-6)
    v = mesh.vertices.copy()
    v["u"] = ((p[:, ax[0]] - lo[ax[0]]) / ext[ax[0]] * 1.5 - 0.25).astype(np.float32)
    v["v"] = ((p[:, ax[1]] - lo[ax[1]]) / ext[ax[1]] * 1.5 - 0.25).astype(np.float32)
    mesh.vertices = v


def textured_cornell(width=96, height=96, spp=8, max_bounces=5, nee=True) -> SceneDesc:
    """Cornell box whose surfaces exercise every texture feature: sRGB RGBA8 + UsdTransform2d + repeat (floor), scalar R32F
    roughness + metallic (back wall), mirrored repeat RGBA32F (left wall), clip + fallback (right wall), tangent-space normal
    map (tall box, OpenPBR), HDR RGBA16F emission with clamp (short box)."""
    d = cornell(width, height, spp, max_bounces)
    for m in d.meshes:
        assign_planar_uv(m)
    d.textures = [procedural_texture(k) for k in ("checker", "roughness", "normal", "emission", "noise")]
    T = TextureBinding
    mats = list(d.materials)
    floor = material_usd_preview_surface("FloorTex", diffuse=(0.5, 0.5, 0.5))
    floor.textures = [T(UPS["diffuse"], 0, (0, 1, 2), WRAP_REPEAT, WRAP_REPEAT, True, uv_rotation=30.0, uv_scale=(3.0, 2.0), uv_translation=(0.13, -0.4))]
    back = material_usd_preview_surface("BackTex", diffuse=(0.7, 0.7, 0.75), metallic=0.0)
    back.textures = [T(UPS["roughness"], 1, (0,), WRAP_CLAMP, WRAP_REPEAT), T(UPS["metallic"], 4, (2,), scale=(1, 1, 0.9, 1), bias=(0, 0, 0.05, 0))]
    left = material_usd_preview_surface("LeftTex", diffuse=(1, 0, 0))
    left.textures = [T(UPS["diffuse"], 4, (1, 0, 2), WRAP_MIRRORED_REPEAT, WRAP_MIRRORED_REPEAT, uv_scale=(2.5, 2.5))]
    right = material_usd_preview_surface("RightTex", diffuse=(0, 1, 0))
    right.textures = [T(UPS["diffuse"], 0, (0, 1, 2), WRAP_CLIP, WRAP_CLIP, True, fallback=(0.05, 0.6, 0.1, 1.0))]
    tall = material_openpbr("TallNormalMapped", base_color=(0.8, 0.8, 0.85), specular_roughness=0.25, base_metalness=0.6)
    tall.textures = [T(BIND_NORMAL, 2, (0, 1, 2), scale=(2, 2, 2, 1), bias=(-1, -1, -1, 0), uv_scale=(2.0, 3.0)),
                     T(OPBR["specular_roughness"], 1, (0,), scale=(0.5, 1, 1, 1), bias=(0.1, 0, 0, 0))]
    short = material_usd_preview_surface("ShortGlow", diffuse=(0.3, 0.3, 0.3))
    short.textures = [T(UPS["emissive"], 3, (0, 1, 2), WRAP_CLAMP, WRAP_CLAMP)]
    base = len(mats)
    mats += [floor, back, left, right, tall, short]
    assign = {"bottomplane": 0, "backplane": 1, "leftplane": 2, "rightplane": 3, "box1": 4, "box2": 5}
    for i, m in enumerate(d.meshes):
        lname = m.name.lower()
        for key, k in assign.items():
            if key in lname:
                d.meshes[i] = dataclasses.replace(m, material=base + k)
    d.materials = mats
    d.name = "textured_cornell"
    if nee:
        d.settings = dataclasses.replace(d.settings, nextEventEstimation=True)
        d.rect_lights.append(RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
    return d


# ---- primvar-driven inputs (R15 scene data test scene) ------------------------------------------
def primvar_cornell(width=96, height=96, spp=8, max_bounces=5) -> SceneDesc:
    """Cornell box shaded by one UsdPreviewSurface whose diffuseColor reads the primvar `displayColor` and whose
    roughness reads `rough`: per-vertex colours (floor), per-face colours (back wall), per-instance colours through
    instancer primvars and shuffled instance ids (short box, 4 instances), a mesh primvar overriding the instancer's
    (tall box), constant primvars, a too-short array and a missing primvar (fallbacks)."""
    d = cornell(width, height, spp, max_bounces)
    pv = material_usd_preview_surface("Primvars", diffuse=(0.5, 0.5, 0.5), roughness=0.5)
    pv.primvars = [PrimvarBinding(UPS["diffuse"], 3, "displayColor", (0.1, 0.2, 0.9, 0.0)), PrimvarBinding(UPS["roughness"], 1, "rough", (0.9, 0.0, 0.0, 0.0))]
    d.materials = list(d.materials) + [pv]
    mat = len(d.materials) - 1
    rnd = lambda seed, n, c: pcg32_uniform(seed, n * c).reshape(n, c).astype(np.float32)
    for i, m in enumerate(d.meshes):
        if m.name == "Light":
            continue
        m = dataclasses.replace(m, material=mat)
        if m.name == "BottomPlane":
            m.primvars = [Primvar("displayColor", INTERP_VERTEX, rnd(11, len(m.vertices), 3)), Primvar("rough", INTERP_CONSTANT, np.array([[0.15]], np.float32))]
        elif m.name == "BackPlane":
            m.primvars = [Primvar("displayColor", INTERP_UNIFORM, rnd(12, len(m.faces), 3)), Primvar("rough", INTERP_UNIFORM, 0.2 + 0.6 * rnd(13, len(m.faces), 1))]
        elif m.name == "LeftPlane":
            m.primvars = [Primvar("displayColor", INTERP_VERTEX, rnd(14, 2, 3))]           # too short for the vertex count: ignored -> fallback
        elif m.name == "Box2":
            # four small copies, well separated and of different heights: overlapping or coplanar instances would z-fight,
            # and which copy wins then depends on the last ulp of the ray (libm vs CUDA sin/cos)
            mt = np.asarray(m.transform, np.float32)
            wp = m.vertices["pos"] @ mt[:3, :3] + mt[3, :3]
            lo, hi = wp.min(0), wp.max(0)
            pivot = np.array([(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, lo[2]], np.float64)
            xf = []
            for k, (tx, ty) in enumerate(((-0.6, -0.55), (0.05, -0.7), (0.6, -0.2), (0.55, 0.45))):
                sc = 0.35 + 0.05 * k
                t = np.eye(4, dtype=np.float64)
                t[0, 0] = t[1, 1] = t[2, 2] = sc
                t[3, :3] = np.array([tx, ty, pivot[2]]) - pivot * sc          # p' = (p - pivot) * sc + target
                xf.append(t.astype(np.float32))
            m.instance_transforms = np.stack(xf)
            m.instance_ids = np.array([3, 1, 0, 2], np.int32)
            m.instancer_primvars = [Primvar("displayColor", INTERP_INSTANCE, rnd(15, 4, 3)), Primvar("rough", INTERP_INSTANCE, 0.25 + 0.6 * rnd(16, 4, 1))]   # (well away from mirror-like: a near-delta GGX lobe amplifies last-ulp differences between CPU and GPU)
        elif m.name == "Box1":
            m.instancer_primvars = [Primvar("displayColor", INTERP_CONSTANT, np.array([[1.0, 0.0, 1.0]], np.float32))]
            m.primvars = [Primvar("displayColor", INTERP_VERTEX, rnd(17, len(m.vertices), 4))]   # vec4 primvar read as colour; overrides the instancer's
        d.meshes[i] = m            # TopPlane / RightPlane: no primvars at all -> fallbacks
    d.settings = dataclasses.replace(d.settings, nextEventEstimation=True)
    d.rect_lights.append(RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
    d.name = "primvar_cornell"
    return d


# ---- the other MDL runtime lookups + per-hit cutouts (R13 / R15 test scene) ----------------------------------
def procedural_volume(size=8, seed=21, channels=4) -> Texture:
    """Smooth RGBA32F (or R32F) 3-D texture, texels[slice, row, col, channel]."""
    z, y, x = np.mgrid[0:size, 0:size, 0:size].astype(np.float32) / np.float32(size)
    r = pcg32_uniform(seed, size ** 3 * 4).reshape(size, size, size, 4).astype(np.float32)
    vol = np.stack([0.5 + 0.5 * np.sin(6.0 * x + 2.0 * z), 0.5 + 0.5 * np.cos(5.0 * y - 3.0 * x), 0.2 + 0.6 * r[..., 2], np.ones_like(x)], -1).astype(np.float32)
    return Texture(np.ascontiguousarray(vol[..., :channels]), depth=size)


def lookup_cornell(width=96, height=96, spp=8, max_bounces=5) -> SceneDesc:
    """Cornell box that exercises what textured_cornell / primvar_cornell leave out: a cutout whose opacity is a TEXTURE
    (alpha mask, filtered; rp_main.ahit:35-61 on both closest and shadow rays) on the tall box, a TEXEL-fetched blocky mask
    on the short box together with a base colour from a filtered 3-D texture of the object-space position, a texel-fetched
    3-D voxel colour on the back wall, a 3-D lookup with CLIP along w and its fallback on the left wall, and INTEGER primvars
    (scene_data_lookup_int / _int3: nearest vertex, no interpolation) on the floor."""
    d = cornell(width, height, spp, max_bounces)
    for m in d.meshes:
        assign_planar_uv(m)
    size = 32
    y, x = np.mgrid[0:size, 0:size]
    holes = (((x // 4 + y // 4) % 3) != 0).astype(np.float32)                                        # 1/3 of the 4x4 blocks are holes
    soft = np.clip(0.5 + 0.5 * np.sin(x * 0.9) * np.sin(y * 0.7), 0.0, 1.0).astype(np.float32)       # a soft-edged alpha in [0, 1]
    mask_rgba = np.stack([np.full_like(soft, 0.8), np.full_like(soft, 0.7), np.full_like(soft, 0.3), soft], -1)
    d.textures = [Texture((mask_rgba * 255 + 0.5).astype(np.uint8)),            # 0: RGBA8, opacity in alpha
                  Texture(holes[::4, ::4][..., None].copy()),                    # 1: 8x8 R32F blocky mask (texel fetch)
                  procedural_volume(8, 21),                                      # 2: 8^3 RGBA32F
                  procedural_volume(4, 22)]                                      # 3: 4^3 RGBA32F (voxels)
    T = TextureBinding
    tall = material_usd_preview_surface("TallAlphaMask", diffuse=(0.2, 0.5, 0.9), opacity=1.0)
    tall.textures = [T(UPS["opacity"], 0, (3,), WRAP_REPEAT, WRAP_REPEAT, uv_scale=(2.0, 2.0), uv_rotation=20.0),
                     T(UPS["diffuse"], 0, (0, 1, 2), WRAP_REPEAT, WRAP_REPEAT, True, uv_scale=(2.0, 2.0), uv_rotation=20.0)]
    short = material_openpbr("ShortBlockyMask", base_color=(0.8, 0.8, 0.8), specular_roughness=0.5)
    short.textures = [T(OPBR["geometry_opacity"], 1, (0,), WRAP_REPEAT, WRAP_REPEAT, lookup=LOOKUP_TEXEL_2D, uv_scale=(0.6, 0.6), uv_translation=(0.2, 0.2)),
                      T(OPBR["base_color"], 2, (0, 1, 2), WRAP_REPEAT, WRAP_MIRRORED_REPEAT, lookup=LOOKUP_FILTERED_3D, wrap_w=WRAP_CLAMP,
                        xyz_scale=(0.9, 1.1, 0.7), xyz_offset=(0.3, 0.1, 0.25))]
    back = material_usd_preview_surface("BackVoxels", diffuse=(0.5, 0.5, 0.5))
    back.textures = [T(UPS["diffuse"], 3, (2, 0, 1), lookup=LOOKUP_TEXEL_3D, xyz_scale=(0.45, 0.45, 0.45), xyz_offset=(0.5, 0.5, 0.5))]     # outside the volume: (0, 0, 0)
    left = material_usd_preview_surface("LeftClipW", diffuse=(1, 0, 0))
    left.textures = [T(UPS["diffuse"], 2, (0, 1, 2), WRAP_REPEAT, WRAP_REPEAT, lookup=LOOKUP_FILTERED_3D, wrap_w=WRAP_CLIP, fallback=(0.9, 0.1, 0.6, 1.0),
                       xyz_scale=(0.7, 0.7, 0.9), xyz_offset=(0.5, 0.5, 0.35))]
    floor = material_usd_preview_surface("FloorIntPrimvars", diffuse=(0.5, 0.5, 0.5), roughness=0.6)
    floor.primvars = [PrimvarBinding(UPS["diffuse"], 3, "colorBits", (0.1, 0.2, 0.9, 0.0)), PrimvarBinding(UPS["metallic"], 1, "metalFlag", (0.0, 0.0, 0.0, 0.0)),
                      PrimvarBinding(UPS["clearcoat"], 1, "missingInt", (0.25, 0.0, 0.0, 0.0))]
    mats = list(d.materials)
    base = len(mats)
    mats += [tall, short, back, left, floor]
    assign = {"box1": 0, "box2": 1, "backplane": 2, "leftplane": 3, "bottomplane": 4}
    for i, m in enumerate(d.meshes):
        lname = m.name.lower()
        for key, k in assign.items():
            if key in lname:
                m = dataclasses.replace(m, material=base + k)
                if key in ("box1", "box2"):                  # see-through boxes must not z-fight with the floor they stand on
                    xf = np.asarray(m.transform, np.float32).copy(); xf[3, 2] += 0.03
                    m.transform = xf
                if key == "bottomplane":
                    nv = len(m.vertices)
                    m.primvars = [Primvar("colorBits", INTERP_VERTEX, (np.arange(nv * 3).reshape(nv, 3) % 2).astype(np.int32)),      # int3 per vertex: nearest vertex wins
                                  Primvar("metalFlag", INTERP_UNIFORM, (np.arange(len(m.faces)) % 2).astype(np.int32)[:, None])]     # int per face
                d.meshes[i] = m
    d.materials = mats
    d.name = "lookup_cornell"
    d.settings = dataclasses.replace(d.settings, nextEventEstimation=True)
    d.rect_lights.append(RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
    return d


def layered_cornell(width=96, height=96, spp=8, max_bounces=5, baked=False) -> SceneDesc:
    """Cornell box whose inputs are built from SEVERAL lookups by the combine word of the bindings (gtlc.h GtlcCombineOp): image x image
    (floor), mix(image, constant, mask image) through a scratch float (back wall), image + image (left wall), mix(image, image, mask)
    plus a product roughness on an OpenPBR box, and a constant SET followed by a multiply (right wall).  All lookups are texel fetches
    of float images, so the same scene can be expressed with ONE pre-combined image per input: `baked=True` builds that form (same
    float32 operations in numpy), and the two must render bit-identically."""
    d = cornell(width, height, spp, max_bounces)
    for m in d.meshes:
        assign_planar_uv(m)
    size = 16
    f32 = np.float32
    A = pcg32_uniform(11, size * size * 4).reshape(size, size, 4).astype(f32)
    B = pcg32_uniform(12, size * size * 4).reshape(size, size, 4).astype(f32)
    y, x = np.mgrid[0:size, 0:size]
    M = (0.5 + 0.5 * np.sin(x * 0.8) * np.cos(y * 0.6)).astype(f32)                        # soft mask in [0, 1]
    K = np.array([0.9, 0.15, 0.1], f32)
    tint = np.array([0.6, 0.8, 1.0], f32)
    one = f32(1.0)
    T = TextureBinding
    tx = dict(lookup=LOOKUP_TEXEL_2D, wrap_u=WRAP_REPEAT, wrap_v=WRAP_REPEAT, uv_scale=(0.7, 0.7), uv_translation=(0.15, 0.15))
    S0, S1 = BIND_SCRATCH, BIND_SCRATCH + 1
    Mc = M[..., None]
    if not baked:
        d.textures = [Texture(A), Texture(B), Texture(M[..., None].copy())]
        floor_t = [T(UPS["diffuse"], 0, (0, 1, 2), **tx), T(UPS["diffuse"], 1, (0, 1, 2), combine=combine(COMBINE_MULTIPLY), **tx)]
        back_t = [T(S0, 2, (0,), **tx), T(UPS["diffuse"], 0, (2, 1, 0), **tx),
                  T(UPS["diffuse"], -1, (0, 1, 2), bias=(float(K[0]), float(K[1]), float(K[2]), 0.0), combine=combine(COMBINE_MIX, S0, constant=True))]
        left_t = [T(UPS["diffuse"], 0, (0, 1, 2), scale=(0.5, 0.5, 0.5, 1.0), **tx), T(UPS["diffuse"], 1, (1, 2, 0), scale=(0.5, 0.5, 0.5, 1.0), combine=combine(COMBINE_ADD), **tx)]
        tall_t = [T(S1, 2, (0,), **tx), T(OPBR["base_color"], 0, (0, 1, 2), **tx), T(OPBR["base_color"], 1, (0, 1, 2), combine=combine(COMBINE_MIX, S1), **tx),
                  T(OPBR["specular_roughness"], 0, (3,), **tx), T(OPBR["specular_roughness"], 1, (3,), combine=combine(COMBINE_MULTIPLY), **tx)]
        right_t = [T(UPS["diffuse"], -1, (0, 1, 2), bias=(float(tint[0]), float(tint[1]), float(tint[2]), 0.0), combine=combine(COMBINE_SET, constant=True)),
                   T(UPS["diffuse"], 1, (0, 0, 0), combine=combine(COMBINE_MULTIPLY), **tx)]
    else:
        rgb = lambda img: np.concatenate([img, np.ones((size, size, 1), f32)], -1).astype(f32)
        floor_i = A[..., :3] * B[..., :3]
        back_i = A[..., [2, 1, 0]] * (one - Mc) + K[None, None, :] * Mc
        left_i = A[..., :3] * f32(0.5) + B[..., [1, 2, 0]] * f32(0.5)
        tall_i = A[..., :3] * (one - Mc) + B[..., :3] * Mc
        rough_i = (A[..., 3] * B[..., 3])[..., None]
        right_i = tint[None, None, :] * B[..., [0, 0, 0]]
        d.textures = [Texture(rgb(floor_i)), Texture(rgb(back_i)), Texture(rgb(left_i)), Texture(rgb(tall_i)), Texture(rough_i.astype(f32).copy()), Texture(rgb(right_i))]
        floor_t = [T(UPS["diffuse"], 0, (0, 1, 2), **tx)]
        back_t = [T(UPS["diffuse"], 1, (0, 1, 2), **tx)]
        left_t = [T(UPS["diffuse"], 2, (0, 1, 2), **tx)]
        tall_t = [T(OPBR["base_color"], 3, (0, 1, 2), **tx), T(OPBR["specular_roughness"], 4, (0,), **tx)]
        right_t = [T(UPS["diffuse"], 5, (0, 1, 2), **tx)]
    floor = material_usd_preview_surface("FloorProduct", diffuse=(0.5, 0.5, 0.5)); floor.textures = floor_t
    back = material_usd_preview_surface("BackMixConstant", diffuse=(0.5, 0.5, 0.5)); back.textures = back_t
    left = material_usd_preview_surface("LeftSum", diffuse=(0.5, 0.5, 0.5)); left.textures = left_t
    tall = material_openpbr("TallMixImages", base_color=(0.8, 0.8, 0.8), specular_roughness=0.4, base_metalness=0.3); tall.textures = tall_t
    right = material_usd_preview_surface("RightConstantTimesImage", diffuse=(0.5, 0.5, 0.5)); right.textures = right_t
    mats = list(d.materials)
    base = len(mats)
    mats += [floor, back, left, tall, right]
    assign = {"bottomplane": 0, "backplane": 1, "leftplane": 2, "box1": 3, "rightplane": 4}
    for i, m in enumerate(d.meshes):
        lname = m.name.lower()
        for key, k in assign.items():
            if key in lname:
                d.meshes[i] = dataclasses.replace(m, material=base + k)
    d.materials = mats
    d.name = "layered_cornell" + ("_baked" if baked else "")
    d.settings = dataclasses.replace(d.settings, nextEventEstimation=True)
    d.rect_lights.append(RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
    return d


# ---- participating media (R14 test scene) ------------------------------------------------------
def volume_cornell(width=96, height=96, spp=8, max_bounces=12, medium_stack_size=2, nee=True) -> SceneDesc:
    """Cornell box with three closed OpenPBR dielectrics: a scattering 'milk glass' sphere with a second, differently
    coloured medium nested inside it (stack depth 2), and a purely absorbing sphere.  medium_stack_size = 0 renders the
    same scene with the reference's Beer-only inside/outside toggle."""
    d = cornell(width, height, spp, max_bounces)
    base = len(d.materials)
    d.materials = list(d.materials) + [
        material_openpbr("Milk", base_color=(0.9, 0.9, 0.9), specular_roughness=0.05, transmission_weight=1.0, transmission_color=(0.85, 0.6, 0.4),
                         transmission_depth=0.25, transmission_scatter=(0.55, 0.5, 0.45), transmission_scatter_anisotropy=0.35),
        material_openpbr("Core", base_color=(0.9, 0.9, 0.9), specular_roughness=0.1, transmission_weight=1.0, transmission_color=(0.3, 0.7, 0.9),
                         transmission_depth=0.1, transmission_scatter=(0.2, 0.3, 0.6), transmission_scatter_anisotropy=-0.2),
        material_openpbr("Amber", base_color=(0.9, 0.9, 0.9), specular_roughness=0.02, transmission_weight=1.0, transmission_color=(0.9, 0.45, 0.1),
                         transmission_depth=0.3)]
    for k, (centre, radius, mat) in enumerate((((0.38, -0.35, -0.6), 0.38, 0), ((0.38, -0.35, -0.6), 0.18, 1), ((-0.45, 0.15, -0.25), 0.28, 2))):
        v, f = sphere_mesh(24, 13, radius)
        d.meshes.append(Mesh(v, f, base + mat, _trs(1.0, 0.3 * k, centre), None, False, False, 20 + k, "sphere%d" % k))
    d.meshes = [m for m in d.meshes if m.name not in ("Box1", "Box2")]
    d.settings = dataclasses.replace(d.settings, mediumStackSize=medium_stack_size, maxVolumeWalkLength=7, nextEventEstimation=nee, rrBounceOffset=4)
    if nee:
        d.rect_lights.append(RectLight(origin=(0.0, 0.0, 0.9), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(10, 10, 10), width=0.5, height=0.5))
    d.name = "volume_cornell"
    return d


# ---- C2: 1M random triangles ------------------------------------------------------------------
def random_triangles(n_tris=1_000_000, width=1920, height=1080, spp=256, max_bounces=8, seed=0x2545F491) -> SceneDesc:
    """Config C2: centre ~ U[-1,1]^3, edges ~ U[-0.02,0.02]^3, flat normals, non-indexed,
    one mesh / one identity instance, Lambert 0.8, camera (0,-4,0) -> +Y, up +Z, vfov 40 deg."""
    r = pcg32_uniform(seed, n_tris * 9).reshape(n_tris, 9)
    c = r[:, 0:3] * np.float32(2) - np.float32(1)
    e1 = (r[:, 3:6] * np.float32(2) - np.float32(1)) * np.float32(0.02)
    e2 = (r[:, 6:9] * np.float32(2) - np.float32(1)) * np.float32(0.02)
    pts = np.empty((n_tris, 3, 3), np.float32)
    pts[:, 0] = c
    pts[:, 1] = c + e1
    pts[:, 2] = c + e2
    nrm = np.cross(e1, e2).astype(np.float32)
    ln = np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    nrm = np.where(ln > 0, nrm / np.maximum(ln, np.float32(1e-30)), np.float32([0, 0, 1])).astype(np.float32)
    normals = np.repeat(nrm, 3, axis=0)
    v = bake_vertices(pts.reshape(-1, 3), normals)
    f = np.arange(n_tris * 3, dtype=np.uint32).reshape(-1, 3)
    mesh = Mesh(v, f, 0, name="soup", object_id=1)
    cam = Camera(position=(0.0, -4.0, 0.0), forward=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0), vfov=float(np.float32(0.6981317)))
    st = RenderSettings(spp=spp, maxBounces=max_bounces)
    return SceneDesc("random_tris_%d" % n_tris, [mesh], [material_diffuse((0.8, 0.8, 0.8), "lambert")], cam, st, width, height)


# ---- parametric meshes (indexed, smooth normals, fallback tangents) --------------------------------
def _grid_faces(nu, nv, wrap_u=True, wrap_v=False):
    """Two triangles per (u, v) cell of an (nu x nv) vertex grid (wrapping adds the seam cell)."""
    cu = nu if wrap_u else nu - 1
    cv = nv if wrap_v else nv - 1
    i, j = np.meshgrid(np.arange(cu), np.arange(cv), indexing="ij")
    i2, j2 = (i + 1) % nu, (j + 1) % nv
    a, b, c, d = i * nv + j, i2 * nv + j, i2 * nv + j2, i * nv + j2
    f = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], -2).reshape(-1, 3)
    return f.astype(np.uint32)


def sphere_mesh(nu, nv, radius=1.0, displacement=0.0, seed=1):
    """UV sphere with 2*nu*(nv-1) triangles (incl. degenerate pole slivers), optional radial displacement."""
    u = (np.arange(nu, dtype=np.float32) / np.float32(nu)) * np.float32(2 * np.pi)
    v = (np.arange(nv, dtype=np.float32) / np.float32(nv - 1)) * np.float32(np.pi)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    n = np.stack([np.sin(vv) * np.cos(uu), np.sin(vv) * np.sin(uu), np.cos(vv)], -1).reshape(-1, 3).astype(np.float32)
    r = np.full(len(n), radius, np.float32)
    if displacement:
        r = r * (np.float32(1) + np.float32(displacement) * (np.sin(np.float32(7 + seed % 5) * uu) * np.sin(np.float32(5 + seed % 3) * vv)).reshape(-1).astype(np.float32))
    # v runs from the +z pole downwards, so (du x dv) points inwards: reverse the winding to face outwards
    return bake_vertices(n * r[:, None], n), np.ascontiguousarray(_grid_faces(nu, nv, True, False)[:, ::-1])


def torus_mesh(nu, nv, R=1.0, r=0.35):
    u = (np.arange(nu, dtype=np.float32) / np.float32(nu)) * np.float32(2 * np.pi)
    v = (np.arange(nv, dtype=np.float32) / np.float32(nv)) * np.float32(2 * np.pi)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    n = np.stack([np.cos(vv) * np.cos(uu), np.cos(vv) * np.sin(uu), np.sin(vv)], -1).reshape(-1, 3).astype(np.float32)
    c = np.stack([np.cos(uu), np.sin(uu), np.zeros_like(uu)], -1).reshape(-1, 3).astype(np.float32) * np.float32(R)
    return bake_vertices(c + n * np.float32(r), n), _grid_faces(nu, nv, True, True)


def cylinder_mesh(nu, nv, radius=0.5, height=1.0):
    """Open-ended tube (side only), 2*nu*(nv-1) triangles, z in [0, height]."""
    u = (np.arange(nu, dtype=np.float32) / np.float32(nu)) * np.float32(2 * np.pi)
    z = (np.arange(nv, dtype=np.float32) / np.float32(nv - 1)) * np.float32(height)
    uu, zz = np.meshgrid(u, z, indexing="ij")
    n = np.stack([np.cos(uu), np.sin(uu), np.zeros_like(uu)], -1).reshape(-1, 3).astype(np.float32)
    p = n * np.float32(radius)
    p[:, 2] = zz.reshape(-1)
    return bake_vertices(p, n), _grid_faces(nu, nv, True, False)


def box_mesh(n, size=(1.0, 1.0, 1.0), inward=False):
    """Axis-aligned box centred at the origin, every face an n x n grid (12 n^2 triangles), flat normals."""
    verts, faces, base = [], [], 0
    s = np.asarray(size, np.float32) * np.float32(0.5)
    lin = (np.arange(n + 1, dtype=np.float32) / np.float32(n)) * np.float32(2) - np.float32(1)
    a, b = np.meshgrid(lin, lin, indexing="ij")
    for axis in range(3):
        for sign in (-1.0, 1.0):
            p = np.zeros(((n + 1) * (n + 1), 3), np.float32)
            p[:, axis] = sign
            p[:, (axis + 1) % 3] = a.reshape(-1)
            p[:, (axis + 2) % 3] = b.reshape(-1)
            nrm = np.zeros_like(p)
            nrm[:, axis] = -sign if inward else sign
            f = _grid_faces(n + 1, n + 1, False, False)
            if (sign < 0) != inward:
                f = f[:, ::-1]                       # keep the winding consistent with the outward/inward normal
            verts.append(bake_vertices(p * s, nrm))
            faces.append(f + np.uint32(base))
            base += len(p)
    return np.concatenate(verts), np.ascontiguousarray(np.concatenate(faces))


def _trs(scale, angle, pos):
    """USD row-vector 4x4: scale, rotate about Z, translate."""
    c, s = math.cos(angle), math.sin(angle)
    m = np.eye(4, dtype=np.float64)
    m[0, :3] = (c * scale, s * scale, 0)
    m[1, :3] = (-s * scale, c * scale, 0)
    m[2, 2] = scale
    m[3, :3] = pos
    return m.astype(np.float32)


_PALETTE = [(0.8, 0.8, 0.8), (0.75, 0.2, 0.15), (0.2, 0.55, 0.25), (0.15, 0.3, 0.7), (0.85, 0.65, 0.2), (0.5, 0.25, 0.6), (0.9, 0.5, 0.3), (0.3, 0.3, 0.3)]


def openpbr_material_table(count, seed=0x9E3779B9):
    """Fixed tables from SURVEY.md 8d: base colours x {rough dielectric, coated glossy dielectric, metal}, every 12th set glass."""
    r = pcg32_uniform(seed, count * 4).reshape(count, 4)
    mats = []
    for i in range(count):
        col = _PALETTE[i % len(_PALETTE)]
        rough = float(np.float32(0.05) + r[i, 0] * np.float32(0.55))
        kind = i % 3
        if i % 12 == 11:
            mats.append(material_openpbr("glass%d" % i, base_color=col, specular_roughness=min(rough, 0.15), transmission_weight=1.0,
                                         transmission_color=(0.9, 0.95, 1.0), transmission_depth=0.5 if i % 24 == 23 else 0.0))
        elif kind == 0:
            mats.append(material_openpbr("diel%d" % i, base_color=col, specular_roughness=rough))
        elif kind == 1:
            mats.append(material_openpbr("coat%d" % i, base_color=col, specular_roughness=rough, coat_weight=1.0, coat_roughness=0.05 + 0.2 * float(r[i, 1]),
                                         coat_color=(1.0, 0.95, 0.9)))
        else:
            mats.append(material_openpbr("metal%d" % i, base_color=col, base_metalness=1.0, specular_roughness=rough))
    return mats


def kitchen_standin(tris_per_instance=12500, instances=400, meshes=120, materials=24, width=1920, height=1080, spp=1024, max_bounces=8,
                    seed=0x9E3779B9) -> SceneDesc:
    """Configs C3/C4 stand-in (the Evermotion kitchen is not redistributable): 8x6x3 m room + instanced boxes / tubes / tori,
    OpenPBR parameter sets, 2 rect lights at the ceiling + 1 distant light, NEE on."""
    n_box = max(1, int(round(math.sqrt(tris_per_instance / 12.0))))
    n_ring = max(3, int(round(math.sqrt(tris_per_instance / 2.0))))
    protos = [box_mesh(n_box, (1.0, 1.0, 1.0)), cylinder_mesh(n_ring, n_ring + 1, 0.5, 1.0), torus_mesh(n_ring, n_ring, 0.6, 0.25)]
    mats = openpbr_material_table(materials, seed)
    room_v, room_f = box_mesh(1, (8.0, 6.0, 3.0), inward=True)
    out = [Mesh(room_v, room_f, 0, _trs(1.0, 0.0, (0.0, 0.0, 1.5)), None, True, False, 1, "room")]
    r = pcg32_uniform(seed ^ 0x5bd1e995, instances * 5).reshape(instances, 5)
    per_mesh = [[] for _ in range(meshes)]
    for k in range(instances):
        scale = 0.2 + 0.3 * float(r[k, 0])
        pos = (-3.6 + 7.2 * float(r[k, 1]), -2.6 + 5.2 * float(r[k, 2]), 0.02 + 1.6 * float(r[k, 3]) * float(r[k, 3]))
        per_mesh[k % meshes].append(_trs(scale, 6.2831853 * float(r[k, 4]), pos))
    for mi in range(meshes):
        if not per_mesh[mi]:
            continue
        v, f = protos[mi % 3]
        out.append(Mesh(v, f, 1 + (mi % (materials - 1)) if materials > 1 else 0, np.eye(4, dtype=np.float32), np.stack(per_mesh[mi]), True, False, 2 + mi, "prop%d" % mi))
    cam = Camera(position=(0.0, -2.8, 1.6), forward=(0.0, 1.0, -0.12), up=(0.0, 0.12, 1.0), vfov=float(np.float32(1.0)))
    st = RenderSettings(spp=spp, maxBounces=max_bounces, nextEventEstimation=True)
    d = SceneDesc("kitchen_standin", out, mats, cam, st, width, height)
    d.rect_lights.append(RectLight(origin=(-2.0, 0.0, 2.95), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(30, 30, 30), width=1.0, height=1.0))
    d.rect_lights.append(RectLight(origin=(2.0, 0.5, 2.95), t0=(1, 0, 0), t1=(0, 1, 0), base_emission=(30, 28, 25), width=1.0, height=1.0))
    d.distant_lights.append(DistantLight(direction=(0.3, 0.5, -0.8), base_emission=(2.0, 2.0, 2.2), angle=0.02))
    return d


def material_stress(tris_per_mesh=31250, meshes=64, width=1920, height=1080, spp=512, max_bounces=8, seed=0x9E3779B9) -> SceneDesc:
    """Config C5 stand-in (GPUOpen matlib is not reachable): an 8x8 wall of displaced spheres, one distinct OpenPBR parameter
    set per mesh (diffuse / coated / metal / glass), constant white dome light + one distant light, NEE on."""
    side = int(math.ceil(math.sqrt(meshes)))
    nu = max(4, int(round(math.sqrt(tris_per_mesh / 2.0))))
    mats = openpbr_material_table(meshes, seed)
    perm = np.argsort(pcg32_stream(seed, meshes), kind="stable")     # neighbouring spheres get unrelated parameter sets
    out = []
    for k in range(meshes):
        v, f = sphere_mesh(nu, nu + 1, 1.0, 0.04, seed=k)
        gx, gz = k % side, k // side
        pos = ((gx - (side - 1) / 2.0) * 2.1, 0.0, (gz - (side - 1) / 2.0) * 2.1)
        out.append(Mesh(v, f, int(perm[k]), _trs(1.0, 0.37 * k, pos), None, True, False, 1 + k, "ball%d" % k))
    dist = side * 2.1 * 0.5 / math.tan(0.35) + 2.0
    cam = Camera(position=(0.0, -dist, 0.0), forward=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0), vfov=float(np.float32(0.7)))
    st = RenderSettings(spp=spp, maxBounces=max_bounces, nextEventEstimation=True)
    d = SceneDesc("material_stress", out, mats, cam, st, width, height)
    d.dome_light = DomeLight(np.ones((2, 4, 4), np.float32))
    d.distant_lights.append(DistantLight(direction=(0.4, 0.7, -0.6), base_emission=(3.0, 3.0, 3.0), angle=0.05))
    return d
