/*
 * gtlc.h -- C-ABI of the B200-native wavefront path-tracing core.
 *
 * This is the drop-in boundary: every entry point below is what a
 * binding of the reference's `gi` renderer API would call instead of the Vulkan-RT
 * implementation in src/gi/impl/Gi.cpp.  Plain pointers, sizes and opaque 64-bit handles
 * (modelled on the POD `struct{uint64_t handle}` handles of src/cgpu/gtl/cgpu/Cgpu.h:118-128);
 * no C++ types, no torch types, no exceptions.  The reference interface each function
 * replaces is cited as (Gi.h:LINE) = /root/reference/src/gi/gtl/gi/Gi.h.
 *
 * The C++ facade with the reference's exact signatures lives in include/gtl/gi/Gi.h and is
 * implemented purely on top of this header (gatling_b200/gi_cpp/Gi.cpp).
 */
#ifndef GTLC_H
#define GTLC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GTLC_API __declspec(dllexport)
#else
#define GTLC_API __attribute__((visibility("default")))
#endif

typedef uint64_t GtlcHandle; /* 0 == invalid (reference: nullptr from giCreate*) */

enum { GTLC_OK = 0, GTLC_ERROR = 1 }; /* GiStatus, Gi.h:32 */

/* GiAovId, Gi.h:34-54.  All AOVs are rendered except the NEE and ClockCycles debug AOVs, which are accepted and
 * left at their clear value.  Formats: Color and every vec3 AOV Float32Vec4, Depth Float32, ObjectId/FaceId/InstanceId Int32. */
enum {
  GTLC_AOV_COLOR = 0, GTLC_AOV_NORMAL, GTLC_AOV_NEE, GTLC_AOV_BARYCENTRICS, GTLC_AOV_TEXCOORDS,
  GTLC_AOV_BOUNCES, GTLC_AOV_CLOCK_CYCLES, GTLC_AOV_OPACITY, GTLC_AOV_TANGENTS, GTLC_AOV_BITANGENTS,
  GTLC_AOV_THIN_WALLED, GTLC_AOV_OBJECT_ID, GTLC_AOV_DEPTH, GTLC_AOV_FACE_ID, GTLC_AOV_INSTANCE_ID,
  GTLC_AOV_DOUBLE_SIDED, GTLC_AOV_ALBEDO, GTLC_AOV_COUNT
};

/* GiRenderBufferFormat, Gi.h:68-73 */
enum { GTLC_FORMAT_INT32 = 0, GTLC_FORMAT_FLOAT32 = 1, GTLC_FORMAT_FLOAT32_VEC4 = 2 };

/* GiVertex, Gi.h:109-117 (48 bytes) */
typedef struct GtlcVertex {
  float pos[3];
  float u;
  float norm[3];
  float v;
  float tangent[3];
  float bitangentSign;
} GtlcVertex;

/* GiFace, Gi.h:119-122 */
typedef struct GtlcFace {
  uint32_t v_i[3];
} GtlcFace;

/* GiPrimvarData, Gi.h:77-93 ("scene data" of the MDL runtime, mdl_interface.glsl:266-479) */
enum GtlcPrimvarType { GTLC_PRIMVAR_FLOAT = 0, GTLC_PRIMVAR_VEC2, GTLC_PRIMVAR_VEC3, GTLC_PRIMVAR_VEC4,
                       GTLC_PRIMVAR_INT, GTLC_PRIMVAR_INT2, GTLC_PRIMVAR_INT3, GTLC_PRIMVAR_INT4 };
enum GtlcPrimvarInterpolation { GTLC_INTERP_CONSTANT = 0, GTLC_INTERP_INSTANCE, GTLC_INTERP_UNIFORM, GTLC_INTERP_VERTEX };
typedef struct GtlcPrimvarDesc {
  const char* name;
  int32_t     type;            /* GtlcPrimvarType */
  int32_t     interpolation;   /* GtlcPrimvarInterpolation: constant = element 0, instance = element instanceId,
                                  uniform = one element per face, vertex = one per vertex (barycentric interpolation) */
  const void* data;
  uint64_t    dataSize;        /* bytes */
} GtlcPrimvarDesc;

/* GiMeshDesc, Gi.h:124-137 with the std::vector references flattened to pointer + count. */
typedef struct GtlcMeshDesc {
  uint32_t          faceCount;
  const GtlcFace*   faces;
  const int32_t*    faceIds;      /* may be NULL */
  int32_t           id;
  int32_t           isDoubleSided;
  int32_t           isLeftHanded;
  const char*       name;         /* may be NULL */
  uint32_t          maxFaceId;
  uint32_t          vertexCount;
  const GtlcVertex* vertices;
  const GtlcPrimvarDesc* primvars;   /* may be NULL */
  uint32_t          primvarCount;
  uint32_t          reserved;
} GtlcMeshDesc;

/* GiCameraDesc, Gi.h:95-107 */
typedef struct GtlcCameraDesc {
  float position[3];
  float forward[3];
  float up[3];
  float vfov;
  float fStop;
  float focusDistance;
  float focalLength;
  float clipStart;
  float clipEnd;
  float exposure;
} GtlcCameraDesc;

/* GiRenderSettings, Gi.h:139-159 (bool -> int32_t) */
typedef struct GtlcRenderSettings {
  int32_t  clippingPlanes;
  int32_t  depthOfField;
  int32_t  domeLightCameraVisible;
  int32_t  filterImportanceSampling;
  float    frame;
  int32_t  jitteredSampling;
  float    lightIntensityMultiplier;
  uint32_t maxBounces;
  float    maxSampleValue;
  uint32_t maxVolumeWalkLength;
  uint32_t mediumStackSize;
  float    metersPerSceneUnit;
  int32_t  nextEventEstimation;
  int32_t  progressiveAccumulation;
  uint32_t rrBounceOffset;
  float    rrInvMinTermProb;
  uint32_t spp;
  float    time;
} GtlcRenderSettings;

/* GiAovBinding, Gi.h:161-166 */
typedef struct GtlcAovBinding {
  int32_t    aovId;
  uint8_t    clearValue[16];
  GtlcHandle renderBuffer;
} GtlcAovBinding;

/* GiRenderParams, Gi.h:168-175 */
typedef struct GtlcRenderParams {
  uint32_t              aovBindingCount;
  const GtlcAovBinding* aovBindings;
  GtlcCameraDesc        camera;
  GtlcHandle            domeLight;   /* 0 = none */
  GtlcRenderSettings    renderSettings;
  GtlcHandle            scene;
} GtlcRenderParams;

/*
 * Materials.  The reference lowers MaterialX/MDL to per-material GLSL through the closed MDL SDK
 * (src/mc, SURVEY.md 8b "inner boundary"); here a material is a closure kind plus a parameter
 * block (class-compiled instead of instance-compiled).  The C++ facade lowers MaterialX XML
 * (constant inputs) to these blocks.
 */
enum {
  GTLC_CLOSURE_DIFFUSE = 0,             /* df::diffuse_reflection_bsdf(tint)           16 B  */
  GTLC_CLOSURE_USD_PREVIEW_SURFACE = 1, /* ND_UsdPreviewSurface_surfaceshader          64 B  */
  GTLC_CLOSURE_OPENPBR = 2              /* src/gi/mtlx/open_pbr_surface.mtlx:99-635   192 B  */
};

#define GTLC_MATERIAL_PARAM_COUNT 48

/* Indices into GtlcMaterialDesc.params for GTLC_CLOSURE_DIFFUSE */
enum { GTLC_DIFFUSE_TINT_R = 0, GTLC_DIFFUSE_TINT_G, GTLC_DIFFUSE_TINT_B, GTLC_DIFFUSE_PAD };

/* ... for GTLC_CLOSURE_USD_PREVIEW_SURFACE (UsdPreviewSurface spec input names) */
enum {
  GTLC_UPS_DIFFUSE_R = 0, GTLC_UPS_DIFFUSE_G, GTLC_UPS_DIFFUSE_B,
  GTLC_UPS_EMISSIVE_R, GTLC_UPS_EMISSIVE_G, GTLC_UPS_EMISSIVE_B,
  GTLC_UPS_USE_SPECULAR_WORKFLOW,
  GTLC_UPS_SPECULAR_R, GTLC_UPS_SPECULAR_G, GTLC_UPS_SPECULAR_B,
  GTLC_UPS_METALLIC, GTLC_UPS_ROUGHNESS, GTLC_UPS_CLEARCOAT, GTLC_UPS_CLEARCOAT_ROUGHNESS,
  GTLC_UPS_OPACITY, GTLC_UPS_IOR,
  GTLC_UPS_PARAM_COUNT /* 16 */
};

/* ... for GTLC_CLOSURE_OPENPBR (input names of open_pbr_surface.mtlx:4-97) */
enum {
  GTLC_OPBR_BASE_WEIGHT = 0,
  GTLC_OPBR_BASE_COLOR_R, GTLC_OPBR_BASE_COLOR_G, GTLC_OPBR_BASE_COLOR_B,
  GTLC_OPBR_BASE_DIFFUSE_ROUGHNESS, GTLC_OPBR_BASE_METALNESS,
  GTLC_OPBR_SPECULAR_WEIGHT,
  GTLC_OPBR_SPECULAR_COLOR_R, GTLC_OPBR_SPECULAR_COLOR_G, GTLC_OPBR_SPECULAR_COLOR_B,
  GTLC_OPBR_SPECULAR_ROUGHNESS, GTLC_OPBR_SPECULAR_IOR, GTLC_OPBR_SPECULAR_ROUGHNESS_ANISOTROPY,
  GTLC_OPBR_TRANSMISSION_WEIGHT,
  GTLC_OPBR_TRANSMISSION_COLOR_R, GTLC_OPBR_TRANSMISSION_COLOR_G, GTLC_OPBR_TRANSMISSION_COLOR_B,
  GTLC_OPBR_TRANSMISSION_DEPTH,
  GTLC_OPBR_COAT_WEIGHT,
  GTLC_OPBR_COAT_COLOR_R, GTLC_OPBR_COAT_COLOR_G, GTLC_OPBR_COAT_COLOR_B,
  GTLC_OPBR_COAT_ROUGHNESS, GTLC_OPBR_COAT_IOR, GTLC_OPBR_COAT_DARKENING,
  GTLC_OPBR_FUZZ_WEIGHT,
  GTLC_OPBR_FUZZ_COLOR_R, GTLC_OPBR_FUZZ_COLOR_G, GTLC_OPBR_FUZZ_COLOR_B,
  GTLC_OPBR_FUZZ_ROUGHNESS,
  GTLC_OPBR_EMISSION_LUMINANCE,
  GTLC_OPBR_EMISSION_COLOR_R, GTLC_OPBR_EMISSION_COLOR_G, GTLC_OPBR_EMISSION_COLOR_B,
  GTLC_OPBR_THIN_FILM_WEIGHT, GTLC_OPBR_THIN_FILM_THICKNESS, GTLC_OPBR_THIN_FILM_IOR,
  GTLC_OPBR_GEOMETRY_OPACITY, GTLC_OPBR_GEOMETRY_THIN_WALLED,
  GTLC_OPBR_SUBSURFACE_WEIGHT,
  GTLC_OPBR_SUBSURFACE_COLOR_R, GTLC_OPBR_SUBSURFACE_COLOR_G, GTLC_OPBR_SUBSURFACE_COLOR_B,
  GTLC_OPBR_TRANSMISSION_SCATTER_R, GTLC_OPBR_TRANSMISSION_SCATTER_G, GTLC_OPBR_TRANSMISSION_SCATTER_B,   /* volume scattering: open_pbr_surface.mtlx:238-245,286-299 */
  GTLC_OPBR_TRANSMISSION_SCATTER_ANISOTROPY,                                                           /* anisotropic_vdf anisotropy = MEDIUM_DIRECTIONAL_BIAS (Frontend.cpp:137-157) */
  GTLC_OPBR_PARAM_COUNT /* 47 <= 48 */
};

/* ---- textures: the reference binds images through MDL resource tables
 * (src/gi/impl/TextureManager.cpp:100-274) and samples them with tex_lookup_float4_2d
 * (src/gi/shaders/mdl_interface.glsl:127-145; sampler linear / Repeat, one mip, Gi.cpp:388-392).  Here a
 * material input is "textured" by a binding that overwrites 1..3 floats of the parameter block per hit.
 * Decoding image files is imgio's job in the reference (out of scope): the C-ABI takes texels. */
enum GtlcTextureFormat {
  GTLC_TEX_RGBA8_UNORM  = 0,   /* CgpuImageFormat::R8G8B8A8Unorm (LDR files) */
  GTLC_TEX_RGBA16_FLOAT = 1,   /* keepHdr images, 8 bytes per texel (TextureManager.cpp:135) */
  GTLC_TEX_R32_FLOAT    = 2,   /* MDL binary float payloads (TextureManager.cpp:196) */
  GTLC_TEX_RGBA32_FLOAT = 3
};
typedef struct GtlcTextureDesc {
  uint32_t    width, height;
  int32_t     format;          /* GtlcTextureFormat */
  uint32_t    depth;           /* 0 or 1: 2-D image; > 1: 3-D texture of `depth` slices (textures_3d[], mdl_interface.glsl:45-125) */
  const void* texels;          /* row 0 = bottom row (v = 0), as imgio delivers images (PngDecoder.cpp:78 flips); slices follow each other */
} GtlcTextureDesc;
/* No Gi.h counterpart: the reference creates images inside giCreateMaterialFrom* through its TextureManager
 * (TextureManager.cpp:100-274); here the caller decodes and hands texels over. */
GTLC_API GtlcHandle gtlc_texture_create(const GtlcTextureDesc* desc);
GTLC_API void       gtlc_texture_destroy(GtlcHandle texture);   /* deferred while a material still references it */

enum GtlcTextureWrap {           /* mdl_types.glsl:117-120 */
  GTLC_WRAP_CLAMP = 0, GTLC_WRAP_REPEAT = 1, GTLC_WRAP_MIRRORED_REPEAT = 2, GTLC_WRAP_CLIP = 3
};
/* Which MDL runtime lookup a binding performs (the reference's generated code picks one per call site):
 *   FILTERED_2D  tex_lookup_float4_2d (mdl_interface.glsl:127-145) at the transformed uv
 *   TEXEL_2D     tex_texel_float4_2d  (:167-185) at ivec2(floor(uv' * resolution)); (0,0,0,0) outside the image
 *   FILTERED_3D  tex_lookup_float4_3d (:45-64) at xyz = object-space hit position * xyzScale + xyzOffset
 *   TEXEL_3D     tex_texel_float4_3d  (:91-110) at ivec3(floor(xyz * resolution)); (0,0,0,0) outside the volume */
enum GtlcTextureLookup { GTLC_LOOKUP_FILTERED_2D = 0, GTLC_LOOKUP_TEXEL_2D = 1, GTLC_LOOKUP_FILTERED_3D = 2, GTLC_LOOKUP_TEXEL_3D = 3 };
#define GTLC_BIND_NORMAL (-1)    /* paramIndex: the value is a tangent-space normal (UsdPreviewSurface `normal`, OpenPBR
                                    `geometry_normal`): shading normal = normalize(T x + B y + N z), then mdl_adapt_normal
                                    (mdl_interface.glsl:238-256) */
typedef struct GtlcTextureBinding {
  int32_t    paramIndex;         /* first float of GtlcMaterialDesc.params that receives the value, or GTLC_BIND_NORMAL */
  int32_t    componentCount;     /* 1..3 floats written */
  int32_t    channels[3];        /* source channel per component: 0 r, 1 g, 2 b, 3 a */
  int32_t    wrapU, wrapV;       /* GtlcTextureWrap.  CLIP returns `fallback` outside [0,1] (UsdUVTexture wrap "black" /
                                    MaterialX addressmode "constant") */
  int32_t    srgb;               /* 1: the r, g, b channels are sRGB-encoded ("srgb_texture", MtlxDocumentPatcher.cpp:244-246) */
  GtlcHandle texture;
  float      fallback[4];        /* UsdUVTexture.fallback / MaterialX image `default` (rgba) */
  float      scale[4], bias[4];  /* UsdUVTexture: value = texel * scale + bias, per source channel */
  float      uvRotation;         /* UsdTransform2d: degrees, counter-clockwise */
  float      uvScale[2], uvTranslation[2];   /* uv' = rotate(uv * scale) + translation */
  int32_t    lookup;             /* GtlcTextureLookup (0 = filtered 2-D, the MaterialX image node) */
  int32_t    wrapW;              /* 3-D lookups: GtlcTextureWrap along the third axis */
  float      xyzScale[3], xyzOffset[3];      /* 3-D lookups: texture coordinate = object-space position * scale + offset */
  int32_t    combine;            /* 0 = the value overwrites the input; otherwise GTLC_COMBINE(op, factorParam) [| GTLC_COMBINE_CONSTANT] */
} GtlcTextureBinding;
/* Bindings are applied in order, so several of them can build one input from several lookups -- what a MaterialX graph with
 * `multiply` / `add` / `mix` nodes between images and a surface-shader input compiles to in the reference (mc generates MDL
 * for the whole graph).  With x = the input's current value (the material's constant, or what earlier bindings left) and
 * v = this binding's value:
 *   GTLC_COMBINE_SET       x = v
 *   GTLC_COMBINE_MULTIPLY  x = x * v
 *   GTLC_COMBINE_ADD       x = x + v
 *   GTLC_COMBINE_MIX       x = x * (1 - f) + v * f     (MDL math::lerp), f = the float at params[factorParam]
 * GTLC_COMBINE_CONSTANT: v = bias, no lookup (`texture` may be 0) -- a constant operand of mix.
 * paramIndex and factorParam may also name one of GTLC_BIND_SCRATCH_COUNT scratch floats behind the parameter block
 * (GTLC_BIND_SCRATCH + k; 0 at every hit before the first binding runs), e.g. a mask:
 *   { paramIndex = GTLC_BIND_SCRATCH, ... mask image, r }                                   scratch0 = mask
 *   { paramIndex = BASE_COLOR, ... image A }                                                base = A
 *   { paramIndex = BASE_COLOR, ... image B, combine = GTLC_COMBINE(GTLC_COMBINE_MIX, GTLC_BIND_SCRATCH) }   base = mix(A, B, mask)
 * The normal binding and the cutout input take GTLC_COMBINE_SET only (the any-hit test evaluates ONE lookup). */
enum GtlcCombineOp { GTLC_COMBINE_SET = 0, GTLC_COMBINE_MULTIPLY = 1, GTLC_COMBINE_ADD = 2, GTLC_COMBINE_MIX = 3 };
#define GTLC_COMBINE_CONSTANT 0x10
#define GTLC_COMBINE(op, factorParam) ((int32_t) (op) | ((int32_t) (factorParam) << 8))
#define GTLC_BIND_SCRATCH GTLC_MATERIAL_PARAM_COUNT
#define GTLC_BIND_SCRATCH_COUNT 4
/* A binding whose paramIndex is the closure's cutout input (GTLC_UPS_OPACITY / GTLC_OPBR_GEOMETRY_OPACITY) makes the
 * cutout per hit: the traversal kernels evaluate it in their any-hit test from the candidate hit's barycentrics, as
 * rp_main.ahit:35-61 evaluates mdl_cutout_opacity on a full shading state. */

/* A material input read from a primvar of the hit mesh (MaterialX geompropvalue / UsdPrimvarReader -> MDL
 * scene_data_lookup_float{,2,3}, mdl_interface.glsl:266-479): value = v0 bc.x + v1 bc.y + v2 bc.z over the elements the
 * primvar's interpolation selects; `fallback` when the mesh has no primvar of that name.  Integer primvars
 * (GTLC_PRIMVAR_INT..INT4) go through scene_data_lookup_int{,2,3} (:417-479): no interpolation, the element of the vertex
 * with the largest barycentric weight, converted to float as it enters the parameter block. */
#define GTLC_MAX_PRIMVAR_BINDINGS 8
typedef struct GtlcPrimvarBinding {
  int32_t     paramIndex;      /* first float of GtlcMaterialDesc.params that receives the value */
  int32_t     componentCount;  /* 1..3 */
  const char* name;            /* primvar name, e.g. "displayColor" */
  float       fallback[4];
} GtlcPrimvarBinding;

typedef struct GtlcMaterialDesc {
  int32_t     closure;                           /* GTLC_CLOSURE_* */
  const char* name;                              /* may be NULL */
  float       params[GTLC_MATERIAL_PARAM_COUNT]; /* see enums above */
  const GtlcTextureBinding* textures;            /* optional; applied in order per shaded hit */
  uint32_t    textureCount;
  uint32_t    primvarCount;                      /* <= GTLC_MAX_PRIMVAR_BINDINGS */
  const GtlcPrimvarBinding* primvars;            /* optional; applied before the textures */
} GtlcMaterialDesc;

/* Fill `desc` with the closure's spec defaults (UsdPreviewSurface spec / open_pbr_surface.mtlx:4-97). */
GTLC_API void gtlc_material_desc_defaults(int32_t closure, GtlcMaterialDesc* desc);

/* ---- lifecycle: giInitialize / giTerminate (Gi.h:199-200) -------------------------------- */
typedef struct GtlcInitParams {
  int32_t deviceIndex;  /* CUDA ordinal; -1 = LOCAL_RANK env or 0 (reference: GTL_DEVICE_INDEX_OVERRIDE, CgpuVk.cpp:975) */
  int32_t reserved;
} GtlcInitParams;

GTLC_API int  gtlc_initialize(const GtlcInitParams* params);
GTLC_API void gtlc_terminate(void);
/* Last error text (reference logs through GB_ERROR; the C-ABI keeps the line for the binding to log). */
GTLC_API const char* gtlc_last_error(void);
/* Error lines logged since the library was loaded: what a test fixture needs to fail on ANY error line the way the
 * reference's does with its log sink (src/hdGatling/main.cpp:78-100,531-549). */
GTLC_API uint64_t gtlc_error_count(void);

/* ---- scene: giCreateScene / giDestroyScene (Gi.h:220-221) -------------------------------- */
GTLC_API GtlcHandle gtlc_scene_create(void);
GTLC_API void       gtlc_scene_destroy(GtlcHandle scene);

/* ---- materials: giCreateMaterialFrom* / giDestroyMaterial (Gi.h:204-207) ----------------- */
GTLC_API GtlcHandle gtlc_material_create(GtlcHandle scene, const GtlcMaterialDesc* desc);
GTLC_API void       gtlc_material_destroy(GtlcHandle material);

/* ---- meshes: Gi.h:209-216 ---------------------------------------------------------------- */
GTLC_API GtlcHandle gtlc_mesh_create(GtlcHandle scene, const GtlcMeshDesc* desc);          /* giCreateMesh */
GTLC_API void gtlc_mesh_set_transform(GtlcHandle mesh, const float* mat4x4);                /* giSetMeshTransform: USD row-major float[16] */
GTLC_API void gtlc_mesh_set_instance_transforms(GtlcHandle mesh, uint32_t count, const float* transforms4x4); /* giSetMeshInstanceTransforms */
GTLC_API void gtlc_mesh_set_instance_ids(GtlcHandle mesh, uint32_t count, const int32_t* ids);               /* giSetMeshInstanceIds */
GTLC_API void gtlc_mesh_set_instancer_primvars(GtlcHandle mesh, uint32_t count, const GtlcPrimvarDesc* primvars);  /* giSetMeshInstancerPrimvars */
GTLC_API void gtlc_mesh_set_material(GtlcHandle mesh, GtlcHandle material);                 /* giSetMeshMaterial */
GTLC_API void gtlc_mesh_set_visibility(GtlcHandle mesh, int32_t visible);                   /* giSetMeshVisibility */
GTLC_API void gtlc_mesh_destroy(GtlcHandle mesh);                                           /* giDestroyMesh */

/* ---- analytic lights: Gi.h:223-251 ------------------------------------------------------- */
/* sphere: Gi.h:223-228 (host encodings Gi.cpp:2573-2661) */
GTLC_API GtlcHandle gtlc_sphere_light_create(GtlcHandle scene);
GTLC_API void gtlc_sphere_light_destroy(GtlcHandle scene, GtlcHandle light);
GTLC_API void gtlc_sphere_light_set_position(GtlcHandle light, const float* position);
GTLC_API void gtlc_sphere_light_set_base_emission(GtlcHandle light, const float* rgb);
GTLC_API void gtlc_sphere_light_set_radius(GtlcHandle light, float radiusX, float radiusY, float radiusZ);
GTLC_API void gtlc_sphere_light_set_diffuse_specular(GtlcHandle light, float diffuse, float specular);

/* distant: Gi.h:230-235 (Gi.cpp:2663-2745) */
GTLC_API GtlcHandle gtlc_distant_light_create(GtlcHandle scene);
GTLC_API void gtlc_distant_light_destroy(GtlcHandle scene, GtlcHandle light);
GTLC_API void gtlc_distant_light_set_direction(GtlcHandle light, const float* direction);
GTLC_API void gtlc_distant_light_set_base_emission(GtlcHandle light, const float* rgb);
GTLC_API void gtlc_distant_light_set_angle(GtlcHandle light, float angle);
GTLC_API void gtlc_distant_light_set_diffuse_specular(GtlcHandle light, float diffuse, float specular);

/* rect: Gi.h:237-243 (Gi.cpp:2747-2843) */
GTLC_API GtlcHandle gtlc_rect_light_create(GtlcHandle scene);
GTLC_API void gtlc_rect_light_destroy(GtlcHandle scene, GtlcHandle light);
GTLC_API void gtlc_rect_light_set_origin(GtlcHandle light, const float* origin);
GTLC_API void gtlc_rect_light_set_tangents(GtlcHandle light, const float* t0, const float* t1);
GTLC_API void gtlc_rect_light_set_base_emission(GtlcHandle light, const float* rgb);
GTLC_API void gtlc_rect_light_set_dimensions(GtlcHandle light, float width, float height);
GTLC_API void gtlc_rect_light_set_diffuse_specular(GtlcHandle light, float diffuse, float specular);

/* disk: Gi.h:245-251 (Gi.cpp:2845-2941) */
GTLC_API GtlcHandle gtlc_disk_light_create(GtlcHandle scene);
GTLC_API void gtlc_disk_light_destroy(GtlcHandle scene, GtlcHandle light);
GTLC_API void gtlc_disk_light_set_origin(GtlcHandle light, const float* origin);
GTLC_API void gtlc_disk_light_set_tangents(GtlcHandle light, const float* t0, const float* t1);
GTLC_API void gtlc_disk_light_set_base_emission(GtlcHandle light, const float* rgb);
GTLC_API void gtlc_disk_light_set_radius(GtlcHandle light, float radiusX, float radiusY);
GTLC_API void gtlc_disk_light_set_diffuse_specular(GtlcHandle light, float diffuse, float specular);

/* ---- dome light: Gi.h:253-257.  The reference loads a lat-long image through imgio (out of
 * scope); the C-ABI takes decoded RGBA float32 texels instead of a file path. ---------------- */
GTLC_API GtlcHandle gtlc_dome_light_create(GtlcHandle scene, uint32_t width, uint32_t height, const float* rgbaTexels);
GTLC_API void gtlc_dome_light_destroy(GtlcHandle light);
GTLC_API void gtlc_dome_light_set_rotation(GtlcHandle light, const float* quatXYZW);
GTLC_API void gtlc_dome_light_set_base_emission(GtlcHandle light, const float* rgb);
GTLC_API void gtlc_dome_light_set_diffuse_specular(GtlcHandle light, float diffuse, float specular);

/* ---- render buffers: Gi.h:259-261 -------------------------------------------------------- */
GTLC_API GtlcHandle gtlc_render_buffer_create(uint32_t width, uint32_t height, int32_t format);  /* giCreateRenderBuffer */
GTLC_API void       gtlc_render_buffer_destroy(GtlcHandle renderBuffer);                         /* giDestroyRenderBuffer */
GTLC_API void*      gtlc_render_buffer_mem(GtlcHandle renderBuffer);                             /* giGetRenderBufferMem: pinned host memory, row 0 = bottom */
/* Device-side copy of the same buffer (what gtlc_render wrote before the D2H copy); the multi-GPU
 * harness hands this pointer to ncclReduce. */
GTLC_API void*      gtlc_render_buffer_device_mem(GtlcHandle renderBuffer);

/* ---- render: giRender (Gi.h:218).  Synchronous: on return every bound AOV is in host memory. */
GTLC_API int gtlc_render(const GtlcRenderParams* params);

/*
 * ---- extensions that have no counterpart in Gi.h (documented in INTEGRATION.md) -------------
 */

/* Multi-GPU framebuffer sharding: this process renders only the tiles with
 * (tileIndex % worldSize == rank); all other pixels of the colour AOV are written as 0 so that a
 * single sum-reduce over ranks yields the full image.  tileSize in pixels (square tiles; 0 = 64).
 * tileSize == GTLC_SHARD_SAMPLES selects the alternative of SURVEY.md 8e, sample-range sharding: the process renders
 * EVERY pixel but only the samples [spp * rank / worldSize, spp * (rank + 1) / worldSize) of each gtlc_render call, scaled
 * by the call's full sample count (every term of rp_main.rgen:498 and :514 is linear in the samples), with w = 1 on rank 0
 * and 0 elsewhere -- the same sum-reduce yields the image.  Tiles give bit-identical pixels (x + 0); sample ranges agree to
 * rounding (a + b summed in another order) and balance ranks on scenes whose cost varies across the frame.  Auxiliary
 * AOVs are not reduced: under sample-range sharding each rank's are those of its own samples. */
#define GTLC_SHARD_SAMPLES 0xFFFFFFFFu
GTLC_API void gtlc_scene_set_shard(GtlcHandle scene, uint32_t rank, uint32_t worldSize, uint32_t tileSize);

/* The one collective of the path: with a communicator in place, gtlc_render() of a sharded scene ends
 * with ncclReduce(sum, float, W x H x 4, root 0) of the HDR colour accumulator on the render stream, BEFORE the
 * device->host copy, so rank 0's render buffer holds the complete image when the call returns (the other ranks' host
 * buffers are not written).  Every rank of the communicator must call gtlc_render() for the same image size, in the
 * same order.  libnccl.so.2 is opened at gtlc_comm_get_unique_id / gtlc_comm_init (no link-time dependency: a
 * single-GPU host never needs NCCL).  Bootstrap: rank 0 creates the id, the host's own control plane (MPI, a TCP
 * store, torch.distributed, a file) carries its GTLC_COMM_ID_BYTES bytes to the other ranks, everybody calls
 * gtlc_comm_init.  The reference has a single device (CgpuVk.cpp:959-1000): no counterpart in Gi.h. */
#define GTLC_COMM_ID_BYTES 128
GTLC_API int  gtlc_comm_get_unique_id(void* id);                                   /* ncclGetUniqueId */
GTLC_API int  gtlc_comm_init(const void* id, int32_t rank, int32_t worldSize);     /* ncclCommInitRank on the library's device (after gtlc_initialize) */
GTLC_API void gtlc_comm_destroy(void);
GTLC_API int  gtlc_comm_world_size(void);                                          /* 0 = no communicator */

/* Render-loop tuning knobs (all have defaults; the bench records the values it used). */
typedef struct GtlcTuning {
  uint32_t maxPathsPerWave;   /* path-state pool size; default 1<<24 */
  int32_t  sortByMaterial;    /* -1 auto (on iff >1 material), 0 off, 1 on */
  int32_t  skipReadback;      /* 1: leave AOVs on the device (kernel-only timing) */
  int32_t  countVisits;       /* 1: traversal kernels also count node/triangle visits (slower; roofline bookkeeping) */
  int32_t  traceVariant;      /* traversal kernel: 0 default (= 7, warp-batched triangle tests); 1 per-lane triangle loop (the walk order of the visit-counting pass and of the oracle); 6 TMA-staged ray tiles */
  int32_t  reserved[3];       /* [0] refill threshold (idle lanes), [1] trace CTAs per SM, [2] unused; 0 = default */
  int32_t  bvhBuilder;        /* 0 host binned-SAH build (default); 1 device LBVH build (fastest rebuilds); 2 device PLOC build (close to the SAH tree's trace speed); takes effect at the next geometry change */
} GtlcTuning;
GTLC_API void gtlc_set_tuning(const GtlcTuning* tuning);
GTLC_API void gtlc_get_tuning(GtlcTuning* tuning);

/* Counters of the last gtlc_render / gtlc_trace_rays call. */
typedef struct GtlcStats {
  uint64_t closestRays;       /* rays traced by trace_closest */
  uint64_t shadowRays;        /* rays traced by trace_any (tMax > 0 only) */
  uint64_t nodeVisits;        /* BVH8 nodes fetched by closest-hit (gtlc_trace_rays: any-hit) rays (valid iff countVisits) */
  uint64_t triVisits;         /* triangles tested by them                                           (valid iff countVisits) */
  uint64_t kernelLaunches;    /* number of __global__ launches issued */
  uint64_t paths;             /* camera paths started */
  float    msTotal;           /* CUDA-event time of the whole call on the render stream */
  float    msTraceClosest;    /* summed CUDA-event time of trace_closest launches */
  float    msTraceAny;
  float    msShade;
  float    msRaygen;
  float    msSort;
  float    msResolve;
  float    msBvhBuild;        /* host BVH8 build + upload of the last (re)build, wall clock */
  uint32_t bvhNodeCount;
  uint32_t bvhTriCount;
  uint32_t bvhMaxDepth;
  uint32_t maxNodeVisitsPerRay; /* valid iff countVisits */
  uint64_t nonFiniteRays;       /* rays with a NaN/Inf origin or direction (answered as misses without traversal) */
  float    msBvhDeviceBuild;    /* bvhBuilder = 1: wall clock of the device build alone (scratch allocation .. final node copy), part of msBvhBuild */
  float    reservedStats;
  uint64_t closestHits;         /* closest-hit rays that found a surface = hits the shade kernel shades (valid iff countVisits) */
  uint64_t shadowNodeVisits;    /* BVH8 nodes fetched / triangles tested by shadow rays (valid iff countVisits) */
  uint64_t shadowTriVisits;
} GtlcStats;
GTLC_API void gtlc_get_stats(GtlcStats* stats);

/* Ray-query entry point used by the traversal parity tests and the roofline bench: traces `count`
 * rays (host arrays) against the scene's BVH8 with the production kernels.
 *   rays : count x 8 floats  (ox, oy, oz, tMin, dx, dy, dz, tMax)          [32 B, SURVEY 8d]
 *   hits : count x 5 x 4 B   (t, u, v, primitiveId(u32), instanceIndex(u32)) [20 B]; miss: t = -1,
 *          prim = inst = 0xFFFFFFFF.  For anyHit != 0 only `t` is meaningful (>=0 occluded, -1 free).
 *   visits : optional count x 2 uint32 (nodes, triangles) per-ray visit counts, or NULL.
 * tMin must be >= 0 (Vulkan's rule for traceRayEXT); a negative tMin is refused with GTLC_ERROR. */
GTLC_API int gtlc_trace_rays(GtlcHandle scene, const float* rays, uint32_t count, int32_t anyHit,
                             void* hits, uint32_t* visits);

/* Access to the host copy of the flattened BVH8 (test/oracle cross-check of the identical structure). */
typedef struct GtlcBvhView {
  const void*  nodes;       /* nodeCount x 80 B compressed BVH8 nodes (layout: DESIGN.md) */
  uint32_t     nodeCount;
  const float* triangles;   /* triCount x 12 floats: 3 x (x, y, z, w); w0 = primId bits, w1 = instance bits */
  uint32_t     triCount;
} GtlcBvhView;
GTLC_API int gtlc_scene_get_bvh(GtlcHandle scene, GtlcBvhView* view);

#ifdef __cplusplus
}
#endif

#endif /* GTLC_H */
