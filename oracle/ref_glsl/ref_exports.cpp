// TEST INFRASTRUCTURE.  C entry points around the reference's own shader functions (see build_ref.py for how the sources get here).
// Everything between the `#include`s of generated files is glue: the globals the fragments expect (ubo, light buffers, the ray
// direction built-in, a recording stand-in for the dome-light texture fetch) and plain-pointer wrappers.
#include "glsl_shim.h"
#undef UINT32_MAX                       // common.glsl declares its own constant of that name
#include "interface/rp_main.h"          // the reference's shared C++/GLSL header (through oracle/ref_glsl/glm/glm.hpp): UniformData, *Light
using namespace gtl::shader_interface::rp_main;

static UniformData ubo;                      // (read-only while a render runs)
static SphereLight sphereLights[16];
static DistantLight distantLights[16];
static RectLight rectLights[16];
static DiskLight diskLights[16];
static thread_local vec3 gl_WorldRayDirectionEXT;
static int tex_sampler;
static inline int sampler2D(int, int) { return 0; }
#define nonuniformEXT(x) (x)

// mdl_shading_state.glsl reads the hit mesh through buffer references and the ray-tracing built-ins; the stand-ins point at test arrays
static thread_local BlasPayload blas_payloads[1];
static thread_local const Face* g_faces; static thread_local const FVertex* g_vertices;
static thread_local BlasPayloadBufferPreamble g_preamble;
struct IndexBuffer { const Face* data; BlasPayloadBufferPreamble preamble; explicit IndexBuffer(uint64_t) : data(g_faces), preamble(g_preamble) {} };
static vec4 ClearValuesF[17]; static ivec4 ClearValuesI[17];
static const int *g_instanceIds, *g_objectIds;            // per instance (orc_instance_ids), set by the AOV render
struct VertexBuffer { const FVertex* data; explicit VertexBuffer(uint64_t) : data(g_vertices) {} };
static thread_local int gl_InstanceCustomIndexEXT, gl_PrimitiveID;
static thread_local mat4x3 gl_ObjectToWorldEXT, gl_WorldToObjectEXT;
static thread_local uvec3 gl_LaunchIDEXT;
static inline void clearAovs(uint) {}          // (the camera-head builds; the whole-pipeline builds compile the real one, rp_main.rgen:132-183)
#define SCENE_DATA_COUNT 0

#include "common.glsl"
#include "colormap.glsl"
#include "mdl_types.glsl"
#include "mdl_wrap.inc"
#include "mdl_adapt_normal.inc"
// mdl_interface.glsl:40-205, the MDL texture runtime: the hardware's part (size query, filtered sample, texel fetch) comes from the caller
namespace texrt {
struct TexCallbacks { const void* desc; void (*sample)(const void*, float, float, float, int, float*); void (*texel)(const void*, int, int, int, float*); int w, h, d; };
static TexCallbacks g_tex;
struct TexArray { int operator[](uint) const { return 0; } };
static const TexArray textures_2d, textures_3d;
struct Sampler2 { }; struct Sampler3 { };
static inline Sampler2 sampler2D(int, int) { return Sampler2(); }
static inline Sampler3 sampler3D(int, int) { return Sampler3(); }
struct SizeQuery { operator ivec2() const { return ivec2(g_tex.w, g_tex.h); } operator ivec3() const { return ivec3(g_tex.w, g_tex.h, g_tex.d); } };
static inline SizeQuery textureSize(int, int) { return SizeQuery(); }
static inline vec4 texture(Sampler2, vec2 c) { float o[4]; g_tex.sample(g_tex.desc, c.x, c.y, 0.0f, 0, o); return vec4(o[0], o[1], o[2], o[3]); }
static inline vec4 texture(Sampler3, vec3 c) { float o[4]; g_tex.sample(g_tex.desc, c.x, c.y, c.z, 1, o); return vec4(o[0], o[1], o[2], o[3]); }
static inline vec4 texelFetch(Sampler2, ivec2 c, int) { float o[4]; g_tex.texel(g_tex.desc, c.x, c.y, 0, o); return vec4(o[0], o[1], o[2], o[3]); }
static inline vec4 texelFetch(Sampler3, ivec3 c, int) { float o[4]; g_tex.texel(g_tex.desc, c.x, c.y, c.z, o); return vec4(o[0], o[1], o[2], o[3]); }
#define TEXTURE_INDEX_OFFSET 0
#include "mdl_textures.inc"
}
// mdl_interface.glsl:266-452, the "scene data" (primvar) runtime: State carries the renderer state here, buffer references become pointers
namespace scenedata {
#undef SCENE_DATA_COUNT
#define SCENE_DATA_COUNT 6
#include "mdl_renderer_state.glsl"
#define RENDERER_STATE_TYPE mdl_renderer_state
#undef H_MDL_TYPES
#include "mdl_types.glsl"                   // again, in this namespace: State with the renderer_state member
struct BufferRefInt { const int* data; explicit BufferRefInt(uint64_t a) : data((const int*) a) {} };
struct BufferRefFloat { const float* data; explicit BufferRefFloat(uint64_t a) : data((const float*) a) {} };
struct BufferRefVec2 { const vec2* data; explicit BufferRefVec2(uint64_t a) : data((const vec2*) a) {} };
struct BufferRefVec4 { const vec4* data; explicit BufferRefVec4(uint64_t a) : data((const vec4*) a) {} };
static const int* InstanceIds; static int gl_InstanceID;
#include "mdl_scene_data.inc"
#undef RENDERER_STATE_TYPE
#undef SCENE_DATA_COUNT
#define SCENE_DATA_COUNT 0
}
// function-level copies for the per-function tests (the whole-pipeline builds further down bring their own, in their namespaces)
namespace fn {
// rp_main.miss:46-53 reads the dome-light image; the stand-in returns the lat-long coordinate it was asked for
static int textures_2d[4];
static inline vec4 textureLod(int, vec2 uv, uint) { return vec4(uv.x, uv.y, 0.0f, 1.0f); }
#include "miss_functions.inc"
#define NEXT_EVENT_ESTIMATION
#include "chit_sample_light.inc"
#undef NEXT_EVENT_ESTIMATION
}
namespace mss2 {
#define MEDIUM_STACK_SIZE 2
#include "rp_main_payload.glsl"
#include "rgen_functions.inc"
#undef MEDIUM_STACK_SIZE
}
namespace mss0 {
#define MEDIUM_STACK_SIZE 0
#include "rp_main_payload.glsl"
#undef MEDIUM_STACK_SIZE
}

#include "mdl_shading_state.glsl"
static mss2::ShadeRayPayload rayPayload;
// rp_main.rgen:191-288 is the head of main(): it is compiled once per combination of the compile-time switches the reference builds
// its pipeline with (GlslShaderGen.cpp:57-72).  The fragment ends inside the sample loop, which the wrapper closes after one pass.
#define CAMERA_VARIANT(NAME)                                                                                         \
  namespace NAME {                                                                                                   \
    using mss2::fisGauss;                                                                                            \
    static void camera(uint px, uint py, float* out8, uint* rngOut)                                                  \
    {                                                                                                                \
      gl_LaunchIDEXT = uvec3(px, py, 0u);
#define CAMERA_VARIANT_END                                                                                           \
        out8[0] = ray_origin.x; out8[1] = ray_origin.y; out8[2] = ray_origin.z; out8[3] = ray_dir.x; out8[4] = ray_dir.y; out8[5] = ray_dir.z; \
        out8[6] = clipRange.x; out8[7] = clipRange.y; *rngOut = rng_state;                                           \
        break;                                                                                                       \
      }                                                                                                              \
    }                                                                                                                \
  }
#define MEDIUM_STACK_SIZE 2
CAMERA_VARIANT(cam_plain)
#include "rgen_camera.inc"
CAMERA_VARIANT_END
#define JITTERED_SAMPLING
CAMERA_VARIANT(cam_jitter)
#include "rgen_camera.inc"
CAMERA_VARIANT_END
#define FILTER_IMPORTANCE_SAMPLING
CAMERA_VARIANT(cam_jitter_fis)
#include "rgen_camera.inc"
CAMERA_VARIANT_END
#define DEPTH_OF_FIELD
CAMERA_VARIANT(cam_jitter_fis_dof)
#include "rgen_camera.inc"
CAMERA_VARIANT_END
#undef JITTERED_SAMPLING
#undef FILTER_IMPORTANCE_SAMPLING
CAMERA_VARIANT(cam_dof)
#include "rgen_camera.inc"
CAMERA_VARIANT_END
#undef DEPTH_OF_FIELD
#undef MEDIUM_STACK_SIZE

// ---- the whole integrator: rp_main.rgen / .chit / .miss / _shadow.miss as C++ (ref_pipeline.inc), driven through a callback table ----
struct RefCallbacks {
  void* scene; const void* renderDesc;
  int (*trace)(void* scene, const float* o, const float* d, float tMin, float tMax, int anyHit, uint seed, float* tuv, uint* primInst);
  void (*views)(void* scene, uint inst, const uint** faces, const void** fverts, uint* bases3, float* o2w12, float* w2o12, uint* instCount);
  uint (*hitBegin)(void* scene, uint inst, uint prim, float u, float v, const float* worldRayDir);
  void (*mappedNormal)(float* n);
  void (*volume)(float* sigma_a, float* sigma_s, float* ior, float* bias);
  void (*bsdfSample)(const float* state18, const float* ior1, const float* ior2, const float* k1, const float* xi, float* k2, float* pdf, float* bsdfOverPdf, int* eventType);
  void (*bsdfEvaluate)(const float* state18, const float* ior1, const float* ior2, const float* k1, const float* k2, float* diffuse, float* glossy, float* pdf);
  void (*bsdfAlbedo)(const float* state18, const float* k1, float* albedo);
  void (*edf)(const float* state18, const float* k1, float* edf, float* pdf, float* cosOut, float* intensity);
  void (*dome)(void* scene, const void* renderDesc, uint domeLightIndex, float u, float v, float* rgb);
};
static RefCallbacks g_cb;
static thread_local int g_domeIndex;
static thread_local uint64_t g_rayCounts[2];
static uint64_t g_totalRayCounts[2];
#define IS_THIN_WALLED
#define IS_EMISSIVE
#define AOV_MASK AOV_BIT_COLOR
#define NEXT_EVENT_ESTIMATION
#define PROGRESSIVE_ACCUMULATION
#define JITTERED_SAMPLING
#define FILTER_IMPORTANCE_SAMPLING
#define DOME_LIGHT_CAMERA_VISIBLE
// the reference's default pipeline (renderDelegate.cpp:93-110) with next-event estimation, Beer-only media ...
#define PIPE_NS pipe_default
#define MEDIUM_STACK_SIZE 0
#include "ref_pipeline.inc"
#undef PIPE_NS
#undef MEDIUM_STACK_SIZE
// ... without next-event estimation (config C2 of BASELINE.json renders that way)
#undef NEXT_EVENT_ESTIMATION
#define PIPE_NS pipe_nonee
#define MEDIUM_STACK_SIZE 0
#include "ref_pipeline.inc"
#undef PIPE_NS
#undef MEDIUM_STACK_SIZE
#define NEXT_EVENT_ESTIMATION
// ... and with a two-deep medium stack (volume random walk)
#define PIPE_NS pipe_media2
#define MEDIUM_STACK_SIZE 2
#include "ref_pipeline.inc"
#undef PIPE_NS
#undef MEDIUM_STACK_SIZE
// ... and the dome light hidden from the camera, depth of field and clipping planes on, no jitter, no progressive accumulation
#undef DOME_LIGHT_CAMERA_VISIBLE
#undef JITTERED_SAMPLING
#undef FILTER_IMPORTANCE_SAMPLING
#undef PROGRESSIVE_ACCUMULATION
#define DEPTH_OF_FIELD
#define CLIPPING_PLANES
#define PIPE_NS pipe_other
#define MEDIUM_STACK_SIZE 0
#include "ref_pipeline.inc"
#undef PIPE_NS
#undef MEDIUM_STACK_SIZE
#undef DEPTH_OF_FIELD
#undef CLIPPING_PLANES
// ... and the default pipeline with every auxiliary AOV bound (all of rp_main.chit:191-290, rgen:132-183,431-436,483-486,517-520) except
// the two whose inputs the oracle does not model: ClockCycles (a timer) and FaceId (byte-packed id table behind the index buffer)
#undef AOV_MASK
#define AOV_MASK (0x1FFFF & ~AOV_BIT_DEBUG_CLOCK_CYCLES & ~AOV_BIT_FACE_ID)
#define PROGRESSIVE_ACCUMULATION
#define JITTERED_SAMPLING
#define FILTER_IMPORTANCE_SAMPLING
#define DOME_LIGHT_CAMERA_VISIBLE
#define PIPE_NS pipe_aovs
#define MEDIUM_STACK_SIZE 0
#include "ref_pipeline.inc"
#undef PIPE_NS
#undef MEDIUM_STACK_SIZE
#undef NEXT_EVENT_ESTIMATION

static inline vec3 ld3(const float* p) { return vec3(p[0], p[1], p[2]); }
static inline void st3(float* o, vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
#define EXPORT extern "C" __attribute__((visibility("default")))

// ---- common.glsl ----
EXPORT uint ref_hash_theironborn(uint x) { return hash_theironborn(x); }
EXPORT uint ref_hash_pcg32(uint* state) { return hash_pcg32(*state); }
EXPORT float ref_uint_as_float(uint v) { return uintAsFloat(v); }
EXPORT uint ref_rng_init(uint pixelIndex, uint sampleIndex) { return rng1d_init(pixelIndex, sampleIndex); }
EXPORT float ref_rng_next1f(uint* state) { return rng1d_next1f(*state); }
EXPORT void ref_rng_next4f(uint* state, float* o) { vec4 r = rng1d_next4f(*state); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
EXPORT void ref_orthonormal_basis(const float* n, float* b1, float* b2) { vec3 a, b; orthonormal_basis(ld3(n), a, b); st3(b1, a); st3(b2, b); }
EXPORT void ref_offset_ray_origin(const float* p, const float* n, float* o) { st3(o, offset_ray_origin(ld3(p), ld3(n))); }
EXPORT uint ref_encode_direction(const float* d) { return encode_direction(ld3(d)); }
EXPORT void ref_decode_direction(uint e, float* o) { st3(o, decode_direction(e)); }
EXPORT void ref_sample_hemisphere(const float* xi, float* o) { st3(o, sample_hemisphere(vec2(xi[0], xi[1]))); }
EXPORT void ref_sample_sphere(const float* xi, const float* radius, float* o) { st3(o, sample_sphere(vec2(xi[0], xi[1]), ld3(radius))); }
EXPORT void ref_sample_disk(const float* xi, const float* radius, float* o) { vec2 r = sample_disk(vec2(xi[0], xi[1]), vec2(radius[0], radius[1])); o[0] = r.x; o[1] = r.y; }
EXPORT float ref_luminance(const float* c) { return luminance(ld3(c)); }
EXPORT float ref_safe_div(float a, float b) { return safe_div(a, b); }
// ---- colormap.glsl ----
EXPORT void ref_colormap(int which, float t, float* o) { st3(o, which == 0 ? colormap_viridis(t) : which == 1 ? colormap_plasma(t) : which == 2 ? colormap_magma(t) : which == 3 ? colormap_inferno(t) : colormap_turbo(t)); }
// ---- rp_main.rgen:48-130 ----
EXPORT float ref_sample_distance(const float* albedo, const float* throughput, const float* sigma_t, float xi, float* pdf) { vec3 p; float r = mss2::sampleDistance(ld3(albedo), ld3(throughput), ld3(sigma_t), xi, p); st3(pdf, p); return r; }
EXPORT float ref_sample_hg_cos(float r, float g) { return mss2::sampleHenyeyGreensteinCos(r, g); }
EXPORT void ref_sample_volume_scattering_direction(const float* xi, float bias, float* dir) { vec3 d = ld3(dir); mss2::sampleVolumeScatteringDirection(vec2(xi[0], xi[1]), bias, d); st3(dir, d); }
EXPORT int ref_russian_roulette(float xi, float* throughput, float rrInvMinTermProb) { ubo.rrInvMinTermProb = rrInvMinTermProb; vec3 t = ld3(throughput); bool r = mss2::russian_roulette(xi, t); st3(throughput, t); return r ? 1 : 0; }
EXPORT void ref_fis_gauss(const float* xi, float* o) { vec2 r = mss2::fisGauss(vec2(xi[0], xi[1])); o[0] = r.x; o[1] = r.y; }
// ---- rp_main_payload.glsl ----
EXPORT uint ref_payload_get_medium_idx(uint bitfield, int mediumStackSize) { if (mediumStackSize == 2) { mss2::ShadeRayPayload p; p.bitfield = bitfield; return mss2::shadeRayPayloadGetMediumIdx(p); } mss0::ShadeRayPayload p; p.bitfield = bitfield; return mss0::shadeRayPayloadGetMediumIdx(p); }
EXPORT uint ref_payload_set_medium_idx(uint bitfield, uint idx) { mss2::ShadeRayPayload p; p.bitfield = bitfield; mss2::shadeRayPayloadSetMediumIdx(p, idx); return p.bitfield; }
EXPORT uint ref_payload_increment_walk(uint bitfield) { mss2::ShadeRayPayload p; p.bitfield = bitfield; mss2::shadeRayPayloadIncrementWalk(p); return p.bitfield; }
EXPORT uint ref_payload_get_walk(uint bitfield) { mss2::ShadeRayPayload p; p.bitfield = bitfield; return mss2::shadeRayPayloadGetWalk(p); }
// ---- rp_main.miss:39-53 ----
EXPORT void ref_quat_rotate_dir(const float* q, const float* dir, float* o) { st3(o, fn::quatRotateDir(vec4(q[0], q[1], q[2], q[3]), ld3(dir))); }
EXPORT void ref_dome_light_uv(const float* rayDir, float* uv) { vec3 r = fn::sampleDomeLight(1u, ld3(rayDir)); uv[0] = r.x; uv[1] = r.y; }
// ---- mdl_interface.glsl:8-38, 238-256 ----
EXPORT float ref_apply_wrap_and_crop(float coord, int wrap, int res) { return apply_wrap_and_crop(coord, wrap, vec2(0.0f, 1.0f), res); }
EXPORT void ref_mdl_adapt_normal(const float* worldRayDir, const float* geomNormal, const float* normal, float* o)
{
  gl_WorldRayDirectionEXT = ld3(worldRayDir);
  State s; s.geom_normal = ld3(geomNormal); s.normal = ld3(normal);
  st3(o, mdl_adapt_normal(s, ld3(normal)));
}
// ---- rp_main.chit:30-129 ----
// counts[4] = sphere, distant, rect, disk; the light arrays hold the interface structs byte for byte (48 / 48 / 48 / 48 bytes each)
EXPORT void ref_sample_light(const uint* counts, const void* spheres, const void* distants, const void* rects, const void* disks, float lightIntensityMultiplier,
                             float sensorExposure, const float* k4, const float* surfacePos, float* out8, uint* diffuseSpecularPacked)
{
  ubo.sphereLightCount = counts[0]; ubo.distantLightCount = counts[1]; ubo.rectLightCount = counts[2]; ubo.diskLightCount = counts[3];
  ubo.totalLightCount = counts[0] + counts[1] + counts[2] + counts[3];
  ubo.lightIntensityMultiplier = lightIntensityMultiplier; ubo.sensorExposure = sensorExposure;
  // (k4.x == 0 enters the sphere branch even when there is no sphere light: the shader then reads element 0 of a buffer that holds no
  //  light -- zeros, as the oracle assumes; the stand-in buffers start from zeros on every call for the same reason)
  memset(sphereLights, 0, sizeof(sphereLights)); memset(distantLights, 0, sizeof(distantLights)); memset(rectLights, 0, sizeof(rectLights)); memset(diskLights, 0, sizeof(diskLights));
  memcpy(sphereLights, spheres, sizeof(SphereLight) * counts[0]); memcpy(distantLights, distants, sizeof(DistantLight) * counts[1]);
  memcpy(rectLights, rects, sizeof(RectLight) * counts[2]); memcpy(diskLights, disks, sizeof(DiskLight) * counts[3]);
  vec3 dirToLight, power; float dist, invPdf; uint dsp;
  fn::sampleLight(vec4(k4[0], k4[1], k4[2], k4[3]), ld3(surfacePos), dirToLight, dist, power, invPdf, dsp);
  st3(out8, dirToLight); out8[3] = dist; st3(out8 + 4, power); out8[7] = invPdf; *diffuseSpecularPacked = dsp;
}
EXPORT int ref_struct_sizes(int which) { return which == 0 ? (int) sizeof(SphereLight) : which == 1 ? (int) sizeof(DistantLight) : which == 2 ? (int) sizeof(RectLight) : (int) sizeof(DiskLight); }
// ---- rp_main.rgen:191-288 ----
// flags: bit 0 JITTERED_SAMPLING, bit 1 FILTER_IMPORTANCE_SAMPLING, bit 2 DEPTH_OF_FIELD.  cam12 / bits2 as orc_camera_ubo delivers them.
EXPORT int ref_camera_ray(int flags, const float* cam12, const uint* bits2, uint sampleIndex, uint px, uint py, float* out8, uint* rngOut)
{
  ubo.cameraPosition = ld3(cam12); ubo.cameraForward = ld3(cam12 + 3); ubo.cameraUp = ld3(cam12 + 6);
  ubo.cameraVFoV = cam12[9]; ubo.focusDistance = cam12[10]; ubo.lensRadius = cam12[11];
  ubo.clipRangePacked = bits2[0]; ubo.imageDims = bits2[1];
  ubo.spp = 1; ubo.sampleOffset = sampleIndex;
  switch (flags) {
    case 0: cam_plain::camera(px, py, out8, rngOut); return 1;
    case 1: cam_jitter::camera(px, py, out8, rngOut); return 1;
    case 3: cam_jitter_fis::camera(px, py, out8, rngOut); return 1;
    case 7: cam_jitter_fis_dof::camera(px, py, out8, rngOut); return 1;
    case 4: cam_dof::camera(px, py, out8, rngOut); return 1;
  }
  return 0;
}
// ---- mdl_shading_state.glsl ----
// faces: the hit mesh's index triples; vertices: the vertex array its vertexOffset indexes; o2w / w2o: row-major 3x4
EXPORT void ref_setup_shading_state(const void* faces, const void* vertices, uint vertexOffset, const float* o2w12, const float* w2o12, uint prim, float u, float v,
                                    const float* rayDir, float* out18, int* frontFace)
{
  g_faces = (const Face*) faces; g_vertices = (const FVertex*) vertices;
  blas_payloads[0].vertexOffset = vertexOffset; blas_payloads[0].bufferAddress = 0; gl_InstanceCustomIndexEXT = 0; gl_PrimitiveID = (int) prim;
  memcpy(gl_ObjectToWorldEXT.m, o2w12, 48); memcpy(gl_WorldToObjectEXT.m, w2o12, 48);
  gl_WorldRayDirectionEXT = ld3(rayDir);
  State st; bool ff = false;
  setup_mdl_shading_state(vec2(u, v), st, ff);
  st3(out18, st.normal); st3(out18 + 3, st.geom_normal); st3(out18 + 6, st.position); st3(out18 + 9, st.tangent_u[0]); st3(out18 + 12, st.tangent_v[0]); st3(out18 + 15, st.text_coords[0]);
  *frontFace = ff ? 1 : 0;
}
// ---- rp_main.rgen + .chit + .miss + _shadow.miss: a whole render ----
// f28 / u16: orc_ubo_values; lights: orc_light_buffers + the counts in u16; u16[11] = medium stack size, u16[12] = switches (bit 0 jitter,
// 1 filter importance sampling, 2 depth of field, 3 NEE, 4 clipping planes, 5 progressive, 6 dome visible to the camera).
// rgba: width * height vec4, holds the previous ColorAov on entry when progressive.  Returns 0 if no pipeline was compiled for the switches.
// aov: NULL, or 17 pointers (by AOV id; vec3 AOVs as tightly packed float[3], ids as int, depth as float; entries 0, 6, 13 unused) that
// selects the all-AOVs pipeline; clearF / clearI: 17 x 4 clear values; instanceIds / objectIds: per instance
// rows: NULL or [begin, end) -- a band of image rows (the bounded samples of bench.py's reference arm); the rows run on an OpenMP team.
EXPORT int ref_render(const RefCallbacks* cb, const float* f28, const uint* u16, const void* spheres, const void* distants, const void* rects, const void* disks,
                      float* rgba, uint64_t* rayCounts, void* const* aov, const float* clearF, const int* clearI, const int* instanceIds, const int* objectIds, const uint* rows)
{
  g_instanceIds = aov ? instanceIds : nullptr; g_objectIds = aov ? objectIds : nullptr;
  if (aov) for (int k = 0; k < 17; k++) { ClearValuesF[k] = vec4(clearF[4 * k], clearF[4 * k + 1], clearF[4 * k + 2], clearF[4 * k + 3]); ClearValuesI[k].x = clearI[4 * k]; ClearValuesI[k].y = clearI[4 * k + 1]; ClearValuesI[k].z = clearI[4 * k + 2]; ClearValuesI[k].w = clearI[4 * k + 3]; }
  else for (int k = 0; k < 17; k++) ClearValuesF[k] = vec4(0.0f);
  g_cb = *cb; g_totalRayCounts[0] = g_totalRayCounts[1] = 0;
  ubo.domeLightRotation = vec4(f28[0], f28[1], f28[2], f28[3]); ubo.domeLightEmissionMultiplier = ld3(f28 + 4); ubo.metersPerSceneUnit = f28[7];
  ubo.cameraPosition = ld3(f28 + 8); ubo.cameraForward = ld3(f28 + 11); ubo.cameraUp = ld3(f28 + 14); ubo.focusDistance = f28[17]; ubo.cameraVFoV = f28[18];
  ubo.lensRadius = f28[19]; ubo.invTotalSampleCount = f28[20]; ubo.maxSampleValue = f28[21]; ubo.rrInvMinTermProb = f28[22]; ubo.lightIntensityMultiplier = f28[23];
  ubo.sensorExposure = f28[24];
  ubo.sphereLightCount = u16[0]; ubo.distantLightCount = u16[1]; ubo.rectLightCount = u16[2]; ubo.diskLightCount = u16[3]; ubo.totalLightCount = u16[4];
  ubo.maxVolumeWalkLength = u16[5]; ubo.imageDims = u16[6]; ubo.sampleOffset = u16[7]; ubo.spp = u16[8]; ubo.maxBouncesAndRrBounceOffset = u16[9]; ubo.clipRangePacked = u16[10];
  memset(sphereLights, 0, sizeof(sphereLights)); memset(distantLights, 0, sizeof(distantLights)); memset(rectLights, 0, sizeof(rectLights)); memset(diskLights, 0, sizeof(diskLights));
  if (u16[0] > 16 || u16[1] > 16 || u16[2] > 16 || u16[3] > 16) return 0;
  memcpy(sphereLights, spheres, sizeof(SphereLight) * u16[0]); memcpy(distantLights, distants, sizeof(DistantLight) * u16[1]);
  memcpy(rectLights, rects, sizeof(RectLight) * u16[2]); memcpy(diskLights, disks, sizeof(DiskLight) * u16[3]);
  const uint width = u16[6] & 0xFFFFu, height = u16[6] >> 16, mss = u16[11], sw = u16[12];
  const uint r0 = rows ? rows[0] : 0u, r1 = rows ? (rows[1] < height ? rows[1] : height) : height;
  int ok = 1;
  if (aov) { if (sw == (1u | 2u | 8u | 32u | 64u) && mss == 0) pipe_aovs::render(width, height, r0, r1, (vec4*) rgba, aov); else ok = 0; }
  else if (sw == (1u | 2u | 8u | 32u | 64u) && mss == 0) pipe_default::render(width, height, r0, r1, (vec4*) rgba);
  else if (sw == (1u | 2u | 32u | 64u) && mss == 0) pipe_nonee::render(width, height, r0, r1, (vec4*) rgba);
  else if (sw == (1u | 2u | 8u | 32u | 64u) && mss == 2) pipe_media2::render(width, height, r0, r1, (vec4*) rgba);
  else if (sw == (4u | 8u | 16u) && mss == 0) pipe_other::render(width, height, r0, r1, (vec4*) rgba);
  else ok = 0;
  rayCounts[0] = g_totalRayCounts[0]; rayCounts[1] = g_totalRayCounts[1];
  return ok;
}
// ---- mdl_interface.glsl:40-205 ----
// mode 0: tex_lookup_float4_2d(coord.xy, wrap_u, wrap_v), 1: tex_lookup_float4_3d, 2: tex_texel_float4_2d(ivec2(coord)), 3: tex_texel_float4_3d; crop = (0, 1)
EXPORT void ref_tex_function(int mode, const void* desc, void (*sample)(const void*, float, float, float, int, float*), void (*texel)(const void*, int, int, int, float*),
                             int w, int h, int d, const float* coord, const int* icoord, int wrapU, int wrapV, int wrapW, float* out)
{
  texrt::g_tex.desc = desc; texrt::g_tex.sample = sample; texrt::g_tex.texel = texel; texrt::g_tex.w = w; texrt::g_tex.h = h; texrt::g_tex.d = d;
  const vec2 crop(0.0f, 1.0f);
  vec4 r;
  if (mode == 0) r = texrt::tex_lookup_float4_2d(1, vec2(coord[0], coord[1]), wrapU, wrapV, crop, crop, 0.0f);
  else if (mode == 1) r = texrt::tex_lookup_float4_3d(1, vec3(coord[0], coord[1], coord[2]), wrapU, wrapV, wrapW, crop, crop, crop, 0.0f);
  else if (mode == 2) r = texrt::tex_texel_float4_2d(1, ivec2(icoord[0], icoord[1]), ivec2(0, 0), 0.0f);
  else r = texrt::tex_texel_float4_3d(1, ivec3(icoord[0], icoord[1], icoord[2]), 0.0f);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
// ---- mdl_interface.glsl:266-452 ----
// One primvar reader at a hit.  word: the reference's packed sceneDataInfo (interpolation << 30 | (stride - 1) << 28 | offset 0) or SCENE_DATA_INVALID;
// data: the primvar's first float; count 1..3 floats -> scene_data_lookup_float / _float2 / _float3, isInt -> scene_data_lookup_int per component.
EXPORT void ref_scene_data_lookup(uint word, const void* data, const uint* hitIndices, float u, float v, int prim, int instanceId, int count, int isInt, const float* fallback, float* out)
{
  using namespace scenedata;
  scenedata::State st;
  st.renderer_state.hitIndices = uvec3(hitIndices[0], hitIndices[1], hitIndices[2]);
  st.renderer_state.hitBarycentrics = vec2(u, v);
  st.renderer_state.sceneDataBufferAddress = (uint64_t) data;
  for (int k = 0; k < 6; k++) st.renderer_state.sceneDataInfos[k] = SCENE_DATA_INVALID;
  st.renderer_state.sceneDataInfos[0] = word;
  int ids[1] = { instanceId }; scenedata::InstanceIds = ids; scenedata::gl_InstanceID = 0; gl_PrimitiveID = prim;
  if (isInt) { for (int c = 0; c < count; c++) out[c] = (float) scenedata::scene_data_lookup_int(st, 1, c, (int) fallback[c], false); return; }
  if (count == 1) out[0] = scenedata::scene_data_lookup_float(st, 1, fallback[0], false);
  else if (count == 2) { vec2 r = scenedata::scene_data_lookup_float2(st, 1, vec2(fallback[0], fallback[1]), false); out[0] = r.x; out[1] = r.y; }
  else { vec3 r = scenedata::scene_data_lookup_color(st, 1, vec3(fallback[0], fallback[1], fallback[2]), false); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
}
