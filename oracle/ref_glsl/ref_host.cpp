// TEST INFRASTRUCTURE.  The reference's HOST-side code for this path, compiled from src/gi/impl/Gi.cpp: the uniform block of giRender (:2375-2425,
// further down) and the light encoders (src/gi/impl/Gi.cpp:287-300 `_EncodeOctahedral` / `_EncodeDirection` and
// :2573-2941, every gi{Create,Set}{Sphere,Distant,Rect,Disk}Light* function) compiled from the reference's own file -- plain C++, so no
// rewriting, only the line ranges cut out by build_ref.py -- against stand-ins for what surrounds them in Gi.cpp: a GiScene holding four
// dense stores, glm (the handful of functions these lines use, written as glm writes them: normalize = v * (1 / sqrt(dot)), packUnorm2x16 =
// round(clamp * 65535)), and glm::packHalf2x16 routed to a caller-supplied function (glm's own half conversion is not in this image; the
// oracle's is used, so the half fields are equal by construction, everything else -- the Knud-Thomsen area, the cone's inverse pdf in
// double precision, the tangent-frame packing, defaults -- is the reference's arithmetic).
// Built WITHOUT -fsingle-precision-constant: these lines mix float and double (M_PI) on purpose.
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "glsl_shim.h"
#include "interface/rp_main.h"

namespace rp = gtl::shader_interface::rp_main;
static uint32_t (*g_packHalf2x16)(float, float);
namespace glm {
  static inline vec3 make_vec3(const float* p) { return vec3(p[0], p[1], p[2]); }
  static inline vec3 normalize(vec3 v) { return ::normalize(v); }
  static inline vec2 abs(vec2 v) { return ::abs(v); }
  static inline uint32_t packUnorm2x16(vec2 v) { return ::packUnorm2x16(v); }
  static inline uint32_t packHalf2x16(vec2 v) { return g_packHalf2x16(v.x, v.y); }
}
enum class GiSceneDirtyFlags : uint32_t { Clean = 0, DirtyFramebuffer = 1 };
static inline GiSceneDirtyFlags& operator|=(GiSceneDirtyFlags& a, GiSceneDirtyFlags b) { a = (GiSceneDirtyFlags) ((uint32_t) a | (uint32_t) b); return a; }
struct DenseStore {                                   // GgpuDenseDataStore: handles are element indices here
  size_t elementSize; std::vector<uint8_t> bytes;
  explicit DenseStore(size_t e) : elementSize(e) {}
  uint64_t allocate() { bytes.resize(bytes.size() + elementSize, 0xCD); return bytes.size() / elementSize - 1; }   // (fresh elements hold garbage, as a recycled GPU slot may)
  void free(uint64_t) {}
  template <typename T> T* write(uint64_t h) { return (T*) &bytes[h * elementSize]; }
};
struct GiScene {
  std::mutex mutex;
  DenseStore sphereLights{ sizeof(rp::SphereLight) }, distantLights{ sizeof(rp::DistantLight) }, rectLights{ sizeof(rp::RectLight) }, diskLights{ sizeof(rp::DiskLight) };
  GiSceneDirtyFlags dirtyFlags = GiSceneDirtyFlags::Clean;
};
struct GiSphereLight { GiScene* scene; uint64_t gpuHandle; };
struct GiDistantLight { GiScene* scene; uint64_t gpuHandle; };
struct GiRectLight { GiScene* scene; uint64_t gpuHandle; };
struct GiDiskLight { GiScene* scene; uint64_t gpuHandle; };

namespace gtl {
#include "gi_encode_direction.inc"      // Gi.cpp:287-300
#include "gi_lights.inc"                // Gi.cpp:2573-2941
}

// ---- Gi.cpp:2374-2423: the uniform block giRender fills for the shaders ----
// The reference's own public header supplies GiCameraDesc / GiRenderSettings / GiRenderParams; stand-ins for the rest of giRender's locals.
#include <gtl/gi/Gi.h>
namespace glm {
  struct quat { float x, y, z, w; quat() : x(0.0f), y(0.0f), z(0.0f), w(1.0f) {} float& operator[](int i) { return (&x)[i]; } };   // (x, y, z, w) in memory; make_quat reads the same order
  static inline vec4 make_vec4(const float* p) { return vec4(p[0], p[1], p[2], p[3]); }
}
namespace ubofill {
  struct DomeLight { glm::quat rotation; vec3 baseEmission; float diffuse = 1.0f, specular = 1.0f; };
  struct CountStore { uint32_t n = 0; uint32_t elementCount() const { return n; } };
  struct Scene { DomeLight* domeLight = nullptr; CountStore sphereLights, distantLights, rectLights, diskLights; uint32_t sampleOffset = 0; };
  struct ShaderCache { uint32_t maxTextureIndex = 0; };
  struct BumpAlloc { rp::UniformData storage; struct Slot { rp::UniformData* cpuPtr; uint32_t bufferOffset; }; template <typename T> Slot alloc() { return Slot{ &storage, 0u }; } };
  static rp::UniformData fill(const gtl::GiRenderParams& params, Scene* scene, uint32_t imageWidth, uint32_t imageHeight)
  {
    using namespace gtl;
    const GiRenderSettings& renderSettings = params.renderSettings;
    ShaderCache cache; ShaderCache* shaderCache = &cache;
    BumpAlloc bump; BumpAlloc* s_bumpAlloc = &bump;
    uint32_t uniformOffset;
    {
#include "gi_uniforms.inc"              // Gi.cpp:2375-2425
    }
    (void) uniformOffset;
    return bump.storage;
  }
}

#define EXPORT extern "C" __attribute__((visibility("default")))
// cam16: position(3) forward(3) up(3) vfov fStop focusDistance focalLength clipStart clipEnd exposure; set8: lightIntensityMultiplier maxSampleValue
// metersPerSceneUnit rrInvMinTermProb (floats) and, as integers, maxBounces maxVolumeWalkLength rrBounceOffset spp; counts4: lights per type;
// dome: NULL or rotation(4) baseEmission(3).  Output in the layout of orc_ubo_values (28 floats, 16 uints).
EXPORT void refhost_uniform_data(const float* cam16, const float* setF4, const uint32_t* setU4, const uint32_t* counts4, const float* dome7, uint32_t sampleOffset,
                                 uint32_t width, uint32_t height, float* f28, uint32_t* u16)
{
  gtl::GiRenderParams params{};
  memcpy(params.camera.position, cam16, 12); memcpy(params.camera.forward, cam16 + 3, 12); memcpy(params.camera.up, cam16 + 6, 12);
  params.camera.vfov = cam16[9]; params.camera.fStop = cam16[10]; params.camera.focusDistance = cam16[11]; params.camera.focalLength = cam16[12];
  params.camera.clipStart = cam16[13]; params.camera.clipEnd = cam16[14]; params.camera.exposure = cam16[15];
  gtl::GiRenderSettings& rs = params.renderSettings;
  rs.lightIntensityMultiplier = setF4[0]; rs.maxSampleValue = setF4[1]; rs.metersPerSceneUnit = setF4[2]; rs.rrInvMinTermProb = setF4[3];
  rs.maxBounces = setU4[0]; rs.maxVolumeWalkLength = setU4[1]; rs.rrBounceOffset = setU4[2]; rs.spp = setU4[3]; rs.frame = 0.0f; rs.time = 0.0f;
  ubofill::Scene scene; ubofill::DomeLight domeLight;
  scene.sphereLights.n = counts4[0]; scene.distantLights.n = counts4[1]; scene.rectLights.n = counts4[2]; scene.diskLights.n = counts4[3]; scene.sampleOffset = sampleOffset;
  if (dome7) { domeLight.rotation.x = dome7[0]; domeLight.rotation.y = dome7[1]; domeLight.rotation.z = dome7[2]; domeLight.rotation.w = dome7[3]; domeLight.baseEmission = vec3(dome7[4], dome7[5], dome7[6]); scene.domeLight = &domeLight; }
  const rp::UniformData u = ubofill::fill(params, &scene, width, height);
  const float f[28] = { u.domeLightRotation.x, u.domeLightRotation.y, u.domeLightRotation.z, u.domeLightRotation.w, u.domeLightEmissionMultiplier.x, u.domeLightEmissionMultiplier.y,
                        u.domeLightEmissionMultiplier.z, u.metersPerSceneUnit, u.cameraPosition.x, u.cameraPosition.y, u.cameraPosition.z, u.cameraForward.x, u.cameraForward.y, u.cameraForward.z,
                        u.cameraUp.x, u.cameraUp.y, u.cameraUp.z, u.focusDistance, u.cameraVFoV, u.lensRadius, u.invTotalSampleCount, u.maxSampleValue, u.rrInvMinTermProb,
                        u.lightIntensityMultiplier, u.sensorExposure, 0.0f, 0.0f, 0.0f };
  const uint32_t v[16] = { u.sphereLightCount, u.distantLightCount, u.rectLightCount, u.diskLightCount, u.totalLightCount, u.maxVolumeWalkLength, u.imageDims, u.sampleOffset, u.spp,
                           u.maxBouncesAndRrBounceOffset, u.clipRangePacked, 0u, 0u, 0u, 0u, 0u };
  memcpy(f28, f, sizeof(f)); memcpy(u16, v, sizeof(v));
}
EXPORT void refhost_set_pack_half(uint32_t (*fn)(float, float)) { g_packHalf2x16 = fn; }
EXPORT uint32_t refhost_encode_direction(const float* v) { return gtl::_EncodeDirection(vec3(v[0], v[1], v[2])); }
// Each function creates the light and applies the setters in the order hdGatling's light adapters call them (src/hdGatling/light.cpp), then
// copies out the 48-byte GPU record.  `set` = 0 leaves the light as giCreate*Light made it (defaults).
EXPORT void refhost_sphere_light(int set, float* pos, float* rgb, const float* radius, float diffuse, float specular, void* out48)
{
  GiScene s; GiSphereLight* l = gtl::giCreateSphereLight(&s);
  if (set) { gtl::giSetSphereLightPosition(l, pos); gtl::giSetSphereLightBaseEmission(l, rgb); gtl::giSetSphereLightRadius(l, radius[0], radius[1], radius[2]); gtl::giSetSphereLightDiffuseSpecular(l, diffuse, specular); }
  memcpy(out48, s.sphereLights.write<rp::SphereLight>(l->gpuHandle), 48); gtl::giDestroySphereLight(&s, l);
}
EXPORT void refhost_distant_light(int set, float* direction, float* rgb, float angle, float diffuse, float specular, void* out48)
{
  GiScene s; GiDistantLight* l = gtl::giCreateDistantLight(&s);
  if (set) { gtl::giSetDistantLightDirection(l, direction); gtl::giSetDistantLightBaseEmission(l, rgb); gtl::giSetDistantLightAngle(l, angle); gtl::giSetDistantLightDiffuseSpecular(l, diffuse, specular); }
  memcpy(out48, s.distantLights.write<rp::DistantLight>(l->gpuHandle), 48); gtl::giDestroyDistantLight(&s, l);
}
EXPORT void refhost_rect_light(int set, float* origin, float* t0, float* t1, float* rgb, float width, float height, float diffuse, float specular, void* out48)
{
  GiScene s; GiRectLight* l = gtl::giCreateRectLight(&s);
  if (set) { gtl::giSetRectLightOrigin(l, origin); gtl::giSetRectLightTangents(l, t0, t1); gtl::giSetRectLightBaseEmission(l, rgb); gtl::giSetRectLightDimensions(l, width, height); gtl::giSetRectLightDiffuseSpecular(l, diffuse, specular); }
  memcpy(out48, s.rectLights.write<rp::RectLight>(l->gpuHandle), 48); gtl::giDestroyRectLight(&s, l);
}
EXPORT void refhost_disk_light(int set, float* origin, float* t0, float* t1, float* rgb, float radiusX, float radiusY, float diffuse, float specular, void* out48)
{
  GiScene s; GiDiskLight* l = gtl::giCreateDiskLight(&s);
  if (set) { gtl::giSetDiskLightOrigin(l, origin); gtl::giSetDiskLightTangents(l, t0, t1); gtl::giSetDiskLightBaseEmission(l, rgb); gtl::giSetDiskLightRadius(l, radiusX, radiusY); gtl::giSetDiskLightDiffuseSpecular(l, diffuse, specular); }
  memcpy(out48, s.diskLights.write<rp::DiskLight>(l->gpuHandle), 48); gtl::giDestroyDiskLight(&s, l);
}

// ---- src/hdGatling/main.cpp:136-143, 465-487: how the reference's graphical tests turn a float AOV into the RGBA8 image they compare ----
#include <algorithm>
#include <cmath>
namespace harness {
#include "harness_srgb.inc"
static void to_bytes(const float* mappedMem, size_t width, size_t height, int compCount, bool isInt, bool gammaEncode, uint8_t* outBytes)
{
    size_t pixelCount = width * height;
    std::vector<uint8_t> byteValues(pixelCount * 4);
#include "harness_bytes.inc"
    memcpy(outBytes, byteValues.data(), byteValues.size());
}
}
EXPORT float refhost_accurate_linear_to_srgb(float x) { return harness::_AccurateLinearToSrgb(x); }
EXPORT void refhost_harness_bytes(const float* pixels, uint32_t width, uint32_t height, int compCount, int isInt, int gammaEncode, uint8_t* out) { harness::to_bytes(pixels, width, height, compCount, isInt != 0, gammaEncode != 0, out); }

// ---- Gi.cpp:845-859: how a mesh's vertices become the rp::FVertex records the shaders read ----
namespace gtl {
static void pack_vertices(const std::vector<GiVertex>& meshVertices, std::vector<rp::FVertex>& vertexData)
{
    std::vector<float> positionData;
    vertexData.resize(meshVertices.size());
    positionData.resize(meshVertices.size() * 3);
#include "gi_pack_vertices.inc"
}
}
EXPORT void refhost_pack_vertices(const void* giVertices, uint32_t count, void* fvertsOut)
{
  std::vector<gtl::GiVertex> v((const gtl::GiVertex*) giVertices, (const gtl::GiVertex*) giVertices + count);
  std::vector<rp::FVertex> out;
  gtl::pack_vertices(v, out);
  memcpy(fvertsOut, out.data(), sizeof(rp::FVertex) * count);
}

// ---- src/hdGatling/mesh.cpp:83-86, 231-261: the tangent frame hdGatling gives a mesh without authored tangents or texture coordinates ----
// (the caller's side of giCreateMesh; gatling_b200/scenes.py prepares its test geometry the same way).  Stand-ins for the few pxr types.
namespace meshprep {
struct GfVec3f { float v[3]; GfVec3f() : v{ 0.0f, 0.0f, 0.0f } {} GfVec3f(float a, float b, float c) : v{ a, b, c } {} float operator[](int i) const { return v[i]; } float& operator[](int i) { return v[i]; } };
static inline float GfDot(const GfVec3f& a, const GfVec3f& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }          // GfVec3f::operator*: x x + y y + z z, left to right
static inline GfVec3f GfCross(const GfVec3f& a, const GfVec3f& b) { return GfVec3f(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }
typedef std::vector<GfVec3f> VtVec3fArray;
typedef std::vector<float> VtFloatArray;
#include "mesh_bitangent_sign.inc"
#include "mesh_fallback_tangents.inc"
}
EXPORT void refhost_fallback_tangents(const float* normals, uint32_t count, float* tangents, float* bitangentSigns)
{
  meshprep::VtVec3fArray n(count), t; meshprep::VtFloatArray s;
  for (uint32_t i = 0; i < count; i++) n[i] = meshprep::GfVec3f(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
  meshprep::_CalculateFallbackTangents(n, t, s);
  for (uint32_t i = 0; i < count; i++) { tangents[3 * i] = t[i][0]; tangents[3 * i + 1] = t[i][1]; tangents[3 * i + 2] = t[i][2]; bitangentSigns[i] = s[i]; }
}
