// TEST INFRASTRUCTURE.  The reference's HOST-side light encoders (src/gi/impl/Gi.cpp:287-300 `_EncodeOctahedral` / `_EncodeDirection` and
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

#define EXPORT extern "C" __attribute__((visibility("default")))
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
