// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's path-tracing hot path, scalar C++ with
// OpenMP over image rows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (gatling_b200/) never does.
//
// PARITY STATUS: PARTLY PINNED.  The reference as a whole cannot be built or run in this environment and its
// golden images are Git-LFS stubs.  Pinned to the reference itself are (1) every PURE function
// of its shader sources -- RNG and hashes, sampling transforms, direction / half packing, ray offsetting,
// sampleLight, the volume / roulette helpers, the payload bit field, the dome-light mapping, texture wrapping,
// mdl_adapt_normal, the whole of mdl_shading_state.glsl and the camera-ray head of rp_main.rgen's main() --
// which oracle/ref_glsl/build_ref.py compiles AS C++ FROM THE REFERENCE'S OWN FILES into oracle/_ref/ and
// tests/test_reference_glsl_cpu.py compares bit for bit with the functions below; (1b) the main() functions of
// rp_main.rgen / .chit / .miss / _shadow.miss themselves, compiled the same way into a CPU renderer that calls
// back here only for ray queries, the material's mdl_* entry points and the dome texture: its images equal
// orc_render's bit for bit on every test scene (test_reference_integrator_*); and (2) the one golden image
// small enough to reconstruct from its LFS pointer (tests/test_reference_pin.py).  UNPINNED remain the any-hit
// decision (ORACLE_DEVIATIONS #10), the ClockCycles / FaceId AOVs and the BSDF arithmetic (closed MDL SDK); the
// other fifteen AOVs are compared as well (test_reference_integrator_auxiliary_aovs).
// The file follows the reference *source text* line by line where the source is in
// tree (integrator: rp_main.rgen / rp_main.chit / rp_main.miss / rp_main_shadow.miss, common.glsl,
// mdl_shading_state.glsl, Gi.cpp host encoders) and restates published definitions where the
// arithmetic lives in third-party binaries (BSDF bodies: orc_bsdf.h; traversal: orc_scene.h).
// Deviations and free choices: ORACLE_DEVIATIONS.md.
//
// Each function cites the reference lines it follows as file:line relative to
// /root/reference/src/gi/shaders/ unless a longer path is given.
#include "orc_scene.h"
#include "orc_openpbr.h"
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

// rp_main_payload.glsl:3-9
static const uint32_t PAYLOAD_VOLUME_WALK_MISS_FLAG = 0x40000000u;
static const uint32_t PAYLOAD_WALK_OFFSET = 12u;
static const uint32_t PAYLOAD_MEDIUM_IDX_MASK = 0x0f000000u;
static const uint32_t PAYLOAD_MEDIUM_IDX_OFFSET = 24;
static const uint32_t PAYLOAD_WALK_MASK = 0x00fff000u;
static const uint32_t PAYLOAD_BOUNCES_MASK = 0x00000fffu;
static const uint32_t PAYLOAD_TERMINATE_FLAG = 0x80000000u;

// rp_main_payload.glsl:20-47 (MEDIUM_STACK_SIZE == 0)
struct Medium { V3 ior, sigma_s, sigma_t; float bias; };                 // rp_main_payload.glsl:11-17
static const uint32_t MAX_MEDIUM_STACK = 4;                               // GtlcRenderSettings.mediumStackSize is limited to 4 here
struct ShadeRayPayload {
  V3 throughput; uint32_t bitfield; V3 radiance; uint32_t rng_state;
  Medium media[MAX_MEDIUM_STACK]; V3 walkSegmentPdf;                      // MEDIUM_STACK_SIZE > 0 (:37-40)
  V3 ray_origin, ray_dir, neeToLight, neeContrib; uint32_t pixelIndex;
};

// rp_main_payload.glsl:76-89: maxMediumIdx = max(1, MEDIUM_STACK_SIZE)
static inline uint32_t payloadGetMediumIdx(const ShadeRayPayload& p, uint32_t mediumStackSize)
{
  uint32_t b = (p.bitfield & PAYLOAD_MEDIUM_IDX_MASK) >> PAYLOAD_MEDIUM_IDX_OFFSET;
  uint32_t maxMediumIdx = mediumStackSize > 1u ? mediumStackSize : 1u;
  return b < maxMediumIdx ? b : maxMediumIdx;
}
// rp_main_payload.glsl:60-74
static inline void payloadIncrementWalk(ShadeRayPayload& p)
{
  uint32_t b = p.bitfield & PAYLOAD_WALK_MASK;
  b = (b + 1u < PAYLOAD_WALK_MASK) ? b + 1u : PAYLOAD_WALK_MASK;          // (sic) the reference adds 1 to the masked, unshifted field
  p.bitfield &= ~PAYLOAD_WALK_MASK;
  p.bitfield |= b;
}
static inline uint32_t payloadGetWalk(const ShadeRayPayload& p) { return (p.bitfield & PAYLOAD_WALK_MASK) >> PAYLOAD_WALK_OFFSET; }
// rp_main_payload.glsl:91-95
static inline void payloadSetMediumIdx(ShadeRayPayload& p, uint32_t idx)
{
  p.bitfield &= ~PAYLOAD_MEDIUM_IDX_MASK;
  uint32_t v = idx << PAYLOAD_MEDIUM_IDX_OFFSET;
  p.bitfield |= (v < PAYLOAD_MEDIUM_IDX_MASK ? v : PAYLOAD_MEDIUM_IDX_MASK);
}

// interface/rp_main.h:25-56 (fields the CPU path reads), filled like Gi.cpp:2391-2423
struct Ubo {
  float domeLightRotation[4]; V3 domeLightEmissionMultiplier;
  uint32_t sphereLightCount, distantLightCount, rectLightCount, diskLightCount, totalLightCount;
  float metersPerSceneUnit;
  V3 cameraPosition; uint32_t imageDims; V3 cameraForward; float focusDistance; V3 cameraUp; float cameraVFoV;
  uint32_t sampleOffset; float lensRadius; uint32_t spp; float invTotalSampleCount; float maxSampleValue;
  uint32_t maxBouncesAndRrBounceOffset; float rrInvMinTermProb; float lightIntensityMultiplier;
  uint32_t clipRangePacked; float sensorExposure; uint32_t maxVolumeWalkLength;
};

struct Flags { // compile-time #defines of the reference, GlslShaderGen.cpp:57-72,95-133
  bool jittered, fis, dof, nee, clipping, progressive, domeCameraVisible;
};

struct RenderCtx { const Scene* s; Ubo ubo; Flags fl; V4 background; uint32_t aovMask; uint32_t mediumStackSize; };

// What the closest-hit shader writes into the auxiliary AOVs on the first bounce (rp_main.chit:191-290),
// captured per sample so that the per-pixel write order of the megakernel can be replayed.
struct AovSample {
  bool hit; uint32_t bounces;
  V3 bary, texcoords, tangent, bitangent, posNormal, albedo;
  bool thinWalled, doubleSided;
  int32_t objectId, faceId, instanceId; float depth;
  float cutoutOpacity;   // < 0: the material has no cutout (chit:202), else ahit:48
  bool neeRan, neeShadowed;   // rgen:431-436 (first bounce, NEXT_EVENT_ESTIMATION builds only)
};

// colormap.glsl:42-53
static inline V3 colormap_inferno(float t)
{
  const V3 c0 = v3(0.0002189403691192265f, 0.001651004631001012f, -0.01948089843709184f);
  const V3 c1 = v3(0.1065134194856116f, 0.5639564367884091f, 3.932712388889277f);
  const V3 c2 = v3(11.60249308247187f, -3.972853965665698f, -15.9423941062914f);
  const V3 c3 = v3(-41.70399613139459f, 17.43639888205313f, 44.35414519872813f);
  const V3 c4 = v3(77.162935699427f, -33.40235894210092f, -81.80730925738993f);
  const V3 c5 = v3(-71.31942824499214f, 32.62606426397723f, 73.20951985803202f);
  const V3 c6 = v3(25.13112622477341f, -12.24266895238567f, -23.07032500287172f);
  return c0 + (c1 + (c2 + (c3 + (c4 + (c5 + c6 * t) * t) * t) * t) * t) * t;
}

// colormap.glsl:3-14
static inline V3 colormap_viridis(float t)
{
  const V3 c0 = v3(0.2777273272234177f, 0.005407344544966578f, 0.3340998053353061f);
  const V3 c1 = v3(0.1050930431085774f, 1.404613529898575f, 1.384590162594685f);
  const V3 c2 = v3(-0.3308618287255563f, 0.214847559468213f, 0.09509516302823659f);
  const V3 c3 = v3(-4.634230498983486f, -5.799100973351585f, -19.33244095627987f);
  const V3 c4 = v3(6.228269936347081f, 14.17993336680509f, 56.69055260068105f);
  const V3 c5 = v3(4.776384997670288f, -13.74514537774601f, -65.35303263337234f);
  const V3 c6 = v3(-5.435455855934631f, 4.645852612178535f, 26.3124352495832f);
  return c0 + (c1 + (c2 + (c3 + (c4 + (c5 + c6 * t) * t) * t) * t) * t) * t;
}

// rp_main.rgen:118-130
static inline V2 fisGauss(V2 xi)
{
  float u1 = fmaxf(1e-38f, xi.x);
  float u2 = xi.y;
  float sigma = 0.375f;
  float r = sigma * sqrtf(-2.0f * logf(u1));
  float phi = 2.0f * PI * u2;
  V2 o; o.x = cosf(phi) * r; o.y = sinf(phi) * r;
  return o;
}

// rp_main.rgen:99-112
static inline bool russian_roulette(float random_float, V3& throughput, float rrInvMinTermProb)
{
  float max_throughput = fmaxf(throughput.x, fmaxf(throughput.y, throughput.z));
  float p = fminf(max_throughput, rrInvMinTermProb);
  if (random_float > p) return true;
  throughput = throughput / p;
  return false;
}

// rp_main.rgen:191-271: primary ray for pixel (x, y), sample s.  Draws 2 (+2 with DoF) numbers.
static inline void camera_ray(const Ubo& ubo, const Flags& fl, uint32_t px, uint32_t py, uint32_t& rng_state, V3& ray_origin, V3& ray_dir)
{
  uint32_t imageWidth = ubo.imageDims & 0xFFFFu;
  uint32_t imageHeight = ubo.imageDims >> 16;
  V3 camera_right = cross(ubo.cameraForward, ubo.cameraUp);              // :199
  float aspect_ratio = (float) imageWidth / (float) imageHeight;         // :200
  float H = 1.0f;
  float W = H * aspect_ratio;
  float d = H / (2.0f * tanf(ubo.cameraVFoV * 0.5f));                    // :204
  float WX = W / (float) imageWidth;
  float HY = H / (float) imageHeight;
  V3 C = ubo.cameraPosition + ubo.cameraForward * d;                     // :209
  V3 L = (C - (camera_right * W) * 0.5f) - (ubo.cameraUp * H) * 0.5f;    // :210

  V2 rand2_xy = rng1d_next2f(rng_state);                                 // :224
  V2 sampleOffset; sampleOffset.x = 0.5f; sampleOffset.y = 0.5f;         // :227
  if (fl.jittered) {
    if (fl.fis) { V2 g = fisGauss(rand2_xy); sampleOffset.x = 0.5f + g.x; sampleOffset.y = 0.5f + g.y; } // :232
    else sampleOffset = rand2_xy;                                        // :235
  }
  V3 P = (L + (camera_right * ((float) px + sampleOffset.x)) * WX) + (ubo.cameraUp * ((float) py + sampleOffset.y)) * HY; // :239-242
  ray_origin = ubo.cameraPosition;
  ray_dir = normalize(P - ray_origin);                                   // :245
  if (fl.dof && ubo.lensRadius > 0.0f) {                                 // :250-267
    V2 rand2_zw = rng1d_next2f(rng_state);
    V3 focalPoint = ray_origin + ray_dir * ubo.focusDistance;
    V3 h = sample_hemisphere(rand2_zw);
    float ax = h.x * ubo.lensRadius, ay = h.y * ubo.lensRadius;
    ray_origin += camera_right * ax;
    ray_origin += ubo.cameraUp * ay;
    ray_dir = normalize(focalPoint - ray_origin);
  }
  ray_dir.x += (ray_dir.x == 0.0f) ? FLOAT_MIN : 0.0f;                   // :271
  ray_dir.y += (ray_dir.y == 0.0f) ? FLOAT_MIN : 0.0f;
  ray_dir.z += (ray_dir.z == 0.0f) ? FLOAT_MIN : 0.0f;
}

// rp_main.rgen:285-288: the clip planes are perpendicular to the camera axis; along this ray they lie further out
static inline V2 clip_range_along(const Ubo& ubo, V3 ray_dir)
{
  float cosConeAngle = fmaxf(1e-5f, dot(ray_dir, ubo.cameraForward));     // :287
  V2 clipRange = unpackHalf2x16(ubo.clipRangePacked);
  clipRange.x = clipRange.x / cosConeAngle; clipRange.y = clipRange.y / cosConeAngle; // :288
  return clipRange;
}

// rp_main.miss:39-44
static inline V3 quatRotateDir(const float* q, V3 dir)
{
  V3 qv = v3(q[0], q[1], q[2]);
  V3 a = cross(qv, dir);
  V3 b = cross(qv, a);
  return dir + ((a * q[3]) + b) * 2.0f;
}

// Bilinear lat-long lookup with sampler state linear/linear, REPEAT (Gi.cpp:388-392), unnormalised
// texel centres at +0.5 (Vulkan spec 16.x "Texel Filtering").
static inline V3 texture_latlong(const Scene& s, float u, float v)
{
  int W = (int) s.domeW, H = (int) s.domeH;
  float x = u * (float) W - 0.5f, y = v * (float) H - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float ax = x - fx, ay = y - fy;
  int x0 = (int) fx, y0 = (int) fy;
  auto wrap = [](int i, int n) { int m = i % n; return m < 0 ? m + n : m; };
  int xa = wrap(x0, W), xb = wrap(x0 + 1, W), ya = wrap(y0, H), yb = wrap(y0 + 1, H);
  auto tex = [&](int xx, int yy) { const float* p = &s.domeTexels[((size_t) yy * W + xx) * 4]; return v3(p[0], p[1], p[2]); };
  V3 top = tex(xa, ya) * (1.0f - ax) + tex(xb, ya) * ax;
  V3 bot = tex(xa, yb) * (1.0f - ax) + tex(xb, yb) * ax;
  return top * (1.0f - ay) + bot * ay;
}

// rp_main.miss:48-49
static inline void dome_latlong_uv(V3 rayDir, float& u, float& v)
{
  u = (atan2f(rayDir.z, rayDir.x) + 0.5f * PI) / (2.0f * PI);
  v = 1.0f - acosf(rayDir.y) / PI;
}

// rp_main.miss:46-53 + texture binding logic Gi.cpp:2183-2199,2334-2337
static inline V3 sampleDomeLight(const RenderCtx& c, uint32_t domeLightIndex, V3 rayDir)
{
  bool haveDome = !c.s->domeTexels.empty();
  if (domeLightIndex == 0 || !haveDome) {
    // 1x1 RGBA8 fallback holding the colour-AOV clear value: glm::u8vec4(backgroundColor * 255.0f)
    V3 bg;
    bg.x = (float)(uint8_t)(c.background.x * 255.0f) / 255.0f;
    bg.y = (float)(uint8_t)(c.background.y * 255.0f) / 255.0f;
    bg.z = (float)(uint8_t)(c.background.z * 255.0f) / 255.0f;
    return bg;
  }
  float u, v;
  dome_latlong_uv(rayDir, u, v);
  return texture_latlong(*c.s, u, v);
}

// rp_main.miss:55-87 (MEDIUM_STACK_SIZE == 0, colour AOV bound)
// rp_main.miss:16-34
static inline void stepVolume(const RenderCtx& c, ShadeRayPayload& p, uint32_t mediumIdx, V3 worldRayDir, float rayTmax)
{
  float distance = rayTmax * c.ubo.metersPerSceneUnit;
  const Medium& m = p.media[mediumIdx - 1];
  V3 transmittance = v3(expf(m.sigma_t.x * -distance), expf(m.sigma_t.y * -distance), expf(m.sigma_t.z * -distance));
  V3 density = m.sigma_t * transmittance;
  float pdf = dot(p.walkSegmentPdf, density);
  p.throughput *= (m.sigma_s * transmittance) / pdf;
  p.ray_origin += worldRayDir * distance;
  p.bitfield |= PAYLOAD_VOLUME_WALK_MISS_FLAG;
  payloadIncrementWalk(p);
}

static inline void miss_shader(const RenderCtx& c, ShadeRayPayload& p, V3 worldRayDir, float rayTmax)
{
  if (c.mediumStackSize > 0) {                                           // :57-66
    uint32_t mediumIdx = payloadGetMediumIdx(p, c.mediumStackSize);
    if (mediumIdx > 0) { stepVolume(c, p, mediumIdx, worldRayDir, rayTmax); return; }
  }
  p.bitfield |= PAYLOAD_TERMINATE_FLAG;                                  // :68
  bool useFallbackDomeLight = false;
  if (!c.fl.domeCameraVisible) {                                         // :76-80
    bool isPrimaryRay = (p.bitfield & PAYLOAD_BOUNCES_MASK) == 0;
    useFallbackDomeLight = isPrimaryRay;
  }
  uint32_t domeLightIndex = useFallbackDomeLight ? 0u : 1u;              // :82
  V3 sampleDir = normalize(quatRotateDir(c.ubo.domeLightRotation, worldRayDir)); // :83
  V3 radiance = sampleDomeLight(c, domeLightIndex, sampleDir) * c.ubo.domeLightEmissionMultiplier; // :84
  p.radiance += p.throughput * radiance;                                 // :86
}

// rp_main.chit:30-129
static inline void sampleLight(const RenderCtx& c, V4 k4, V3 surfacePos, V3& dirToLight, float& dist, V3& power, float& invPdf, uint32_t& diffuseSpecularPacked)
{
  const Ubo& ubo = c.ubo; const Scene& s = *c.s;
  static const SphereLight zeroSphere = {};
  float total = (float) ubo.totalLightCount;
  if ((k4.x * total) <= (float) ubo.sphereLightCount) {                                        // :32
    uint32_t n = ubo.sphereLightCount;
    uint32_t lightIndex = (uint32_t)(k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u; // :34 (n==0 wraps to UINT_MAX)
    const SphereLight& light = (n == 0 || lightIndex >= s.sphereLights.size()) ? zeroSphere : s.sphereLights[lightIndex];
    V3 lpos = v3(light.pos[0], light.pos[1], light.pos[2]);
    V2 zw; zw.x = k4.z; zw.y = k4.w;
    V3 samplePos = lpos + sample_sphere(zw, v3(light.radiusXYZ[0], light.radiusXYZ[1], light.radiusXYZ[2])); // :40
    V3 dir = samplePos - surfacePos;
    dist = length(dir);
    dirToLight = safe_div(dir, dist);
    V3 lightNormal = normalize(samplePos - lpos);                                              // :46
    float cosTheta = fmaxf(0.0f, dot(-dirToLight, lightNormal));
    invPdf = safe_div((light.area > 0.0f) ? (light.area * cosTheta) : 1.0f, dist * dist);    // :48
    power = v3(light.baseEmission[0], light.baseEmission[1], light.baseEmission[2]) * ubo.lightIntensityMultiplier;
    diffuseSpecularPacked = light.diffuseSpecularPacked;
  } else if ((k4.x * total) <= (float)(ubo.sphereLightCount + ubo.distantLightCount)) {        // :53
    uint32_t n = ubo.distantLightCount;
    uint32_t lightIndex = (uint32_t)(k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u;
    const DistantLight& light = s.distantLights[lightIndex];
    dist = 100000.0f;
    dirToLight = -v3(light.direction[0], light.direction[1], light.direction[2]);
    power = v3(light.baseEmission[0], light.baseEmission[1], light.baseEmission[2]) * ubo.lightIntensityMultiplier;
    invPdf = light.invPdf;
    diffuseSpecularPacked = light.diffuseSpecularPacked;
    if (light.angle > 0.0f) {                                                                  // :65-74
      V3 t1, t2;
      orthonormal_basis(dirToLight, t1, t2);
      float phi = ((k4.z * 2.0f) * PI) - PI;
      float theta = k4.w * light.angle;
      dirToLight = normalize(((t1 * cosf(phi) + t2 * sinf(phi)) * sinf(theta)) + dirToLight * cosf(theta));
    }
  } else if ((k4.x * total) <= (float)(ubo.sphereLightCount + ubo.distantLightCount + ubo.rectLightCount)) { // :76
    uint32_t n = ubo.rectLightCount;
    uint32_t lightIndex = (uint32_t)(k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u;
    const RectLight& light = s.rectLights[lightIndex];
    float sx = (k4.z - 0.5f) * light.width, sy = (k4.w - 0.5f) * light.height;                 // :83
    V3 t0 = decode_direction(light.tangentFramePacked[0]);
    V3 t1 = decode_direction(light.tangentFramePacked[1]);
    V3 samplePos = (v3(light.origin[0], light.origin[1], light.origin[2]) + t0 * sx) + t1 * sy; // :87
    V3 dir = samplePos - surfacePos;
    dist = length(dir);
    dirToLight = safe_div(dir, dist);
    V3 lightNormal = cross(t1, t0);                                                            // :93
    float cosTheta = fmaxf(0.0f, dot(-dirToLight, lightNormal));
    float area = light.width * light.height;
    invPdf = safe_div((area > 0.0f) ? (area * cosTheta) : 1.0f, dist * dist);
    power = v3(light.baseEmission[0], light.baseEmission[1], light.baseEmission[2]) * ubo.lightIntensityMultiplier;
    diffuseSpecularPacked = light.diffuseSpecularPacked;
  } else {                                                                                     // :101
    uint32_t n = ubo.diskLightCount;
    uint32_t lightIndex = (uint32_t)(k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u;
    const DiskLight& light = s.diskLights[lightIndex];
    V2 radiusXY; radiusXY.x = light.radiusX; radiusXY.y = light.radiusY;
    V2 zw; zw.x = k4.z; zw.y = k4.w;
    V2 sampleOnDisk = sample_disk(zw, radiusXY);
    V3 t0 = decode_direction(light.tangentFramePacked[0]);
    V3 t1 = decode_direction(light.tangentFramePacked[1]);
    V3 samplePos = (v3(light.origin[0], light.origin[1], light.origin[2]) + t0 * sampleOnDisk.x) + t1 * sampleOnDisk.y;
    V3 dir = samplePos - surfacePos;
    dist = length(dir);
    dirToLight = safe_div(dir, dist);
    V3 lightNormal = cross(t1, t0);
    float cosTheta = fmaxf(0.0f, dot(-dirToLight, lightNormal));
    float area = (radiusXY.x * radiusXY.y) * PI;
    invPdf = safe_div((area > 0.0f) ? (area * cosTheta) : 1.0f, dist * dist);
    power = v3(light.baseEmission[0], light.baseEmission[1], light.baseEmission[2]) * ubo.lightIntensityMultiplier;
    diffuseSpecularPacked = light.diffuseSpecularPacked;
  }
  power *= exp2f(ubo.sensorExposure);                                                          // :127
  invPdf *= (float) ubo.totalLightCount;                                                       // :128
}

// R15: textured / primvar-driven inputs are evaluated once per hit into a private copy of the parameter block (the reference
// re-evaluates them inside every generated mdl_* function; same values).  Returns the mapped shading normal when a normal map is bound.
static inline void material_at_hit(const Scene& s, const InstanceDesc& in, const Hit& hit, const State& st, V3 worldRayDir, const Material& base,
                                   Material& mat, V3& mappedNormal, bool& hasMappedNormal)
{
  mat = base;
  mappedNormal = st.normal; hasMappedNormal = false;
  if (!s.materialPrimvars.empty() && !s.materialPrimvars[in.material].empty() && !s.primvarInfos.empty()) {   // scene_data_lookup_float{,2,3}
    const std::vector<GtlcPrimvarBinding>& pbs = s.materialPrimvars[in.material];
    const V3 bc = v3((1.0f - hit.u) - hit.v, hit.u, hit.v);                                     // mdl_interface.glsl:318
    const uint32_t* f = &s.faces[(size_t) (in.faceBase + hit.prim) * 3];
    for (size_t k = 0; k < pbs.size(); k++) {
      const uint32_t* info = &s.primvarInfos[((size_t) hit.inst * GTLC_MAX_PRIMVAR_BINDINGS + k) * 2];
      if (!(info[1] & 0x100u)) { for (int cidx = 0; cidx < pbs[k].componentCount; cidx++) mat.p[pbs[k].paramIndex + cidx] = pbs[k].fallback[cidx]; continue; }
      const uint32_t stride = info[1] & 15u, interpolation = (info[1] >> 4) & 15u;
      uint32_t idx[3] = { f[0], f[1], f[2] };                                                   // :281-301
      if (interpolation == GTLC_INTERP_UNIFORM) idx[0] = idx[1] = idx[2] = hit.prim;
      else if (interpolation == GTLC_INTERP_INSTANCE) idx[0] = idx[1] = idx[2] = in.instanceId > 0 ? (uint32_t) in.instanceId : 0u;
      else if (interpolation == GTLC_INTERP_CONSTANT) idx[0] = idx[1] = idx[2] = 0u;
      const float* dta = &s.primvarData[info[0]];
      if (info[1] & 0x200u) {                                                                   // scene_data_lookup_int{,2,3} :417-479: nearest vertex, no interpolation
        const uint32_t pick = bc.x > bc.y ? (bc.x > bc.z ? idx[0] : idx[2]) : (bc.y > bc.z ? idx[1] : idx[2]);
        for (int cidx = 0; cidx < pbs[k].componentCount; cidx++)
          mat.p[pbs[k].paramIndex + cidx] = cidx < (int) stride ? (float) (int32_t) floatBitsToUint(dta[(size_t) pick * stride + cidx]) : pbs[k].fallback[cidx];
        continue;
      }
      for (int cidx = 0; cidx < pbs[k].componentCount; cidx++) {
        float v[3];
        for (int q = 0; q < 3; q++) v[q] = cidx < (int) stride ? dta[(size_t) idx[q] * stride + cidx] : 0.0f;
        mat.p[pbs[k].paramIndex + cidx] = (v[0] * bc.x + v[1] * bc.y) + v[2] * bc.z;
      }
    }
  }
  if (!s.materialTextures.empty() && !s.materialTextures[in.material].empty())
    hasMappedNormal = apply_texture_bindings(s.textures, s.materialTextures[in.material], st.text_coords.x, st.text_coords.y, st.local_position, worldRayDir,
                                             st.geom_normal, st.normal, st.tangent_u, st.tangent_v, mat.p, mappedNormal);

}

// rp_main.chit:132-493 (MEDIUM_STACK_SIZE == 0; AOV section :191-290 handled by the caller for Color only)
static inline void closest_hit_shader(const RenderCtx& c, ShadeRayPayload& p, const Hit& hit, V3 worldRayDir, AovSample* aov)
{
  const Scene& s = *c.s;
  const InstanceDesc& in = s.instances[hit.inst];
  bool isLeftHanded = (in.flags & 1u) != 0;                               // :137
  bool isDoubleSided = (in.flags & 2u) != 0;                              // :138
  const Material& base = s.materials[in.material];

  State st; bool isFrontFace;
  setup_shading_state(s, hit, worldRayDir, st, isFrontFace);              // :143
  const V3 normal = st.normal;                                            // :147
  Material mat; V3 mappedNormal; bool hasMappedNormal;
  material_at_hit(s, in, hit, st, worldRayDir, base, mat, mappedNormal, hasMappedNormal);

  V3 throughput = p.throughput;
  V3 radiance = p.radiance;
  uint32_t mediumIdx = payloadGetMediumIdx(p, c.mediumStackSize);         // :153
  bool thinWalled = mat_is_thin_walled(base);                             // :155-158 (compile-time flag of the material)

  V3 prevMediumIor = v3(1.0f), nextMediumIor = v3(1.0f);                  // :161-162
  if (mediumIdx > 0) {                                                    // :164-186
    float distance = hit.t * c.ubo.metersPerSceneUnit;
    V3 transmittance;
    if (c.mediumStackSize == 0) {
      V3 sigma_a = mat_volume_absorption(mat);
      transmittance = v3(expf(-sigma_a.x * distance), expf(-sigma_a.y * distance), expf(-sigma_a.z * distance));
    } else {
      const Medium& m = p.media[mediumIdx - 1];
      prevMediumIor = m.ior;
      if (mediumIdx > 1) nextMediumIor = p.media[mediumIdx - 2].ior;
      transmittance = v3(expf(-m.sigma_t.x * distance), expf(-m.sigma_t.y * distance), expf(-m.sigma_t.z * distance));
    }
    throughput *= transmittance;
  }
  const float USE_MATERIAL_IOR = -1.0f;                                   // :188-189
  V3 iorCurrent = (isFrontFace || thinWalled) ? prevMediumIor : v3(USE_MATERIAL_IOR);
  V3 iorOther = (isFrontFace || thinWalled) ? v3(USE_MATERIAL_IOR) : nextMediumIor;

  if (aov && (p.bitfield & PAYLOAD_BOUNCES_MASK) == 0) {                  // :192-290 auxiliary AOVs on the first bounce
    aov->hit = true;
    aov->tangent = (st.tangent_u + v3(1.0f)) * 0.5f;                      // :208
    aov->bitangent = (st.tangent_v + v3(1.0f)) * 0.5f;                    // :211
    aov->bary = v3((1.0f - hit.u) - hit.v, hit.u, hit.v);                 // :214
    aov->texcoords = st.text_coords;                                      // :217
    aov->thinWalled = thinWalled;                                         // :220
    aov->objectId = in.objectId;                                          // :223-224
    V2 clipRange = unpackHalf2x16(c.ubo.clipRangePacked);                 // :227
    aov->depth = 2.0f * logf(hit.t / clipRange.x) / logf(clipRange.y / clipRange.x) - 1.0f;   // :228
    aov->faceId = s.faceIds.empty() ? (int32_t) hit.prim : s.faceIds[in.faceBase + hit.prim];   // :232-240
    aov->instanceId = in.instanceId;                                      // :243
    aov->doubleSided = isDoubleSided;                                     // :247
    aov->posNormal = (normal + v3(1.0f)) * 0.5f;                          // :252
    aov->cutoutOpacity = (mat_cutout_opacity(base) < 1.0f || s.materialOpacityBinding[in.material] >= 0 || s.materialOpacityPrimvar[in.material] >= 0) ? mat_cutout_opacity(mat) : -1.0f;
    aov->albedo = bsdf_albedo(mat, -worldRayDir, st);                     // :264-280 (front-face BSDF; no backface BSDFs exist here)
  }

  if (mat_is_emissive(mat)) {                                             // :293-343
    if (isDoubleSided) st.geom_normal *= (isLeftHanded ? -1.0f : 1.0f);
    EdfEvaluateData e; e.k1 = -worldRayDir; e.pdf = 0.0f; e.edf = v3(0.0f); e.cos = 0.0f;
    if (isFrontFace || !isDoubleSided) edf_evaluate(e, st);
    if (e.pdf > 0.0f) {
      V3 emission_intensity = v3(0.0f);
      if (isFrontFace || !isDoubleSided) emission_intensity = mat_emission_intensity(mat);
      emission_intensity *= exp2f(c.ubo.sensorExposure);
      radiance += (throughput * e.edf) * emission_intensity;
    }
    if (isDoubleSided) st.geom_normal *= (isLeftHanded ? -1.0f : 1.0f);
  }
  st.normal = normal;                                                     // :346
  if (hasMappedNormal) st.normal = mappedNormal;                          // what *_init() leaves in the state (:349-357)

  uint32_t eventType;
  {                                                                       // :361-389
    BsdfSampleData d;
    d.ior1 = iorCurrent; d.ior2 = iorOther; d.k1 = -worldRayDir;
    d.xi = rng1d_next4f(p.rng_state);                                     // :370
    bsdf_sample(mat, d, st);
    eventType = (uint32_t) d.event_type;
    throughput *= d.bsdf_over_pdf;                                        // :386
    p.ray_dir = d.k2;                                                     // :388
  }
  bool isTransmissionEvent = (eventType & BSDF_EVENT_TRANSMISSION) != 0;  // :391

  if (c.fl.nee && (eventType & (BSDF_EVENT_DIFFUSE | BSDF_EVENT_GLOSSY)) != 0) { // :394-444
    st.normal = normal;
    V4 k4 = rng1d_next4f(p.rng_state);                                    // :404
    V3 dirToLight, lightPower; float lightDist, invLightSamplePdf; uint32_t diffuseSpecularPacked;
    sampleLight(c, k4, st.position, dirToLight, lightDist, lightPower, invLightSamplePdf, diffuseSpecularPacked);
    V3 neeContrib = v3(0.0f);
    bool neeValid = (lightDist > 0.0f) && dot(dirToLight, st.geom_normal) > 0.0f; // :415
    if (neeValid) {
      BsdfEvaluateData e;
      e.ior1 = iorCurrent; e.ior2 = iorOther; e.k1 = -worldRayDir; e.k2 = dirToLight;
      bsdf_evaluate(mat, e, st);                                          // :425 (always the front-face BSDF)
      neeValid = (e.pdf > 0.0f);
      if (neeValid) {
        V2 diffuseSpecular = unpackHalf2x16(diffuseSpecularPacked);
        V3 neeRadiance = lightPower * invLightSamplePdf;
        V3 weight = throughput * neeRadiance;
        neeContrib += (weight * e.bsdf_diffuse) * diffuseSpecular.x;      // :436
        neeContrib += (weight * e.bsdf_glossy) * diffuseSpecular.y;       // :437
      }
    }
    p.neeToLight = dirToLight * lightDist;                                // :441
    p.neeContrib = neeContrib;
  }

  if (!thinWalled && isTransmissionEvent) {                               // :447-480
    if (c.mediumStackSize > 0) {
      if (isFrontFace) {                                                  // push, :450-467
        mediumIdx++;
        if (mediumIdx <= c.mediumStackSize) {
          V3 sigma_a = mat_volume_absorption(mat), sigma_s = mat_volume_scattering(mat);
          Medium m; m.ior = mat_ior(mat); m.sigma_s = sigma_s; m.sigma_t = sigma_a + sigma_s; m.bias = mat_directional_bias(mat);
          p.media[mediumIdx - 1] = m;
        }
      } else if (mediumIdx > 0) mediumIdx--;                              // pop, :468-471
    } else
    mediumIdx = (1u - mediumIdx);                                         // :473
    p.bitfield &= ~PAYLOAD_WALK_MASK;
    payloadSetMediumIdx(p, mediumIdx);
  }
  if (eventType == BSDF_EVENT_ABSORB) p.bitfield |= PAYLOAD_TERMINATE_FLAG; // :483-486

  V3 geomNormal = st.geom_normal * (isTransmissionEvent ? -1.0f : 1.0f);  // :488
  p.ray_origin = offset_ray_origin(st.position, geomNormal);              // :489
  p.throughput = throughput;
  p.radiance = radiance;
}

struct RayCounters { uint64_t closest, shadow; };

// rp_main.rgen:48-97
static inline float sampleDistance(V3 albedo, V3 throughput, V3 sigma_t, float xi, V3& pdf)
{
  V3 weights = throughput * albedo;
  float sum = (weights.x + weights.y) + weights.z;
  pdf = (sum > 1e-9f) ? (weights / sum) : v3(1.0f / 3.0f);
  if (xi < pdf.x) return sigma_t.x;
  else if (xi < (pdf.x + pdf.y)) return sigma_t.y;
  else return sigma_t.z;
}
static inline float sampleHenyeyGreensteinCos(float r, float g)
{
  if (fabsf(g) < 1e-3f) return 1.0f - 2.0f * r;
  float s = (1.0f - g * g) / ((1.0f - g) + (2.0f * g) * r);
  return ((1.0f + g * g) - s * s) / (2.0f * g);
}
static inline void sampleVolumeScatteringDirection(V2 xi, float bias, V3& dir)
{
  float cos_theta = sampleHenyeyGreensteinCos(xi.x, bias);
  float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
  float phi = (2.0f * PI) * xi.y;
  V3 t, b;
  orthonormal_basis(dir, t, b);
  dir = ((t * sin_theta) * cosf(phi) + (b * sin_theta) * sinf(phi)) + dir * cos_theta;
}

// rp_main.rgen:215-499: one sample of one pixel; returns the clamped sample colour.
static inline V3 trace_sample(const RenderCtx& c, uint32_t px, uint32_t py, uint32_t pixelIndex, uint32_t s, RayCounters& rc, AovSample* aov = nullptr)
{
  if (aov) { aov->hit = false; aov->bounces = 0; aov->neeRan = false; aov->neeShadowed = false; }
  const Ubo& ubo = c.ubo;
  uint32_t sampleIndex = ubo.sampleOffset + s;                            // :217
  uint32_t rng_state = rng1d_init(pixelIndex, sampleIndex);               // :223
  V3 ray_origin, ray_dir;
  camera_ray(ubo, c.fl, px, py, rng_state, ray_origin, ray_dir);

  ShadeRayPayload p;                                                      // :274-283
  p.throughput = v3(1.0f); p.bitfield = 0; p.radiance = v3(0.0f); p.rng_state = rng_state;
  p.ray_origin = ray_origin; p.ray_dir = ray_dir; p.pixelIndex = pixelIndex;
  p.neeToLight = v3(0.0f); p.neeContrib = v3(0.0f);
  p.walkSegmentPdf = v3(1.0f);                                            // :280-282

  V2 clipRange = clip_range_along(ubo, ray_dir);                          // :287-288

  uint32_t rrBounceOffset = ubo.maxBouncesAndRrBounceOffset & 0xFFFFu;    // :291
  uint32_t maxBounces = ubo.maxBouncesAndRrBounceOffset >> 16; if (maxBounces > PAYLOAD_BOUNCES_MASK) maxBounces = PAYLOAD_BOUNCES_MASK;

  while (true) {                                                          // :294
    uint32_t bounce = (p.bitfield & PAYLOAD_BOUNCES_MASK);
    bool terminate = (p.bitfield & PAYLOAD_TERMINATE_FLAG) != 0;
    if (bounce >= maxBounces || terminate) break;                         // :301
    float tMin = 0.0f, tMax = FLOAT_MAX;
    if (c.fl.clipping && bounce == 0) { tMin = clipRange.x; tMax = clipRange.y; } // :308-314
    uint32_t mediumIdx = 0;
    if (c.mediumStackSize > 0) {                                          // :317-346
      mediumIdx = payloadGetMediumIdx(p, c.mediumStackSize);
      if (mediumIdx > 0) {
        const Medium& m = p.media[mediumIdx - 1];
        p.walkSegmentPdf = v3(1.0f);
        uint32_t walkLength = payloadGetWalk(p);
        bool hasScattering = m.sigma_s.x > 0.0f || m.sigma_s.y > 0.0f || m.sigma_s.z > 0.0f;
        if (hasScattering && walkLength <= ubo.maxVolumeWalkLength) {
          V3 albedo = v3(safe_div(m.sigma_s.x, m.sigma_t.x), safe_div(m.sigma_s.y, m.sigma_t.y), safe_div(m.sigma_s.z, m.sigma_t.z));
          V2 xi = rng1d_next2f(p.rng_state);
          float sgm = sampleDistance(albedo, p.throughput, m.sigma_t, xi.x, p.walkSegmentPdf);
          sgm *= ubo.metersPerSceneUnit;
          tMax = -logf(1.0f - xi.y) / sgm;                                 // collision free distance
        }
      }
    }
    p.neeContrib = v3(0.0f);                                              // :349

    Hit hit;
    V3 d = p.ray_dir, o = p.ray_origin;
    rc.closest++;
    if (scene_trace(*c.s, o, d, tMin, tMax, false, hit, p.rng_state)) closest_hit_shader(c, p, hit, d, aov); // :381-393 (any-hit seed: ahit:159)
    else miss_shader(c, p, d, tMax);

    if (c.fl.nee) {                                                       // :397-438
      V3 shadowray_origin = p.ray_origin;
      float lightDist = length(p.neeToLight);
      V3 shadowray_dir = safe_div(p.neeToLight, lightDist);
      bool traceRay = luminance(p.neeContrib) > 1e-6f && lightDist > 1e-9f; // :407
      float stMin = traceRay ? 0.01f : 0.0f;
      float stMax = traceRay ? lightDist : 0.0f;
      bool shadowed = true;
      if (traceRay) {
        rc.shadow++;
        Hit sh;
        shadowed = scene_trace(*c.s, shadowray_origin, shadowray_dir, stMin, stMax, true, sh, p.rng_state);   // :399 shadow payload rng
      }
      else shadowed = false;              // tMin = tMax = 0: no traversal, the shadow miss shader still runs (rp_main_shadow.miss)
      bool addContrib = traceRay && !shadowed;                            // :426
      if (addContrib) p.radiance += p.neeContrib;                         // :429
      if (aov && bounce == 0) { aov->neeRan = true; aov->neeShadowed = shadowed; }   // :431-436
    }
    if (length(p.throughput) < 1e-9f) p.bitfield |= PAYLOAD_TERMINATE_FLAG; // :441-444
    if (bounce > rrBounceOffset) {                                        // :447-459
      float k1 = rng1d_next1f(p.rng_state);
      if (russian_roulette(k1, p.throughput, ubo.rrInvMinTermProb)) p.bitfield |= PAYLOAD_TERMINATE_FLAG;
    }
    if (c.mediumStackSize > 0 && (p.bitfield & PAYLOAD_VOLUME_WALK_MISS_FLAG) != 0) {   // :462-477
      V2 xi = rng1d_next2f(p.rng_state);
      float bias = p.media[mediumIdx - 1].bias;
      sampleVolumeScatteringDirection(xi, bias, p.ray_dir);
      p.bitfield &= ~PAYLOAD_VOLUME_WALK_MISS_FLAG;
    }
    p.bitfield++;                                                         // :480
  }
  if (aov) aov->bounces = (p.bitfield & PAYLOAD_BOUNCES_MASK);            // :483-486
  V3 radiance = p.radiance;                                               // :489-496
  float maxValue = fmaxf(radiance.x, fmaxf(radiance.y, radiance.z));
  if (maxValue > ubo.maxSampleValue) radiance *= (ubo.maxSampleValue / maxValue);
  return vmax(radiance, 0.0f);
}

// --------------------------------------------------------------------------------------------
// C API (ctypes)
// --------------------------------------------------------------------------------------------
extern "C" {

struct OrcSceneDesc {
  const GtlcVertex* vertices; uint32_t vertexCount;
  const uint32_t* faces; uint32_t faceCount;
  const InstanceDesc* instances; uint32_t instanceCount;
  const GtlcMaterialDesc* materials; uint32_t materialCount;
  const SphereLight* sphereLights; uint32_t sphereLightCount;
  const DistantLight* distantLights; uint32_t distantLightCount;
  const RectLight* rectLights; uint32_t rectLightCount;
  const DiskLight* diskLights; uint32_t diskLightCount;
  const float* domeTexels; uint32_t domeW, domeH;
  float domeRotation[4]; float domeEmission[3]; uint32_t pad;
  const int32_t* faceIds;   // optional: one per face of the global face array (GiMeshDesc.faceIds)
  const GtlcTextureDesc* textures; uint32_t textureCount; uint32_t pad2;   // R15; GtlcTextureBinding.texture = 1-based index
  // R15 scene data, resolved by the caller: column k of an instance's row belongs to binding k of its material
  const uint32_t* primvarInfos;   // instanceCount x GTLC_MAX_PRIMVAR_BINDINGS x (float offset, stride | interpolation << 4 | valid << 8)
  const float* primvarData; uint64_t primvarDataCount;
};

struct OrcRenderDesc {
  GtlcCameraDesc camera;
  GtlcRenderSettings settings;
  float clearColor[4];
  uint32_t width, height;
  uint32_t sampleOffset;
  uint32_t rowBegin, rowEnd;      // bounded samples for the CPU baseline: render rows [rowBegin, rowEnd)
  uint32_t shardRank, shardCount, shardTile; // shardCount <= 1: everything; shardTile == GTLC_SHARD_SAMPLES: every pixel, samples [spp*rank/count, spp*(rank+1)/count)
};

void* orc_scene_create(const OrcSceneDesc* d)
{
  Scene* s = new Scene();
  s->vertices.assign(d->vertices, d->vertices + d->vertexCount);
  s->faces.assign(d->faces, d->faces + (size_t) d->faceCount * 3);
  s->instances.assign(d->instances, d->instances + d->instanceCount);
  for (uint32_t i = 0; i < d->materialCount; i++) {
    Material m; m.closure = d->materials[i].closure; memcpy(m.p, d->materials[i].params, sizeof(m.p));
    s->materials.push_back(m);
    const GtlcMaterialDesc& md = d->materials[i];
    s->materialTextures.emplace_back(md.textures, md.textures + (md.textures ? md.textureCount : 0));
    s->materialPrimvars.emplace_back(md.primvars, md.primvars + (md.primvars ? md.primvarCount : 0));
    int ob = -1;                                          // the binding that feeds the cutout input (the last one wins, as in shading)
    const int opacityParam = m.closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE ? (int) GTLC_UPS_OPACITY : (m.closure == GTLC_CLOSURE_OPENPBR ? (int) GTLC_OPBR_GEOMETRY_OPACITY : -2);
    for (uint32_t b = 0; md.textures && b < md.textureCount; b++)
      if (md.textures[b].componentCount == 1 && md.textures[b].paramIndex == opacityParam) ob = (int) b;
    s->materialOpacityBinding.push_back(ob);
    int op = -1;
    for (uint32_t b = 0; md.primvars && b < md.primvarCount; b++)
      if (md.primvars[b].componentCount == 1 && md.primvars[b].paramIndex == opacityParam) op = (int) b;
    s->materialOpacityPrimvar.push_back(op);
  }
  for (uint32_t i = 0; i < d->textureCount; i++) {
    const GtlcTextureDesc& td = d->textures[i];
    Texture t; t.width = td.width; t.height = td.height; t.depth = td.depth > 1u ? td.depth : 1u; t.format = td.format;
    const size_t bpt = td.format == GTLC_TEX_RGBA8_UNORM ? 4 : td.format == GTLC_TEX_RGBA16_FLOAT ? 8 : td.format == GTLC_TEX_R32_FLOAT ? 4 : 16;
    t.data.assign((const uint8_t*) td.texels, (const uint8_t*) td.texels + (size_t) td.width * td.height * t.depth * bpt);
    s->textures.push_back(t);
  }
  if (d->primvarInfos) s->primvarInfos.assign(d->primvarInfos, d->primvarInfos + (size_t) d->instanceCount * GTLC_MAX_PRIMVAR_BINDINGS * 2);
  if (d->primvarData) s->primvarData.assign(d->primvarData, d->primvarData + d->primvarDataCount);
  if (d->sphereLightCount) s->sphereLights.assign(d->sphereLights, d->sphereLights + d->sphereLightCount);
  if (d->distantLightCount) s->distantLights.assign(d->distantLights, d->distantLights + d->distantLightCount);
  if (d->rectLightCount) s->rectLights.assign(d->rectLights, d->rectLights + d->rectLightCount);
  if (d->diskLightCount) s->diskLights.assign(d->diskLights, d->diskLights + d->diskLightCount);
  if (d->domeTexels && d->domeW && d->domeH) {
    s->domeTexels.assign(d->domeTexels, d->domeTexels + (size_t) d->domeW * d->domeH * 4);
    s->domeW = d->domeW; s->domeH = d->domeH;
    memcpy(s->domeRotation, d->domeRotation, 16); memcpy(s->domeEmission, d->domeEmission, 12);
  }
  if (d->faceIds) s->faceIds.assign(d->faceIds, d->faceIds + d->faceCount);
  scene_build(*s);
  return s;
}
void orc_scene_destroy(void* h) { delete (Scene*) h; }

static void fill_ctx(RenderCtx& c, const Scene* s, const OrcRenderDesc* r)
{
  const GtlcRenderSettings& rs = r->settings;
  c.s = s;
  Ubo& u = c.ubo;
  bool dome = !s->domeTexels.empty();
  // Gi.cpp:2384-2386: no dome -> rotation "doesn't matter", emission multiplier 1
  if (dome) { memcpy(u.domeLightRotation, s->domeRotation, 16); u.domeLightEmissionMultiplier = v3(s->domeEmission[0], s->domeEmission[1], s->domeEmission[2]); }
  else { u.domeLightRotation[0] = u.domeLightRotation[1] = u.domeLightRotation[2] = 0.0f; u.domeLightRotation[3] = 1.0f; u.domeLightEmissionMultiplier = v3(1.0f); }
  u.sphereLightCount = (uint32_t) s->sphereLights.size(); u.distantLightCount = (uint32_t) s->distantLights.size();
  u.rectLightCount = (uint32_t) s->rectLights.size(); u.diskLightCount = (uint32_t) s->diskLights.size();
  u.totalLightCount = u.sphereLightCount + u.distantLightCount + u.rectLightCount + u.diskLightCount;
  u.metersPerSceneUnit = rs.metersPerSceneUnit;
  u.cameraPosition = v3(r->camera.position[0], r->camera.position[1], r->camera.position[2]);
  u.imageDims = (r->height << 16) | r->width;
  u.cameraForward = normalize(v3(r->camera.forward[0], r->camera.forward[1], r->camera.forward[2])); // Gi.cpp:2375
  u.cameraUp = normalize(v3(r->camera.up[0], r->camera.up[1], r->camera.up[2]));
  u.focusDistance = r->camera.focusDistance;
  u.cameraVFoV = r->camera.vfov;
  u.sampleOffset = r->sampleOffset;
  u.lensRadius = (r->camera.fStop > 0.0f) ? (r->camera.focalLength / (2.0f * r->camera.fStop)) : 0.0f; // Gi.cpp:2378-2382
  u.spp = rs.spp;
  u.invTotalSampleCount = 1.0f / (float)(r->sampleOffset + rs.spp);
  u.maxSampleValue = rs.maxSampleValue;
  u.maxBouncesAndRrBounceOffset = (rs.maxBounces << 16) | rs.rrBounceOffset;
  u.rrInvMinTermProb = rs.rrInvMinTermProb;
  u.lightIntensityMultiplier = rs.lightIntensityMultiplier;
  u.clipRangePacked = packHalf2x16(r->camera.clipStart, r->camera.clipEnd);
  u.sensorExposure = r->camera.exposure;
  u.maxVolumeWalkLength = rs.maxVolumeWalkLength;                          // Gi.cpp:2391-2423
  c.mediumStackSize = rs.mediumStackSize < MAX_MEDIUM_STACK ? rs.mediumStackSize : MAX_MEDIUM_STACK;
  c.aovMask = 0;
  c.fl.jittered = rs.jitteredSampling != 0; c.fl.fis = rs.filterImportanceSampling != 0; c.fl.dof = rs.depthOfField != 0;
  c.fl.nee = rs.nextEventEstimation != 0; c.fl.clipping = rs.clippingPlanes != 0; c.fl.progressive = rs.progressiveAccumulation != 0;
  c.fl.domeCameraVisible = rs.domeLightCameraVisible != 0;
  c.background.x = r->clearColor[0]; c.background.y = r->clearColor[1]; c.background.z = r->clearColor[2]; c.background.w = r->clearColor[3];
}

// rgba: width*height*4 floats, row 0 = bottom (rp_main.rgen:210,239-242).  On entry (sampleOffset > 0
// and progressive) it must hold the previous ColorAov; on exit the new one (rp_main.rgen:506-515).
// rayCounts (optional): [closest, shadow].  Returns seconds spent in the sample loop.
double orc_render(void* h, const OrcRenderDesc* r, float* rgba, uint64_t* rayCounts)
{
  const Scene* s = (const Scene*) h;
  RenderCtx c; fill_ctx(c, s, r);
  uint32_t W = r->width, H = r->height, spp = r->settings.spp;
  uint32_t rowEnd = r->rowEnd ? r->rowEnd : H;
  float invSpp = 1.0f / (float) spp;                                      // :212
  // Sample-range sharding (the alternative to tiles): this rank's share of the call's samples, every term of
  // :498 and :514 linear in the samples, so that the sum over ranks is the image (w: 1 on rank 0, 0 elsewhere).
  const bool sampleShard = r->shardCount > 1 && r->shardTile == GTLC_SHARD_SAMPLES;
  const uint32_t sBegin = sampleShard ? (uint32_t) ((uint64_t) spp * r->shardRank / r->shardCount) : 0u;
  const uint32_t sEnd = sampleShard ? (uint32_t) ((uint64_t) spp * (r->shardRank + 1) / r->shardCount) : spp;
  uint64_t nClosest = 0, nShadow = 0;
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nClosest, nShadow)
  for (int y = (int) r->rowBegin; y < (int) rowEnd; y++) {
    RayCounters rc = { 0, 0 };
    for (uint32_t x = 0; x < W; x++) {
      uint32_t pixelIndex = x + (uint32_t) y * W;                         // :195
      float* out = &rgba[(size_t) pixelIndex * 4];
      if (r->shardCount > 1 && !sampleShard) {
        uint32_t tilesX = (W + r->shardTile - 1) / r->shardTile;
        uint32_t tile = ((uint32_t) y / r->shardTile) * tilesX + (x / r->shardTile);
        if (tile % r->shardCount != r->shardRank) { out[0] = out[1] = out[2] = out[3] = 0.0f; continue; }
      }
      V3 pixel_color = v3(0.0f);
      for (uint32_t sIdx = sBegin; sIdx < sEnd; sIdx++) {
        V3 sample_color = trace_sample(c, x, (uint32_t) y, pixelIndex, sIdx, rc);
        pixel_color += sample_color * invSpp;                             // :498
      }
      V3 prevColor = pixel_color;                                         // :507-514
      if (c.fl.progressive && c.ubo.sampleOffset > 0) prevColor = v3(out[0], out[1], out[2]);
      V3 o = (prevColor * (float) c.ubo.sampleOffset + pixel_color * (float) c.ubo.spp) * c.ubo.invTotalSampleCount;
      out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = (sampleShard && r->shardRank != 0) ? 0.0f : 1.0f;
    }
    nClosest += rc.closest; nShadow += rc.shadow;
  }
  auto t1 = std::chrono::steady_clock::now();
  if (rayCounts) { rayCounts[0] = nClosest; rayCounts[1] = nShadow; }
  return std::chrono::duration<double>(t1 - t0).count();
}

// Same as orc_render, additionally producing the auxiliary AOVs with the write order of the megakernel:
// clearAovs (rp_main.rgen:132-183), per-sample first-bounce writes of the closest-hit shader in sample order
// (rp_main.chit:191-290: the last sample that hit geometry wins for the non-accumulating AOVs, Normal/Albedo
// follow the progressive recurrence), Bounces after every sample (rgen:483-486), final normal re-normalisation
// (rgen:517-520).  NEE and ClockCycles debug AOVs are not produced (left at their clear value).
// buffers[id]: Float32Vec4 (vec3 AOVs, w = clear value's w), Float32 (Depth), Int32 (ObjectId/FaceId/InstanceId).
struct OrcAovDesc { uint32_t mask; uint32_t pad; void* buffers[GTLC_AOV_COUNT]; uint8_t clear[GTLC_AOV_COUNT][16]; };

double orc_render_aovs(void* h, const OrcRenderDesc* r, float* rgba, uint64_t* rayCounts, const OrcAovDesc* a)
{
  const Scene* s = (const Scene*) h;
  RenderCtx c; fill_ctx(c, s, r);
  c.aovMask = a->mask;
  const uint32_t W = r->width, H = r->height, spp = r->settings.spp;
  const float invSpp = 1.0f / (float) spp;
  uint32_t maxBounces = c.ubo.maxBouncesAndRrBounceOffset >> 16; if (maxBounces > PAYLOAD_BOUNCES_MASK) maxBounces = PAYLOAD_BOUNCES_MASK;
  uint64_t nClosest = 0, nShadow = 0;
  auto has = [&](int id) { return (a->mask & (1u << id)) != 0 && a->buffers[id] != nullptr; };
  auto clearF = [&](int id, int k) { float v; memcpy(&v, &a->clear[id][4 * k], 4); return v; };
  auto clearI = [&](int id) { int32_t v; memcpy(&v, &a->clear[id][0], 4); return v; };
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nClosest, nShadow)
  for (int y = 0; y < (int) H; y++) {
    RayCounters rc = { 0, 0 };
    for (uint32_t x = 0; x < W; x++) {
      const uint32_t pixelIndex = x + (uint32_t) y * W;
      float* out = &rgba[(size_t) pixelIndex * 4];
      auto v3buf = [&](int id) { return &((float*) a->buffers[id])[(size_t) pixelIndex * 4]; };
      // clearAovs
      const int lastWins[] = { GTLC_AOV_BARYCENTRICS, GTLC_AOV_TEXCOORDS, GTLC_AOV_OPACITY, GTLC_AOV_TANGENTS, GTLC_AOV_BITANGENTS, GTLC_AOV_THIN_WALLED, GTLC_AOV_DOUBLE_SIDED, GTLC_AOV_BOUNCES, GTLC_AOV_NEE, GTLC_AOV_CLOCK_CYCLES };
      for (int id : lastWins) if (has(id)) { float* b = v3buf(id); for (int k = 0; k < 4; k++) b[k] = clearF(id, k); }
      if (has(GTLC_AOV_OBJECT_ID)) ((int32_t*) a->buffers[GTLC_AOV_OBJECT_ID])[pixelIndex] = clearI(GTLC_AOV_OBJECT_ID);
      if (has(GTLC_AOV_FACE_ID)) ((int32_t*) a->buffers[GTLC_AOV_FACE_ID])[pixelIndex] = clearI(GTLC_AOV_FACE_ID);
      if (has(GTLC_AOV_INSTANCE_ID)) ((int32_t*) a->buffers[GTLC_AOV_INSTANCE_ID])[pixelIndex] = clearI(GTLC_AOV_INSTANCE_ID);
      if (has(GTLC_AOV_DEPTH)) ((float*) a->buffers[GTLC_AOV_DEPTH])[pixelIndex] = clearF(GTLC_AOV_DEPTH, 0);
      if (c.ubo.sampleOffset == 0) {
        if (has(GTLC_AOV_NORMAL)) { float* b = v3buf(GTLC_AOV_NORMAL); for (int k = 0; k < 4; k++) b[k] = clearF(GTLC_AOV_NORMAL, k); }
        if (has(GTLC_AOV_ALBEDO)) { float* b = v3buf(GTLC_AOV_ALBEDO); for (int k = 0; k < 4; k++) b[k] = clearF(GTLC_AOV_ALBEDO, k); }
      }
      V3 pixel_color = v3(0.0f);
      for (uint32_t sIdx = 0; sIdx < spp; sIdx++) {
        AovSample as;
        V3 sample_color = trace_sample(c, x, (uint32_t) y, pixelIndex, sIdx, rc, &as);
        pixel_color += sample_color * invSpp;
        if (as.hit) {
          auto put3 = [&](int id, V3 v) { if (has(id)) { float* b = v3buf(id); b[0] = v.x; b[1] = v.y; b[2] = v.z; } };
          put3(GTLC_AOV_OPACITY, as.cutoutOpacity < 0.0f ? v3(1.0f, 0.0f, 0.0f) : (as.cutoutOpacity == 0.0f ? v3(1.0f) : colormap_viridis(as.cutoutOpacity)));   // chit:202 / ahit:48
          put3(GTLC_AOV_TANGENTS, as.tangent); put3(GTLC_AOV_BITANGENTS, as.bitangent);
          put3(GTLC_AOV_BARYCENTRICS, as.bary); put3(GTLC_AOV_TEXCOORDS, as.texcoords);
          put3(GTLC_AOV_THIN_WALLED, as.thinWalled ? v3(1.0f, 0.0f, 0.0f) : v3(0.0f, 1.0f, 0.0f));   // chit:220
          put3(GTLC_AOV_DOUBLE_SIDED, as.doubleSided ? v3(0.0f, 1.0f, 0.0f) : v3(1.0f, 0.0f, 0.0f)); // chit:247
          if (has(GTLC_AOV_OBJECT_ID)) ((int32_t*) a->buffers[GTLC_AOV_OBJECT_ID])[pixelIndex] = as.objectId;
          if (has(GTLC_AOV_FACE_ID)) ((int32_t*) a->buffers[GTLC_AOV_FACE_ID])[pixelIndex] = as.faceId;
          if (has(GTLC_AOV_INSTANCE_ID)) ((int32_t*) a->buffers[GTLC_AOV_INSTANCE_ID])[pixelIndex] = as.instanceId;
          if (has(GTLC_AOV_DEPTH)) ((float*) a->buffers[GTLC_AOV_DEPTH])[pixelIndex] = as.depth;
          auto accumulate = [&](int id, V3 cur) {                                        // chit:251-261, 263-289
            if (!has(id)) return;
            float* b = v3buf(id);
            V3 prev = cur;
            if (c.fl.progressive && c.ubo.sampleOffset > 0) prev = v3(b[0], b[1], b[2]);
            V3 o = (prev * (float) c.ubo.sampleOffset + cur * (float) c.ubo.spp) * c.ubo.invTotalSampleCount;
            b[0] = o.x; b[1] = o.y; b[2] = o.z;
          };
          accumulate(GTLC_AOV_NORMAL, as.posNormal);
          accumulate(GTLC_AOV_ALBEDO, as.albedo);
        }
        if (has(GTLC_AOV_NEE) && as.neeRan) { float* b = v3buf(GTLC_AOV_NEE); b[0] = as.neeShadowed ? 1.0f : 0.0f; b[1] = as.neeShadowed ? 0.0f : 1.0f; b[2] = 0.0f; }   // rgen:431-436
        if (has(GTLC_AOV_BOUNCES)) { V3 cm = colormap_inferno((float) as.bounces / (float) maxBounces); float* b = v3buf(GTLC_AOV_BOUNCES); b[0] = cm.x; b[1] = cm.y; b[2] = cm.z; }   // rgen:483-486
      }
      V3 prevColor = pixel_color;
      if (c.fl.progressive && c.ubo.sampleOffset > 0) prevColor = v3(out[0], out[1], out[2]);
      V3 o = (prevColor * (float) c.ubo.sampleOffset + pixel_color * (float) c.ubo.spp) * c.ubo.invTotalSampleCount;
      out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = 1.0f;
      if (has(GTLC_AOV_NORMAL)) {                                                        // rgen:517-520
        float* b = v3buf(GTLC_AOV_NORMAL);
        V3 n = v3(b[0], b[1], b[2]) * 2.0f - v3(1.0f);
        V3 q = (normalize(n) + v3(1.0f)) * 0.5f;
        b[0] = q.x; b[1] = q.y; b[2] = q.z;
      }
    }
    nClosest += rc.closest; nShadow += rc.shadow;
  }
  auto t1 = std::chrono::steady_clock::now();
  if (rayCounts) { rayCounts[0] = nClosest; rayCounts[1] = nShadow; }
  return std::chrono::duration<double>(t1 - t0).count();
}

// rays: n x (ox, oy, oz, tMin, dx, dy, dz, tMax); hits: n x (t, u, v, prim, inst) 20 B records
void orc_trace(void* h, const float* rays, uint32_t n, int anyHit, void* hitsOut)
{
  const Scene* s = (const Scene*) h;
  Hit* hits = (Hit*) hitsOut;
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < (int) n; i++) {
    const float* r = &rays[(size_t) i * 8];
    Hit hit;
    scene_trace(*s, v3(r[0], r[1], r[2]), v3(r[4], r[5], r[6]), r[3], r[7], anyHit != 0, hit);
    hits[i] = hit;
  }
}

// First bounce path state for raygen parity: out n x 8 floats (ox, oy, oz, rng bits, dx, dy, dz, 0)
void orc_camera_rays(const OrcRenderDesc* r, uint32_t sampleIndex, float* out)
{
  Scene dummy; RenderCtx c; fill_ctx(c, &dummy, r);
  for (uint32_t y = 0; y < r->height; y++)
    for (uint32_t x = 0; x < r->width; x++) {
      uint32_t pixelIndex = x + y * r->width;
      uint32_t rng = rng1d_init(pixelIndex, sampleIndex);
      V3 o, d; camera_ray(c.ubo, c.fl, x, y, rng, o, d);
      float* p = &out[(size_t) pixelIndex * 8];
      p[0] = o.x; p[1] = o.y; p[2] = o.z; p[3] = uintBitsToFloat(rng); p[4] = d.x; p[5] = d.y; p[6] = d.z; p[7] = 0.0f;
    }
}

// Closure-level access for white-furnace / reciprocity tests.
// state: normal(3) geom_normal(3) tangent_u(3); k1(3); xi(4) -> k2(3) bsdf_over_pdf(3) event
void orc_bsdf_sample(const GtlcMaterialDesc* md, const float* state9, const float* k1, const float* xi, float* k2, float* bsdfOverPdf, int32_t* eventType)
{
  Material m; m.closure = md->closure; memcpy(m.p, md->params, sizeof(m.p));
  State st; st.normal = v3(state9[0], state9[1], state9[2]); st.geom_normal = v3(state9[3], state9[4], state9[5]);
  st.tangent_u = v3(state9[6], state9[7], state9[8]); st.tangent_v = cross(st.normal, st.tangent_u); st.position = v3(0.0f); st.text_coords = v3(0.0f);
  BsdfSampleData d; d.ior1 = v3(1.0f); d.ior2 = v3(-1.0f); d.k1 = v3(k1[0], k1[1], k1[2]);
  d.xi.x = xi[0]; d.xi.y = xi[1]; d.xi.z = xi[2]; d.xi.w = xi[3];
  bsdf_sample(m, d, st);
  k2[0] = d.k2.x; k2[1] = d.k2.y; k2[2] = d.k2.z;
  bsdfOverPdf[0] = d.bsdf_over_pdf.x; bsdfOverPdf[1] = d.bsdf_over_pdf.y; bsdfOverPdf[2] = d.bsdf_over_pdf.z;
  *eventType = d.event_type;
}
void orc_bsdf_evaluate(const GtlcMaterialDesc* md, const float* state9, const float* k1, const float* k2, float* diffuse, float* glossy, float* pdf)
{
  Material m; m.closure = md->closure; memcpy(m.p, md->params, sizeof(m.p));
  State st; st.normal = v3(state9[0], state9[1], state9[2]); st.geom_normal = v3(state9[3], state9[4], state9[5]);
  st.tangent_u = v3(state9[6], state9[7], state9[8]); st.tangent_v = cross(st.normal, st.tangent_u); st.position = v3(0.0f); st.text_coords = v3(0.0f);
  BsdfEvaluateData d; d.ior1 = v3(1.0f); d.ior2 = v3(-1.0f); d.k1 = v3(k1[0], k1[1], k1[2]); d.k2 = v3(k2[0], k2[1], k2[2]);
  bsdf_evaluate(m, d, st);
  diffuse[0] = d.bsdf_diffuse.x; diffuse[1] = d.bsdf_diffuse.y; diffuse[2] = d.bsdf_diffuse.z;
  glossy[0] = d.bsdf_glossy.x; glossy[1] = d.bsdf_glossy.y; glossy[2] = d.bsdf_glossy.z;
  *pdf = d.pdf;
}

// ---- known-answer access to the pure functions (common.glsl) ---------------------------------
uint32_t orc_hash_theironborn(uint32_t x) { return hash_theironborn(x); }
uint32_t orc_hash_pcg32(uint32_t* state) { return hash_pcg32(*state); }
float orc_uint_as_float(uint32_t v) { return uintAsFloat(v); }
uint32_t orc_rng_init(uint32_t pixelIndex, uint32_t sampleIndex) { return rng1d_init(pixelIndex, sampleIndex); }
float orc_rng_next1f(uint32_t* state) { return rng1d_next1f(*state); }
void orc_orthonormal_basis(const float* n, float* b1, float* b2) { V3 a, b; orthonormal_basis(v3(n[0], n[1], n[2]), a, b); b1[0] = a.x; b1[1] = a.y; b1[2] = a.z; b2[0] = b.x; b2[1] = b.y; b2[2] = b.z; }
void orc_offset_ray_origin(const float* p, const float* n, float* o) { V3 r = offset_ray_origin(v3(p[0], p[1], p[2]), v3(n[0], n[1], n[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
uint32_t orc_encode_direction(const float* d) { return encode_direction(v3(d[0], d[1], d[2])); }
uint32_t orc_host_encode_direction(const float* d) { return host_encode_direction(v3(d[0], d[1], d[2])); }
void orc_decode_direction(uint32_t e, float* o) { V3 r = decode_direction(e); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void orc_sample_hemisphere(const float* xi, float* o) { V2 x; x.x = xi[0]; x.y = xi[1]; V3 r = sample_hemisphere(x); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void orc_sample_sphere(const float* xi, const float* radius, float* o) { V2 x; x.x = xi[0]; x.y = xi[1]; V3 r = sample_sphere(x, v3(radius[0], radius[1], radius[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void orc_sample_disk(const float* xi, const float* radius, float* o) { V2 x, rr; x.x = xi[0]; x.y = xi[1]; rr.x = radius[0]; rr.y = radius[1]; V2 r = sample_disk(x, rr); o[0] = r.x; o[1] = r.y; }
float orc_luminance(const float* c) { return luminance(v3(c[0], c[1], c[2])); }
void orc_fis_gauss(const float* xi, float* o) { V2 x; x.x = xi[0]; x.y = xi[1]; V2 r = fisGauss(x); o[0] = r.x; o[1] = r.y; }
// One camera ray with everything rp_main.rgen:191-288 derives for it: origin(3) dir(3) clipRange(2), and the RNG state after the draws
void orc_camera_ray_full(const OrcRenderDesc* r, uint32_t sampleIndex, uint32_t px, uint32_t py, float* out8, uint32_t* rngState)
{
  Scene dummy; RenderCtx c; fill_ctx(c, &dummy, r);
  uint32_t rng = rng1d_init(px + py * r->width, sampleIndex);
  V3 o, d; camera_ray(c.ubo, c.fl, px, py, rng, o, d);
  V2 clip = clip_range_along(c.ubo, d);
  out8[0] = o.x; out8[1] = o.y; out8[2] = o.z; out8[3] = d.x; out8[4] = d.y; out8[5] = d.z; out8[6] = clip.x; out8[7] = clip.y; *rngState = rng;
}
// The uniform-block values Gi.cpp:2373-2423 derives from the camera: position(3) forward(3) up(3) vfov focusDistance lensRadius, then the
// bit patterns of clipRangePacked and imageDims
void orc_camera_ubo(const OrcRenderDesc* r, float* out12, uint32_t* bits2)
{
  Scene dummy; RenderCtx c; fill_ctx(c, &dummy, r);
  const Ubo& u = c.ubo;
  const float v[12] = { u.cameraPosition.x, u.cameraPosition.y, u.cameraPosition.z, u.cameraForward.x, u.cameraForward.y, u.cameraForward.z,
                        u.cameraUp.x, u.cameraUp.y, u.cameraUp.z, u.cameraVFoV, u.focusDistance, u.lensRadius };
  memcpy(out12, v, sizeof(v)); bits2[0] = u.clipRangePacked; bits2[1] = u.imageDims;
}
// Raw views of a scene's packed geometry (the buffers setup_mdl_shading_state reads) and the oracle's shading state at a hit:
// normal(3) geom_normal(3) position(3) tangent_u(3) tangent_v(3) text_coords(3)
void orc_scene_views(void* h, uint32_t inst, const uint32_t** faces, const void** fverts, uint32_t* bases3, float* o2w12, float* w2o12, uint32_t* instCount)
{
  const Scene* s = (const Scene*) h;
  *instCount = (uint32_t) s->instances.size();
  if (inst >= s->instances.size()) return;
  const InstanceDesc& in = s->instances[inst];
  *faces = s->faces.data(); *fverts = s->fverts.data();
  bases3[0] = in.vertexBase; bases3[1] = in.faceBase; bases3[2] = in.faceCount;
  memcpy(o2w12, in.objectToWorld, 48); memcpy(w2o12, &s->worldToObject[(size_t) inst * 12], 48);
}
void orc_setup_shading_state(void* h, uint32_t inst, uint32_t prim, float u, float v, const float* rayDir, float* out18, int32_t* frontFace)
{
  const Scene* s = (const Scene*) h;
  Hit hit; hit.t = 1.0f; hit.u = u; hit.v = v; hit.prim = prim; hit.inst = inst;
  State st; bool ff = false;
  setup_shading_state(*s, hit, v3(rayDir[0], rayDir[1], rayDir[2]), st, ff);
  const V3 f[6] = { st.normal, st.geom_normal, st.position, st.tangent_u, st.tangent_v, st.text_coords };
  for (int k = 0; k < 6; k++) { out18[3 * k] = f[k].x; out18[3 * k + 1] = f[k].y; out18[3 * k + 2] = f[k].z; }
  *frontFace = ff ? 1 : 0;
}
// ---- entry points for the reference's integrator compiled from its own shader sources (oracle/ref_glsl/ref_pipeline.inc) ----
// That build runs rp_main.rgen / .chit / .miss as C++ and calls back here for everything the reference gets from outside its shader
// text: ray queries against the scene (the Vulkan driver), the material's generated mdl_* functions (MDL SDK) and the dome texture.
struct HitCtx { const Scene* s; const InstanceDesc* in; Material mat; State st; V3 mappedNormal; bool hasMappedNormal; };
static thread_local HitCtx g_hitCtx;
static inline State state_from(const float* f, const HitCtx& h)
{
  State st = h.st;                       // local_position etc. from the oracle; the six vectors the shader code owns from the caller
  st.normal = v3(f[0], f[1], f[2]); st.geom_normal = v3(f[3], f[4], f[5]); st.position = v3(f[6], f[7], f[8]);
  st.tangent_u = v3(f[9], f[10], f[11]); st.tangent_v = v3(f[12], f[13], f[14]); st.text_coords = v3(f[15], f[16], f[17]);
  return st;
}
int orc_trace_one(void* h, const float* o, const float* d, float tMin, float tMax, int anyHit, uint32_t seed, float* tuv, uint32_t* primInst)
{
  Hit hit;
  const bool found = scene_trace(*(const Scene*) h, v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), tMin, tMax, anyHit != 0, hit, seed);
  tuv[0] = hit.t; tuv[1] = hit.u; tuv[2] = hit.v; primInst[0] = hit.prim; primInst[1] = hit.inst;
  return found ? 1 : 0;
}
// flags: bit 0 thin-walled, 1 emissive, 2 left-handed instance, 3 double-sided instance, 4 a normal map is bound
uint32_t orc_hit_begin(void* h, uint32_t inst, uint32_t prim, float u, float v, const float* worldRayDir)
{
  HitCtx& c = g_hitCtx;
  c.s = (const Scene*) h; c.in = &c.s->instances[inst];
  Hit hit; hit.t = 1.0f; hit.u = u; hit.v = v; hit.prim = prim; hit.inst = inst;
  bool ff; const V3 wd = v3(worldRayDir[0], worldRayDir[1], worldRayDir[2]);
  setup_shading_state(*c.s, hit, wd, c.st, ff);
  const Material& base = c.s->materials[c.in->material];
  material_at_hit(*c.s, *c.in, hit, c.st, wd, base, c.mat, c.mappedNormal, c.hasMappedNormal);
  return (mat_is_thin_walled(base) ? 1u : 0u) | (mat_is_emissive(c.mat) ? 2u : 0u) | ((c.in->flags & 1u) ? 4u : 0u) | ((c.in->flags & 2u) ? 8u : 0u) | (c.hasMappedNormal ? 16u : 0u);
}
// The parameter block at the current hit, after the primvar readers and texture bindings ran (orc_hit_begin)
void orc_hit_params(float* params) { memcpy(params, g_hitCtx.mat.p, sizeof(g_hitCtx.mat.p)); }
// The scene-data table entry of primvar reader k at instance `inst` (the reference's sceneDataInfos word, unpacked): info[0] = first float of the
// primvar in the data array, info[1] = stride in floats | interpolation << 4 | 0x100 present | 0x200 integer data
void orc_primvar_entry(void* h, uint32_t inst, uint32_t k, uint32_t* info2, const float** data)
{
  const Scene* s = (const Scene*) h;
  info2[0] = info2[1] = 0; *data = nullptr;
  if (s->primvarInfos.empty()) return;
  const uint32_t* info = &s->primvarInfos[((size_t) inst * GTLC_MAX_PRIMVAR_BINDINGS + k) * 2];
  info2[0] = info[0]; info2[1] = info[1]; *data = s->primvarData.data() + info[0];
}
void orc_hit_mapped_normal(float* n) { n[0] = g_hitCtx.mappedNormal.x; n[1] = g_hitCtx.mappedNormal.y; n[2] = g_hitCtx.mappedNormal.z; }
void orc_hit_volume(float* sigma_a, float* sigma_s, float* ior, float* bias)
{
  const Material& m = g_hitCtx.mat;
  const V3 a = mat_volume_absorption(m), sc = mat_volume_scattering(m), i = mat_ior(m);
  sigma_a[0] = a.x; sigma_a[1] = a.y; sigma_a[2] = a.z; sigma_s[0] = sc.x; sigma_s[1] = sc.y; sigma_s[2] = sc.z; ior[0] = i.x; ior[1] = i.y; ior[2] = i.z;
  *bias = mat_directional_bias(m);
}
void orc_hit_bsdf_sample(const float* state18, const float* ior1, const float* ior2, const float* k1, const float* xi, float* k2, float* pdf, float* bsdfOverPdf, int32_t* eventType)
{
  BsdfSampleData d; d.ior1 = v3(ior1[0], ior1[1], ior1[2]); d.ior2 = v3(ior2[0], ior2[1], ior2[2]); d.k1 = v3(k1[0], k1[1], k1[2]);
  d.xi = v4(xi[0], xi[1], xi[2], xi[3]); d.k2 = v3(0.0f); d.pdf = 0.0f; d.bsdf_over_pdf = v3(0.0f); d.event_type = 0;
  bsdf_sample(g_hitCtx.mat, d, state_from(state18, g_hitCtx));
  k2[0] = d.k2.x; k2[1] = d.k2.y; k2[2] = d.k2.z; *pdf = d.pdf; bsdfOverPdf[0] = d.bsdf_over_pdf.x; bsdfOverPdf[1] = d.bsdf_over_pdf.y; bsdfOverPdf[2] = d.bsdf_over_pdf.z; *eventType = d.event_type;
}
void orc_hit_bsdf_evaluate(const float* state18, const float* ior1, const float* ior2, const float* k1, const float* k2, float* diffuse, float* glossy, float* pdf)
{
  BsdfEvaluateData e; e.ior1 = v3(ior1[0], ior1[1], ior1[2]); e.ior2 = v3(ior2[0], ior2[1], ior2[2]); e.k1 = v3(k1[0], k1[1], k1[2]); e.k2 = v3(k2[0], k2[1], k2[2]);
  e.bsdf_diffuse = v3(0.0f); e.bsdf_glossy = v3(0.0f); e.pdf = 0.0f;
  bsdf_evaluate(g_hitCtx.mat, e, state_from(state18, g_hitCtx));
  diffuse[0] = e.bsdf_diffuse.x; diffuse[1] = e.bsdf_diffuse.y; diffuse[2] = e.bsdf_diffuse.z; glossy[0] = e.bsdf_glossy.x; glossy[1] = e.bsdf_glossy.y; glossy[2] = e.bsdf_glossy.z; *pdf = e.pdf;
}
void orc_hit_bsdf_albedo(const float* state18, const float* k1, float* albedo)
{
  const V3 a = bsdf_albedo(g_hitCtx.mat, v3(k1[0], k1[1], k1[2]), state_from(state18, g_hitCtx));
  albedo[0] = a.x; albedo[1] = a.y; albedo[2] = a.z;
}
void orc_hit_edf(const float* state18, const float* k1, float* edf, float* pdf, float* cosOut, float* intensity)
{
  EdfEvaluateData e; e.k1 = v3(k1[0], k1[1], k1[2]); e.pdf = 0.0f; e.edf = v3(0.0f); e.cos = 0.0f;
  if (mat_is_emissive(g_hitCtx.mat)) edf_evaluate(e, state_from(state18, g_hitCtx));
  const V3 i = mat_emission_intensity(g_hitCtx.mat);
  edf[0] = e.edf.x; edf[1] = e.edf.y; edf[2] = e.edf.z; *pdf = e.pdf; *cosOut = e.cos; intensity[0] = i.x; intensity[1] = i.y; intensity[2] = i.z;
}
// The image behind textures_2d[domeLightIndex] at (u, v): index 0 = the 1x1 fallback holding the clear colour, 1 = the dome light's lat-long image
void orc_dome_fetch(void* h, const OrcRenderDesc* r, uint32_t domeLightIndex, float u, float v, float* rgb)
{
  const Scene* s = (const Scene*) h;
  V3 c;
  if (domeLightIndex == 0 || s->domeTexels.empty()) {
    c.x = (float)(uint8_t)(r->clearColor[0] * 255.0f) / 255.0f; c.y = (float)(uint8_t)(r->clearColor[1] * 255.0f) / 255.0f; c.z = (float)(uint8_t)(r->clearColor[2] * 255.0f) / 255.0f;
  } else c = texture_latlong(*s, u, v);
  rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
// Every UniformData value the shaders read, as Gi.cpp:2373-2423 fills them: 28 floats then 16 uints (see ref_exports.cpp ref_set_ubo)
void orc_ubo_values(void* h, const OrcRenderDesc* r, float* f, uint32_t* u)
{
  RenderCtx c; fill_ctx(c, (const Scene*) h, r);
  const Ubo& b = c.ubo;
  const float fv[28] = { b.domeLightRotation[0], b.domeLightRotation[1], b.domeLightRotation[2], b.domeLightRotation[3],
                         b.domeLightEmissionMultiplier.x, b.domeLightEmissionMultiplier.y, b.domeLightEmissionMultiplier.z, b.metersPerSceneUnit,
                         b.cameraPosition.x, b.cameraPosition.y, b.cameraPosition.z, b.cameraForward.x, b.cameraForward.y, b.cameraForward.z,
                         b.cameraUp.x, b.cameraUp.y, b.cameraUp.z, b.focusDistance, b.cameraVFoV, b.lensRadius, b.invTotalSampleCount, b.maxSampleValue,
                         b.rrInvMinTermProb, b.lightIntensityMultiplier, b.sensorExposure, 0.0f, 0.0f, 0.0f };
  const uint32_t uv[16] = { b.sphereLightCount, b.distantLightCount, b.rectLightCount, b.diskLightCount, b.totalLightCount, b.maxVolumeWalkLength, b.imageDims,
                            b.sampleOffset, b.spp, b.maxBouncesAndRrBounceOffset, b.clipRangePacked, c.mediumStackSize,
                            (uint32_t) ((c.fl.jittered ? 1 : 0) | (c.fl.fis ? 2 : 0) | (c.fl.dof ? 4 : 0) | (c.fl.nee ? 8 : 0) | (c.fl.clipping ? 16 : 0) | (c.fl.progressive ? 32 : 0) | (c.fl.domeCameraVisible ? 64 : 0)), 0u, 0u, 0u };
  memcpy(f, fv, sizeof(fv)); memcpy(u, uv, sizeof(uv));
}
// The hit-independent values the OpenPBR closure derives from a parameter block (orc_openpbr.h opbr_lobes, orc_bsdf.h): what
// open_pbr_surface.mtlx feeds into its BSDF / EDF / VDF nodes.  out: a, a_c, eta, metal, spec_w, coat_w, trans_w, pad, m_f0(3), m_f82(3), trans_tint(3),
// coat_att(3), Kd(3), sigma_a(3), sigma_s(3), emission intensity(3)  (32 floats)
void orc_openpbr_lowered(const float* params, float* out)
{
  Material m; m.closure = GTLC_CLOSURE_OPENPBR; memcpy(m.p, params, sizeof(m.p));
  const OpbrLobes o = opbr_lobes(m);
  const V3 sa = mat_volume_absorption(m), ss = mat_volume_scattering(m), em = mat_emission_intensity(m);
  const float v[32] = { o.a, o.a_c, o.eta, o.metal, o.spec_w, o.coat_w, o.trans_w, 0.0f, o.m_f0.x, o.m_f0.y, o.m_f0.z, o.m_f82.x, o.m_f82.y, o.m_f82.z,
                        o.trans_tint.x, o.trans_tint.y, o.trans_tint.z, o.coat_att.x, o.coat_att.y, o.coat_att.z, o.Kd.x, o.Kd.y, o.Kd.z,
                        sa.x, sa.y, sa.z, ss.x, ss.y, ss.z, em.x, em.y, em.z };
  memcpy(out, v, sizeof(v));
}
// Host-side light arithmetic, Gi.cpp:2635-2640 (Knud Thomsen's approximation of an ellipsoid's area; the last factor is a double) and
// Gi.cpp:2724-2725 (solid angle of the distant light's cone; 2 pi in double)
float orc_sphere_light_area(float radiusX, float radiusY, float radiusZ)
{
  float ab = powf(radiusX * radiusY, 1.6f);
  float ac = powf(radiusX * radiusZ, 1.6f);
  float bc = powf(radiusY * radiusZ, 1.6f);
  return float(powf((ab + ac + bc) / 3.0f, 1.0f / 1.6f) * 4.0f * M_PI);
}
float orc_distant_light_inv_pdf(float angle)
{
  float halfAngle = 0.5f * angle;
  return (halfAngle > 0.0f) ? float(2.0f * M_PI * (1.0f - cosf(halfAngle))) : 1.0f;
}
// InstanceIds[] (rp_main_descriptors.glsl:20) and the objectId of every instance's mesh preamble (rp_main.h:142)
void orc_instance_ids(void* h, int32_t* instanceIds, int32_t* objectIds)
{
  const Scene* s = (const Scene*) h;
  for (size_t i = 0; i < s->instances.size(); i++) { instanceIds[i] = s->instances[i].instanceId; objectIds[i] = s->instances[i].objectId; }
}
// The light buffers as the shaders see them (interface/rp_main.h layouts, 48 bytes per light)
void orc_light_buffers(void* h, const void** spheres, const void** distants, const void** rects, const void** disks)
{
  const Scene* s = (const Scene*) h;
  *spheres = s->sphereLights.data(); *distants = s->distantLights.data(); *rects = s->rectLights.data(); *disks = s->diskLights.data();
}
// (function-level entry points for tests/test_reference_glsl_cpu.py: the same functions, compiled from the reference's own sources)
void orc_rng_next4f(uint32_t* state, float* o) { V4 r = rng1d_next4f(*state); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
float orc_safe_div(float a, float b) { return safe_div(a, b); }
void orc_colormap(int which, float t, float* o) { V3 r = which == 0 ? colormap_viridis(t) : colormap_inferno(t); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
float orc_sample_distance(const float* albedo, const float* throughput, const float* sigma_t, float xi, float* pdf)
{
  V3 p; float r = sampleDistance(v3(albedo[0], albedo[1], albedo[2]), v3(throughput[0], throughput[1], throughput[2]), v3(sigma_t[0], sigma_t[1], sigma_t[2]), xi, p);
  pdf[0] = p.x; pdf[1] = p.y; pdf[2] = p.z; return r;
}
float orc_sample_hg_cos(float r, float g) { return sampleHenyeyGreensteinCos(r, g); }
void orc_sample_volume_scattering_direction(const float* xi, float bias, float* dir) { V2 x; x.x = xi[0]; x.y = xi[1]; V3 d = v3(dir[0], dir[1], dir[2]); sampleVolumeScatteringDirection(x, bias, d); dir[0] = d.x; dir[1] = d.y; dir[2] = d.z; }
int orc_russian_roulette(float xi, float* throughput, float rrInvMinTermProb) { V3 t = v3(throughput[0], throughput[1], throughput[2]); bool r = russian_roulette(xi, t, rrInvMinTermProb); throughput[0] = t.x; throughput[1] = t.y; throughput[2] = t.z; return r ? 1 : 0; }
uint32_t orc_payload_get_medium_idx(uint32_t bitfield, int mediumStackSize) { ShadeRayPayload p; p.bitfield = bitfield; return payloadGetMediumIdx(p, (uint32_t) mediumStackSize); }
uint32_t orc_payload_set_medium_idx(uint32_t bitfield, uint32_t idx) { ShadeRayPayload p; p.bitfield = bitfield; payloadSetMediumIdx(p, idx); return p.bitfield; }
uint32_t orc_payload_increment_walk(uint32_t bitfield) { ShadeRayPayload p; p.bitfield = bitfield; payloadIncrementWalk(p); return p.bitfield; }
uint32_t orc_payload_get_walk(uint32_t bitfield) { ShadeRayPayload p; p.bitfield = bitfield; return payloadGetWalk(p); }
void orc_quat_rotate_dir(const float* q, const float* dir, float* o) { V3 r = quatRotateDir(q, v3(dir[0], dir[1], dir[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void orc_dome_light_uv(const float* rayDir, float* uv) { dome_latlong_uv(v3(rayDir[0], rayDir[1], rayDir[2]), uv[0], uv[1]); }
float orc_apply_wrap_and_crop(float coord, int wrap, int res) { return wrap_coord(coord, wrap, res); }
void orc_mdl_adapt_normal(const float* worldRayDir, const float* geomNormal, const float* normal, float* o)
{
  V3 r = mdl_adapt_normal(v3(worldRayDir[0], worldRayDir[1], worldRayDir[2]), v3(geomNormal[0], geomNormal[1], geomNormal[2]), v3(normal[0], normal[1], normal[2]));
  o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void orc_sample_light(const uint32_t* counts, const void* spheres, const void* distants, const void* rects, const void* disks, float lightIntensityMultiplier,
                      float sensorExposure, const float* k4, const float* surfacePos, float* out8, uint32_t* diffuseSpecularPacked)
{
  Scene s;
  s.sphereLights.assign((const SphereLight*) spheres, (const SphereLight*) spheres + counts[0]);
  s.distantLights.assign((const DistantLight*) distants, (const DistantLight*) distants + counts[1]);
  s.rectLights.assign((const RectLight*) rects, (const RectLight*) rects + counts[2]);
  s.diskLights.assign((const DiskLight*) disks, (const DiskLight*) disks + counts[3]);
  RenderCtx c; memset(&c.ubo, 0, sizeof(c.ubo)); c.s = &s;
  c.ubo.sphereLightCount = counts[0]; c.ubo.distantLightCount = counts[1]; c.ubo.rectLightCount = counts[2]; c.ubo.diskLightCount = counts[3];
  c.ubo.totalLightCount = counts[0] + counts[1] + counts[2] + counts[3];
  c.ubo.lightIntensityMultiplier = lightIntensityMultiplier; c.ubo.sensorExposure = sensorExposure;
  V3 dirToLight, power; float dist, invPdf; uint32_t dsp;
  sampleLight(c, v4(k4[0], k4[1], k4[2], k4[3]), v3(surfacePos[0], surfacePos[1], surfacePos[2]), dirToLight, dist, power, invPdf, dsp);
  out8[0] = dirToLight.x; out8[1] = dirToLight.y; out8[2] = dirToLight.z; out8[3] = dist; out8[4] = power.x; out8[5] = power.y; out8[6] = power.z; out8[7] = invPdf;
  *diffuseSpecularPacked = dsp;
}
uint32_t orc_pack_half2x16(float a, float b) { return packHalf2x16(a, b); }
void orc_unpack_half2x16(uint32_t p, float* o) { V2 r = unpackHalf2x16(p); o[0] = r.x; o[1] = r.y; }
void orc_invert_affine(const float* m, float* o) { invert_affine(m, o); }
int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// Watertight triangle test (KAT + parity of the GPU intersector)
int orc_tri_intersect(const float* o, const float* d, const float* p0, const float* p1, const float* p2, float* tuv)
{
  RayPre rp = ray_precompute(v3(d[0], d[1], d[2]));
  float t, u, v;
  bool ok = tri_intersect(rp, v3(o[0], o[1], o[2]), v3(p0[0], p0[1], p0[2]), v3(p1[0], p1[1], p1[2]), v3(p2[0], p2[1], p2[2]), t, u, v);
  if (ok) { tuv[0] = t; tuv[1] = u; tuv[2] = v; }
  return ok ? 1 : 0;
}

} // extern "C"

#include "orc_bvh8.h"
#include "orc_lbvh8.h"

// CPU restatement of the device BVH8 builder; see orc_lbvh8.h.  Returns the node count.
extern "C" uint32_t orc_lbvh8_build(const float* tris, uint32_t n, uint32_t* nodes, float* outTris, uint32_t* depth, int mode)
{
  return orc_lbvh8::build(tris, n, nodes, outTris, depth, mode);
}

// KAT access to the texture restatement: one lookup of `tex` at (u, v); out = rgba, returns 1 when clipped.
extern "C" int orc_tex_lookup(const GtlcTextureDesc* tex, float u, float v, int wrapU, int wrapV, float* out)
{
  Texture t; t.width = tex->width; t.height = tex->height; t.format = tex->format;
  const size_t bpt = tex->format == GTLC_TEX_RGBA8_UNORM ? 4 : tex->format == GTLC_TEX_RGBA16_FLOAT ? 8 : tex->format == GTLC_TEX_R32_FLOAT ? 4 : 16;
  t.data.assign((const uint8_t*) tex->texels, (const uint8_t*) tex->texels + (size_t) tex->width * tex->height * bpt);
  bool clipped;
  V4 r = tex_lookup_float4_2d(t, u, v, wrapU, wrapV, clipped);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  return clipped ? 1 : 0;
}
static Texture texture_from_desc(const GtlcTextureDesc* tex)
{
  Texture t; t.width = tex->width; t.height = tex->height; t.depth = tex->depth > 1u ? tex->depth : 1u; t.format = tex->format;
  const size_t bpt = tex->format == GTLC_TEX_RGBA8_UNORM ? 4 : tex->format == GTLC_TEX_RGBA16_FLOAT ? 8 : tex->format == GTLC_TEX_R32_FLOAT ? 4 : 16;
  t.data.assign((const uint8_t*) tex->texels, (const uint8_t*) tex->texels + (size_t) tex->width * tex->height * t.depth * bpt);
  return t;
}
// ... the 3-D lookup (mdl_interface.glsl:45-64) and the texel fetches (:91-110, :167-185)
extern "C" int orc_tex_lookup_3d(const GtlcTextureDesc* tex, float u, float v, float w, int wrapU, int wrapV, int wrapW, float* out)
{
  const Texture t = texture_from_desc(tex);
  bool clipped;
  V4 r = tex_lookup_float4_3d(t, u, v, w, wrapU, wrapV, wrapW, clipped);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  return clipped ? 1 : 0;
}
extern "C" void orc_tex_texel(const GtlcTextureDesc* tex, int x, int y, int z, float* out)
{
  const Texture t = texture_from_desc(tex);
  V4 r = tex_texel_float4(t, x, y, z);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
extern "C" float orc_srgb_texture_to_linear(float c) { return srgb_texture_to_linear(c); }
// What `texture(sampler, coord)` / `texelFetch` return for the reference's texture functions when they run as C++ (ref_exports.cpp): the
// hardware's part of a lookup -- linear filter with Repeat addressing / a raw texel -- with the wrap, crop and clip logic left to the caller.
extern "C" void orc_tex_hw_sample(const GtlcTextureDesc* tex, float u, float v, float w, int is3d, float* out)
{
  const Texture t = texture_from_desc(tex);
  const V4 r = is3d ? sample_linear_repeat_3d(t, u, v, w) : sample_linear_repeat(t, u, v);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
extern "C" void orc_tex_hw_texel(const GtlcTextureDesc* tex, int x, int y, int z, float* out)
{
  const Texture t = texture_from_desc(tex);
  const V4 r = texel(t, x, y, z);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
