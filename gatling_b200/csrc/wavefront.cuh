// Wavefront path-tracing kernels: ray generation, shading (closest-hit + miss + NEE + path control),
// and sample resolve.  Together with traverse.cuh they replace the reference's single Vulkan-RT
// megakernel (/root/reference/src/gi/shaders/rp_main.{rgen,chit,miss}, rp_main_shadow.miss); every
// function cites the shader lines it restates (paths relative to /root/reference/src/gi/shaders/).
//
// Data layout (all queues are SoA arrays of 16-byte words, read and written with 128-bit accesses):
//   path queue entry i : rayO[i] = (origin.xyz, tMin)      rayD[i] = (dir.xyz, tMax)
//                        thr[i]  = (throughput.xyz, bitfield)   rad[i] = (radiance.xyz, rng_state)
//                        slot[i] = sample slot (sampleInWave * wavePixels + pixelInWave)
//   hit record i       : hitTUVP[i] = (t, u, v, primitiveId)  hitInst[i] = instance index
//   shadow queue entry : shO = (origin.xyz, 0.01)  shD = (dir.xyz, lightDist)  shC = (neeContrib.xyz, target)
//   sampleRad[slot]    : final (unclamped) radiance of a finished sample
// The path queue is double buffered: the shade kernel compacts surviving paths into the "out"
// queue with warp-ballot + one atomic per warp, so the next bounce reads densely packed state
// (this is ShadeRayPayload, rp_main_payload.glsl:20-47, turned inside out).
#pragma once
#include "closures.cuh"
#include "textures.cuh"
#include "traverse.cuh"

namespace gtlc {

// rp_main_payload.glsl:3-9
constexpr uint32_t kPayloadMediumIdxMask = 0x0f000000u;
constexpr uint32_t kPayloadMediumIdxOffset = 24;
constexpr uint32_t kPayloadWalkMask = 0x00fff000u;
constexpr uint32_t kPayloadBouncesMask = 0x00000fffu;
constexpr uint32_t kPayloadTerminateFlag = 0x80000000u;
constexpr uint32_t kPayloadVolumeWalkMissFlag = 0x40000000u;
constexpr uint32_t kPayloadWalkOffset = 12;
constexpr uint32_t kMaxMediumStack = 4;   // GtlcRenderSettings.mediumStackSize is limited to 4: the stack travels as one uint4 of material indices

// interface/rp_main.h:25-56 UniformData (T1), as kernel launch parameters instead of a UBO.
struct UniformData {
  float domeLightRotation[4];
  float domeLightEmissionMultiplier[3];
  uint32_t sphereLightCount, distantLightCount, rectLightCount, diskLightCount, totalLightCount;
  float metersPerSceneUnit;
  float cameraPosition[3]; uint32_t imageDims;
  float cameraForward[3]; float focusDistance;
  float cameraUp[3]; float cameraVFoV;
  uint32_t sampleOffset; float lensRadius; uint32_t spp; float invTotalSampleCount;
  float maxSampleValue; uint32_t maxBouncesAndRrBounceOffset; float rrInvMinTermProb;
  float lightIntensityMultiplier; uint32_t clipRangePacked; float sensorExposure;
  // compile-time #defines of the reference (GlslShaderGen.cpp:57-72,95-133) as run-time flags
  uint32_t flags;
  float background[3];   // 1x1 fallback dome texel = u8(clearValue * 255) / 255 (Gi.cpp:2183-2199)
  uint32_t mediumStackSize, maxVolumeWalkLength;   // MEDIUM_STACK_SIZE (GlslShaderGen.cpp:67), rp_main.h maxVolumeWalkLength
};
enum : uint32_t {
  FLAG_JITTERED = 1u, FLAG_FIS = 2u, FLAG_DOF = 4u, FLAG_NEE = 8u, FLAG_CLIPPING = 16u,
  FLAG_PROGRESSIVE = 32u, FLAG_DOME_CAMERA_VISIBLE = 64u, FLAG_HAS_DOME = 128u
};

// (InstanceInfo, the per-instance record, lives in traverse.cuh: the any-hit test of a textured cutout reads it too)
constexpr uint32_t kMatFlagEmissive = 1u, kMatFlagThinWalled = 2u, kMatFlagCutout = 4u;

struct DevicePrimvarBinding { int32_t paramIndex; int32_t componentCount; float fallback[4]; float pad[2]; };

// Shading records derived ONCE per geometry upload from the reference's packed FVertex / Face arrays (which stay in HBM as
// the canonical input, and as what the device BVH builders read):
//   shading vertex, 64 bytes = two 256-bit loads: (pos.xyz, bitangentSign, normal.xyz, u) (tangent.xyz, v, -, -, -, -) -- the
//     octahedral normal / tangent already decoded (decode_direction: 2 IEEE divisions + a normalize each, six per hit);
//   shading face, 32 bytes = one 256-bit load: (i0, i1, i2, faceId, object-space geometric normal.xyz, -).
// Same device functions, evaluated per vertex / face instead of per hit: the bits a hit sees do not change.
// (kShadeVertexWords = 16, kShadeFaceWords = 8: traverse.cuh)

struct DeviceScene {
  const uint32_t* __restrict__ sverts;   // kShadeVertexWords per vertex
  const uint32_t* __restrict__ sfaces;   // kShadeFaceWords per face
  const float* __restrict__ matLowered;  // kLobeFloats per material (closure_openpbr.cuh), valid for materials without per-hit inputs
  const float4* __restrict__ instO2W;    // 3 x float4 per instance (gl_ObjectToWorldEXT rows)
  const float4* __restrict__ instW2O;    // 3 x float4 per instance (gl_WorldToObjectEXT rows)
  const InstanceInfo* __restrict__ instInfo;
  const int2* __restrict__ instIds;       // (objectId, instanceId) per instance (preamble.objectId, InstanceIds[], rp_main.chit:223,243)
  const float* __restrict__ matParams;   // GTLC_MATERIAL_PARAM_COUNT floats per material
  const int32_t* __restrict__ matClosure;
  const uint32_t* __restrict__ matFlags;
  const float4* __restrict__ sphereLights;   // 3 x float4 per light (rp_main.h:73-113)
  const float4* __restrict__ distantLights;
  const float4* __restrict__ rectLights;
  const float4* __restrict__ diskLights;
  const float4* __restrict__ domeTexels; uint32_t domeW, domeH;
  const DeviceTexture* __restrict__ textures;        // R15: texture table
  const DeviceTexBinding* __restrict__ texBindings;  // all materials' bindings, material-major
  const uint2* __restrict__ matTexRange;             // (first binding, binding count) per material
  // R15 scene data: column k of a mesh's row belongs to binding k of the mesh's material
  const DevicePrimvarBinding* __restrict__ primvarBindings;   // all materials' bindings, material-major
  const uint2* __restrict__ matPrimvarRange;                  // (first binding, binding count) per material
  const uint2* __restrict__ primvarInfos;                     // [meshSlot * GTLC_MAX_PRIMVAR_BINDINGS + k] = (float offset, stride | interpolation << 4 | valid << 8)
  const float* __restrict__ primvarData;
};

// med / walk exist only when mediumStackSize > 0 (R14): the medium stack as material indices (the reference stores the
// 40-byte Medium values, rp_main_payload.glsl:11-17; they are per-material constants) and walkSegmentPdf.
struct PathQueue { float4* rayO; float4* rayD; float4* thr; float4* rad; uint32_t* slot; uint4* med; float4* walk; };

struct WaveParams {
  uint32_t wavePixelBase;   // first local pixel of this wave
  uint32_t wavePixels;      // pixels in this wave
  uint32_t waveSamples;     // samples per pixel in this wave
  uint32_t waveSampleBase;  // index of the wave's first sample within this giRender call (s in rgen:215)
  const uint32_t* __restrict__ pixelMap;  // local pixel -> global pixel index (NULL = identity)
  uint32_t firstWave;       // this pixel range's first wave of the call (the sample sum starts at 0)
  float alpha;              // w of the colour AOV: 1, or 0 on the ranks > 0 of a sample-range shard (gtlc.h)
};

// Per-sample capture of what the closest-hit shader writes into the auxiliary AOVs on the first bounce
// (rp_main.chit:191-290); resolve_aov_kernel replays the megakernel's per-pixel write order from it.
//   a[0] = (barycentrics.xyz, flags: bit0 hit, bit1 thinWalled, bit2 doubleSided, bit3 NEE shadow ray occluded (set by the shadow kernel), bit4 NEE block ran)   a[1] = (texcoords.xyz, depth)
//   a[2] = (tangent.xyz, objectId)   a[3] = (bitangent.xyz, faceId)   a[4] = (posNormal.xyz, instanceId)   a[5] = (albedo.xyz, cutout opacity or -1)
//   clk  = clock cycles this sample's path has spent in the kernels so far (ClockCycles AOV, rgen:187-189,501-504), or NULL
struct AovSampleBuffers { float4* a[6]; float* clk; };

struct Counters {          // device-resident
  uint32_t numPaths[2];    // path-queue lengths (double buffered)
  uint32_t numShadow;
  uint32_t workCounter;         // ray fetch counter of trace_closest: zeroed by raygen / the shade kernel in front of it
  uint32_t workCounterShadow;   // ... of trace_shadow: zeroed by the shade kernel in front of it
  uint32_t repackErrors;        // bvh8_repack_nodes_kernel
  unsigned long long stats[10]; // [0] closest rays, [1] node visits, [2] tri tests (closest), [3] shadow rays, [4] max node visits of one ray, [5] non-finite rays; counting pass: [6] closest hits, [7] node visits, [8] tri tests (shadow)
};

// ---- rp_main.rgen:118-130 ---------------------------------------------------------------------
GTLC_D V2 fisGauss(V2 xi)
{
  float u1 = fmaxf(1e-38f, xi.x);
  float u2 = xi.y;
  float sigma = 0.375f;
  float r = sigma * sqrtf(-2.0f * logf(u1));
  float phi = 2.0f * kPi * u2;
  const SinCos sc = sincos_of(phi);
  V2 o; o.x = sc.c * r; o.y = sc.s * r;
  return o;
}

GTLC_D V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }

// ---- raygen: rp_main.rgen:191-288 -------------------------------------------------------------
__global__ void __launch_bounds__(256) raygen_kernel(UniformData ubo, WaveParams wp, PathQueue q, Counters* ctr, float* clk)
{
  const uint32_t total = wp.wavePixels * wp.waveSamples;
  const uint32_t imageWidth = ubo.imageDims & 0xFFFFu;
  const uint32_t imageHeight = ubo.imageDims >> 16;
  const V3 camPos = ld3(ubo.cameraPosition), camFwd = ld3(ubo.cameraForward), camUp = ld3(ubo.cameraUp);
  const V3 camera_right = cross(camFwd, camUp);                          // :199
  const float aspect_ratio = (float) imageWidth / (float) imageHeight;   // :200
  const float H = 1.0f;
  const float W = H * aspect_ratio;
  const float d = H / (2.0f * tanf(ubo.cameraVFoV * 0.5f));              // :204
  const float WX = W / (float) imageWidth;
  const float HY = H / (float) imageHeight;
  const V3 C = camPos + camFwd * d;                                      // :209
  const V3 L = (C - (camera_right * W) * 0.5f) - (camUp * H) * 0.5f;     // :210

  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x) {
    const uint32_t clk0 = clk ? (uint32_t) clock() : 0u;
    const uint32_t j = slot / wp.wavePixels;
    const uint32_t lp = wp.wavePixelBase + (slot - j * wp.wavePixels);
    const uint32_t pixelIndex = wp.pixelMap ? wp.pixelMap[lp] : lp;      // :195
    const uint32_t px = pixelIndex % imageWidth, py = pixelIndex / imageWidth;
    const uint32_t sampleIndex = ubo.sampleOffset + wp.waveSampleBase + j;  // :217
    uint32_t rng_state = rng1d_init(pixelIndex, sampleIndex);            // :223
    V2 rand2_xy = rng1d_next2f(rng_state);                               // :224
    V2 so; so.x = 0.5f; so.y = 0.5f;                                     // :227
    if (ubo.flags & FLAG_JITTERED) {
      if (ubo.flags & FLAG_FIS) { V2 g = fisGauss(rand2_xy); so.x = 0.5f + g.x; so.y = 0.5f + g.y; }  // :232
      else so = rand2_xy;                                                // :235
    }
    V3 P = (L + (camera_right * ((float) px + so.x)) * WX) + (camUp * ((float) py + so.y)) * HY;  // :239-242
    V3 ray_origin = camPos;
    V3 ray_dir = normalize(P - ray_origin);                              // :245
    if ((ubo.flags & FLAG_DOF) && ubo.lensRadius > 0.0f) {               // :250-267
      V2 rand2_zw = rng1d_next2f(rng_state);
      V3 focalPoint = ray_origin + ray_dir * ubo.focusDistance;
      V3 h = sample_hemisphere(rand2_zw);
      float ax = h.x * ubo.lensRadius, ay = h.y * ubo.lensRadius;
      ray_origin += camera_right * ax;
      ray_origin += camUp * ay;
      ray_dir = normalize(focalPoint - ray_origin);
    }
    ray_dir.x += (ray_dir.x == 0.0f) ? kFloatMin : 0.0f;                 // :271
    ray_dir.y += (ray_dir.y == 0.0f) ? kFloatMin : 0.0f;
    ray_dir.z += (ray_dir.z == 0.0f) ? kFloatMin : 0.0f;
    float tMin = 0.0f, tMax = kFloatMax;                                 // :306-314 (bounce 0)
    if (ubo.flags & FLAG_CLIPPING) {
      float cosConeAngle = fmaxf(1e-5f, dot(ray_dir, camFwd));           // :287
      V2 clipRange = unpackHalf2x16(ubo.clipRangePacked);
      tMin = clipRange.x / cosConeAngle; tMax = clipRange.y / cosConeAngle;   // :288
    }
    q.rayO[slot] = make_float4(ray_origin.x, ray_origin.y, ray_origin.z, tMin);
    q.rayD[slot] = make_float4(ray_dir.x, ray_dir.y, ray_dir.z, tMax);
    q.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, u2f(0u));               // :274-275
    q.rad[slot] = make_float4(0.0f, 0.0f, 0.0f, u2f(rng_state));        // :276-277
    q.slot[slot] = slot;
    if (q.med) { q.med[slot] = make_uint4(0u, 0u, 0u, 0u); q.walk[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f); }   // :280-282 (MEDIUM_STACK_SIZE > 0)
    if (clk) clk[slot] = (float) ((uint32_t) clock() - clk0);            // the ClockCycles AOV starts counting here (rgen:187-189)
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { ctr->numPaths[0] = total; ctr->numPaths[1] = 0; ctr->numShadow = 0; ctr->workCounter = 0; ctr->workCounterShadow = 0; }
}

// Per-bounce bookkeeping needs no kernel of its own: trace_closest's prologue zeroes the OUT queue length and the shadow
// queue length and adds its ray count to the statistics (TraceParams.zeroA / zeroB / rayStat), trace_shadow's adds its
// own, and the shade kernel -- which runs between the two -- zeroes both ray-fetch counters for the launches behind it.

// ---- rp_main.miss:39-53 -------------------------------------------------------------------------
GTLC_D V3 quatRotateDir(const float* q, V3 dir)
{
  V3 qv = v3(q[0], q[1], q[2]);
  V3 a = cross(qv, dir);
  V3 b = cross(qv, a);
  return dir + ((a * q[3]) + b) * 2.0f;
}

GTLC_D V3 texture_latlong(const DeviceScene& s, float u, float v)
{
  // sampler: linear / linear, REPEAT (Gi.cpp:388-392); one mip
  const int W = (int) s.domeW, H = (int) s.domeH;
  float x = u * (float) W - 0.5f, y = v * (float) H - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float ax = x - fx, ay = y - fy;
  int x0 = (int) fx, y0 = (int) fy;
  int xa = x0 % W; if (xa < 0) xa += W;
  int xb = (x0 + 1) % W; if (xb < 0) xb += W;
  int ya = y0 % H; if (ya < 0) ya += H;
  int yb = (y0 + 1) % H; if (yb < 0) yb += H;
  float4 t00 = __ldg(&s.domeTexels[(size_t) ya * W + xa]), t10 = __ldg(&s.domeTexels[(size_t) ya * W + xb]);
  float4 t01 = __ldg(&s.domeTexels[(size_t) yb * W + xa]), t11 = __ldg(&s.domeTexels[(size_t) yb * W + xb]);
  V3 top = v3(t00.x, t00.y, t00.z) * (1.0f - ax) + v3(t10.x, t10.y, t10.z) * ax;
  V3 bot = v3(t01.x, t01.y, t01.z) * (1.0f - ax) + v3(t11.x, t11.y, t11.z) * ax;
  return top * (1.0f - ay) + bot * ay;
}

GTLC_D V3 sampleDomeLight(const UniformData& ubo, const DeviceScene& s, uint32_t domeLightIndex, V3 rayDir)
{
  if (domeLightIndex == 0 || !(ubo.flags & FLAG_HAS_DOME)) return ld3(ubo.background);
  float u = (atan2f(rayDir.z, rayDir.x) + 0.5f * kPi) / (2.0f * kPi);    // miss:48
  float v = 1.0f - acosf(rayDir.y) / kPi;                                // miss:49
  return texture_latlong(s, u, v);
}

struct LightSample { V3 dirToLight; float dist; V3 power; float invPdf; uint32_t diffuseSpecularPacked; };

// ---- rp_main.chit:30-129 ------------------------------------------------------------------------
GTLC_D void sampleLight(const UniformData& ubo, const DeviceScene& s, V4 k4, V3 surfacePos, LightSample& ls)
{
  const float total = (float) ubo.totalLightCount;
  if ((k4.x * total) <= (float) ubo.sphereLightCount) {                                      // :32
    const uint32_t n = ubo.sphereLightCount;
    uint32_t lightIndex = (uint32_t) (k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u;   // :34
    float4 l0 = make_float4(0, 0, 0, 0), l1 = l0, l2 = l0;   // no lights: zero-filled buffer
    if (n != 0u) { l0 = __ldg(&s.sphereLights[lightIndex * 3 + 0]); l1 = __ldg(&s.sphereLights[lightIndex * 3 + 1]); l2 = __ldg(&s.sphereLights[lightIndex * 3 + 2]); }
    const V3 lpos = v3(l0.x, l0.y, l0.z);
    V2 zw; zw.x = k4.z; zw.y = k4.w;
    V3 samplePos = lpos + sample_sphere(zw, v3(l2.x, l2.y, l2.z));                           // :40
    V3 dir = samplePos - surfacePos;
    ls.dist = length(dir);
    ls.dirToLight = safe_div(dir, ls.dist);
    V3 lightNormal = normalize(samplePos - lpos);                                            // :46
    float cosTheta = fmaxf(0.0f, dot(-ls.dirToLight, lightNormal));
    const float area = l1.w;
    ls.invPdf = safe_div((area > 0.0f) ? (area * cosTheta) : 1.0f, ls.dist * ls.dist);      // :48
    ls.power = v3(l1.x, l1.y, l1.z) * ubo.lightIntensityMultiplier;
    ls.diffuseSpecularPacked = f2u(l0.w);
  } else if ((k4.x * total) <= (float) (ubo.sphereLightCount + ubo.distantLightCount)) {     // :53
    const uint32_t n = ubo.distantLightCount;
    uint32_t lightIndex = (uint32_t) (k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u;
    const float4 l0 = __ldg(&s.distantLights[lightIndex * 3 + 0]), l1 = __ldg(&s.distantLights[lightIndex * 3 + 1]), l2 = __ldg(&s.distantLights[lightIndex * 3 + 2]);
    ls.dist = 100000.0f;
    ls.dirToLight = -v3(l0.x, l0.y, l0.z);
    ls.power = v3(l1.x, l1.y, l1.z) * ubo.lightIntensityMultiplier;
    ls.invPdf = l2.w;
    ls.diffuseSpecularPacked = f2u(l1.w);
    const float angle = l0.w;
    if (angle > 0.0f) {                                                                      // :65-74
      V3 t1, t2;
      orthonormal_basis(ls.dirToLight, t1, t2);
      float phi = ((k4.z * 2.0f) * kPi) - kPi;
      float theta = k4.w * angle;
      const SinCos sp = sincos_of(phi), st = sincos_of(theta);
      ls.dirToLight = normalize(((t1 * sp.c + t2 * sp.s) * st.s) + ls.dirToLight * st.c);
    }
  } else if ((k4.x * total) <= (float) (ubo.sphereLightCount + ubo.distantLightCount + ubo.rectLightCount)) {   // :76
    const uint32_t n = ubo.rectLightCount;
    uint32_t lightIndex = (uint32_t) (k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u;
    const float4 l0 = __ldg(&s.rectLights[lightIndex * 3 + 0]), l1 = __ldg(&s.rectLights[lightIndex * 3 + 1]), l2 = __ldg(&s.rectLights[lightIndex * 3 + 2]);
    const float width = l0.w, height = l1.w;
    float sx = (k4.z - 0.5f) * width, sy = (k4.w - 0.5f) * height;                           // :83
    V3 t0 = decode_direction(f2u(l2.x));
    V3 t1 = decode_direction(f2u(l2.y));
    V3 samplePos = (v3(l0.x, l0.y, l0.z) + t0 * sx) + t1 * sy;                               // :87
    V3 dir = samplePos - surfacePos;
    ls.dist = length(dir);
    ls.dirToLight = safe_div(dir, ls.dist);
    V3 lightNormal = cross(t1, t0);                                                          // :93
    float cosTheta = fmaxf(0.0f, dot(-ls.dirToLight, lightNormal));
    float area = width * height;
    ls.invPdf = safe_div((area > 0.0f) ? (area * cosTheta) : 1.0f, ls.dist * ls.dist);
    ls.power = v3(l1.x, l1.y, l1.z) * ubo.lightIntensityMultiplier;
    ls.diffuseSpecularPacked = f2u(l2.z);
  } else {                                                                                   // :101
    const uint32_t n = ubo.diskLightCount;
    uint32_t lightIndex = (uint32_t) (k4.y * (float) n); if (lightIndex > n - 1u) lightIndex = n - 1u;
    const float4 l0 = __ldg(&s.diskLights[lightIndex * 3 + 0]), l1 = __ldg(&s.diskLights[lightIndex * 3 + 1]), l2 = __ldg(&s.diskLights[lightIndex * 3 + 2]);
    V2 radiusXY; radiusXY.x = l0.w; radiusXY.y = l1.w;
    V2 zw; zw.x = k4.z; zw.y = k4.w;
    V2 sampleOnDisk = sample_disk(zw, radiusXY);
    V3 t0 = decode_direction(f2u(l2.x));
    V3 t1 = decode_direction(f2u(l2.y));
    V3 samplePos = (v3(l0.x, l0.y, l0.z) + t0 * sampleOnDisk.x) + t1 * sampleOnDisk.y;
    V3 dir = samplePos - surfacePos;
    ls.dist = length(dir);
    ls.dirToLight = safe_div(dir, ls.dist);
    V3 lightNormal = cross(t1, t0);
    float cosTheta = fmaxf(0.0f, dot(-ls.dirToLight, lightNormal));
    float area = (radiusXY.x * radiusXY.y) * kPi;
    ls.invPdf = safe_div((area > 0.0f) ? (area * cosTheta) : 1.0f, ls.dist * ls.dist);
    ls.power = v3(l1.x, l1.y, l1.z) * ubo.lightIntensityMultiplier;
    ls.diffuseSpecularPacked = f2u(l2.z);
  }
  ls.power *= exp2f(ubo.sensorExposure);                                                     // :127
  ls.invPdf *= (float) ubo.totalLightCount;                                                  // :128
}

// ---- shading records (see DeviceScene) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) shading_vertices_kernel(const float4* __restrict__ fverts, uint32_t count, uint4* __restrict__ out)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const float4 f1 = fverts[(size_t) i * 2], f2 = fverts[(size_t) i * 2 + 1];      // FVertex, rp_main.h:58-64
    const V3 n = decode_direction(f2u(f2.x)), t = decode_direction(f2u(f2.y));      // mdl_shading_state.glsl:30-32, 47-49
    uint4* o = out + (size_t) i * 4;
    o[0] = make_uint4(f2u(f1.x), f2u(f1.y), f2u(f1.z), f2u(f1.w));
    o[1] = make_uint4(f2u(n.x), f2u(n.y), f2u(n.z), f2u(f2.z));
    o[2] = make_uint4(f2u(t.x), f2u(t.y), f2u(t.z), f2u(f2.w));
    o[3] = make_uint4(0u, 0u, 0u, 0u);
  }
}
// one launch per mesh: faces [faceBase, faceBase + faceCount) index the vertices from vertexBase on
__global__ void __launch_bounds__(256) shading_faces_kernel(const uint32_t* __restrict__ faces, const int32_t* __restrict__ faceIds, const float4* __restrict__ fverts,
                                                            uint32_t faceBase, uint32_t faceCount, uint32_t vertexBase, uint4* __restrict__ out)
{
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < faceCount; k += gridDim.x * blockDim.x) {
    const uint32_t f = faceBase + k;
    const uint32_t i0 = faces[(size_t) f * 3], i1 = faces[(size_t) f * 3 + 1], i2 = faces[(size_t) f * 3 + 2];
    const float4 a = fverts[(size_t) (vertexBase + i0) * 2], b = fverts[(size_t) (vertexBase + i1) * 2], c = fverts[(size_t) (vertexBase + i2) * 2];
    const V3 p0 = v3(a.x, a.y, a.z), p1 = v3(b.x, b.y, b.z), p2 = v3(c.x, c.y, c.z);
    const V3 ng = normalize(cross(p1 - p0, p2 - p0));                                // mdl_shading_state.glsl:26
    out[(size_t) f * 2] = make_uint4(i0, i1, i2, (uint32_t) faceIds[f]);
    out[(size_t) f * 2 + 1] = make_uint4(f2u(ng.x), f2u(ng.y), f2u(ng.z), 0u);
  }
}

// ---- mdl_shading_state.glsl:4-98 (R4) -----------------------------------------------------------
GTLC_D void setup_shading_state(const DeviceScene& s, const InstanceInfo in, uint32_t inst, uint32_t prim, float hu, float hv,
                                V3 rayDir, ShadingState& st, bool& isFrontFace, int32_t& faceId)
{
  const float4 o0 = __ldg(&s.instO2W[inst * 3 + 0]), o1 = __ldg(&s.instO2W[inst * 3 + 1]), o2 = __ldg(&s.instO2W[inst * 3 + 2]);
  const float4 w0 = __ldg(&s.instW2O[inst * 3 + 0]), w1 = __ldg(&s.instW2O[inst * 3 + 1]), w2 = __ldg(&s.instW2O[inst * 3 + 2]);
  uint32_t fr[8];
  ldg256(s.sfaces + (size_t) (in.faceBase + prim) * kShadeFaceWords, fr);              // :10 (+ the face's geometric normal)
  const uint32_t i0 = fr[0], i1 = fr[1], i2 = fr[2];
  faceId = (int32_t) fr[3];
  const uint32_t* vb = s.sverts + (size_t) in.vertexBase * kShadeVertexWords;
  uint32_t a1[8], a2[8], b1[8], b2[8], c1[8], c2[8];                                   // :12-14
  ldg256(vb + (size_t) i0 * kShadeVertexWords, a1); ldg256(vb + (size_t) i0 * kShadeVertexWords + 8, a2);
  ldg256(vb + (size_t) i1 * kShadeVertexWords, b1); ldg256(vb + (size_t) i1 * kShadeVertexWords + 8, b2);
  ldg256(vb + (size_t) i2 * kShadeVertexWords, c1); ldg256(vb + (size_t) i2 * kShadeVertexWords + 8, c2);
  const V3 bc = v3((1.0f - hu) - hv, hu, hv);                                          // :16
  const V3 p0 = v3(u2f(a1[0]), u2f(a1[1]), u2f(a1[2])), p1 = v3(u2f(b1[0]), u2f(b1[1]), u2f(b1[2])), p2 = v3(u2f(c1[0]), u2f(c1[1]), u2f(c1[2]));
  const V3 localPos = (p0 * bc.x + p1 * bc.y) + p2 * bc.z;                             // :23
  const V3 pos = xform_point(o0, o1, o2, localPos);                                    // :24
  V3 geomNormal = v3(u2f(fr[4]), u2f(fr[5]), u2f(fr[6]));                              // :26 normalize(cross(p1 - p0, p2 - p0)), per face
  geomNormal = normalize(xform_dir_transposed(w0, w1, w2, geomNormal));                // :27
  const V3 n0 = v3(u2f(a1[4]), u2f(a1[5]), u2f(a1[6])), n1 = v3(u2f(b1[4]), u2f(b1[5]), u2f(b1[6])), n2 = v3(u2f(c1[4]), u2f(c1[5]), u2f(c1[6]));  // :30-32 (decoded per vertex)
  const V3 localNormal = normalize((n0 * bc.x + n1 * bc.y) + n2 * bc.z);               // :34
  V3 normal = normalize(xform_dir_transposed(w0, w1, w2, localNormal));                // :35
  isFrontFace = dot(geomNormal, -rayDir) >= 0.0f;                                      // :38
  if (!isFrontFace) { geomNormal = -geomNormal; normal = -normal; }                    // :40-44
  const V3 t0 = v3(u2f(a2[0]), u2f(a2[1]), u2f(a2[2])), t1 = v3(u2f(b2[0]), u2f(b2[1]), u2f(b2[2])), t2 = v3(u2f(c2[0]), u2f(c2[1]), u2f(c2[2]));  // :47-49
  const V3 localTangent = normalize((t0 * bc.x + t1 * bc.y) + t2 * bc.z);              // :51
  V3 tangent = normalize(xform_dir(o0, o1, o2, localTangent));                         // :52
  tangent = normalize(tangent - normal * dot(tangent, normal));                        // :55
  const float bitangentSign = (bc.x * u2f(a1[3]) + bc.y * u2f(b1[3])) + bc.z * u2f(c1[3]);   // :57
  const V3 bitangent = cross(normal, tangent) * bitangentSign;                         // :58
  st.u = (bc.x * u2f(a1[7]) + bc.y * u2f(b1[7])) + bc.z * u2f(c1[7]);                  // :61-64
  st.v = (bc.x * u2f(a2[3]) + bc.y * u2f(b2[3])) + bc.z * u2f(c2[3]);
  st.normal = normal; st.geom_normal = geomNormal; st.position = pos;                  // :82-90
  st.local_position = localPos;                                                        // object space: the coordinate of 3-D texture bindings
  st.tangent_u = tangent; st.tangent_v = bitangent;
  st.hitIndices[0] = i0; st.hitIndices[1] = i1; st.hitIndices[2] = i2;                  // mdl_shading_state.glsl:66-79 (renderer state)
}

// ---- volume random walk (R14): rp_main.rgen:48-97 ------------------------------------------------------
GTLC_D float sampleDistance(V3 albedo, V3 throughput, V3 sigma_t, float xi, V3& pdf)
{
  const V3 weights = throughput * albedo;
  const float sum = (weights.x + weights.y) + weights.z;
  pdf = (sum > 1e-9f) ? (weights / sum) : v3(1.0f / 3.0f);
  if (xi < pdf.x) return sigma_t.x;
  else if (xi < (pdf.x + pdf.y)) return sigma_t.y;
  else return sigma_t.z;
}
GTLC_D float sampleHenyeyGreensteinCos(float r, float g)
{
  if (fabsf(g) < 1e-3f) return 1.0f - 2.0f * r;
  const float s = (1.0f - g * g) / ((1.0f - g) + (2.0f * g) * r);
  return ((1.0f + g * g) - s * s) / (2.0f * g);
}
GTLC_D void sampleVolumeScatteringDirection(V2 xi, float bias, V3& dir)
{
  const float cos_theta = sampleHenyeyGreensteinCos(xi.x, bias);
  const float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
  const float phi = (2.0f * kPi) * xi.y;
  V3 t, b;
  orthonormal_basis(dir, t, b);
  const SinCos sc = sincos_of(phi);
  dir = ((t * sin_theta) * sc.c + (b * sin_theta) * sc.s) + dir * cos_theta;
}
// rp_main_payload.glsl:76-89
GTLC_D uint32_t payload_medium_idx(uint32_t bitfield, uint32_t mediumStackSize)
{
  const uint32_t b = (bitfield & kPayloadMediumIdxMask) >> kPayloadMediumIdxOffset;
  const uint32_t maxMediumIdx = mediumStackSize > 1u ? mediumStackSize : 1u;
  return b < maxMediumIdx ? b : maxMediumIdx;
}
// the Medium a material defines (rp_main.chit:455-465)
GTLC_D void medium_of(const DeviceScene& s, uint32_t material, V3& sigma_s, V3& sigma_t)
{
  const float* p = s.matParams + (size_t) material * GTLC_MATERIAL_PARAM_COUNT;
  const int closure = s.matClosure[material];
  sigma_s = mat_volume_scattering(closure, p);
  sigma_t = mat_volume_absorption(closure, p) + sigma_s;
}

// ---- primvar-driven material inputs (R15): scene_data_lookup_float{,2,3}, mdl_interface.glsl:281-417 ----------
GTLC_D void apply_primvar_bindings(const DeviceScene& s, const InstanceInfo in, uint32_t inst, uint32_t prim, float hu, float hv,
                                   const ShadingState& st, float* lp)
{
  const uint2 range = s.matPrimvarRange[in.material];
  const uint32_t meshSlot = in.flags >> 8;
  const V3 bc = v3((1.0f - hu) - hv, hu, hv);
  for (uint32_t k = 0; k < range.y; k++) {
    const DevicePrimvarBinding& pb = s.primvarBindings[range.x + k];
    const uint2 info = s.primvarInfos[(size_t) meshSlot * GTLC_MAX_PRIMVAR_BINDINGS + k];
    if (!(info.y & 0x100u)) {                                        // scene_data_isvalid: default_value
      for (int c = 0; c < pb.componentCount; c++) lp[pb.paramIndex + c] = pb.fallback[c];
      continue;
    }
    const uint32_t stride = info.y & 15u, interpolation = (info.y >> 4) & 15u;
    uint32_t i0, i1, i2;                                             // get_scene_data_indices :281-301
    if (interpolation == GTLC_INTERP_UNIFORM) i0 = i1 = i2 = prim;
    else if (interpolation == GTLC_INTERP_INSTANCE) { const int id = s.instIds[inst].y; i0 = i1 = i2 = id > 0 ? (uint32_t) id : 0u; }   // InstanceIds[gl_InstanceID]
    else if (interpolation == GTLC_INTERP_CONSTANT) i0 = i1 = i2 = 0u;
    else { i0 = st.hitIndices[0]; i1 = st.hitIndices[1]; i2 = st.hitIndices[2]; }
    const float* d = s.primvarData + info.x;
    if (info.y & 0x200u) {                                           // scene_data_lookup_int{,2,3} :417-479: "nearest interpolation"
      const uint32_t pick = bc.x > bc.y ? (bc.x > bc.z ? i0 : i2) : (bc.y > bc.z ? i1 : i2);
      for (int c = 0; c < pb.componentCount; c++)
        lp[pb.paramIndex + c] = c < (int) stride ? (float) (int32_t) f2u(__ldg(d + (size_t) pick * stride + c)) : pb.fallback[c];
      continue;
    }
    for (int c = 0; c < pb.componentCount; c++) {
      const float v0 = c < (int) stride ? __ldg(d + (size_t) i0 * stride + c) : 0.0f;
      const float v1 = c < (int) stride ? __ldg(d + (size_t) i1 * stride + c) : 0.0f;
      const float v2 = c < (int) stride ? __ldg(d + (size_t) i2 * stride + c) : 0.0f;
      lp[pb.paramIndex + c] = (v0 * bc.x + v1 * bc.y) + v2 * bc.z;
    }
  }
}

// ---- textured material inputs (R15) ------------------------------------------------------------------
// Evaluates the material's texture bindings at the hit's uv and writes the results into `lp`, the path's private
// copy of the parameter block.  A GTLC_BIND_NORMAL binding yields the mapped shading normal (returned through
// `mappedNormal`; the caller installs it where the reference's *_init() would, rp_main.chit:346-357).
GTLC_D bool apply_texture_bindings(const DeviceScene& s, uint32_t material, V3 worldRayDir, const ShadingState& st, float* lp, V3& mappedNormal)
{
  const uint2 range = s.matTexRange[material];
  bool hasNormal = false;
  for (uint32_t b = 0; b < range.y; b++) {
    const DeviceTexBinding& tb = s.texBindings[range.x + b];
    float ch[4];
    eval_tex_binding(tb, s.textures, st.u, st.v, st.local_position, ch);
    const int count = (int) (tb.packed & 15u);
    float val[3] = { 0.0f, 0.0f, 0.0f };
    for (int k = 0; k < count && k < 3; k++) val[k] = ch[(tb.packed >> (4 + 2 * k)) & 3u];
    if (tb.paramIndex == GTLC_BIND_NORMAL) {
      const V3 n = normalize((st.tangent_u * val[0] + st.tangent_v * val[1]) + st.normal * val[2]);
      mappedNormal = mdl_adapt_normal(worldRayDir, st.geom_normal, n);
      hasNormal = true;
    } else {
      const uint32_t op = tb.combine & 15u;                                           // gtlc.h GtlcCombineOp
      const float f = (op == GTLC_COMBINE_MIX) ? lp[(tb.combine >> 8) & 255u] : 0.0f;
      for (int k = 0; k < count && k < 3; k++) {
        float& x = lp[tb.paramIndex + k];
        if (op == GTLC_COMBINE_SET) x = val[k];
        else if (op == GTLC_COMBINE_MULTIPLY) x = x * val[k];
        else if (op == GTLC_COMBINE_ADD) x = x + val[k];
        else x = x * (1.0f - f) + val[k] * f;
      }
    }
  }
  return hasNormal;
}

// ---- shade: rp_main.chit:132-493 + rp_main.miss:55-87 + rp_main.rgen:397-480 ----------------------
// One thread per live path.  Survivors are compacted into `qout`, shadow rays appended to the shadow
// queue, finished samples written to sampleRad[slot].
// GTLC_SHADE_PHASE_SYNC: the body is cut into five phases (state set-up + emission | BSDF sample | light sample |
// BSDF evaluate | roulette + compaction) and the 16 warps of a 512-thread CTA -- one CTA per SM at 90 registers -- enter
// each phase together, so that the SM's instruction cache holds one phase at a time.  The OpenPBR + NEE instantiation is
// ~190 KB of SASS and was measured stalling on instruction fetch (smsp__average_warps_issue_stalled_no_instruction ~17
// per issued instruction, issue slots 18 % busy); with the barriers the kernel takes 36 ms instead of 49 ms on the C3
// stand-in.  The Lambert-only instantiation is small and keeps 256-thread CTAs without barriers.
#ifndef GTLC_SHADE_PHASE_SYNC
#define GTLC_SHADE_PHASE_SYNC 1
#endif
#ifndef GTLC_SHADE_PHASE_MASK
#define GTLC_SHADE_PHASE_MASK 2   // bit k: barrier in front of phase k + 2
#endif
#ifndef GTLC_SHADE_SYNC_MIN_CLOSURE
#define GTLC_SHADE_SYNC_MIN_CLOSURE GTLC_CLOSURE_OPENPBR
#endif
#ifndef GTLC_SHADE_BIG_BLOCK
#define GTLC_SHADE_BIG_BLOCK 512
#endif
// Resident CTAs per SM asked of the compiler for the OpenPBR instantiations without per-hit inputs (the ones a production frame
// spends its time in): 2 x 512 threads = 64 registers.  ptxas then spills ~360 bytes per thread, and the kernel is still 11 %
// (C3 stand-in) to 18 % (C5 stand-in) faster than one 512-thread CTA at 108 registers: it waits on dependent IEEE division /
// square-root sequences (issue slots 50 % busy at 16 warps per SM), so twice the warps win more than the spills cost.  Measured
// on the C3 stand-in, shade ms per 173 M hits: 1 x 512 @ 108 regs 27.9 | 1 x 640 @ 96 30.2 | 1 x 768 @ 80 27.5 | 1 x 896 @ 72 31.3 |
// 1 x 1024 @ 64 25.9 | 2 x 512 @ 64 24.7.
#ifndef GTLC_SHADE_MIN_BLOCKS
#define GTLC_SHADE_MIN_BLOCKS 2
#endif
constexpr int shade_min_blocks(int maxClosure, bool aov, bool tex, bool vol)
{
  return (GTLC_SHADE_PHASE_SYNC && maxClosure >= GTLC_SHADE_SYNC_MIN_CLOSURE && !aov && !tex && !vol) ? GTLC_SHADE_MIN_BLOCKS : 1;
}
constexpr int shade_block_size(int maxClosure) { return (GTLC_SHADE_PHASE_SYNC && maxClosure >= GTLC_SHADE_SYNC_MIN_CLOSURE) ? GTLC_SHADE_BIG_BLOCK : 256; }   // Lambert / UsdPreviewSurface: small code, no barrier
template <int MAXC, bool NEE, bool AOV, bool TEX, bool VOL>
__global__ void __launch_bounds__(shade_block_size(MAXC), shade_min_blocks(MAXC, AOV, TEX, VOL)) shade_kernel(UniformData ubo, DeviceScene s, PathQueue qin, PathQueue qout,
                                                    const float4* __restrict__ hitTUVP, const uint32_t* __restrict__ hitInst,
                                                    const uint32_t* __restrict__ perm,
                                                    float4* shO, float4* shD, float4* shC, uint32_t* shSeed, float4* sampleRad,
                                                    Counters* ctr, int inIdx, AovSampleBuffers aov)
{
  const unsigned FULL = 0xFFFFFFFFu;
  const uint32_t n = ctr->numPaths[inIdx];
  const uint32_t nRounded = ((n + blockDim.x - 1u) / blockDim.x) * blockDim.x;   // CTA-uniform trip count (warp ballots and the phase barrier need whole CTAs)
  const int lane = threadIdx.x & 31;
  const uint32_t rrBounceOffset = ubo.maxBouncesAndRrBounceOffset & 0xFFFFu;           // rgen:291
  uint32_t maxBounces = ubo.maxBouncesAndRrBounceOffset >> 16; if (maxBounces > kPayloadBouncesMask) maxBounces = kPayloadBouncesMask;
  constexpr bool kPhaseSync = GTLC_SHADE_PHASE_SYNC && MAXC >= GTLC_SHADE_SYNC_MIN_CLOSURE;
  const bool nee = NEE;   // NEXT_EVENT_ESTIMATION is a compile-time permutation in the reference too (GlslShaderGen.cpp:107,247)
  if (blockIdx.x == 0 && threadIdx.x == 0) { ctr->workCounter = 0; ctr->workCounterShadow = 0; }   // trace_closest of this bounce is done, trace_shadow not started

  for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < nRounded; it += gridDim.x * blockDim.x) {
    bool valid = it < n;
    const uint32_t clk0 = (AOV && aov.clk) ? (uint32_t) clock() : 0u;
    bool survives = false, wantShadow = false;
    V3 ray_origin = v3(0.0f), ray_dir = v3(0.0f), throughput = v3(0.0f), radiance = v3(0.0f);
    V3 neeToLight = v3(0.0f), neeContrib = v3(0.0f);
    uint32_t bitfield = 0, rng_state = 0, slot = 0, shadowSeed = 0;
    float nextTMax = kFloatMax; V3 nextWalkPdf = v3(1.0f); uint4 nextMedia = make_uint4(0u, 0u, 0u, 0u);
    // state that lives across the two phases of a hit (BSDF sample | NEE)
    ShadingState st; V3 normal = v3(0.0f), iorCurrent = v3(1.0f), iorOther = v3(1.0f), worldRayDir = v3(0.0f);
    const float* mp = nullptr; const float* lowered = nullptr; int closure = 0; bool doNee = false, isHit = false, isFrontFace = false, thinWalled = false, hasMappedNormal = false;
    V3 mappedNormal = v3(0.0f); uint32_t mediumIdx = 0, hitMaterial = 0;
    bool neeValid = false; V3 neeDir = v3(0.0f), neeRadiance = v3(0.0f); uint32_t neeDiffuseSpecular = 0;
    float memo0 = -1.0f, memo1 = -1.0f;                                                // BsdfSample.memo -> BsdfEval.memo
    float lp[TEX ? GTLC_MATERIAL_PARAM_COUNT + GTLC_BIND_SCRATCH_COUNT : 1];                                     // textured inputs: private parameter block
    uint32_t bounce = 0, loopMediumIdx = 0;                                            // rgen:298, :318 (read before the trace)
    uint32_t media[kMaxMediumStack] = { 0u, 0u, 0u, 0u };
    if (valid) {
      const uint32_t i = perm ? perm[it] : it;
      const float4 ro = qin.rayO[i], rd = qin.rayD[i], th = qin.thr[i], ra = qin.rad[i];
      const float4 h = hitTUVP[i];
      const uint32_t inst = hitInst[i];
      slot = qin.slot[i];
      ray_origin = v3(ro.x, ro.y, ro.z); ray_dir = v3(rd.x, rd.y, rd.z);
      worldRayDir = ray_dir;
      throughput = v3(th.x, th.y, th.z); bitfield = f2u(th.w);
      radiance = v3(ra.x, ra.y, ra.z); rng_state = f2u(ra.w);
      bounce = bitfield & kPayloadBouncesMask;                                         // rgen:298
      V3 walkSegmentPdf = v3(1.0f);
      if (VOL) {
        const uint4 m4 = qin.med[i]; media[0] = m4.x; media[1] = m4.y; media[2] = m4.z; media[3] = m4.w;
        const float4 w4 = qin.walk[i]; walkSegmentPdf = v3(w4.x, w4.y, w4.z);
        loopMediumIdx = payload_medium_idx(bitfield, ubo.mediumStackSize);
      }

      if (inst == 0xFFFFFFFFu && VOL && loopMediumIdx > 0) {
        // ---- rp_main.miss:16-34,57-66: the sampled free-flight distance ended inside the medium ----
        const float distance = rd.w * ubo.metersPerSceneUnit;                          // gl_RayTmaxEXT
        V3 sigma_s, sigma_t; medium_of(s, media[loopMediumIdx - 1], sigma_s, sigma_t);
        const V3 transmittance = v3(expf(sigma_t.x * -distance), expf(sigma_t.y * -distance), expf(sigma_t.z * -distance));
        const V3 density = sigma_t * transmittance;
        const float pdf = dot(walkSegmentPdf, density);
        throughput *= (sigma_s * transmittance) / pdf;
        ray_origin += worldRayDir * distance;
        bitfield |= kPayloadVolumeWalkMissFlag;
        uint32_t wb = bitfield & kPayloadWalkMask;                                     // rp_main_payload.glsl:60-69 (sic: +1 on the unshifted field)
        wb = (wb + 1u < kPayloadWalkMask) ? wb + 1u : kPayloadWalkMask;
        bitfield &= ~kPayloadWalkMask; bitfield |= wb;
      } else if (inst == 0xFFFFFFFFu) {
        // ---- rp_main.miss:55-87 ----
        bitfield |= kPayloadTerminateFlag;                                             // miss:68
        bool useFallbackDomeLight = false;
        if (!(ubo.flags & FLAG_DOME_CAMERA_VISIBLE)) useFallbackDomeLight = (bounce == 0);   // miss:76-80
        const uint32_t domeLightIndex = useFallbackDomeLight ? 0u : 1u;                // miss:82
        const V3 sampleDir = normalize(quatRotateDir(ubo.domeLightRotation, worldRayDir));   // miss:83
        const V3 rdn = sampleDomeLight(ubo, s, domeLightIndex, sampleDir) * ld3(ubo.domeLightEmissionMultiplier);  // miss:84
        radiance += throughput * rdn;                                                  // miss:86
        if (AOV && bounce == 0) aov.a[0][slot] = make_float4(0.0f, 0.0f, 0.0f, u2f(NEE ? 16u : 0u));
      } else {
        // ---- rp_main.chit:132-493 ----
        const uint32_t prim = f2u(h.w);
        const InstanceInfo in = s.instInfo[inst];
        const bool isLeftHanded = (in.flags & 1u) != 0;                                // chit:137
        const bool isDoubleSided = (in.flags & 2u) != 0;                               // chit:138
        mp = s.matParams + (size_t) in.material * GTLC_MATERIAL_PARAM_COUNT;
        closure = s.matClosure[in.material];
        const uint32_t mflags = s.matFlags[in.material];
        isHit = true; hitMaterial = in.material;
        int32_t faceId;
        setup_shading_state(s, in, inst, prim, h.y, h.z, worldRayDir, st, isFrontFace, faceId);   // chit:143
        lowered = s.matLowered + (size_t) in.material * kLobeFloats;
        normal = st.normal;                                                            // chit:147
        mappedNormal = normal;
        if (TEX && (s.matTexRange[in.material].y | s.matPrimvarRange[in.material].y) != 0u) {
          for (int k = 0; k < GTLC_MATERIAL_PARAM_COUNT; k++) lp[k] = mp[k];
          for (int k = 0; k < GTLC_BIND_SCRATCH_COUNT; k++) lp[GTLC_MATERIAL_PARAM_COUNT + k] = 0.0f;
          apply_primvar_bindings(s, in, inst, prim, h.y, h.z, st, lp);
          hasMappedNormal = apply_texture_bindings(s, in.material, worldRayDir, st, lp, mappedNormal);
          mp = lp; lowered = nullptr;                                                  // per-hit inputs: the closure evaluates its constants itself
        }
        mediumIdx = VOL ? loopMediumIdx : payload_medium_idx(bitfield, 0u);            // chit:153
        thinWalled = (mflags & kMatFlagThinWalled) != 0;                               // chit:155-158
        V3 prevMediumIor = v3(1.0f), nextMediumIor = v3(1.0f);                         // chit:161-162
        if (mediumIdx > 0) {                                                           // chit:164-186
          const float distance = h.x * ubo.metersPerSceneUnit;
          if (!VOL) {
            const V3 sigma_a = mat_volume_absorption(closure, mp);
            throughput *= v3(expf(-sigma_a.x * distance), expf(-sigma_a.y * distance), expf(-sigma_a.z * distance));
          } else {
            const uint32_t top = media[mediumIdx - 1];
            V3 sigma_s, sigma_t; medium_of(s, top, sigma_s, sigma_t);
            prevMediumIor = mat_ior(s.matClosure[top], s.matParams + (size_t) top * GTLC_MATERIAL_PARAM_COUNT);
            if (mediumIdx > 1) { const uint32_t below = media[mediumIdx - 2]; nextMediumIor = mat_ior(s.matClosure[below], s.matParams + (size_t) below * GTLC_MATERIAL_PARAM_COUNT); }
            throughput *= v3(expf(-sigma_t.x * distance), expf(-sigma_t.y * distance), expf(-sigma_t.z * distance));
          }
        }
        const float kUseMaterialIor = -1.0f;                                           // mdl_types.glsl:155
        iorCurrent = (isFrontFace || thinWalled) ? prevMediumIor : v3(kUseMaterialIor);            // chit:188
        iorOther = (isFrontFace || thinWalled) ? v3(kUseMaterialIor) : nextMediumIor;              // chit:189
        if (AOV && bounce == 0) {                                                      // chit:192-290
          const V2 clipRange = unpackHalf2x16(ubo.clipRangePacked);                    // chit:227
          const float logDepth = 2.0f * logf(h.x / clipRange.x) / logf(clipRange.y / clipRange.x) - 1.0f;   // chit:228
          const int2 ids = s.instIds[inst];
          const V3 tanAov = (st.tangent_u + v3(1.0f)) * 0.5f, bitAov = (st.tangent_v + v3(1.0f)) * 0.5f;   // chit:208,211
          const V3 posNormal = (normal + v3(1.0f)) * 0.5f;                             // chit:252
          const V3 albedo = bsdf_albedo<MAXC>(closure, mp, lowered, -worldRayDir, st);         // chit:264-280
          const uint32_t fl = 1u | (thinWalled ? 2u : 0u) | (isDoubleSided ? 4u : 0u) | (NEE ? 16u : 0u);
          aov.a[0][slot] = make_float4((1.0f - h.y) - h.z, h.y, h.z, u2f(fl));         // chit:214
          aov.a[1][slot] = make_float4(st.u, st.v, 0.0f, logDepth);                    // chit:217
          aov.a[2][slot] = make_float4(tanAov.x, tanAov.y, tanAov.z, u2f((uint32_t) ids.x));
          aov.a[3][slot] = make_float4(bitAov.x, bitAov.y, bitAov.z, u2f((uint32_t) faceId));
          aov.a[4][slot] = make_float4(posNormal.x, posNormal.y, posNormal.z, u2f((uint32_t) ids.y));
          aov.a[5][slot] = make_float4(albedo.x, albedo.y, albedo.z, (mflags & kMatFlagCutout) ? mat_cutout_opacity(closure, mp) : -1.0f);
        }
        if (mflags & kMatFlagEmissive) {                                               // chit:293-343
          if (isDoubleSided) st.geom_normal *= (isLeftHanded ? -1.0f : 1.0f);
          float edfPdf = 0.0f;
          if (isFrontFace || !isDoubleSided) edfPdf = fmaxf(dot(-worldRayDir, st.normal), 0.0f) * kInvPi;   // df::diffuse_edf
          if (edfPdf > 0.0f) {
            V3 emission_intensity = mat_emission_intensity(closure, mp);
            emission_intensity *= exp2f(ubo.sensorExposure);
            radiance += (throughput * v3(kInvPi)) * emission_intensity;                // chit:335
          }
          if (isDoubleSided) st.geom_normal *= (isLeftHanded ? -1.0f : 1.0f);
        }
      }
    }
    if (kPhaseSync && (GTLC_SHADE_PHASE_MASK & 1)) __syncthreads();                    // phase 2: BSDF importance sampling
    if (valid && isHit) {
      st.normal = normal;                                                            // chit:346
      if (TEX && hasMappedNormal) st.normal = mappedNormal;                          // what *_init() leaves in the state (chit:349-357)
      uint32_t eventType;
      {                                                                              // chit:361-389
        BsdfSample d;
        d.ior1 = iorCurrent; d.ior2 = iorOther; d.k1 = -worldRayDir;
        d.xi = rng1d_next4f(rng_state);                                              // chit:370
        bsdf_sample<MAXC>(closure, mp, lowered, d, st);
        eventType = (uint32_t) d.event_type;
        if (!(TEX && hasMappedNormal)) { memo0 = d.memo[0]; memo1 = d.memo[1]; }      // the NEE evaluation below uses the unmapped normal (chit:397)
        throughput *= d.bsdf_over_pdf;                                               // chit:386
        ray_dir = d.k2;                                                              // chit:388
      }
      const bool isTransmissionEvent = (eventType & BSDF_EVENT_TRANSMISSION) != 0;   // chit:391
      doNee = nee && (eventType & (BSDF_EVENT_DIFFUSE | BSDF_EVENT_GLOSSY)) != 0;    // chit:394 (the NEE block is phases 3 and 4)
      if (!thinWalled && isTransmissionEvent) {                                      // chit:447-480
        if (VOL) {
          if (isFrontFace) {                                                         // push (chit:450-467)
            mediumIdx++;
            if (mediumIdx <= ubo.mediumStackSize) media[mediumIdx - 1] = hitMaterial;
          } else if (mediumIdx > 0) mediumIdx--;                                     // pop (chit:468-471)
        } else mediumIdx = 1u - mediumIdx;                                           // chit:473
        bitfield &= ~kPayloadWalkMask;
        bitfield &= ~kPayloadMediumIdxMask;
        const uint32_t mv = mediumIdx << kPayloadMediumIdxOffset;                    // rp_main_payload.glsl:91-95
        bitfield |= (mv < kPayloadMediumIdxMask ? mv : kPayloadMediumIdxMask);
      }
      if (eventType == BSDF_EVENT_ABSORB) bitfield |= kPayloadTerminateFlag;         // chit:483-486
      const V3 geomNormal = st.geom_normal * (isTransmissionEvent ? -1.0f : 1.0f);   // chit:488
      ray_origin = offset_ray_origin(st.position, geomNormal);                       // chit:489
    }
    if (kPhaseSync && (GTLC_SHADE_PHASE_MASK & 2)) __syncthreads();                    // phase 3: light sampling
    if (valid && doNee) {                                                              // chit:394-444
      st.normal = normal;
      const V4 k4 = rng1d_next4f(rng_state);                                         // chit:404
      LightSample ls;
      sampleLight(ubo, s, k4, st.position, ls);
      neeValid = (ls.dist > 0.0f) && dot(ls.dirToLight, st.geom_normal) > 0.0f;        // chit:415
      neeToLight = ls.dirToLight * ls.dist;                                          // chit:441
      neeDir = ls.dirToLight; neeRadiance = ls.power * ls.invPdf; neeDiffuseSpecular = ls.diffuseSpecularPacked;
    }
    if (kPhaseSync && (GTLC_SHADE_PHASE_MASK & 4)) __syncthreads();                    // phase 4: BSDF evaluation towards the light
    if (valid && doNee) {
      if (neeValid) {
        BsdfEval e;
        e.ior1 = iorCurrent; e.ior2 = iorOther; e.k1 = -worldRayDir; e.k2 = neeDir; e.memo[0] = memo0; e.memo[1] = memo1;
        bsdf_evaluate<MAXC>(closure, mp, lowered, e, st);                                         // chit:425
        neeValid = (e.pdf > 0.0f);
        if (neeValid) {
          const V2 diffuseSpecular = unpackHalf2x16(neeDiffuseSpecular);
          const V3 weight = throughput * neeRadiance;
          neeContrib += (weight * e.bsdf_diffuse) * diffuseSpecular.x;             // chit:436
          neeContrib += (weight * e.bsdf_glossy) * diffuseSpecular.y;              // chit:437
        }
      }
    }
    if (kPhaseSync && (GTLC_SHADE_PHASE_MASK & 8)) __syncthreads();                    // phase 5: shadow-ray set-up, roulette, compaction
    if (valid) {
      // ---- rp_main.rgen:397-480 ----
      shadowSeed = rng_state;                                                          // rgen:399 (before the roulette draw)
      if (nee) {
        const float lightDist = length(neeToLight);                                    // rgen:403
        wantShadow = luminance(neeContrib) > 1e-6f && lightDist > 1e-9f;               // rgen:407
      }
      if (length(throughput) < 1e-9f) bitfield |= kPayloadTerminateFlag;               // rgen:441-444
      if (bounce > rrBounceOffset) {                                                   // rgen:447-459
        const float k1 = rng1d_next1f(rng_state);
        const float max_throughput = fmaxf(throughput.x, fmaxf(throughput.y, throughput.z));   // rgen:101
        const float p = fminf(max_throughput, ubo.rrInvMinTermProb);
        if (k1 > p) bitfield |= kPayloadTerminateFlag;
        else throughput = throughput / p;
      }
      if (VOL && (bitfield & kPayloadVolumeWalkMissFlag) != 0) {                       // rgen:462-477
        const V2 xi = rng1d_next2f(rng_state);
        const uint32_t top = media[loopMediumIdx - 1];
        const float bias = mat_directional_bias(s.matClosure[top], s.matParams + (size_t) top * GTLC_MATERIAL_PARAM_COUNT);
        sampleVolumeScatteringDirection(xi, bias, ray_dir);
        bitfield &= ~kPayloadVolumeWalkMissFlag;
      }
      bitfield++;                                                                      // rgen:480
      const uint32_t nextBounce = bitfield & kPayloadBouncesMask;
      survives = !(bitfield & kPayloadTerminateFlag) && nextBounce < maxBounces;       // rgen:301
      if (VOL && survives) {                                                           // rgen:317-346 of the NEXT loop iteration
        nextWalkPdf = v3(1.0f);
        const uint32_t mi = payload_medium_idx(bitfield, ubo.mediumStackSize);
        if (mi > 0) {
          V3 sigma_s, sigma_t; medium_of(s, media[mi - 1], sigma_s, sigma_t);
          const uint32_t walkLength = (bitfield & kPayloadWalkMask) >> kPayloadWalkOffset;
          const bool hasScattering = sigma_s.x > 0.0f || sigma_s.y > 0.0f || sigma_s.z > 0.0f;
          if (hasScattering && walkLength <= ubo.maxVolumeWalkLength) {
            const V3 albedo = v3(safe_div(sigma_s.x, sigma_t.x), safe_div(sigma_s.y, sigma_t.y), safe_div(sigma_s.z, sigma_t.z));
            const V2 xi = rng1d_next2f(rng_state);
            float sg = sampleDistance(albedo, throughput, sigma_t, xi.x, nextWalkPdf);
            sg *= ubo.metersPerSceneUnit;
            nextTMax = -logf(1.0f - xi.y) / sg;                                        // collision-free distance
          }
        }
        nextMedia = make_uint4(media[0], media[1], media[2], media[3]);
      }
    }

    // ---- compaction: warp ballot + one atomic per warp -------------------------------------------
    const unsigned liveMask = __ballot_sync(FULL, survives);
    uint32_t outPos = 0;
    if (liveMask) {
      const int leader = __ffs(liveMask) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&ctr->numPaths[inIdx ^ 1], (uint32_t) __popc(liveMask));
      base = __shfl_sync(FULL, base, leader);
      outPos = base + (uint32_t) __popc(liveMask & ((1u << lane) - 1u));
    }
    const unsigned shMask = __ballot_sync(FULL, wantShadow);
    uint32_t shPos = 0;
    if (shMask) {
      const int leader = __ffs(shMask) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&ctr->numShadow, (uint32_t) __popc(shMask));
      base = __shfl_sync(FULL, base, leader);
      shPos = base + (uint32_t) __popc(shMask & ((1u << lane) - 1u));
    }
    if (valid) {
      if (survives) {
        qout.rayO[outPos] = make_float4(ray_origin.x, ray_origin.y, ray_origin.z, 0.0f);   // rgen:306
        qout.rayD[outPos] = make_float4(ray_dir.x, ray_dir.y, ray_dir.z, nextTMax);        // rgen:307, :344
        if (VOL) { qout.med[outPos] = nextMedia; qout.walk[outPos] = make_float4(nextWalkPdf.x, nextWalkPdf.y, nextWalkPdf.z, 0.0f); }
        qout.thr[outPos] = make_float4(throughput.x, throughput.y, throughput.z, u2f(bitfield));
        qout.rad[outPos] = make_float4(radiance.x, radiance.y, radiance.z, u2f(rng_state));
        qout.slot[outPos] = slot;
      } else {
        sampleRad[slot] = make_float4(radiance.x, radiance.y, radiance.z, u2f(bitfield & kPayloadBouncesMask));   // .w: bounces (rgen:483-486)
      }
      if (wantShadow) {
        const float lightDist = length(neeToLight);
        const V3 sdir = safe_div(neeToLight, lightDist);                               // rgen:404
        shO[shPos] = make_float4(ray_origin.x, ray_origin.y, ray_origin.z, 0.01f);     // rgen:402,409
        shD[shPos] = make_float4(sdir.x, sdir.y, sdir.z, lightDist);                   // rgen:410
        const uint32_t target = survives ? outPos : (0x80000000u | slot);
        shC[shPos] = make_float4(neeContrib.x, neeContrib.y, neeContrib.z, u2f(target));
        shSeed[shPos] = shadowSeed;                                                    // rgen:399 shadowRayPayload.rng_state
      }
      if (AOV && aov.clk) aov.clk[slot] += (float) ((uint32_t) clock() - clk0);          // ClockCycles AOV: this bounce's shading
    }
  }
}

// ---- resolve: rp_main.rgen:488-498 (+ :506-515 on the last wave) ----------------------------------
__global__ void __launch_bounds__(256) resolve_kernel(UniformData ubo, WaveParams wp, const float4* __restrict__ sampleRad,
                                                      float4* accum, float4* colorAov, int lastWave)
{
  const float invSpp = 1.0f / (float) ubo.spp;                                         // :212
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < wp.wavePixels; q += gridDim.x * blockDim.x) {
    const uint32_t lp = wp.wavePixelBase + q;
    float4 a4 = wp.firstWave ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : accum[lp];
    V3 pixel_color = v3(a4.x, a4.y, a4.z);
    for (uint32_t j = 0; j < wp.waveSamples; j++) {
      const float4 r4 = sampleRad[(size_t) j * wp.wavePixels + q];
      V3 radiance = v3(r4.x, r4.y, r4.z);                                              // :489
      const float maxValue = fmaxf(radiance.x, fmaxf(radiance.y, radiance.z));
      if (maxValue > ubo.maxSampleValue) radiance *= (ubo.maxSampleValue / maxValue);  // :491-494
      const V3 sample_color = vmax(radiance, 0.0f);                                    // :496
      pixel_color += sample_color * invSpp;                                            // :498
    }
    if (!lastWave) { accum[lp] = make_float4(pixel_color.x, pixel_color.y, pixel_color.z, 0.0f); continue; }
    if (!colorAov) continue;                                                           // colour AOV not bound (AOV_MASK & AOV_BIT_COLOR == 0)
    const uint32_t pixelIndex = wp.pixelMap ? wp.pixelMap[lp] : lp;
    V3 prevColor = pixel_color;                                                        // :507
    if ((ubo.flags & FLAG_PROGRESSIVE) && ubo.sampleOffset > 0) { const float4 p4 = colorAov[pixelIndex]; prevColor = v3(p4.x, p4.y, p4.z); }  // :508-513
    const V3 o = (prevColor * (float) ubo.sampleOffset + pixel_color * (float) ubo.spp) * ubo.invTotalSampleCount;   // :514
    colorAov[pixelIndex] = make_float4(o.x, o.y, o.z, wp.alpha);
  }
}

// colormap.glsl:42-53
GTLC_D V3 colormap_inferno(float t)
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
GTLC_D V3 colormap_viridis(float t)
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

// Destination render buffers of the auxiliary AOVs (device pointers, NULL = not bound) and their clear values.
struct AovTargets { void* buf[GTLC_AOV_COUNT]; float4 clear[GTLC_AOV_COUNT]; };

// ---- auxiliary AOV resolve: replays, per pixel and in sample order, the writes the megakernel performs:
// clearAovs (rgen:132-183, only in the first wave of a giRender call), first-bounce writes of every sample
// (chit:191-290; the last sample that hit wins, Normal/Albedo follow the progressive recurrence), Bounces after
// every sample (rgen:483-486), and the final re-normalisation of the normal (rgen:517-520, last wave only).
__global__ void __launch_bounds__(256) resolve_aov_kernel(UniformData ubo, WaveParams wp, const float4* __restrict__ sampleRad,
                                                          AovSampleBuffers smp, AovTargets tg, int firstWave, int lastWave)
{
  uint32_t maxBounces = ubo.maxBouncesAndRrBounceOffset >> 16; if (maxBounces > kPayloadBouncesMask) maxBounces = kPayloadBouncesMask;
  const bool progressive = (ubo.flags & FLAG_PROGRESSIVE) && ubo.sampleOffset > 0;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < wp.wavePixels; q += gridDim.x * blockDim.x) {
    const uint32_t lp = wp.wavePixelBase + q;
    const uint32_t pix = wp.pixelMap ? wp.pixelMap[lp] : lp;
    float4* const* vb = (float4* const*) tg.buf;
    if (firstWave) {
      const int lastWins[10] = { GTLC_AOV_BARYCENTRICS, GTLC_AOV_TEXCOORDS, GTLC_AOV_OPACITY, GTLC_AOV_TANGENTS, GTLC_AOV_BITANGENTS, GTLC_AOV_THIN_WALLED,
                                 GTLC_AOV_DOUBLE_SIDED, GTLC_AOV_BOUNCES, GTLC_AOV_NEE, GTLC_AOV_CLOCK_CYCLES };
#pragma unroll
      for (int k = 0; k < 10; k++) if (tg.buf[lastWins[k]]) vb[lastWins[k]][pix] = tg.clear[lastWins[k]];
      if (tg.buf[GTLC_AOV_OBJECT_ID]) ((int32_t*) tg.buf[GTLC_AOV_OBJECT_ID])[pix] = (int32_t) f2u(tg.clear[GTLC_AOV_OBJECT_ID].x);
      if (tg.buf[GTLC_AOV_FACE_ID]) ((int32_t*) tg.buf[GTLC_AOV_FACE_ID])[pix] = (int32_t) f2u(tg.clear[GTLC_AOV_FACE_ID].x);
      if (tg.buf[GTLC_AOV_INSTANCE_ID]) ((int32_t*) tg.buf[GTLC_AOV_INSTANCE_ID])[pix] = (int32_t) f2u(tg.clear[GTLC_AOV_INSTANCE_ID].x);
      if (tg.buf[GTLC_AOV_DEPTH]) ((float*) tg.buf[GTLC_AOV_DEPTH])[pix] = tg.clear[GTLC_AOV_DEPTH].x;
      if (ubo.sampleOffset == 0) {
        if (tg.buf[GTLC_AOV_NORMAL]) vb[GTLC_AOV_NORMAL][pix] = tg.clear[GTLC_AOV_NORMAL];
        if (tg.buf[GTLC_AOV_ALBEDO]) vb[GTLC_AOV_ALBEDO][pix] = tg.clear[GTLC_AOV_ALBEDO];
      }
    }
    if (smp.clk && tg.buf[GTLC_AOV_CLOCK_CYCLES]) {                                       // rgen:501-504: cycles this pixel's invocation took
      float cycles = firstWave ? 0.0f : vb[GTLC_AOV_CLOCK_CYCLES][pix].x;                // (summed over the waves of one giRender call)
      for (uint32_t j = 0; j < wp.waveSamples; j++) cycles += smp.clk[(size_t) j * wp.wavePixels + q];
      vb[GTLC_AOV_CLOCK_CYCLES][pix] = make_float4(cycles, 0.0f, 0.0f, tg.clear[GTLC_AOV_CLOCK_CYCLES].w);
    }
    for (uint32_t j = 0; j < wp.waveSamples; j++) {
      const size_t slot = (size_t) j * wp.wavePixels + q;
      const float4 a0 = smp.a[0][slot];
      const uint32_t fl = f2u(a0.w);
      if (fl & 1u) {
        const float4 a1 = smp.a[1][slot], a2 = smp.a[2][slot], a3 = smp.a[3][slot], a4 = smp.a[4][slot], a5 = smp.a[5][slot];
        auto put3 = [&](int id, float x, float y, float z) { if (tg.buf[id]) { float4 c = vb[id][pix]; c.x = x; c.y = y; c.z = z; vb[id][pix] = c; } };
        if (a5.w < 0.0f) put3(GTLC_AOV_OPACITY, 1.0f, 0.0f, 0.0f);                    // chit:202 (no cutout)
        else { const V3 vc = (a5.w == 0.0f) ? v3(1.0f) : colormap_viridis(a5.w); put3(GTLC_AOV_OPACITY, vc.x, vc.y, vc.z); }   // ahit:48 for the accepted hit
        put3(GTLC_AOV_TANGENTS, a2.x, a2.y, a2.z); put3(GTLC_AOV_BITANGENTS, a3.x, a3.y, a3.z);
        put3(GTLC_AOV_BARYCENTRICS, a0.x, a0.y, a0.z); put3(GTLC_AOV_TEXCOORDS, a1.x, a1.y, a1.z);
        if (fl & 2u) put3(GTLC_AOV_THIN_WALLED, 1.0f, 0.0f, 0.0f); else put3(GTLC_AOV_THIN_WALLED, 0.0f, 1.0f, 0.0f);     // chit:220
        if (fl & 4u) put3(GTLC_AOV_DOUBLE_SIDED, 0.0f, 1.0f, 0.0f); else put3(GTLC_AOV_DOUBLE_SIDED, 1.0f, 0.0f, 0.0f);   // chit:247
        if (tg.buf[GTLC_AOV_OBJECT_ID]) ((int32_t*) tg.buf[GTLC_AOV_OBJECT_ID])[pix] = (int32_t) f2u(a2.w);
        if (tg.buf[GTLC_AOV_FACE_ID]) ((int32_t*) tg.buf[GTLC_AOV_FACE_ID])[pix] = (int32_t) f2u(a3.w);
        if (tg.buf[GTLC_AOV_INSTANCE_ID]) ((int32_t*) tg.buf[GTLC_AOV_INSTANCE_ID])[pix] = (int32_t) f2u(a4.w);
        if (tg.buf[GTLC_AOV_DEPTH]) ((float*) tg.buf[GTLC_AOV_DEPTH])[pix] = a1.w;
        auto accumulate = [&](int id, V3 cur) {                                        // chit:251-261, 263-289
          if (!tg.buf[id]) return;
          float4 c = vb[id][pix];
          V3 prev = cur;
          if (progressive) prev = v3(c.x, c.y, c.z);
          const V3 o = (prev * (float) ubo.sampleOffset + cur * (float) ubo.spp) * ubo.invTotalSampleCount;
          c.x = o.x; c.y = o.y; c.z = o.z; vb[id][pix] = c;
        };
        accumulate(GTLC_AOV_NORMAL, v3(a4.x, a4.y, a4.z));
        accumulate(GTLC_AOV_ALBEDO, v3(a5.x, a5.y, a5.z));
      }
      if ((fl & 16u) && tg.buf[GTLC_AOV_NEE]) {                                        // rgen:431-436
        float4 c = vb[GTLC_AOV_NEE][pix]; c.x = (fl & 8u) ? 1.0f : 0.0f; c.y = (fl & 8u) ? 0.0f : 1.0f; c.z = 0.0f; vb[GTLC_AOV_NEE][pix] = c;
      }
      if (tg.buf[GTLC_AOV_BOUNCES]) {                                                  // rgen:483-486
        const V3 cm = colormap_inferno((float) f2u(sampleRad[slot].w) / (float) maxBounces);
        float4 c = vb[GTLC_AOV_BOUNCES][pix]; c.x = cm.x; c.y = cm.y; c.z = cm.z; vb[GTLC_AOV_BOUNCES][pix] = c;
      }
    }
    if (lastWave && tg.buf[GTLC_AOV_NORMAL]) {                                         // rgen:517-520
      float4 c = vb[GTLC_AOV_NORMAL][pix];
      const V3 n = v3(c.x, c.y, c.z) * 2.0f - v3(1.0f);
      const V3 r = (normalize(n) + v3(1.0f)) * 0.5f;
      c.x = r.x; c.y = r.y; c.z = r.z; vb[GTLC_AOV_NORMAL][pix] = c;
    }
  }
}

} // namespace gtlc
