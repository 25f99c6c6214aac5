// C-ABI implementation (include/gtlc.h): object store, scene flattening + BVH8 build/upload, and the
// host side of the wavefront render loop.  Replaces the bodies of the reference's gi* functions
// (/root/reference/src/gi/impl/Gi.cpp) and the cgpu/ggpu Vulkan layers underneath them with CUDA
// runtime calls: cudaMalloc'd SoA queues, one stream, CUDA events for per-kernel timing.
#include "../../include/gtlc.h"
#include "bvh8.h"
#include "wavefront.cuh"
#include "sort.cuh"
#include "bvh8_gpu.cuh"

static_assert(gtlc::kStackSize == (int) gtlc::kBvh8MaxDepth, "traversal stack and builder depth limit must agree");

#include <dlfcn.h>
#include <nccl.h>   // types only: the library is dlopen'ed (gtlc_comm_*)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

using namespace gtlc;

namespace {

// ---------------------------------------------------------------------------------------------
// error handling (reference: GB_ERROR + GiStatus::Error, no exceptions across the boundary)
// ---------------------------------------------------------------------------------------------
std::mutex g_errMutex;
std::string g_lastError;
uint64_t g_errorCount = 0;

void setError(const char* fmt, ...)
{
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  std::lock_guard<std::mutex> g(g_errMutex);
  g_lastError = buf;
  g_errorCount++;
  fprintf(stderr, "[gtlc] error: %s\n", buf);
}

#define CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { setError("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); return false; } } while (0)

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
enum : uint32_t { TAG_SCENE = 0x53434e45, TAG_MATERIAL = 0x4d41544c, TAG_MESH = 0x4d455348, TAG_SPHERE = 0x5350484c,
                  TAG_DISTANT = 0x44495354, TAG_RECT = 0x52454354, TAG_DISK = 0x4449534b, TAG_DOME = 0x444f4d45, TAG_RB = 0x52424620, TAG_TEXTURE = 0x54455854 };

struct Object { uint32_t tag; };

struct Scene;

// Textures are shared between materials (and scenes): reference counted, the creator holds one reference.
struct Texture : Object { uint32_t width, height, depth; int32_t format; void* dTexels; int refs; };
struct Material : Object { Scene* scene; int32_t closure; float params[GTLC_MATERIAL_PARAM_COUNT]; std::string name;
                           std::vector<GtlcTextureBinding> bindings; std::vector<Texture*> boundTextures;
                           struct PrimvarBinding { int32_t paramIndex, componentCount; std::string name; float fallback[4]; };
                           std::vector<PrimvarBinding> primvarBindings;
                           int opacityBinding = -1;   // index into `bindings` of the binding that feeds the cutout input (textured cutout), or -1
                           int opacityPrimvar = -1;   // index into `primvarBindings` of the reader that feeds the cutout input (e.g. displayOpacity), or -1
                         };

struct Mesh : Object {
  Scene* scene;
  std::vector<GtlcVertex> vertices;
  std::vector<GtlcFace> faces;
  std::vector<int32_t> faceIds;
  int32_t id; bool doubleSided, flipFacing, visible;
  std::string name;
  float transform[12];                       // row-major 3x4 (Gi.cpp:641-650)
  std::vector<float> instanceTransforms;     // 12 per instance
  std::vector<int32_t> instanceIds;
  Material* material;
  struct Primvar { std::string name; int32_t type, interpolation; std::vector<uint8_t> data; };   // GiPrimvarData
  std::vector<Primvar> primvars, instancerPrimvars;
};

// rp_main.h:73-113, 48 bytes each
struct SphereLightData { float pos[3]; uint32_t diffuseSpecularPacked; float baseEmission[3]; float area; float radiusXYZ[3]; float padding; };
struct DistantLightData { float direction[3]; float angle; float baseEmission[3]; uint32_t diffuseSpecularPacked; float padding[3]; float invPdf; };
struct RectLightData { float origin[3]; float width; float baseEmission[3]; float height; uint32_t tangentFramePacked[2]; uint32_t diffuseSpecularPacked; float padding; };
struct DiskLightData { float origin[3]; float radiusX; float baseEmission[3]; float radiusY; uint32_t tangentFramePacked[2]; uint32_t diffuseSpecularPacked; float padding; };
static_assert(sizeof(SphereLightData) == 48 && sizeof(DistantLightData) == 48 && sizeof(RectLightData) == 48 && sizeof(DiskLightData) == 48, "light structs are 48 bytes");

template <typename T> struct Light : Object { Scene* scene; T data; };
using SphereLight = Light<SphereLightData>;
using DistantLight = Light<DistantLightData>;
using RectLight = Light<RectLightData>;
using DiskLight = Light<DiskLightData>;

struct DomeLight : Object { Scene* scene; uint32_t width, height; std::vector<float> texels; float rotation[4]; float baseEmission[3]; float diffuse, specular; float4* dTexels; };

struct RenderBuffer : Object { uint32_t width, height; int32_t format; size_t size; void* hostMem; void* devMem; };

template <typename T> struct DeviceArray {
  T* ptr = nullptr; size_t capacity = 0;
  bool reserve(size_t n) { if (n <= capacity) return true; release(); CUDA_OK(cudaMalloc(&ptr, std::max<size_t>(n, 1) * sizeof(T))); capacity = std::max<size_t>(n, 1); return true; }
  bool upload(const T* src, size_t n, cudaStream_t st) { if (!reserve(n)) return false; if (n) CUDA_OK(cudaMemcpyAsync(ptr, src, n * sizeof(T), cudaMemcpyHostToDevice, st)); return true; }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; capacity = 0; }
};

struct Scene : Object {
  std::mutex mutex;                            // Gi.cpp:224
  std::unordered_set<Mesh*> meshSet;           // ownership only
  std::vector<Mesh*> meshes;                   // creation order (deterministic flattening)
  std::vector<Material*> materials;
  std::vector<SphereLight*> sphereLights;
  std::vector<DistantLight*> distantLights;
  std::vector<RectLight*> rectLights;
  std::vector<DiskLight*> diskLights;
  bool geometryDirty = true, lightsDirty = true, materialsDirty = true, framebufferDirty = true;
  uint32_t sampleOffset = 0;                   // Gi.cpp:230
  // multi-GPU shard
  uint32_t shardRank = 0, shardCount = 1, shardTile = 64;
  // flattened host copies
  Bvh8 bvh; bool bvhValid = false;        // host copy of the BVH8 (stale while bvhHostStale: built on the device, not yet read back)
  bool bvhHostStale = false; uint32_t nodeCount = 0, triCount = 0, maxDepth = 0, totalTris = 0;
  std::vector<float> hInstOpacity; DeviceArray<float> dInstOpacity;
  DeviceArray<uint32_t> dMatOpacityBinding; DeviceArray<float> dMatOpacityConst; DeviceArray<uint2> dMatOpacityPrimvar; DeviceArray<float> dPrimvarFallback; DeviceArray<CutoutScene> dCutoutScene;   // textured cutouts
  std::vector<uint2> hPrimvarInfos; std::vector<float> hPrimvarData; DeviceArray<uint2> dPrimvarInfos; DeviceArray<float> dPrimvarData;   // R15 scene data
  DeviceArray<DevicePrimvarBinding> dPrimvarBindings; DeviceArray<uint2> dMatPrimvarRange;
  std::vector<float> hFverts; std::vector<uint32_t> hFaces;
  std::vector<float> hO2W, hW2O; std::vector<InstanceInfo> hInstInfo; std::vector<int2> hInstIds; std::vector<int32_t> hFaceIds; std::vector<uint32_t> hRankBase;
  std::vector<Material*> flatMaterials;
  // device
  DeviceArray<uint4> dNodes; DeviceArray<float4> dTris;     // canonical BVH8 (bvh8.h): what the builders write and gtlc_scene_get_bvh returns
  DeviceArray<uint4> dNodesT; DeviceArray<float4> dTrisT;   // traversal layout (traverse.cuh): 96-byte nodes, 64-byte triangles
  DeviceArray<float4> dFverts; DeviceArray<uint32_t> dFaces;   // the reference's packed FVertex / Face arrays (canonical; input of the device BVH builders)
  DeviceArray<uint4> dSverts, dSfaces; DeviceArray<float> dMatLowered;   // shading records + lowered materials (wavefront.cuh: DeviceScene)
  struct MeshRange { uint32_t faceBase, faceCount, vertexBase; }; std::vector<MeshRange> meshRanges;
  DeviceArray<float4> dO2W, dW2O; DeviceArray<InstanceInfo> dInstInfo; DeviceArray<int2> dInstIds; DeviceArray<int32_t> dFaceIds; DeviceArray<uint32_t> dRankBase;
  DeviceArray<float> dMatParams; DeviceArray<int32_t> dMatClosure; DeviceArray<uint32_t> dMatFlags;
  DeviceArray<DeviceTexture> dTextures; DeviceArray<DeviceTexBinding> dTexBindings; DeviceArray<uint2> dMatTexRange; bool hasTextures = false, hasTexturedCutout = false;
  DeviceArray<float4> dSphere, dDistant, dRect, dDisk;
  DeviceArray<uint32_t> dPixelMap; uint32_t pixelMapW = 0, pixelMapH = 0, pixelMapRank = 0, pixelMapCount = 1, pixelMapTile = 0, localPixels = 0;
  double lastBuildMs = 0.0;
  GtlcRenderParams oldParams; bool haveOldParams = false; std::vector<GtlcAovBinding> oldBindings;
};

struct Global {
  bool initialized = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  int smCount = 148;
  uint32_t currentTriCount = 0;   // of the scene about to be traced (syncSceneToDevice)
  GtlcTuning tuning;
  GtlcStats stats;
  // path-state pool
  size_t poolCapacity = 0, volumeCapacity = 0;
  PathQueue q[2];
  float4* hitTUVP = nullptr; uint32_t* hitInst = nullptr;
  float4 *shO = nullptr, *shD = nullptr, *shC = nullptr; uint32_t* shSeed = nullptr;
  float4* sampleRad = nullptr;
  float4* accum = nullptr; size_t accumCapacity = 0;
  Counters* counters = nullptr;
  uint32_t* visits = nullptr; size_t visitsCapacity = 0;
  uint32_t* perm = nullptr; uint32_t* sortHist = nullptr; int sortBlocks = 0;
  AovSampleBuffers aovSamples; size_t aovCapacity = 0;
  std::vector<cudaEvent_t> events; size_t eventsUsed = 0;
  // multi-GPU: one NCCL communicator per process, entry points resolved from libnccl.so.2 at gtlc_comm_init
  struct Nccl {
    void* lib = nullptr; ncclComm_t comm = nullptr; int rank = 0, world = 0;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
  } nccl;
  float4* reduced = nullptr; size_t reducedCapacity = 0;   // root: the reduced colour AOV (the shard buffers keep foreign pixels at 0)
  Global() { memset(&aovSamples, 0, sizeof(aovSamples)); memset(&tuning, 0, sizeof(tuning)); tuning.maxPathsPerWave = 1u << 24; tuning.sortByMaterial = -1; memset(&stats, 0, sizeof(stats)); memset(q, 0, sizeof(q)); }
};
Global g;
std::mutex g_destroyMutex;   // Gi.cpp:249
std::mutex g_renderMutex;    // the path-state pools, counters and the stream are per process: renders / ray queries of different scenes serialise

template <typename T> T* as(GtlcHandle h, uint32_t tag)
{
  Object* o = reinterpret_cast<Object*>(static_cast<uintptr_t>(h));
  if (!o || o->tag != tag) { if (h) setError("invalid handle (expected tag %08x)", tag); return nullptr; }
  return static_cast<T*>(o);
}
template <typename T> GtlcHandle handleOf(T* p) { return static_cast<GtlcHandle>(reinterpret_cast<uintptr_t>(static_cast<Object*>(p))); }

// ---------------------------------------------------------------------------------------------
// host encoders: Gi.cpp:287-300, glm::packHalf2x16
// ---------------------------------------------------------------------------------------------
uint32_t packUnorm2x16(float x, float y)
{
  auto q = [](float v) { v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); return (uint32_t) roundf(v * 65535.0f); };
  return q(x) | (q(y) << 16);
}
uint32_t encodeDirection(const float* d)   // _EncodeDirection, Gi.cpp:294-300
{
  float len = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
  float inv = 1.0f / len;
  float v[3] = { d[0] * inv, d[1] * inv, d[2] * inv };
  float s = (fabsf(v[0]) + fabsf(v[1])) + fabsf(v[2]);
  v[0] /= s; v[1] /= s; v[2] /= s;
  float ex, ey;
  if (v[2] < 0.0f) { ex = (1.0f - fabsf(v[1])) * (v[0] >= 0.0f ? 1.0f : -1.0f); ey = (1.0f - fabsf(v[0])) * (v[1] >= 0.0f ? 1.0f : -1.0f); }
  else { ex = v[0]; ey = v[1]; }
  return packUnorm2x16(ex * 0.5f + 0.5f, ey * 0.5f + 0.5f);
}
uint16_t floatToHalf(float f)
{
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u, mant = x & 0x007FFFFFu; int32_t exp = (int32_t) ((x >> 23) & 0xFF);
  if (exp == 255) return (uint16_t) (sign | 0x7C00u | (mant ? 0x200u : 0u));
  int32_t e = exp - 127 + 15;
  if (e >= 31) return (uint16_t) (sign | 0x7C00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t) sign;
    mant |= 0x00800000u;
    uint32_t shift = (uint32_t) (14 - e), h = mant >> shift, rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) h++;
    return (uint16_t) (sign | h);
  }
  uint32_t h = ((uint32_t) e << 10) | (mant >> 13), rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
  return (uint16_t) (sign | h);
}
uint32_t packHalf2x16(float a, float b) { return (uint32_t) floatToHalf(a) | ((uint32_t) floatToHalf(b) << 16); }

// USD row-major float[16] (row-vector convention) -> row-major 3x4 for column vectors (Gi.cpp:641-650)
void usdToRow3x4(const float* m, float* out)
{
  for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) out[r * 4 + c] = m[c * 4 + r];
}
// world = instance o mesh (Gi.cpp:1191), evaluated in float64 in a fixed order, rounded once
void composeTransforms(const float* meshT, const float* instT, float* out)
{
  double a[16], b[16], r[16];
  auto expand = [](const float* t, double* m) { for (int i = 0; i < 12; i++) m[i] = t[i]; m[12] = m[13] = m[14] = 0.0; m[15] = 1.0; };
  expand(instT, a); expand(meshT, b);   // column-vector convention: world = I * M
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
    double s = a[i * 4 + 0] * b[0 * 4 + j];
    s = s + a[i * 4 + 1] * b[1 * 4 + j];
    s = s + a[i * 4 + 2] * b[2 * 4 + j];
    s = s + a[i * 4 + 3] * b[3 * 4 + j];
    r[i * 4 + j] = s;
  }
  for (int i = 0; i < 12; i++) out[i] = (float) r[i];
}
void invertAffine(const float* m, float* out)
{
  double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], gg = m[8], h = m[9], i = m[10];
  double A = e * i - f * h, B = -(d * i - f * gg), C = d * h - e * gg;
  double det = a * A + b * B + c * C;
  double id = 1.0 / det;
  double r[9] = { A * id, -(b * i - c * h) * id, (b * f - c * e) * id,
                  B * id, (a * i - c * gg) * id, -(a * f - c * d) * id,
                  C * id, -(a * h - b * gg) * id, (a * e - b * d) * id };
  double tx = m[3], ty = m[7], tz = m[11];
  for (int k = 0; k < 3; k++) {
    out[k * 4 + 0] = (float) r[k * 3 + 0]; out[k * 4 + 1] = (float) r[k * 3 + 1]; out[k * 4 + 2] = (float) r[k * 3 + 2];
    out[k * 4 + 3] = (float) (-(r[k * 3 + 0] * tx + r[k * 3 + 1] * ty + r[k * 3 + 2] * tz));
  }
}

// ---------------------------------------------------------------------------------------------
// scene flattening + BVH8 (replaces _giBuildGeometryStructures/_giCreateBvh, Gi.cpp:784-1315)
// ---------------------------------------------------------------------------------------------
bool buildHostGeometry(Scene* s, bool hostBvh = true)
{
  auto t0 = std::chrono::steady_clock::now();
  s->meshRanges.clear(); s->hFverts.clear(); s->hFaces.clear(); s->hO2W.clear(); s->hW2O.clear(); s->hInstInfo.clear(); s->hInstIds.clear(); s->hFaceIds.clear(); s->hRankBase.clear(); s->hInstOpacity.clear(); s->hPrimvarInfos.clear(); s->hPrimvarData.clear(); s->flatMaterials.clear();
  // world-space triangles for the host builder: one job per (instance, block of faces), filled by the whole team below
  struct FlattenJob { const Mesh* mesh; uint32_t inst, firstFace, lastFace, firstTri; float opacity; float o2w[12]; };
  std::vector<FlattenJob> jobs;
  uint64_t wanted = 0;
  for (Mesh* m : s->meshes) if (m->visible && m->material) wanted += (uint64_t) m->faces.size() * (m->instanceTransforms.size() / 12);
  if (wanted >= 0x80000000ull) { setError("scene flattens to %llu triangles; indices are 32-bit (2^31 at most)", (unsigned long long) wanted); return false; }
  uint32_t rank = 0;
  for (Mesh* m : s->meshes) {
    if (!m->visible) continue;                                          // Gi.cpp:801-804
    if (!m->material) { setError("invalid BLAS material (mesh '%s' has no material)", m->name.c_str()); continue; }   // Gi.cpp:819-823
    if (m->faces.empty()) continue;                                     // Gi.cpp:833-836
    uint32_t matIndex = 0;
    for (; matIndex < s->flatMaterials.size(); matIndex++) if (s->flatMaterials[matIndex] == m->material) break;
    if (matIndex == s->flatMaterials.size()) s->flatMaterials.push_back(m->material);
    const uint32_t vertexBase = (uint32_t) (s->hFverts.size() / 8);
    const uint32_t faceBase = (uint32_t) (s->hFaces.size() / 3);
    s->meshRanges.push_back({ faceBase, (uint32_t) m->faces.size(), vertexBase });
    {                                                                   // FVertex packing, Gi.cpp:845-859
      const size_t nv = m->vertices.size(), nf = m->faces.size(), fv0 = s->hFverts.size(), ff0 = s->hFaces.size(), fi0 = s->hFaceIds.size();
      s->hFverts.resize(fv0 + nv * 8); s->hFaces.resize(ff0 + nf * 3); s->hFaceIds.resize(fi0 + nf);
      float* fv = s->hFverts.data() + fv0; uint32_t* ff = s->hFaces.data() + ff0; int32_t* fi = s->hFaceIds.data() + fi0;
      const bool ownIds = m->faceIds.size() == nf;
#pragma omp parallel for schedule(static) if (nv > 50000)
      for (int64_t i = 0; i < (int64_t) nv; i++) {
        const GtlcVertex& v = m->vertices[i];
        const uint32_t en = encodeDirection(v.norm), et = encodeDirection(v.tangent);
        float* f = fv + i * 8;
        f[0] = v.pos[0]; f[1] = v.pos[1]; f[2] = v.pos[2]; f[3] = v.bitangentSign; f[6] = v.u; f[7] = v.v;
        memcpy(&f[4], &en, 4); memcpy(&f[5], &et, 4);
      }
#pragma omp parallel for schedule(static) if (nf > 50000)
      for (int64_t i = 0; i < (int64_t) nf; i++) {
        ff[i * 3 + 0] = m->faces[i].v_i[0]; ff[i * 3 + 1] = m->faces[i].v_i[1]; ff[i * 3 + 2] = m->faces[i].v_i[2];
        fi[i] = ownIds ? m->faceIds[i] : (int32_t) i;                    // rp_main.chit:232-240
      }
    }
    // scene data of this mesh for its material's primvar bindings (Gi.cpp:905-1018): mesh primvars override instancer primvars
    const uint32_t meshSlot = (uint32_t) (s->hPrimvarInfos.size() / GTLC_MAX_PRIMVAR_BINDINGS);
    for (int k = 0; k < GTLC_MAX_PRIMVAR_BINDINGS; k++) {
      uint2 info = make_uint2(0u, 0u);
      if ((size_t) k < m->material->primvarBindings.size()) {
        const std::string& want = m->material->primvarBindings[k].name;
        const Mesh::Primvar* found = nullptr;
        for (const Mesh::Primvar& p : m->instancerPrimvars) if (p.name == want && !p.data.empty()) { found = &p; break; }
        for (const Mesh::Primvar& p : m->primvars) if (p.name == want && !p.data.empty()) { found = &p; break; }
        if (found && found->type >= GTLC_PRIMVAR_FLOAT && found->type <= GTLC_PRIMVAR_INT4) {
          const bool isInt = found->type >= GTLC_PRIMVAR_INT;        // read through scene_data_lookup_int (no interpolation), stored as its 32-bit patterns
          const uint32_t stride = (uint32_t) (isInt ? found->type - GTLC_PRIMVAR_INT : found->type) + 1u;
          size_t need = 1;                                          // elements the lookup may index (get_scene_data_indices)
          if (found->interpolation == GTLC_INTERP_VERTEX) need = m->vertices.size();
          else if (found->interpolation == GTLC_INTERP_UNIFORM) need = m->faces.size();
          else if (found->interpolation == GTLC_INTERP_INSTANCE) {
            const size_t ic = m->instanceTransforms.size() / 12;
            for (size_t q = 0; q < ic; q++) { const int32_t id = q < m->instanceIds.size() ? m->instanceIds[q] : (int32_t) q; if (id >= 0) need = std::max(need, (size_t) id + 1); }
          }
          if (found->data.size() < need * stride * 4) { setError("primvar '%s' of mesh '%s' holds %zu bytes, its interpolation needs %zu: ignored", want.c_str(), m->name.c_str(), found->data.size(), need * stride * 4); s->hPrimvarInfos.push_back(info); continue; }
          info = make_uint2((uint32_t) s->hPrimvarData.size(), stride | ((uint32_t) found->interpolation << 4) | 0x100u | (isInt ? 0x200u : 0u));
          const float* src = reinterpret_cast<const float*>(found->data.data());
          s->hPrimvarData.insert(s->hPrimvarData.end(), src, src + found->data.size() / 4);
        }
      }
      s->hPrimvarInfos.push_back(info);
    }
    const uint32_t flags = (m->flipFacing ? 1u : 0u) | (m->doubleSided ? 2u : 0u) | (meshSlot << 8);   // Gi.cpp:1135-1143; bits 8.. = row of the scene-data table
    const size_t instCount = m->instanceTransforms.size() / 12;
    for (size_t k = 0; k < instCount; k++) {                            // Gi.cpp:1188-1202
      float o2w[12], w2o[12];
      composeTransforms(m->transform, &m->instanceTransforms[k * 12], o2w);
      invertAffine(o2w, w2o);
      const uint32_t inst = (uint32_t) s->hInstInfo.size();
      s->hO2W.insert(s->hO2W.end(), o2w, o2w + 12);
      s->hW2O.insert(s->hW2O.end(), w2o, w2o + 12);
      s->hInstInfo.push_back({ vertexBase, faceBase, flags, matIndex });
      s->hRankBase.push_back(rank);   // rank of (inst, prim) = rankBase[inst] + prim, ascending in (inst, prim)
      s->hInstOpacity.push_back((m->material->opacityBinding >= 0 || m->material->opacityPrimvar >= 0) ? -1.0f : mat_cutout_opacity(m->material->closure, m->material->params));   // < 0: evaluated per hit
      rank += (uint32_t) m->faces.size();
      s->hInstIds.push_back(make_int2(m->id, k < m->instanceIds.size() ? m->instanceIds[k] : (int32_t) k));
      if (!hostBvh) continue;   // the device builder flattens on the GPU (lbvh_flatten_kernel, same arithmetic)
      const uint32_t faceCount = (uint32_t) m->faces.size(), firstTri = rank - faceCount;
      for (uint32_t f0 = 0; f0 < faceCount; f0 += 65536u) {
        FlattenJob j; j.mesh = m; j.inst = inst; j.firstFace = f0; j.lastFace = std::min(faceCount, f0 + 65536u); j.firstTri = firstTri + f0;
        j.opacity = s->hInstOpacity.back(); memcpy(j.o2w, o2w, sizeof(o2w));
        jobs.push_back(j);
      }
    }
  }
  s->totalTris = rank;
  if (hostBvh) {
    RawVector<BuildTri> tris((size_t) rank);      // triangle index = rankBase[inst] + prim
#pragma omp parallel for schedule(dynamic, 1) if (rank > 50000)
    for (int64_t ji = 0; ji < (int64_t) jobs.size(); ji++) {
      const FlattenJob& j = jobs[(size_t) ji];
      const float* o2w = j.o2w;
      for (uint32_t f = j.firstFace; f < j.lastFace; f++) {
        BuildTri t; t.prim = f; t.inst = j.inst; t.opacity = j.opacity;
        for (int c = 0; c < 3; c++) {
          const float* p = j.mesh->vertices[j.mesh->faces[f].v_i[c]].pos;
          t.v[c][0] = ((o2w[0] * p[0] + o2w[1] * p[1]) + o2w[2] * p[2]) + o2w[3];
          t.v[c][1] = ((o2w[4] * p[0] + o2w[5] * p[1]) + o2w[6] * p[2]) + o2w[7];
          t.v[c][2] = ((o2w[8] * p[0] + o2w[9] * p[1]) + o2w[10] * p[2]) + o2w[11];
        }
        tris[(size_t) j.firstTri + (f - j.firstFace)] = t;
      }
    }
    bvh8_build(tris.data(), tris.size(), s->bvh);
    if (s->bvh.maxDepth > (uint32_t) kStackSize) { setError("BVH8 depth %u exceeds the traversal stack (%d entries)", s->bvh.maxDepth, kStackSize); return false; }
    s->nodeCount = (uint32_t) s->bvh.nodes.size(); s->triCount = (uint32_t) (s->bvh.tris.size() / 12); s->maxDepth = s->bvh.maxDepth;
    s->bvhHostStale = false;
  }
  s->bvhValid = true;
  s->lastBuildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("GTLC_BUILD_TIMING")) fprintf(stderr, "[gtlc] flatten + host BVH %.0f ms (BVH build alone %.0f ms) for %u triangles\n", s->lastBuildMs, hostBvh ? s->bvh.buildSeconds * 1e3 : 0.0, s->totalTris);
  return true;
}

bool buildBvhOnDevice(Scene* s);

// Canonical BVH8 in HBM (host-built and uploaded, or built on the device) -> the layout the traversal kernels read.
bool repackBvh(Scene* s)
{
  cudaStream_t st = g.stream;
  const uint32_t nodeCount = s->nodeCount, triCount = s->triCount;
  if (!s->dNodesT.reserve((size_t) std::max(nodeCount, 1u) * 6) || !s->dTrisT.reserve((size_t) std::max(triCount, 1u) * 4)) return false;
  uint32_t* dErr = &g.counters->repackErrors;
  CUDA_OK(cudaMemsetAsync(dErr, 0, sizeof(uint32_t), st));
  if (nodeCount) bvh8_repack_nodes_kernel<<<std::min<uint32_t>((nodeCount + 255u) / 256u, (uint32_t) g.smCount * 16u), 256, 0, st>>>(s->dNodes.ptr, nodeCount, s->dNodesT.ptr, dErr);
  if (triCount) bvh8_repack_tris_kernel<<<std::min<uint32_t>((triCount + 255u) / 256u, (uint32_t) g.smCount * 16u), 256, 0, st>>>(s->dTris.ptr, triCount, s->dRankBase.ptr, s->dTrisT.ptr);
  uint32_t errors = 0;
  CUDA_OK(cudaMemcpyAsync(&errors, dErr, sizeof(errors), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
  if (errors) { setError("BVH8 re-pack: %u nodes do not number their triangles in slot order", errors); return false; }
  g.stats.kernelLaunches += 2;
  return true;
}

bool uploadGeometry(Scene* s, bool hostBvh = true)
{
  cudaStream_t st = g.stream;
  auto t0 = std::chrono::steady_clock::now();
  if (hostBvh) {
    if (!s->dNodes.upload((const uint4*) s->bvh.nodes.data(), s->bvh.nodes.size() * 5, st)) return false;
    if (!s->dTris.upload((const float4*) s->bvh.tris.data(), s->bvh.tris.size() / 4, st)) return false;
  }
  if (!s->dInstOpacity.upload(s->hInstOpacity.data(), s->hInstOpacity.size(), st)) return false;
  if (!s->dPrimvarInfos.upload(s->hPrimvarInfos.data(), s->hPrimvarInfos.size(), st)) return false;
  if (!s->dPrimvarData.upload(s->hPrimvarData.data(), s->hPrimvarData.size(), st)) return false;
  if (!s->dFverts.upload((const float4*) s->hFverts.data(), s->hFverts.size() / 4, st)) return false;
  if (!s->dFaces.upload(s->hFaces.data(), s->hFaces.size(), st)) return false;
  if (!s->dO2W.upload((const float4*) s->hO2W.data(), s->hO2W.size() / 4, st)) return false;
  if (!s->dW2O.upload((const float4*) s->hW2O.data(), s->hW2O.size() / 4, st)) return false;
  if (!s->dInstInfo.upload(s->hInstInfo.data(), s->hInstInfo.size(), st)) return false;
  if (!s->dInstIds.upload(s->hInstIds.data(), s->hInstIds.size(), st)) return false;
  if (!s->dFaceIds.upload(s->hFaceIds.data(), s->hFaceIds.size(), st)) return false;
  if (!s->dRankBase.upload(s->hRankBase.data(), s->hRankBase.size(), st)) return false;
  // shading records: decoded normals / tangents per vertex, indices + face id + geometric normal per face
  const uint32_t vertexCount = (uint32_t) (s->hFverts.size() / 8), faceCount = (uint32_t) (s->hFaces.size() / 3);
  if (!s->dSverts.reserve((size_t) std::max(vertexCount, 1u) * 4) || !s->dSfaces.reserve((size_t) std::max(faceCount, 1u) * 2)) return false;
  if (vertexCount) shading_vertices_kernel<<<std::min<uint32_t>((vertexCount + 255u) / 256u, (uint32_t) g.smCount * 16u), 256, 0, st>>>(s->dFverts.ptr, vertexCount, s->dSverts.ptr);
  for (const Scene::MeshRange& r : s->meshRanges)
    if (r.faceCount) shading_faces_kernel<<<std::min<uint32_t>((r.faceCount + 255u) / 256u, (uint32_t) g.smCount * 16u), 256, 0, st>>>(s->dFaces.ptr, s->dFaceIds.ptr, s->dFverts.ptr, r.faceBase, r.faceCount, r.vertexBase, s->dSfaces.ptr);
  g.stats.kernelLaunches += 1 + s->meshRanges.size();
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
  if (!hostBvh && !buildBvhOnDevice(s)) return false;
  if (!repackBvh(s)) return false;
  s->lastBuildMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

// Device BVH8 build (bvh8_gpu.cuh).  Inputs are the geometry arrays uploadGeometry() just placed in HBM.
template <typename T> struct ScratchArray {
  T* ptr = nullptr;
  ~ScratchArray() { if (ptr) cudaFree(ptr); }
  bool alloc(size_t n) { return cudaMalloc(&ptr, std::max<size_t>(n, 1) * sizeof(T)) == cudaSuccess; }
};

bool buildBvhOnDevice(Scene* s)
{
  const uint32_t n = s->totalTris;
  cudaStream_t st = g.stream;
  auto t0 = std::chrono::steady_clock::now();
  g.stats.msBvhDeviceBuild = 0.0f;
  if (n == 0) {                     // empty scene: one node without children (as the host builder)
    Bvh8Node empty; memset(&empty, 0, sizeof(empty));
    if (!s->dNodes.upload((const uint4*) &empty, 5, st)) return false;
    if (!s->dTris.reserve(3)) return false;
    CUDA_OK(cudaStreamSynchronize(st));
    s->nodeCount = 1; s->triCount = 0; s->maxDepth = 0; s->bvhHostStale = true;
    return true;
  }
  const uint32_t grid = (n + kLbvhBlock - 1) / kLbvhBlock;
  const uint32_t sortGrid = std::min<uint32_t>(grid, (uint32_t) g.smCount * 4u);
  ScratchArray<float4> inTris, boxLo, boxHi; ScratchArray<unsigned long long> keys0, keys1; ScratchArray<uint32_t> vals0, vals1, hist, parent, visit, bounds;
  ScratchArray<uint32_t> total, clusters0, clusters1, nearest, mergeFlag, mergeScan, keepScan, tileSums;
  ScratchArray<uint2> children, range, queue0, queue1; ScratchArray<LbvhDp> dp; ScratchArray<uint32_t> childList, childCount, cntNodes, cntTris, totals; ScratchArray<uint4> nodesTmp;
  const size_t refs = (size_t) 2 * n;
  if (!inTris.alloc((size_t) n * 3) || !boxLo.alloc(refs) || !boxHi.alloc(refs) || !keys0.alloc(n) || !keys1.alloc(n) || !vals0.alloc(n) || !vals1.alloc(n) ||
      !hist.alloc((size_t) 256 * sortGrid) || !parent.alloc(refs) || !visit.alloc(n) || !bounds.alloc(8) || !children.alloc(n) || !range.alloc(n) ||
      !queue0.alloc(n) || !queue1.alloc(n) || !dp.alloc(n) || !childList.alloc((size_t) n * 8) || !childCount.alloc(n) || !cntNodes.alloc(n) || !cntTris.alloc(n) ||
      !totals.alloc(2) || !nodesTmp.alloc((size_t) n * 5) || !total.alloc(n)) { setError("device BVH build: out of memory for %u triangles", n); return false; }
  if (!s->dTris.reserve((size_t) n * 3)) return false;

  lbvh_flatten_kernel<<<grid, kLbvhBlock, 0, st>>>(n, (uint32_t) s->hInstInfo.size(), s->dRankBase.ptr, s->dInstInfo.ptr, s->dO2W.ptr, s->dFverts.ptr, s->dFaces.ptr, s->dInstOpacity.ptr, inTris.ptr);
  const uint32_t initBounds[8] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u };
  CUDA_OK(cudaMemcpyAsync(bounds.ptr, initBounds, sizeof(initBounds), cudaMemcpyHostToDevice, st));
  lbvh_bounds_kernel<<<std::min<uint32_t>(grid, (uint32_t) g.smCount * 8u), kLbvhBlock, 0, st>>>(n, inTris.ptr, bounds.ptr);
  lbvh_morton_kernel<<<grid, kLbvhBlock, 0, st>>>(n, inTris.ptr, bounds.ptr, keys0.ptr, vals0.ptr);
  unsigned long long* k[2] = { keys0.ptr, keys1.ptr }; uint32_t* v[2] = { vals0.ptr, vals1.ptr };
  for (int pass = 0; pass < 8; pass++) {
    lbvh_radix_count_kernel<<<sortGrid, kLbvhBlock, 0, st>>>(n, k[0], pass * 8, hist.ptr);
    lbvh_scan_kernel<<<1, 1024, 0, st>>>(hist.ptr, 256u * sortGrid, nullptr);
    lbvh_radix_scatter_kernel<<<sortGrid, kLbvhBlock, 0, st>>>(n, k[0], v[0], pass * 8, hist.ptr, k[1], v[1]);
    std::swap(k[0], k[1]); std::swap(v[0], v[1]);
  }
  g.stats.kernelLaunches += 3 + 8 * 3;
  LbvhTree T; memset(&T, 0, sizeof(T));
  T.n = n; T.inTris = inTris.ptr; T.keys = k[0]; T.vals = v[0]; T.children = children.ptr; T.parent = parent.ptr; T.range = range.ptr;
  T.boxLo = boxLo.ptr; T.boxHi = boxHi.ptr; T.visit = visit.ptr; T.dp = dp.ptr; T.total = total.ptr;
  if (n > 1 && g.tuning.bvhBuilder == 2) {
    // PLOC: iterate nearest-neighbour search / mutual merges / order-preserving compaction until one cluster is left
    const size_t tiles = (size_t) n / kScanTile + 1;
    if (!clusters0.alloc(n) || !clusters1.alloc(n) || !nearest.alloc(n) || !mergeFlag.alloc(n) || !mergeScan.alloc(n) || !keepScan.alloc(n) || !tileSums.alloc(tiles)) { setError("device BVH build: out of memory for %u triangles", n); return false; }
    auto scanLarge = [&](uint32_t* a, uint32_t cnt, uint32_t* totalOut) {
      const uint32_t nt = (cnt + kScanTile - 1) / kScanTile;
      lbvh_scan_tiles_kernel<<<nt, 1024, 0, st>>>(a, cnt, tileSums.ptr);
      lbvh_scan_kernel<<<1, 1024, 0, st>>>(tileSums.ptr, nt, totalOut);
      if (nt > 1) lbvh_scan_add_kernel<<<(cnt + kLbvhBlock - 1) / kLbvhBlock, kLbvhBlock, 0, st>>>(a, cnt, tileSums.ptr);
      g.stats.kernelLaunches += nt > 1 ? 3 : 2;
    };
    lbvh_leaf_box_kernel<<<grid, kLbvhBlock, 0, st>>>(T, clusters0.ptr);
    uint32_t* cl[2] = { clusters0.ptr, clusters1.ptr };
    uint32_t count = n, nextId = n - 2, rounds = 0;
    while (count > 1) {
      if (++rounds > 4096) { setError("device BVH build: PLOC did not converge"); return false; }
      const uint32_t cg = (count + kLbvhBlock - 1) / kLbvhBlock;
      ploc_nearest_kernel<<<cg, kLbvhBlock, 0, st>>>(T, count, cl[0], nearest.ptr);
      ploc_flags_kernel<<<cg, kLbvhBlock, 0, st>>>(count, nearest.ptr, mergeFlag.ptr, mergeScan.ptr, keepScan.ptr);
      scanLarge(mergeScan.ptr, count, totals.ptr);
      scanLarge(keepScan.ptr, count, nullptr);
      ploc_apply_kernel<<<cg, kLbvhBlock, 0, st>>>(T, count, cl[0], nearest.ptr, mergeFlag.ptr, mergeScan.ptr, keepScan.ptr, nextId, cl[1]);
      g.stats.kernelLaunches += 3;
      uint32_t merges = 0;
      CUDA_OK(cudaMemcpyAsync(&merges, totals.ptr, sizeof(merges), cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
      if (merges == 0 || merges > count / 2) { setError("device BVH build: PLOC round merged %u of %u clusters", merges, count); return false; }
      nextId -= merges; count -= merges;
      std::swap(cl[0], cl[1]);
    }
  } else if (n > 1) lbvh_karras_kernel<<<grid, kLbvhBlock, 0, st>>>(T);
  lbvh_refit_kernel<<<grid, kLbvhBlock, 0, st>>>(T);
  g.stats.kernelLaunches += 2;

  const uint2 rootEntry = make_uint2(0u, 0u);
  CUDA_OK(cudaMemcpyAsync(queue0.ptr, &rootEntry, sizeof(rootEntry), cudaMemcpyHostToDevice, st));
  LbvhLevel Lv; memset(&Lv, 0, sizeof(Lv));
  Lv.childList = childList.ptr; Lv.childCount = childCount.ptr; Lv.cntNodes = cntNodes.ptr; Lv.cntTris = cntTris.ptr;
  uint2* q[2] = { queue0.ptr, queue1.ptr };
  uint32_t count = 1, nodeBase = 1, triBase = 0, depth = 0;
  while (count > 0) {
    depth++;
    if (depth > (uint32_t) kStackSize) { setError("device BVH build: wide tree deeper than the traversal stack (%d)", kStackSize); return false; }
    Lv.count = count; Lv.queue = q[0]; Lv.nextQueue = q[1]; Lv.nodeBase = nodeBase; Lv.triBase = triBase;
    const uint32_t lg = (count + kLbvhBlock - 1) / kLbvhBlock;
    lbvh_collapse_count_kernel<<<lg, kLbvhBlock, 0, st>>>(T, Lv);
    lbvh_scan_kernel<<<1, 1024, 0, st>>>(cntNodes.ptr, count, totals.ptr + 0);
    lbvh_scan_kernel<<<1, 1024, 0, st>>>(cntTris.ptr, count, totals.ptr + 1);
    lbvh_collapse_emit_kernel<<<lg, kLbvhBlock, 0, st>>>(T, Lv, nodesTmp.ptr, s->dTris.ptr);
    g.stats.kernelLaunches += 4;
    uint32_t tot[2];
    CUDA_OK(cudaMemcpyAsync(tot, totals.ptr, sizeof(tot), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    nodeBase += tot[0]; triBase += tot[1]; count = tot[0];
    std::swap(q[0], q[1]);
  }
  if (triBase != n) { setError("device BVH build: emitted %u of %u triangles", triBase, n); return false; }
  if (!s->dNodes.reserve((size_t) nodeBase * 5)) return false;
  CUDA_OK(cudaMemcpyAsync(s->dNodes.ptr, nodesTmp.ptr, (size_t) nodeBase * 80, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));
  s->nodeCount = nodeBase; s->triCount = n; s->maxDepth = depth; s->bvhHostStale = true;
  g.stats.msBvhDeviceBuild = (float) std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

bool uploadMaterials(Scene* s)
{
  std::vector<float> params; std::vector<int32_t> closure; std::vector<uint32_t> flags;
  std::vector<DeviceTexture> textures; std::vector<Texture*> textureIndex; std::vector<DeviceTexBinding> bindings; std::vector<uint2> ranges;
  std::vector<DevicePrimvarBinding> pbindings; std::vector<uint2> pranges;
  std::vector<uint32_t> opacityBinding; std::vector<float> opacityConst; std::vector<uint2> opacityPrimvar; std::vector<float> pfallback;
  const double kPi = 3.14159265358979323846;
  for (Material* m : s->flatMaterials) {
    params.insert(params.end(), m->params, m->params + GTLC_MATERIAL_PARAM_COUNT);
    closure.push_back(m->closure);
    flags.push_back((mat_is_emissive(m->closure, m->params) || !m->bindings.empty() || !m->primvarBindings.empty() ? kMatFlagEmissive : 0u) |   // a textured input may be the emission
                    (mat_is_thin_walled(m->closure, m->params) ? kMatFlagThinWalled : 0u) |
                    (mat_cutout_opacity(m->closure, m->params) < 1.0f || m->opacityBinding >= 0 || m->opacityPrimvar >= 0 ? kMatFlagCutout : 0u));
    opacityBinding.push_back(m->opacityBinding >= 0 ? (uint32_t) (bindings.size() + (size_t) m->opacityBinding) : 0xFFFFFFFFu);
    opacityConst.push_back(mat_cutout_opacity(m->closure, m->params));
    opacityPrimvar.push_back(m->opacityPrimvar >= 0 ? make_uint2((uint32_t) (pbindings.size() + (size_t) m->opacityPrimvar), (uint32_t) m->opacityPrimvar) : make_uint2(0xFFFFFFFFu, 0u));
    ranges.push_back(make_uint2((uint32_t) bindings.size(), (uint32_t) m->bindings.size()));
    pranges.push_back(make_uint2((uint32_t) pbindings.size(), (uint32_t) m->primvarBindings.size()));
    for (const Material::PrimvarBinding& pb : m->primvarBindings) { DevicePrimvarBinding d; memset(&d, 0, sizeof(d)); d.paramIndex = pb.paramIndex; d.componentCount = pb.componentCount; memcpy(d.fallback, pb.fallback, 16); pbindings.push_back(d); pfallback.push_back(pb.fallback[0]); }
    for (size_t b = 0; b < m->bindings.size(); b++) {
      const GtlcTextureBinding& src = m->bindings[b];
      Texture* t = m->boundTextures[b];
      uint32_t ti = 0;
      for (; t && ti < textureIndex.size(); ti++) if (textureIndex[ti] == t) break;     // (t == NULL: a constant operand, no lookup)
      if (t && ti == textureIndex.size()) { textureIndex.push_back(t); textures.push_back({ t->dTexels, t->width, t->height, (uint32_t) t->format, t->depth }); }
      DeviceTexBinding d; memset(&d, 0, sizeof(d));
      d.paramIndex = src.paramIndex; d.texture = ti; d.combine = (uint32_t) src.combine;
      d.packed = (uint32_t) src.componentCount | ((uint32_t) src.channels[0] << 4) | ((uint32_t) src.channels[1] << 6) | ((uint32_t) src.channels[2] << 8) |
                 ((uint32_t) src.wrapU << 10) | ((uint32_t) src.wrapV << 12) | ((src.srgb ? 1u : 0u) << 14) | ((uint32_t) src.lookup << 15) | (((uint32_t) src.wrapW & 3u) << 17);
      for (int k = 0; k < 3; k++) { d.xyzScale[k] = src.xyzScale[k]; d.xyzOffset[k] = src.xyzOffset[k]; }
      memcpy(d.fallback, src.fallback, 16); memcpy(d.scale, src.scale, 16); memcpy(d.bias, src.bias, 16);
      d.uvCos = (float) cos((double) src.uvRotation * kPi / 180.0); d.uvSin = (float) sin((double) src.uvRotation * kPi / 180.0);
      d.uvScale[0] = src.uvScale[0]; d.uvScale[1] = src.uvScale[1]; d.uvTranslation[0] = src.uvTranslation[0]; d.uvTranslation[1] = src.uvTranslation[1];
      bindings.push_back(d);
    }
  }
  s->hasTextures = !bindings.empty() || !pbindings.empty();
  s->hasTexturedCutout = false;
  for (uint32_t ob : opacityBinding) s->hasTexturedCutout = s->hasTexturedCutout || ob != 0xFFFFFFFFu;
  for (const uint2& op : opacityPrimvar) s->hasTexturedCutout = s->hasTexturedCutout || op.x != 0xFFFFFFFFu;
  if (!s->dPrimvarBindings.upload(pbindings.data(), pbindings.size(), g.stream)) return false;
  if (!s->dMatPrimvarRange.upload(pranges.data(), pranges.size(), g.stream)) return false;
  if (!s->dMatParams.upload(params.data(), params.size(), g.stream)) return false;
  if (!s->dMatClosure.upload(closure.data(), closure.size(), g.stream)) return false;
  if (!s->dMatFlags.upload(flags.data(), flags.size(), g.stream)) return false;
  if (!s->dTextures.upload(textures.data(), textures.size(), g.stream)) return false;
  if (!s->dTexBindings.upload(bindings.data(), bindings.size(), g.stream)) return false;
  if (!s->dMatTexRange.upload(ranges.data(), ranges.size(), g.stream)) return false;
  if (!s->dMatOpacityBinding.upload(opacityBinding.data(), opacityBinding.size(), g.stream)) return false;
  if (!s->dMatOpacityConst.upload(opacityConst.data(), opacityConst.size(), g.stream)) return false;
  if (!s->dMatOpacityPrimvar.upload(opacityPrimvar.data(), opacityPrimvar.size(), g.stream)) return false;
  if (!s->dPrimvarFallback.upload(pfallback.data(), pfallback.size(), g.stream)) return false;
  CutoutScene cut;                                    // what a textured cutout's any-hit test reads (geometry arrays are uploaded before the materials)
  cut.sverts = (const uint32_t*) s->dSverts.ptr; cut.sfaces = (const uint32_t*) s->dSfaces.ptr; cut.instInfo = s->dInstInfo.ptr;
  cut.texBindings = s->dTexBindings.ptr; cut.textures = s->dTextures.ptr;
  cut.matOpacityBinding = s->dMatOpacityBinding.ptr; cut.matOpacityConst = s->dMatOpacityConst.ptr;
  cut.matOpacityPrimvar = s->dMatOpacityPrimvar.ptr; cut.primvarFallback = s->dPrimvarFallback.ptr;
  cut.primvarInfos = s->dPrimvarInfos.ptr; cut.primvarData = s->dPrimvarData.ptr; cut.instIds = s->dInstIds.ptr;
  if (!s->dCutoutScene.upload(&cut, 1, g.stream)) return false;
  const uint32_t matCount = (uint32_t) closure.size();
  if (!s->dMatLowered.reserve((size_t) std::max(matCount, 1u) * kLobeFloats)) return false;
  if (matCount) lower_materials_kernel<<<(matCount + 63u) / 64u, 64, 0, g.stream>>>(s->dMatParams.ptr, s->dMatClosure.ptr, matCount, s->dMatLowered.ptr);
  CUDA_OK(cudaStreamSynchronize(g.stream));
  CUDA_OK(cudaGetLastError());
  return true;
}

template <typename L, typename D> bool uploadLightArray(const std::vector<L*>& lights, DeviceArray<float4>& dst)
{
  std::vector<D> host; host.reserve(lights.size());
  for (L* l : lights) host.push_back(l->data);
  return dst.upload((const float4*) host.data(), host.size() * 3, g.stream);
}
bool uploadLights(Scene* s)
{
  if (!uploadLightArray<SphereLight, SphereLightData>(s->sphereLights, s->dSphere)) return false;
  if (!uploadLightArray<DistantLight, DistantLightData>(s->distantLights, s->dDistant)) return false;
  if (!uploadLightArray<RectLight, RectLightData>(s->rectLights, s->dRect)) return false;
  if (!uploadLightArray<DiskLight, DiskLightData>(s->diskLights, s->dDisk)) return false;
  CUDA_OK(cudaStreamSynchronize(g.stream));
  return true;
}

bool ensurePool(size_t capacity)
{
  if (capacity <= g.poolCapacity) return true;
  auto freeQ = [](PathQueue& q) { cudaFree(q.rayO); cudaFree(q.rayD); cudaFree(q.thr); cudaFree(q.rad); cudaFree(q.slot); cudaFree(q.med); cudaFree(q.walk); memset(&q, 0, sizeof(q)); };
  g.volumeCapacity = 0;
  freeQ(g.q[0]); freeQ(g.q[1]);
  cudaFree(g.hitTUVP); cudaFree(g.hitInst); cudaFree(g.shO); cudaFree(g.shD); cudaFree(g.shC); cudaFree(g.shSeed); cudaFree(g.sampleRad);
  g.hitTUVP = nullptr; g.hitInst = nullptr; g.shO = g.shD = g.shC = nullptr; g.shSeed = nullptr; g.sampleRad = nullptr; g.poolCapacity = 0;
  for (int k = 0; k < 2; k++) {
    CUDA_OK(cudaMalloc(&g.q[k].rayO, capacity * 16)); CUDA_OK(cudaMalloc(&g.q[k].rayD, capacity * 16));
    CUDA_OK(cudaMalloc(&g.q[k].thr, capacity * 16)); CUDA_OK(cudaMalloc(&g.q[k].rad, capacity * 16));
    CUDA_OK(cudaMalloc(&g.q[k].slot, capacity * 4));
  }
  CUDA_OK(cudaMalloc(&g.hitTUVP, capacity * 16)); CUDA_OK(cudaMalloc(&g.hitInst, capacity * 4));
  CUDA_OK(cudaMalloc(&g.shO, capacity * 16)); CUDA_OK(cudaMalloc(&g.shD, capacity * 16)); CUDA_OK(cudaMalloc(&g.shC, capacity * 16));
  CUDA_OK(cudaMalloc(&g.shSeed, capacity * 4));
  CUDA_OK(cudaMalloc(&g.sampleRad, capacity * 16));
  cudaFree(g.perm); g.perm = nullptr;
  CUDA_OK(cudaMalloc(&g.perm, capacity * 4));
  if (!g.sortHist) { g.sortBlocks = g.smCount * 4; CUDA_OK(cudaMalloc(&g.sortHist, ((size_t) kSortBins * g.sortBlocks + kSortBins) * 4)); }   // per-CTA counts + the 256 bin totals
  g.poolCapacity = capacity;
  return true;
}

// medium stack + walkSegmentPdf of every path (only with mediumStackSize > 0)
bool ensureVolumePool()
{
  if (g.volumeCapacity >= g.poolCapacity && g.q[0].med) return true;
  for (int k = 0; k < 2; k++) { cudaFree(g.q[k].med); cudaFree(g.q[k].walk); g.q[k].med = nullptr; g.q[k].walk = nullptr; }
  g.volumeCapacity = 0;
  for (int k = 0; k < 2; k++) { CUDA_OK(cudaMalloc(&g.q[k].med, g.poolCapacity * 16)); CUDA_OK(cudaMalloc(&g.q[k].walk, g.poolCapacity * 16)); }
  g.volumeCapacity = g.poolCapacity;
  return true;
}

bool ensureAovSamples(size_t capacity)
{
  if (capacity <= g.aovCapacity) return true;
  for (int k = 0; k < 6; k++) { cudaFree(g.aovSamples.a[k]); g.aovSamples.a[k] = nullptr; }
  cudaFree(g.aovSamples.clk); g.aovSamples.clk = nullptr;
  g.aovCapacity = 0;
  for (int k = 0; k < 6; k++) CUDA_OK(cudaMalloc(&g.aovSamples.a[k], capacity * 16));
  CUDA_OK(cudaMalloc(&g.aovSamples.clk, capacity * 4));
  g.aovCapacity = capacity;
  return true;
}

cudaEvent_t nextEvent()
{
  if (g.eventsUsed == g.events.size()) { cudaEvent_t e; cudaEventCreate(&e); g.events.push_back(e); }
  return g.events[g.eventsUsed++];
}

struct TimedSpan { cudaEvent_t a, b; int kind; };
enum { K_RAYGEN, K_TRACE, K_SHADE, K_SHADOW, K_RESOLVE, K_SORT, K_COUNT };

bool loadNccl()
{
  if (g.nccl.lib) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // an already loaded copy with this SONAME (e.g. the host application's) is reused
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { setError("multi-GPU: cannot open libnccl.so.2 (%s)", dlerror()); return false; }
  g.nccl.GetUniqueId = (decltype(g.nccl.GetUniqueId)) dlsym(lib, "ncclGetUniqueId");
  g.nccl.CommInitRank = (decltype(g.nccl.CommInitRank)) dlsym(lib, "ncclCommInitRank");
  g.nccl.Reduce = (decltype(g.nccl.Reduce)) dlsym(lib, "ncclReduce");
  g.nccl.CommDestroy = (decltype(g.nccl.CommDestroy)) dlsym(lib, "ncclCommDestroy");
  g.nccl.GetErrorString = (decltype(g.nccl.GetErrorString)) dlsym(lib, "ncclGetErrorString");
  if (!g.nccl.GetUniqueId || !g.nccl.CommInitRank || !g.nccl.Reduce || !g.nccl.CommDestroy || !g.nccl.GetErrorString) { setError("multi-GPU: libnccl.so.2 lacks an entry point"); dlclose(lib); return false; }
  g.nccl.lib = lib;
  return true;
}
#define NCCL_OK(expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) { setError("%s failed: %s", #expr, g.nccl.GetErrorString(_r)); return false; } } while (0)

bool requireInit()
{
  if (!g.initialized) { setError("gtlc_initialize() has not been called (or failed: no CUDA device) -- there is no CPU fallback"); return false; }
  return true;
}

bool buildPixelMap(Scene* s, uint32_t W, uint32_t H)
{
  if (s->shardCount <= 1 || s->shardTile == GTLC_SHARD_SAMPLES) { s->localPixels = W * H; s->pixelMapCount = 1; s->pixelMapW = W; s->pixelMapH = H; return true; }
  if (s->pixelMapW == W && s->pixelMapH == H && s->pixelMapRank == s->shardRank && s->pixelMapCount == s->shardCount && s->pixelMapTile == s->shardTile) return true;
  std::vector<uint32_t> map;
  const uint32_t T = s->shardTile, tilesX = (W + T - 1) / T;
  for (uint32_t y = 0; y < H; y++) for (uint32_t x = 0; x < W; x++) {
    uint32_t tile = (y / T) * tilesX + (x / T);
    if (tile % s->shardCount == s->shardRank) map.push_back(x + y * W);
  }
  if (!s->dPixelMap.upload(map.data(), map.size(), g.stream)) return false;
  CUDA_OK(cudaStreamSynchronize(g.stream));
  s->localPixels = (uint32_t) map.size();
  s->pixelMapW = W; s->pixelMapH = H; s->pixelMapRank = s->shardRank; s->pixelMapCount = s->shardCount; s->pixelMapTile = s->shardTile;
  return true;
}

TraceParams makeTraceParams(Scene* s)
{
  TraceParams P; memset(&P, 0, sizeof(P));
  P.nodes = (const uint32_t*) s->dNodesT.ptr; P.tris = (const uint32_t*) s->dTrisT.ptr; P.biasWord = 0x47000000u;
  P.stats = g.counters->stats;
  P.refillThreshold = g.tuning.reserved[0] > 0 ? g.tuning.reserved[0] : kRefillThreshold;
  P.triThreshold = g.tuning.reserved[2] > 0 ? g.tuning.reserved[2] : 8;
  P.cut = s->hasTexturedCutout ? s->dCutoutScene.ptr : nullptr;
  return P;
}

template <int MODE, bool TEXCUT> void launchTraceT(const TraceParams& P, bool count, int blocks)
{
  if (count) { trace_kernel<MODE, true, TEXCUT><<<blocks, kTraceBlock, 0, g.stream>>>(P); return; }   // visit counting: variant 0 order = the oracle's walk
  if (P.aovFlags || P.clk) { trace_kernel<MODE, false, TEXCUT><<<blocks, kTraceBlock, 0, g.stream>>>(P); return; }   // only the per-lane kernel writes the NEE debug AOV / the ClockCycles AOV
  switch (g.tuning.traceVariant) {
    case 1: trace_kernel<MODE, false, TEXCUT><<<blocks, kTraceBlock, 0, g.stream>>>(P); break;          // per-lane triangle loop (the order the oracle's walker restates)
    case 6: if (TEXCUT) trace_kernel<MODE, false, TEXCUT><<<blocks, kTraceBlock, 0, g.stream>>>(P);      // (the TMA variant has no textured-cutout form)
            else trace_kernel_tma<MODE><<<blocks, kTraceBlock, 0, g.stream>>>(P);
            break;
    default:                                                                                    // 0 (default) and 7: warp-batched triangle tests
      // its queue entries pack (owner lane << 27 | triangle index): beyond 2^27 triangles the per-lane kernel takes over
      if (g.currentTriCount >= (1u << 27)) trace_kernel<MODE, false, TEXCUT><<<blocks, kTraceBlock, 0, g.stream>>>(P);
      else trace_kernel_batched<MODE, TEXCUT><<<blocks, kTraceBlock, 0, g.stream>>>(P);
      break;
  }
}
// Scenes with a textured cutout (TraceParams.cut set) run the instantiations whose any-hit test can evaluate a material input.
template <int MODE> void launchTrace(const TraceParams& P, bool count, int blocks)
{
  if (P.cut) launchTraceT<MODE, true>(P, count, blocks); else launchTraceT<MODE, false>(P, count, blocks);
}

// _CalcDirtyFlagsForRenderParams (Gi.cpp:1859-1987), reduced to "does the accumulated image survive?"
bool paramsResetAccumulation(Scene* s, const GtlcRenderParams* p)
{
  if (!s->haveOldParams) return true;
  const GtlcRenderParams& o = s->oldParams;
  if (memcmp(&o.camera, &p->camera, sizeof(GtlcCameraDesc)) != 0) return true;
  GtlcRenderSettings a = o.renderSettings, b = p->renderSettings;
  a.spp = b.spp = 0;   // the sample count per call does not invalidate the image
  if (memcmp(&a, &b, sizeof(a)) != 0) return true;
  if (o.domeLight != p->domeLight) return true;
  if (s->oldBindings.size() != p->aovBindingCount) return true;
  for (uint32_t i = 0; i < p->aovBindingCount; i++) {
    const GtlcAovBinding &x = s->oldBindings[i], &y = p->aovBindings[i];
    if (x.aovId != y.aovId || x.renderBuffer != y.renderBuffer || memcmp(x.clearValue, y.clearValue, 16) != 0) return true;
  }
  return false;
}

} // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

uint64_t gtlc_error_count(void) { std::lock_guard<std::mutex> l(g_errMutex); return g_errorCount; }
const char* gtlc_last_error(void) { std::lock_guard<std::mutex> l(g_errMutex); static thread_local std::string copy; copy = g_lastError; return copy.c_str(); }

void gtlc_material_desc_defaults(int32_t closure, GtlcMaterialDesc* d)
{
  memset(d, 0, sizeof(*d)); d->closure = closure;
  float* p = d->params;
  if (closure == GTLC_CLOSURE_DIFFUSE) { p[0] = p[1] = p[2] = 0.8f; }
  else if (closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) {   // UsdPreviewSurface specification defaults
    p[GTLC_UPS_DIFFUSE_R] = p[GTLC_UPS_DIFFUSE_G] = p[GTLC_UPS_DIFFUSE_B] = 0.18f;
    p[GTLC_UPS_ROUGHNESS] = 0.5f; p[GTLC_UPS_CLEARCOAT_ROUGHNESS] = 0.01f; p[GTLC_UPS_OPACITY] = 1.0f; p[GTLC_UPS_IOR] = 1.5f;
  } else if (closure == GTLC_CLOSURE_OPENPBR) {             // open_pbr_surface.mtlx:11-84
    p[GTLC_OPBR_BASE_WEIGHT] = 1.0f; p[GTLC_OPBR_BASE_COLOR_R] = p[GTLC_OPBR_BASE_COLOR_G] = p[GTLC_OPBR_BASE_COLOR_B] = 0.8f;
    p[GTLC_OPBR_SPECULAR_WEIGHT] = 1.0f; p[GTLC_OPBR_SPECULAR_COLOR_R] = p[GTLC_OPBR_SPECULAR_COLOR_G] = p[GTLC_OPBR_SPECULAR_COLOR_B] = 1.0f;
    p[GTLC_OPBR_SPECULAR_ROUGHNESS] = 0.3f; p[GTLC_OPBR_SPECULAR_IOR] = 1.5f;
    p[GTLC_OPBR_TRANSMISSION_COLOR_R] = p[GTLC_OPBR_TRANSMISSION_COLOR_G] = p[GTLC_OPBR_TRANSMISSION_COLOR_B] = 1.0f;
    p[GTLC_OPBR_COAT_COLOR_R] = p[GTLC_OPBR_COAT_COLOR_G] = p[GTLC_OPBR_COAT_COLOR_B] = 1.0f; p[GTLC_OPBR_COAT_IOR] = 1.6f; p[GTLC_OPBR_COAT_DARKENING] = 1.0f;
    p[GTLC_OPBR_FUZZ_COLOR_R] = p[GTLC_OPBR_FUZZ_COLOR_G] = p[GTLC_OPBR_FUZZ_COLOR_B] = 1.0f; p[GTLC_OPBR_FUZZ_ROUGHNESS] = 0.5f;
    p[GTLC_OPBR_EMISSION_COLOR_R] = p[GTLC_OPBR_EMISSION_COLOR_G] = p[GTLC_OPBR_EMISSION_COLOR_B] = 1.0f;
    p[GTLC_OPBR_THIN_FILM_THICKNESS] = 0.5f; p[GTLC_OPBR_THIN_FILM_IOR] = 1.4f; p[GTLC_OPBR_GEOMETRY_OPACITY] = 1.0f;
    p[GTLC_OPBR_SUBSURFACE_COLOR_R] = p[GTLC_OPBR_SUBSURFACE_COLOR_G] = p[GTLC_OPBR_SUBSURFACE_COLOR_B] = 0.8f;
  }
}

int gtlc_initialize(const GtlcInitParams* params)
{
  if (g.initialized) return GTLC_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) { setError("no CUDA device available (%s); the B200 path has no CPU fallback", cudaGetErrorString(e)); return GTLC_ERROR; }
  int dev = params ? params->deviceIndex : -1;
  if (dev < 0) { const char* lr = getenv("LOCAL_RANK"); dev = lr ? atoi(lr) % count : 0; }
  if (dev >= count) { setError("device index %d out of range (%d devices)", dev, count); return GTLC_ERROR; }
  if (cudaSetDevice(dev) != cudaSuccess) { setError("cudaSetDevice(%d) failed", dev); return GTLC_ERROR; }
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  g.device = dev; g.smCount = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking) != cudaSuccess) { setError("cudaStreamCreate failed"); return GTLC_ERROR; }
  if (cudaMalloc(&g.counters, sizeof(Counters)) != cudaSuccess) { setError("cudaMalloc(counters) failed"); return GTLC_ERROR; }
  cudaMemset(g.counters, 0, sizeof(Counters));
  g.initialized = true;
  return GTLC_OK;
}

void gtlc_terminate(void)
{
  if (!g.initialized) return;
  cudaSetDevice(g.device);
  cudaDeviceSynchronize();
  for (int k = 0; k < 2; k++) { cudaFree(g.q[k].rayO); cudaFree(g.q[k].rayD); cudaFree(g.q[k].thr); cudaFree(g.q[k].rad); cudaFree(g.q[k].slot); cudaFree(g.q[k].med); cudaFree(g.q[k].walk); }
  memset(g.q, 0, sizeof(g.q));
  g.volumeCapacity = 0;
  cudaFree(g.hitTUVP); cudaFree(g.hitInst); cudaFree(g.shO); cudaFree(g.shD); cudaFree(g.shC); cudaFree(g.shSeed); cudaFree(g.sampleRad); cudaFree(g.accum); cudaFree(g.counters); cudaFree(g.visits); cudaFree(g.perm); cudaFree(g.sortHist); g.perm = nullptr; g.sortHist = nullptr;
  for (int k = 0; k < 6; k++) { cudaFree(g.aovSamples.a[k]); g.aovSamples.a[k] = nullptr; }
  cudaFree(g.aovSamples.clk); g.aovSamples.clk = nullptr;
  g.aovCapacity = 0;
  g.hitTUVP = nullptr; g.hitInst = nullptr; g.shO = g.shD = g.shC = nullptr; g.shSeed = nullptr; g.sampleRad = nullptr; g.accum = nullptr; g.counters = nullptr; g.visits = nullptr;
  g.poolCapacity = 0; g.accumCapacity = 0; g.visitsCapacity = 0;
  for (cudaEvent_t e : g.events) cudaEventDestroy(e);
  g.events.clear(); g.eventsUsed = 0;
  if (g.nccl.comm) { g.nccl.CommDestroy(g.nccl.comm); g.nccl.comm = nullptr; g.nccl.world = 0; }
  cudaFree(g.reduced); g.reduced = nullptr; g.reducedCapacity = 0;
  cudaStreamDestroy(g.stream); g.stream = nullptr;
  g.initialized = false;
}

// ---- multi-GPU communicator ----------------------------------------------------------------
int gtlc_comm_get_unique_id(void* id)
{
  if (!id) { setError("gtlc_comm_get_unique_id: null id"); return GTLC_ERROR; }
  if (!loadNccl()) return GTLC_ERROR;
  ncclUniqueId u;
  ncclResult_t r = g.nccl.GetUniqueId(&u);
  if (r != ncclSuccess) { setError("ncclGetUniqueId failed: %s", g.nccl.GetErrorString(r)); return GTLC_ERROR; }
  static_assert(sizeof(u) == GTLC_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(id, &u, sizeof(u));
  return GTLC_OK;
}
int gtlc_comm_init(const void* id, int32_t rank, int32_t worldSize)
{
  if (!requireInit()) return GTLC_ERROR;
  if (!id || worldSize < 1 || rank < 0 || rank >= worldSize) { setError("gtlc_comm_init: invalid arguments (rank %d of %d)", rank, worldSize); return GTLC_ERROR; }
  std::lock_guard<std::mutex> rl(g_renderMutex);
  if (g.nccl.comm) { setError("gtlc_comm_init: a communicator exists already"); return GTLC_ERROR; }
  if (!loadNccl()) return GTLC_ERROR;
  cudaSetDevice(g.device);
  ncclUniqueId u; memcpy(&u, id, sizeof(u));
  ncclResult_t r = g.nccl.CommInitRank(&g.nccl.comm, worldSize, u, rank);
  if (r != ncclSuccess) { setError("ncclCommInitRank failed: %s", g.nccl.GetErrorString(r)); g.nccl.comm = nullptr; return GTLC_ERROR; }
  g.nccl.rank = rank; g.nccl.world = worldSize;
  return GTLC_OK;
}
void gtlc_comm_destroy(void)
{
  std::lock_guard<std::mutex> rl(g_renderMutex);
  if (!g.nccl.comm) return;
  cudaSetDevice(g.device); cudaStreamSynchronize(g.stream);
  g.nccl.CommDestroy(g.nccl.comm); g.nccl.comm = nullptr; g.nccl.world = 0;
}
int gtlc_comm_world_size(void) { return g.nccl.comm ? g.nccl.world : 0; }

void gtlc_set_tuning(const GtlcTuning* t) { if (t) { g.tuning = *t; if (g.tuning.maxPathsPerWave == 0) g.tuning.maxPathsPerWave = 1u << 24; } }
void gtlc_get_tuning(GtlcTuning* t) { if (t) *t = g.tuning; }
void gtlc_get_stats(GtlcStats* s) { if (s) *s = g.stats; }

// ---- scene ---------------------------------------------------------------------------------
GtlcHandle gtlc_scene_create(void)
{
  static std::once_flag warm;
  std::call_once(warm, [] { bvh8_warm_up_threads(); });    // see bvh8.h: the builder's thread team exists before its first build
  Scene* s = new Scene(); s->tag = TAG_SCENE; return handleOf(s);
}

void gtlc_scene_destroy(GtlcHandle h)
{
  Scene* s = as<Scene>(h, TAG_SCENE); if (!s) return;
  if (g.initialized) { cudaSetDevice(g.device); cudaStreamSynchronize(g.stream); }
  s->dNodes.release(); s->dTris.release(); s->dNodesT.release(); s->dTrisT.release(); s->dSverts.release(); s->dSfaces.release(); s->dMatLowered.release(); s->dFverts.release(); s->dFaces.release(); s->dO2W.release(); s->dW2O.release(); s->dInstInfo.release(); s->dInstIds.release(); s->dFaceIds.release(); s->dRankBase.release(); s->dInstOpacity.release(); s->dMatOpacityBinding.release(); s->dMatOpacityConst.release(); s->dMatOpacityPrimvar.release(); s->dPrimvarFallback.release(); s->dCutoutScene.release();
  s->dMatParams.release(); s->dTextures.release(); s->dTexBindings.release(); s->dMatTexRange.release(); s->dPrimvarInfos.release(); s->dPrimvarData.release(); s->dPrimvarBindings.release(); s->dMatPrimvarRange.release(); s->dMatClosure.release(); s->dMatFlags.release(); s->dSphere.release(); s->dDistant.release(); s->dRect.release(); s->dDisk.release(); s->dPixelMap.release();
  s->tag = 0;
  delete s;
}

void gtlc_scene_set_shard(GtlcHandle h, uint32_t rank, uint32_t worldSize, uint32_t tileSize)
{
  Scene* s = as<Scene>(h, TAG_SCENE); if (!s) return;
  std::lock_guard<std::mutex> l(s->mutex);
  s->shardRank = rank; s->shardCount = worldSize ? worldSize : 1; s->shardTile = tileSize ? tileSize : 64;
  s->framebufferDirty = true;
}

// ---- materials -----------------------------------------------------------------------------
// ---- textures (R15) ------------------------------------------------------------------------
static size_t texelBytes(int32_t format) { return format == GTLC_TEX_RGBA8_UNORM ? 4 : format == GTLC_TEX_RGBA16_FLOAT ? 8 : format == GTLC_TEX_R32_FLOAT ? 4 : 16; }
static void releaseTexture(Texture* t)
{
  if (--t->refs > 0) return;
  cudaSetDevice(g.device); cudaStreamSynchronize(g.stream); cudaFree(t->dTexels);
  t->tag = 0; delete t;
}
GtlcHandle gtlc_texture_create(const GtlcTextureDesc* d)
{
  if (!requireInit()) return 0;
  if (!d || !d->texels || d->width == 0 || d->height == 0) { setError("gtlc_texture_create: empty texture"); return 0; }
  if (d->format < GTLC_TEX_RGBA8_UNORM || d->format > GTLC_TEX_RGBA32_FLOAT) { setError("gtlc_texture_create: unknown format %d", d->format); return 0; }
  cudaSetDevice(g.device);
  Texture* t = new Texture(); t->tag = TAG_TEXTURE; t->width = d->width; t->height = d->height; t->depth = d->depth > 1u ? d->depth : 1u; t->format = d->format; t->refs = 1; t->dTexels = nullptr;
  const size_t bytes = (size_t) d->width * d->height * t->depth * texelBytes(d->format);
  if (cudaMalloc(&t->dTexels, bytes) != cudaSuccess || cudaMemcpy(t->dTexels, d->texels, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    setError("gtlc_texture_create: device allocation of %zu bytes failed", bytes); cudaFree(t->dTexels); delete t; return 0;
  }
  return handleOf(t);
}
void gtlc_texture_destroy(GtlcHandle h)
{
  Texture* t = as<Texture>(h, TAG_TEXTURE); if (!t) return;
  std::lock_guard<std::mutex> dl(g_destroyMutex);
  releaseTexture(t);
}

GtlcHandle gtlc_material_create(GtlcHandle sh, const GtlcMaterialDesc* d)
{
  Scene* s = as<Scene>(sh, TAG_SCENE); if (!s || !d) return 0;
  if (d->closure < GTLC_CLOSURE_DIFFUSE || d->closure > GTLC_CLOSURE_OPENPBR) { setError("unknown closure %d", d->closure); return 0; }
  if (d->textureCount && !d->textures) { setError("gtlc_material_create: null texture bindings"); return 0; }
  if (d->primvarCount > GTLC_MAX_PRIMVAR_BINDINGS || (d->primvarCount && !d->primvars)) { setError("gtlc_material_create: %u primvar bindings (at most %d)", d->primvarCount, GTLC_MAX_PRIMVAR_BINDINGS); return 0; }
  for (uint32_t b = 0; b < d->primvarCount; b++) {
    const GtlcPrimvarBinding& pb = d->primvars[b];
    if (pb.componentCount < 1 || pb.componentCount > 3 || pb.paramIndex < 0 || pb.paramIndex + pb.componentCount > GTLC_MATERIAL_PARAM_COUNT) { setError("gtlc_material_create: primvar binding %u writes outside the parameter block", b); return 0; }
  }
  std::vector<Texture*> bound;
  for (uint32_t b = 0; b < d->textureCount; b++) {
    const GtlcTextureBinding& tb = d->textures[b];
    const int op = tb.combine & 15, factorParam = (tb.combine >> 8) & 255;
    const bool constant = (tb.combine & GTLC_COMBINE_CONSTANT) != 0;
    if (op > GTLC_COMBINE_MIX || (tb.combine & ~(0xFF1F)) != 0 || (op == GTLC_COMBINE_MIX && factorParam >= GTLC_MATERIAL_PARAM_COUNT + GTLC_BIND_SCRATCH_COUNT)) { setError("gtlc_material_create: binding %u has an invalid combine word 0x%x", b, (unsigned) tb.combine); return 0; }
    if (tb.paramIndex == GTLC_BIND_NORMAL && tb.combine != 0) { setError("gtlc_material_create: binding %u: the normal takes one plain lookup (combine must be 0)", b); return 0; }
    Texture* t = (constant && !tb.texture) ? nullptr : as<Texture>(tb.texture, TAG_TEXTURE);
    if (!t && !constant) { setError("gtlc_material_create: binding %u has no valid texture", b); return 0; }
    const bool okIndex = tb.paramIndex == GTLC_BIND_NORMAL || (tb.paramIndex >= 0 && tb.paramIndex + tb.componentCount <= GTLC_MATERIAL_PARAM_COUNT) ||
                         (tb.paramIndex >= GTLC_BIND_SCRATCH && tb.paramIndex + tb.componentCount <= GTLC_BIND_SCRATCH + GTLC_BIND_SCRATCH_COUNT);   // inside the block or inside the scratch floats, not across
    if (!okIndex || tb.componentCount < 1 || tb.componentCount > 3 || (tb.paramIndex == GTLC_BIND_NORMAL && tb.componentCount != 3)) { setError("gtlc_material_create: binding %u writes outside the parameter block", b); return 0; }
    for (int k = 0; k < 3; k++) if (tb.channels[k] < 0 || tb.channels[k] > 3) { setError("gtlc_material_create: binding %u has an invalid channel", b); return 0; }
    if (tb.wrapU < GTLC_WRAP_CLAMP || tb.wrapU > GTLC_WRAP_CLIP || tb.wrapV < GTLC_WRAP_CLAMP || tb.wrapV > GTLC_WRAP_CLIP) { setError("gtlc_material_create: binding %u has an invalid wrap mode", b); return 0; }
    if (tb.lookup < GTLC_LOOKUP_FILTERED_2D || tb.lookup > GTLC_LOOKUP_TEXEL_3D) { setError("gtlc_material_create: binding %u has an invalid lookup mode %d", b, tb.lookup); return 0; }
    if (tb.lookup >= GTLC_LOOKUP_FILTERED_3D && (tb.wrapW < GTLC_WRAP_CLAMP || tb.wrapW > GTLC_WRAP_CLIP)) { setError("gtlc_material_create: binding %u has an invalid wrap mode", b); return 0; }
    if (t && tb.lookup < GTLC_LOOKUP_FILTERED_3D && t->depth > 1u) { setError("gtlc_material_create: binding %u reads a 3-D texture with a 2-D lookup", b); return 0; }
    bound.push_back(t);
  }
  Material* m = new Material(); m->tag = TAG_MATERIAL; m->scene = s; m->closure = d->closure; memcpy(m->params, d->params, sizeof(m->params));
  if (d->name) m->name = d->name;
  if (d->textureCount) m->bindings.assign(d->textures, d->textures + d->textureCount);
  {                                                   // a binding that feeds the cutout input makes the any-hit test evaluate it per hit (the last one wins, as in shading)
    const int opacityParam = d->closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE ? (int) GTLC_UPS_OPACITY : (d->closure == GTLC_CLOSURE_OPENPBR ? (int) GTLC_OPBR_GEOMETRY_OPACITY : -2);
    for (uint32_t b = 0; b < d->textureCount; b++) {
      const GtlcTextureBinding& tb = d->textures[b];
      if (tb.paramIndex >= 0 && tb.paramIndex <= opacityParam && opacityParam < tb.paramIndex + tb.componentCount) {
        if (tb.componentCount != 1) { setError("gtlc_material_create: binding %u writes the cutout input as part of a %d-component value; bind it on its own", b, tb.componentCount); delete m; return 0; }
        if (tb.combine != 0 || m->opacityBinding >= 0) { setError("gtlc_material_create: binding %u: the cutout input takes ONE plain lookup (the any-hit test evaluates it at every candidate hit)", b); delete m; return 0; }
        m->opacityBinding = (int) b;
      }
    }
    for (uint32_t b = 0; b < d->primvarCount; b++)
      if (d->primvars[b].componentCount == 1 && d->primvars[b].paramIndex == opacityParam) m->opacityPrimvar = (int) b;
  }
  for (uint32_t b = 0; b < d->primvarCount; b++) { Material::PrimvarBinding pb; pb.paramIndex = d->primvars[b].paramIndex; pb.componentCount = d->primvars[b].componentCount; pb.name = d->primvars[b].name ? d->primvars[b].name : ""; memcpy(pb.fallback, d->primvars[b].fallback, 16); m->primvarBindings.push_back(pb); }
  {
    std::lock_guard<std::mutex> dl(g_destroyMutex);
    for (Texture* t : bound) if (t) t->refs++;
  }
  m->boundTextures = bound;
  std::lock_guard<std::mutex> l(s->mutex);
  s->materials.push_back(m);
  return handleOf(m);
}
void gtlc_material_destroy(GtlcHandle h)
{
  Material* m = as<Material>(h, TAG_MATERIAL); if (!m) return;
  std::lock_guard<std::mutex> dl(g_destroyMutex);
  Scene* s = m->scene;
  {
    std::lock_guard<std::mutex> l(s->mutex);
    s->materials.erase(std::remove(s->materials.begin(), s->materials.end(), m), s->materials.end());
    for (Mesh* mesh : s->meshes) if (mesh->material == m) mesh->material = nullptr;
    s->geometryDirty = true; s->framebufferDirty = true;
  }
  for (Texture* t : m->boundTextures) if (t) releaseTexture(t);
  m->tag = 0; delete m;
}

// ---- meshes --------------------------------------------------------------------------------
static bool copyPrimvars(const GtlcPrimvarDesc* src, uint32_t count, std::vector<Mesh::Primvar>& dst)
{
  dst.clear();
  if (count && !src) return false;
  for (uint32_t i = 0; i < count; i++) {
    if (!src[i].name || src[i].type < GTLC_PRIMVAR_FLOAT || src[i].type > GTLC_PRIMVAR_INT4 || src[i].interpolation < GTLC_INTERP_CONSTANT || src[i].interpolation > GTLC_INTERP_VERTEX) return false;
    Mesh::Primvar p; p.name = src[i].name; p.type = src[i].type; p.interpolation = src[i].interpolation;
    if (src[i].data && src[i].dataSize) p.data.assign((const uint8_t*) src[i].data, (const uint8_t*) src[i].data + src[i].dataSize);
    dst.push_back(std::move(p));
  }
  return true;
}
GtlcHandle gtlc_mesh_create(GtlcHandle sh, const GtlcMeshDesc* d)
{
  Scene* s = as<Scene>(sh, TAG_SCENE); if (!s || !d) return 0;
  if ((d->faceCount && (!d->faces || !d->vertices)) || (d->vertexCount && !d->vertices)) { setError("gtlc_mesh_create: null geometry"); return 0; }
  for (uint32_t f = 0; f < d->faceCount; f++) for (int c = 0; c < 3; c++)
    if (d->faces[f].v_i[c] >= d->vertexCount) { setError("gtlc_mesh_create: face %u references vertex %u >= %u", f, d->faces[f].v_i[c], d->vertexCount); return 0; }
  Mesh* m = new Mesh(); m->tag = TAG_MESH; m->scene = s;
  m->vertices.assign(d->vertices, d->vertices + d->vertexCount);      // giCreateMesh copies (Gi.cpp:620-631)
  m->faces.assign(d->faces, d->faces + d->faceCount);
  if (d->faceIds) m->faceIds.assign(d->faceIds, d->faceIds + d->faceCount);
  m->id = d->id; m->doubleSided = d->isDoubleSided != 0; m->flipFacing = d->isLeftHanded != 0; m->visible = true;
  if (d->name) m->name = d->name;
  if (d->primvarCount && (!d->primvars || !copyPrimvars(d->primvars, d->primvarCount, m->primvars))) { setError("gtlc_mesh_create: invalid primvar description"); delete m; return 0; }
  const float ident[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
  memcpy(m->transform, ident, sizeof(ident));
  m->instanceTransforms.assign(ident, ident + 12);
  m->instanceIds.assign(1, 0);
  m->material = nullptr;
  std::lock_guard<std::mutex> l(s->mutex);
  s->meshSet.insert(m); s->meshes.push_back(m);
  s->geometryDirty = true; s->framebufferDirty = true;
  return handleOf(m);
}
void gtlc_mesh_set_transform(GtlcHandle h, const float* m16)
{
  Mesh* m = as<Mesh>(h, TAG_MESH); if (!m || !m16) return;
  std::lock_guard<std::mutex> l(m->scene->mutex);
  usdToRow3x4(m16, m->transform);
  m->scene->geometryDirty = true; m->scene->framebufferDirty = true;
}
void gtlc_mesh_set_instance_transforms(GtlcHandle h, uint32_t count, const float* t16)
{
  Mesh* m = as<Mesh>(h, TAG_MESH); if (!m || (count && !t16)) return;
  std::lock_guard<std::mutex> l(m->scene->mutex);
  m->instanceTransforms.resize((size_t) count * 12);
  for (uint32_t i = 0; i < count; i++) usdToRow3x4(t16 + (size_t) i * 16, &m->instanceTransforms[(size_t) i * 12]);
  if (m->instanceIds.size() != count) { m->instanceIds.resize(count); for (uint32_t i = 0; i < count; i++) m->instanceIds[i] = (int32_t) i; }
  m->scene->geometryDirty = true; m->scene->framebufferDirty = true;
}
void gtlc_mesh_set_instancer_primvars(GtlcHandle h, uint32_t count, const GtlcPrimvarDesc* primvars)
{
  Mesh* m = as<Mesh>(h, TAG_MESH); if (!m) return;
  std::lock_guard<std::mutex> l(m->scene->mutex);
  if (!copyPrimvars(primvars, count, m->instancerPrimvars)) setError("gtlc_mesh_set_instancer_primvars: invalid primvar description");
  m->scene->geometryDirty = true; m->scene->framebufferDirty = true;
}
void gtlc_mesh_set_instance_ids(GtlcHandle h, uint32_t count, const int32_t* ids)
{
  Mesh* m = as<Mesh>(h, TAG_MESH); if (!m || (count && !ids)) return;
  std::lock_guard<std::mutex> l(m->scene->mutex);
  m->instanceIds.assign(ids, ids + count);
  m->scene->geometryDirty = true; m->scene->framebufferDirty = true;
}
void gtlc_mesh_set_material(GtlcHandle h, GtlcHandle mat)
{
  Mesh* m = as<Mesh>(h, TAG_MESH); if (!m) return;
  Material* mm = mat ? as<Material>(mat, TAG_MATERIAL) : nullptr;
  std::lock_guard<std::mutex> l(m->scene->mutex);
  m->material = mm;
  m->scene->geometryDirty = true; m->scene->framebufferDirty = true;
}
void gtlc_mesh_set_visibility(GtlcHandle h, int32_t visible)
{
  Mesh* m = as<Mesh>(h, TAG_MESH); if (!m) return;
  std::lock_guard<std::mutex> l(m->scene->mutex);
  m->visible = visible != 0;
  m->scene->geometryDirty = true; m->scene->framebufferDirty = true;
}
void gtlc_mesh_destroy(GtlcHandle h)
{
  Mesh* m = as<Mesh>(h, TAG_MESH); if (!m) return;
  std::lock_guard<std::mutex> dl(g_destroyMutex);
  Scene* s = m->scene;
  {
    std::lock_guard<std::mutex> l(s->mutex);
    s->meshSet.erase(m);
    s->meshes.erase(std::remove(s->meshes.begin(), s->meshes.end(), m), s->meshes.end());
    s->geometryDirty = true; s->framebufferDirty = true;
  }
  m->tag = 0; delete m;
}

// ---- lights (Gi.cpp:2573-2941) ---------------------------------------------------------------
#define LIGHT_COMMON(TYPE, TAGV, VEC)                                                                        \
  static TYPE* new_##TYPE(Scene* s) { TYPE* l = new TYPE(); l->tag = TAGV; l->scene = s; memset(&l->data, 0, sizeof(l->data)); return l; } \
  static void destroy_##TYPE(Scene* s, TYPE* l) { std::lock_guard<std::mutex> g2(s->mutex); s->VEC.erase(std::remove(s->VEC.begin(), s->VEC.end(), l), s->VEC.end()); s->lightsDirty = true; s->framebufferDirty = true; l->tag = 0; delete l; }
LIGHT_COMMON(SphereLight, TAG_SPHERE, sphereLights)
LIGHT_COMMON(DistantLight, TAG_DISTANT, distantLights)
LIGHT_COMMON(RectLight, TAG_RECT, rectLights)
LIGHT_COMMON(DiskLight, TAG_DISK, diskLights)
// setters write the light's record under the scene lock: uploadLights / gtlc_render read it under the same lock
#define LIGHT_LOCK(l) std::lock_guard<std::mutex> g2((l)->scene->mutex)
#define LIGHT_MARK(l) do { (l)->scene->lightsDirty = true; (l)->scene->framebufferDirty = true; } while (0)

GtlcHandle gtlc_sphere_light_create(GtlcHandle sh)
{
  Scene* s = as<Scene>(sh, TAG_SCENE); if (!s) return 0;
  SphereLight* l = new_SphereLight(s);                                  // defaults Gi.cpp:2584-2594
  l->data.diffuseSpecularPacked = packHalf2x16(1.0f, 1.0f); l->data.area = 1.0f;
  l->data.radiusXYZ[0] = l->data.radiusXYZ[1] = l->data.radiusXYZ[2] = 0.5f;
  std::lock_guard<std::mutex> g2(s->mutex); s->sphereLights.push_back(l); s->lightsDirty = true; s->framebufferDirty = true;
  return handleOf(l);
}
void gtlc_sphere_light_destroy(GtlcHandle sh, GtlcHandle h) { Scene* s = as<Scene>(sh, TAG_SCENE); SphereLight* l = as<SphereLight>(h, TAG_SPHERE); if (s && l) destroy_SphereLight(s, l); }
void gtlc_sphere_light_set_position(GtlcHandle h, const float* p) { SphereLight* l = as<SphereLight>(h, TAG_SPHERE); if (!l || !p) return; LIGHT_LOCK(l); memcpy(l->data.pos, p, 12); LIGHT_MARK(l); }
void gtlc_sphere_light_set_base_emission(GtlcHandle h, const float* c) { SphereLight* l = as<SphereLight>(h, TAG_SPHERE); if (!l || !c) return; LIGHT_LOCK(l); memcpy(l->data.baseEmission, c, 12); LIGHT_MARK(l); }
void gtlc_sphere_light_set_radius(GtlcHandle h, float rx, float ry, float rz)
{
  SphereLight* l = as<SphereLight>(h, TAG_SPHERE); if (!l) return;
  float ab = powf(rx * ry, 1.6f), ac = powf(rx * rz, 1.6f), bc = powf(ry * rz, 1.6f);    // Knud Thomsen, Gi.cpp:2637-2640
  float area = float(powf((ab + ac + bc) / 3.0f, 1.0f / 1.6f) * 4.0f * M_PI);
  LIGHT_LOCK(l);
  l->data.radiusXYZ[0] = rx; l->data.radiusXYZ[1] = ry; l->data.radiusXYZ[2] = rz; l->data.area = area; LIGHT_MARK(l);
}
void gtlc_sphere_light_set_diffuse_specular(GtlcHandle h, float d, float sp) { SphereLight* l = as<SphereLight>(h, TAG_SPHERE); if (!l) return; LIGHT_LOCK(l); l->data.diffuseSpecularPacked = packHalf2x16(d, sp); LIGHT_MARK(l); }

GtlcHandle gtlc_distant_light_create(GtlcHandle sh)
{
  Scene* s = as<Scene>(sh, TAG_SCENE); if (!s) return 0;
  DistantLight* l = new_DistantLight(s);                                // Gi.cpp:2680-2688
  l->data.diffuseSpecularPacked = packHalf2x16(1.0f, 1.0f); l->data.invPdf = 1.0f;
  std::lock_guard<std::mutex> g2(s->mutex); s->distantLights.push_back(l); s->lightsDirty = true; s->framebufferDirty = true;
  return handleOf(l);
}
void gtlc_distant_light_destroy(GtlcHandle sh, GtlcHandle h) { Scene* s = as<Scene>(sh, TAG_SCENE); DistantLight* l = as<DistantLight>(h, TAG_DISTANT); if (s && l) destroy_DistantLight(s, l); }
void gtlc_distant_light_set_direction(GtlcHandle h, const float* d) { DistantLight* l = as<DistantLight>(h, TAG_DISTANT); if (!l || !d) return; LIGHT_LOCK(l); memcpy(l->data.direction, d, 12); LIGHT_MARK(l); }
void gtlc_distant_light_set_base_emission(GtlcHandle h, const float* c) { DistantLight* l = as<DistantLight>(h, TAG_DISTANT); if (!l || !c) return; LIGHT_LOCK(l); memcpy(l->data.baseEmission, c, 12); LIGHT_MARK(l); }
void gtlc_distant_light_set_angle(GtlcHandle h, float angle)
{
  DistantLight* l = as<DistantLight>(h, TAG_DISTANT); if (!l) return;
  float halfAngle = 0.5f * angle;                                       // Gi.cpp:2723-2726
  float invPdf = (halfAngle > 0.0f) ? float(2.0f * M_PI * (1.0f - cosf(halfAngle))) : 1.0f;
  LIGHT_LOCK(l);
  l->data.angle = angle; l->data.invPdf = invPdf; LIGHT_MARK(l);
}
void gtlc_distant_light_set_diffuse_specular(GtlcHandle h, float d, float sp) { DistantLight* l = as<DistantLight>(h, TAG_DISTANT); if (!l) return; LIGHT_LOCK(l); l->data.diffuseSpecularPacked = packHalf2x16(d, sp); LIGHT_MARK(l); }

GtlcHandle gtlc_rect_light_create(GtlcHandle sh)
{
  Scene* s = as<Scene>(sh, TAG_SCENE); if (!s) return 0;
  RectLight* l = new_RectLight(s);                                      // Gi.cpp:2757-2771
  const float t0[3] = { 1, 0, 0 }, t1[3] = { 0, 1, 0 };
  l->data.width = 1.0f; l->data.height = 1.0f; l->data.tangentFramePacked[0] = encodeDirection(t0); l->data.tangentFramePacked[1] = encodeDirection(t1);
  l->data.diffuseSpecularPacked = packHalf2x16(1.0f, 1.0f);
  std::lock_guard<std::mutex> g2(s->mutex); s->rectLights.push_back(l); s->lightsDirty = true; s->framebufferDirty = true;
  return handleOf(l);
}
void gtlc_rect_light_destroy(GtlcHandle sh, GtlcHandle h) { Scene* s = as<Scene>(sh, TAG_SCENE); RectLight* l = as<RectLight>(h, TAG_RECT); if (s && l) destroy_RectLight(s, l); }
void gtlc_rect_light_set_origin(GtlcHandle h, const float* o) { RectLight* l = as<RectLight>(h, TAG_RECT); if (!l || !o) return; LIGHT_LOCK(l); memcpy(l->data.origin, o, 12); LIGHT_MARK(l); }
void gtlc_rect_light_set_tangents(GtlcHandle h, const float* t0, const float* t1) { RectLight* l = as<RectLight>(h, TAG_RECT); if (!l || !t0 || !t1) return; LIGHT_LOCK(l); l->data.tangentFramePacked[0] = encodeDirection(t0); l->data.tangentFramePacked[1] = encodeDirection(t1); LIGHT_MARK(l); }
void gtlc_rect_light_set_base_emission(GtlcHandle h, const float* c) { RectLight* l = as<RectLight>(h, TAG_RECT); if (!l || !c) return; LIGHT_LOCK(l); memcpy(l->data.baseEmission, c, 12); LIGHT_MARK(l); }
void gtlc_rect_light_set_dimensions(GtlcHandle h, float w, float hh) { RectLight* l = as<RectLight>(h, TAG_RECT); if (!l) return; LIGHT_LOCK(l); l->data.width = w; l->data.height = hh; LIGHT_MARK(l); }
void gtlc_rect_light_set_diffuse_specular(GtlcHandle h, float d, float sp) { RectLight* l = as<RectLight>(h, TAG_RECT); if (!l) return; LIGHT_LOCK(l); l->data.diffuseSpecularPacked = packHalf2x16(d, sp); LIGHT_MARK(l); }

GtlcHandle gtlc_disk_light_create(GtlcHandle sh)
{
  Scene* s = as<Scene>(sh, TAG_SCENE); if (!s) return 0;
  DiskLight* l = new_DiskLight(s);                                      // Gi.cpp:2852-2866
  const float t0[3] = { 1, 0, 0 }, t1[3] = { 0, 1, 0 };
  l->data.radiusX = 0.5f; l->data.radiusY = 0.5f; l->data.tangentFramePacked[0] = encodeDirection(t0); l->data.tangentFramePacked[1] = encodeDirection(t1);
  l->data.diffuseSpecularPacked = packHalf2x16(1.0f, 1.0f);
  std::lock_guard<std::mutex> g2(s->mutex); s->diskLights.push_back(l); s->lightsDirty = true; s->framebufferDirty = true;
  return handleOf(l);
}
void gtlc_disk_light_destroy(GtlcHandle sh, GtlcHandle h) { Scene* s = as<Scene>(sh, TAG_SCENE); DiskLight* l = as<DiskLight>(h, TAG_DISK); if (s && l) destroy_DiskLight(s, l); }
void gtlc_disk_light_set_origin(GtlcHandle h, const float* o) { DiskLight* l = as<DiskLight>(h, TAG_DISK); if (!l || !o) return; LIGHT_LOCK(l); memcpy(l->data.origin, o, 12); LIGHT_MARK(l); }
void gtlc_disk_light_set_tangents(GtlcHandle h, const float* t0, const float* t1) { DiskLight* l = as<DiskLight>(h, TAG_DISK); if (!l || !t0 || !t1) return; LIGHT_LOCK(l); l->data.tangentFramePacked[0] = encodeDirection(t0); l->data.tangentFramePacked[1] = encodeDirection(t1); LIGHT_MARK(l); }
void gtlc_disk_light_set_base_emission(GtlcHandle h, const float* c) { DiskLight* l = as<DiskLight>(h, TAG_DISK); if (!l || !c) return; LIGHT_LOCK(l); memcpy(l->data.baseEmission, c, 12); LIGHT_MARK(l); }
void gtlc_disk_light_set_radius(GtlcHandle h, float rx, float ry) { DiskLight* l = as<DiskLight>(h, TAG_DISK); if (!l) return; LIGHT_LOCK(l); l->data.radiusX = rx; l->data.radiusY = ry; LIGHT_MARK(l); }
void gtlc_disk_light_set_diffuse_specular(GtlcHandle h, float d, float sp) { DiskLight* l = as<DiskLight>(h, TAG_DISK); if (!l) return; LIGHT_LOCK(l); l->data.diffuseSpecularPacked = packHalf2x16(d, sp); LIGHT_MARK(l); }

// ---- dome light (Gi.cpp:2943-3000) -----------------------------------------------------------
GtlcHandle gtlc_dome_light_create(GtlcHandle sh, uint32_t w, uint32_t hgt, const float* rgba)
{
  Scene* s = as<Scene>(sh, TAG_SCENE); if (!s || !rgba || !w || !hgt) { setError("gtlc_dome_light_create: invalid arguments"); return 0; }
  if (!requireInit()) return 0;
  DomeLight* d = new DomeLight(); d->tag = TAG_DOME; d->scene = s; d->width = w; d->height = hgt;
  d->texels.assign(rgba, rgba + (size_t) w * hgt * 4);
  d->rotation[0] = d->rotation[1] = d->rotation[2] = 0.0f; d->rotation[3] = 1.0f;
  d->baseEmission[0] = d->baseEmission[1] = d->baseEmission[2] = 1.0f; d->diffuse = d->specular = 1.0f; d->dTexels = nullptr;
  cudaSetDevice(g.device);
  if (cudaMalloc(&d->dTexels, (size_t) w * hgt * 16) != cudaSuccess || cudaMemcpy(d->dTexels, rgba, (size_t) w * hgt * 16, cudaMemcpyHostToDevice) != cudaSuccess) {
    setError("dome light upload failed"); delete d; return 0;
  }
  return handleOf(d);
}
void gtlc_dome_light_destroy(GtlcHandle h)
{
  DomeLight* d = as<DomeLight>(h, TAG_DOME); if (!d) return;
  std::lock_guard<std::mutex> rl(g_renderMutex);      // a render in flight may be sampling the texels (gtlc_render holds this lock until its stream is idle)
  { std::lock_guard<std::mutex> l(d->scene->mutex); d->scene->framebufferDirty = true; }
  if (d->dTexels) { cudaSetDevice(g.device); cudaFree(d->dTexels); }
  d->tag = 0; delete d;
}
void gtlc_dome_light_set_rotation(GtlcHandle h, const float* q) { DomeLight* d = as<DomeLight>(h, TAG_DOME); if (!d || !q) return; std::lock_guard<std::mutex> l(d->scene->mutex); memcpy(d->rotation, q, 16); d->scene->framebufferDirty = true; }
void gtlc_dome_light_set_base_emission(GtlcHandle h, const float* c) { DomeLight* d = as<DomeLight>(h, TAG_DOME); if (!d || !c) return; std::lock_guard<std::mutex> l(d->scene->mutex); memcpy(d->baseEmission, c, 12); d->scene->framebufferDirty = true; }
void gtlc_dome_light_set_diffuse_specular(GtlcHandle h, float df, float sp) { DomeLight* d = as<DomeLight>(h, TAG_DOME); if (!d) return; std::lock_guard<std::mutex> l(d->scene->mutex); d->diffuse = df; d->specular = sp; d->scene->framebufferDirty = true; }

// ---- render buffers (Gi.cpp:2995-3006) -------------------------------------------------------
GtlcHandle gtlc_render_buffer_create(uint32_t w, uint32_t h, int32_t format)
{
  if (!requireInit()) return 0;
  if (format < GTLC_FORMAT_INT32 || format > GTLC_FORMAT_FLOAT32_VEC4) { setError("bad render buffer format %d", format); return 0; }
  cudaSetDevice(g.device);
  RenderBuffer* rb = new RenderBuffer(); rb->tag = TAG_RB; rb->width = w; rb->height = h; rb->format = format;
  rb->size = (size_t) w * h * (format == GTLC_FORMAT_FLOAT32_VEC4 ? 16 : 4);
  rb->hostMem = nullptr; rb->devMem = nullptr;
  if (cudaHostAlloc(&rb->hostMem, std::max<size_t>(rb->size, 16), cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&rb->devMem, std::max<size_t>(rb->size, 16)) != cudaSuccess) {
    setError("render buffer allocation failed (%zu bytes)", rb->size); if (rb->hostMem) cudaFreeHost(rb->hostMem); delete rb; return 0;
  }
  memset(rb->hostMem, 0, rb->size); cudaMemset(rb->devMem, 0, rb->size);
  return handleOf(rb);
}
void gtlc_render_buffer_destroy(GtlcHandle h) { RenderBuffer* rb = as<RenderBuffer>(h, TAG_RB); if (!rb) return; cudaSetDevice(g.device); cudaStreamSynchronize(g.stream); cudaFreeHost(rb->hostMem); cudaFree(rb->devMem); rb->tag = 0; delete rb; }
void* gtlc_render_buffer_mem(GtlcHandle h) { RenderBuffer* rb = as<RenderBuffer>(h, TAG_RB); return rb ? rb->hostMem : nullptr; }
void* gtlc_render_buffer_device_mem(GtlcHandle h) { RenderBuffer* rb = as<RenderBuffer>(h, TAG_RB); return rb ? rb->devMem : nullptr; }

// ---- BVH access ------------------------------------------------------------------------------
static bool syncSceneToDevice(Scene* s);
int gtlc_scene_get_bvh(GtlcHandle h, GtlcBvhView* view)
{
  Scene* s = as<Scene>(h, TAG_SCENE); if (!s || !view) return GTLC_ERROR;
  std::unique_lock<std::mutex> rl(g_renderMutex, std::defer_lock);
  if (g.tuning.bvhBuilder >= 1) rl.lock();     // the device builders run on the shared stream
  std::lock_guard<std::mutex> l(s->mutex);
  if (g.tuning.bvhBuilder >= 1) {           // device builders: build if needed, then read the result back once
    if (!requireInit()) return GTLC_ERROR;
    cudaSetDevice(g.device);
    if (!syncSceneToDevice(s)) return GTLC_ERROR;
    if (s->bvhHostStale) {
      s->bvh.nodes.resize(s->nodeCount); s->bvh.tris.resize((size_t) s->triCount * 12); s->bvh.maxDepth = s->maxDepth;
      if (cudaMemcpy(s->bvh.nodes.data(), s->dNodes.ptr, (size_t) s->nodeCount * 80, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(s->bvh.tris.data(), s->dTris.ptr, (size_t) s->triCount * 48, cudaMemcpyDeviceToHost) != cudaSuccess) { setError("BVH read-back failed"); return GTLC_ERROR; }
      s->bvhHostStale = false;
    }
  } else if (s->geometryDirty || !s->bvhValid || s->bvhHostStale) {
    // (stale = the tree in HBM came from a device builder and the tuning went back to the host builder: rebuild here and
    // drop the device copy so that the next render / ray query uploads the tree this call returns)
    const bool wasStale = s->bvhHostStale;
    if (!buildHostGeometry(s)) return GTLC_ERROR;          // device upload happens lazily in render/trace
    if (wasStale) { cudaSetDevice(g.device); cudaStreamSynchronize(g.stream); s->dNodes.release(); }
  }
  view->nodes = s->bvh.nodes.data(); view->nodeCount = (uint32_t) s->bvh.nodes.size();
  view->triangles = s->bvh.tris.data(); view->triCount = (uint32_t) (s->bvh.tris.size() / 12);
  return GTLC_OK;
}

static bool syncSceneToDevice(Scene* s)
{
  if (s->geometryDirty || !s->bvhValid || !s->dNodes.ptr) {
    const bool hostBvh = g.tuning.bvhBuilder < 1;
    if (s->geometryDirty || !s->bvhValid) { if (!buildHostGeometry(s, hostBvh)) return false; }
    if (!uploadGeometry(s, hostBvh)) return false;
    s->geometryDirty = false; s->materialsDirty = true;
    g.stats.msBvhBuild = (float) s->lastBuildMs;
  }
  if (s->materialsDirty) { if (!uploadMaterials(s)) return false; s->materialsDirty = false; }
  if (s->lightsDirty) { if (!uploadLights(s)) return false; s->lightsDirty = false; }
  g.stats.bvhNodeCount = s->nodeCount; g.stats.bvhTriCount = s->triCount; g.stats.bvhMaxDepth = s->maxDepth;
  g.currentTriCount = s->triCount;
  return true;
}

// ---- ray queries (parity tests / roofline bench) ---------------------------------------------
int gtlc_trace_rays(GtlcHandle h, const float* rays, uint32_t count, int32_t anyHit, void* hitsOut, uint32_t* visitsOut)
{
  if (!requireInit()) return GTLC_ERROR;
  Scene* s = as<Scene>(h, TAG_SCENE); if (!s) return GTLC_ERROR;
  if ((count && !rays) || !hitsOut) { setError("gtlc_trace_rays: null ray / hit array"); return GTLC_ERROR; }
  // tMin >= 0 is Vulkan's rule for traceRayEXT (VUID-RuntimeSpirv-OpRayQueryInitializeKHR-06351 / -RayTmin-06351), and the
  // batched kernel relies on it: it orders hits by the raw bits of t, which equals float order only for t >= 0.
  for (uint32_t i = 0; i < count; i++) {
    if (rays[(size_t) i * 8 + 3] < 0.0f) { setError("gtlc_trace_rays: ray %u has tMin %g < 0 (tMin must be non-negative, as for traceRayEXT)", i, (double) rays[(size_t) i * 8 + 3]); return GTLC_ERROR; }
  }
  std::lock_guard<std::mutex> rl(g_renderMutex);
  std::lock_guard<std::mutex> l(s->mutex);
  cudaSetDevice(g.device);
  if (!syncSceneToDevice(s)) return GTLC_ERROR;
  auto body = [&]() -> bool {
    if (!ensurePool(std::max<size_t>(count, 1))) return false;
    std::vector<float4> o(count), d(count);
    for (uint32_t i = 0; i < count; i++) { const float* r = rays + (size_t) i * 8; o[i] = make_float4(r[0], r[1], r[2], r[3]); d[i] = make_float4(r[4], r[5], r[6], r[7]); }
    CUDA_OK(cudaMemcpyAsync(g.q[0].rayO, o.data(), (size_t) count * 16, cudaMemcpyHostToDevice, g.stream));
    CUDA_OK(cudaMemcpyAsync(g.q[0].rayD, d.data(), (size_t) count * 16, cudaMemcpyHostToDevice, g.stream));
    Counters c; memset(&c, 0, sizeof(c)); c.numPaths[0] = count;
    CUDA_OK(cudaMemcpyAsync(g.counters, &c, sizeof(c), cudaMemcpyHostToDevice, g.stream));
    const bool countVisits = visitsOut != nullptr || g.tuning.countVisits;
    if (visitsOut) {
      if (g.visitsCapacity < count) { cudaFree(g.visits); g.visits = nullptr; CUDA_OK(cudaMalloc(&g.visits, (size_t) count * 8)); g.visitsCapacity = count; }
    }
    TraceParams P = makeTraceParams(s);
    P.rayO = g.q[0].rayO; P.rayD = g.q[0].rayD; P.hitTUVP = g.hitTUVP; P.hitInst = g.hitInst;
    P.numRays = &g.counters->numPaths[0]; P.workCounter = &g.counters->workCounter; P.visitsPerRay = visitsOut ? g.visits : nullptr;
    cudaEvent_t e0 = nextEvent(), e1 = nextEvent();
    cudaEventRecord(e0, g.stream);
    const int blocks = g.smCount * 8;
    if (anyHit) launchTrace<TRACE_ANY>(P, countVisits, blocks); else launchTrace<TRACE_CLOSEST>(P, countVisits, blocks);
    cudaEventRecord(e1, g.stream);
    CUDA_OK(cudaGetLastError());
    std::vector<float4> tuvp(count); std::vector<uint32_t> inst(count);
    CUDA_OK(cudaMemcpyAsync(tuvp.data(), g.hitTUVP, (size_t) count * 16, cudaMemcpyDeviceToHost, g.stream));
    CUDA_OK(cudaMemcpyAsync(inst.data(), g.hitInst, (size_t) count * 4, cudaMemcpyDeviceToHost, g.stream));
    if (visitsOut) CUDA_OK(cudaMemcpyAsync(visitsOut, g.visits, (size_t) count * 8, cudaMemcpyDeviceToHost, g.stream));
    CUDA_OK(cudaMemcpyAsync(&c, g.counters, sizeof(c), cudaMemcpyDeviceToHost, g.stream));
    CUDA_OK(cudaStreamSynchronize(g.stream));
    float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1);
    struct HitRec { float t, u, v; uint32_t prim, inst; };
    HitRec* out = (HitRec*) hitsOut;
    for (uint32_t i = 0; i < count; i++) { out[i].t = tuvp[i].x; out[i].u = tuvp[i].y; out[i].v = tuvp[i].z; memcpy(&out[i].prim, &tuvp[i].w, 4); out[i].inst = inst[i]; }
    memset(&g.stats, 0, offsetof(GtlcStats, msBvhBuild));
    if (anyHit) { g.stats.shadowRays = count; g.stats.msTraceAny = ms; } else { g.stats.closestRays = count; g.stats.msTraceClosest = ms; }
    g.stats.nodeVisits = c.stats[1]; g.stats.triVisits = c.stats[2]; g.stats.kernelLaunches = 1; g.stats.msTotal = ms;
    g.stats.maxNodeVisitsPerRay = (uint32_t) c.stats[4]; g.stats.nonFiniteRays = c.stats[5];
    return true;
  };
  g.eventsUsed = 0;
  return body() ? GTLC_OK : GTLC_ERROR;
}

// ---- giRender (Gi.cpp:1989-2524) -------------------------------------------------------------
int gtlc_render(const GtlcRenderParams* params)
{
  if (!requireInit()) return GTLC_ERROR;
  if (!params) { setError("gtlc_render: null params"); return GTLC_ERROR; }
  Scene* s = as<Scene>(params->scene, TAG_SCENE); if (!s) return GTLC_ERROR;
  std::lock_guard<std::mutex> rl(g_renderMutex);
  std::lock_guard<std::mutex> l(s->mutex);
  cudaSetDevice(g.device);
  const GtlcRenderSettings& rs = params->renderSettings;
  if (rs.mediumStackSize > kMaxMediumStack) { setError("mediumStackSize %u exceeds the %u nested media this build tracks per path", rs.mediumStackSize, kMaxMediumStack); return GTLC_ERROR; }
  if (rs.spp == 0) { setError("spp must be > 0"); return GTLC_ERROR; }

  // colour AOV + dimensions (Gi.cpp:2265-2267: the first bound render buffer defines the image size)
  RenderBuffer* colorRb = nullptr; uint32_t W = 0, H = 0; float clear[4] = { 0, 0, 0, 0 };
  for (uint32_t i = 0; i < params->aovBindingCount; i++) {
    const GtlcAovBinding& b = params->aovBindings[i];
    RenderBuffer* rb = as<RenderBuffer>(b.renderBuffer, TAG_RB); if (!rb) return GTLC_ERROR;
    if (!W) { W = rb->width; H = rb->height; }
    if (rb->width != W || rb->height != H) { setError("AOV render buffers differ in size"); return GTLC_ERROR; }
    if (b.aovId == GTLC_AOV_COLOR) {
      if (rb->format != GTLC_FORMAT_FLOAT32_VEC4) { setError("colour AOV needs a Float32Vec4 render buffer"); return GTLC_ERROR; }
      colorRb = rb; memcpy(clear, b.clearValue, 16);
    }
  }
  if (!W || !H) { setError("gtlc_render: no AOV bound"); return GTLC_ERROR; }
  if (W > 65535 || H > 65535) { setError("image dimensions exceed 65535 (imageDims packing, rp_main.h:39)"); return GTLC_ERROR; }

  // dirty-flag evaluation (Gi.cpp:2048-2052, 2125-2129)
  if (paramsResetAccumulation(s, params)) s->framebufferDirty = true;
  if (s->geometryDirty || s->materialsDirty || s->lightsDirty) s->framebufferDirty = true;
  if (!syncSceneToDevice(s)) return GTLC_ERROR;
  if (s->framebufferDirty) { s->sampleOffset = 0; s->framebufferDirty = false; }
  s->oldParams = *params; s->oldBindings.assign(params->aovBindings, params->aovBindings + params->aovBindingCount); s->oldParams.aovBindings = nullptr; s->haveOldParams = true;

  DomeLight* dome = params->domeLight ? as<DomeLight>(params->domeLight, TAG_DOME) : nullptr;

  // UniformData fill (Gi.cpp:2373-2426)
  UniformData ubo; memset(&ubo, 0, sizeof(ubo));
  auto norm3 = [](const float* v, float* o) { float len = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); float inv = 1.0f / len; o[0] = v[0] * inv; o[1] = v[1] * inv; o[2] = v[2] * inv; };
  if (dome) { memcpy(ubo.domeLightRotation, dome->rotation, 16); memcpy(ubo.domeLightEmissionMultiplier, dome->baseEmission, 12); }
  else { ubo.domeLightRotation[3] = 1.0f; ubo.domeLightEmissionMultiplier[0] = ubo.domeLightEmissionMultiplier[1] = ubo.domeLightEmissionMultiplier[2] = 1.0f; }
  ubo.sphereLightCount = (uint32_t) s->sphereLights.size(); ubo.distantLightCount = (uint32_t) s->distantLights.size();
  ubo.rectLightCount = (uint32_t) s->rectLights.size(); ubo.diskLightCount = (uint32_t) s->diskLights.size();
  ubo.totalLightCount = ubo.sphereLightCount + ubo.distantLightCount + ubo.rectLightCount + ubo.diskLightCount;
  ubo.metersPerSceneUnit = rs.metersPerSceneUnit;
  memcpy(ubo.cameraPosition, params->camera.position, 12);
  ubo.imageDims = (H << 16) | W;
  norm3(params->camera.forward, ubo.cameraForward); norm3(params->camera.up, ubo.cameraUp);
  ubo.focusDistance = params->camera.focusDistance; ubo.cameraVFoV = params->camera.vfov;
  ubo.sampleOffset = s->sampleOffset;
  ubo.lensRadius = (params->camera.fStop > 0.0f) ? (params->camera.focalLength / (2.0f * params->camera.fStop)) : 0.0f;
  ubo.spp = rs.spp; ubo.invTotalSampleCount = 1.0f / float(s->sampleOffset + rs.spp);
  ubo.maxSampleValue = rs.maxSampleValue; ubo.maxBouncesAndRrBounceOffset = (rs.maxBounces << 16) | rs.rrBounceOffset;
  ubo.rrInvMinTermProb = rs.rrInvMinTermProb; ubo.lightIntensityMultiplier = rs.lightIntensityMultiplier;
  ubo.clipRangePacked = packHalf2x16(params->camera.clipStart, params->camera.clipEnd); ubo.sensorExposure = params->camera.exposure;
  ubo.mediumStackSize = rs.mediumStackSize; ubo.maxVolumeWalkLength = rs.maxVolumeWalkLength;
  ubo.flags = (rs.jitteredSampling ? FLAG_JITTERED : 0u) | (rs.filterImportanceSampling ? FLAG_FIS : 0u) | (rs.depthOfField ? FLAG_DOF : 0u) |
              (rs.nextEventEstimation ? FLAG_NEE : 0u) | (rs.clippingPlanes ? FLAG_CLIPPING : 0u) | (rs.progressiveAccumulation ? FLAG_PROGRESSIVE : 0u) |
              (rs.domeLightCameraVisible ? FLAG_DOME_CAMERA_VISIBLE : 0u) | (dome ? FLAG_HAS_DOME : 0u);
  for (int k = 0; k < 3; k++) ubo.background[k] = (float) (uint8_t) (clear[k] * 255.0f) / 255.0f;   // Gi.cpp:2196

  auto body = [&]() -> bool {
    if (!buildPixelMap(s, W, H)) return false;
    const uint32_t localPixels = s->localPixels;
    // tiles: this rank's pixels, all samples.  sample ranges: all pixels, this rank's share of the call's samples (gtlc.h)
    const bool sampleShard = s->shardCount > 1 && s->shardTile == GTLC_SHARD_SAMPLES, tileShard = s->shardCount > 1 && !sampleShard;
    const uint32_t sBegin = sampleShard ? (uint32_t) ((uint64_t) rs.spp * s->shardRank / s->shardCount) : 0u;
    const uint32_t sEnd = sampleShard ? (uint32_t) ((uint64_t) rs.spp * (s->shardRank + 1) / s->shardCount) : rs.spp;
    const uint32_t maxPaths = std::max<uint32_t>(g.tuning.maxPathsPerWave, 1024u);
    uint32_t pixelsPerWave = std::min(localPixels, maxPaths);
    uint32_t samplesPerWave = std::max<uint32_t>(1u, std::min<uint32_t>(std::max(sEnd - sBegin, 1u), pixelsPerWave ? maxPaths / pixelsPerWave : 1u));
    if (pixelsPerWave < localPixels) samplesPerWave = 1;
    if (!ensurePool((size_t) pixelsPerWave * samplesPerWave)) return false;
    const bool volumes = rs.mediumStackSize > 0;
    if (volumes && !ensureVolumePool()) return false;
    if (g.accumCapacity < localPixels) { cudaFree(g.accum); g.accum = nullptr; CUDA_OK(cudaMalloc(&g.accum, std::max<size_t>(localPixels, 1) * 16)); g.accumCapacity = localPixels; }

    DeviceScene ds; memset(&ds, 0, sizeof(ds));
    ds.sverts = (const uint32_t*) s->dSverts.ptr; ds.sfaces = (const uint32_t*) s->dSfaces.ptr; ds.matLowered = s->dMatLowered.ptr;
    ds.instO2W = s->dO2W.ptr; ds.instW2O = s->dW2O.ptr; ds.instInfo = s->dInstInfo.ptr; ds.instIds = s->dInstIds.ptr;
    ds.matParams = s->dMatParams.ptr; ds.matClosure = s->dMatClosure.ptr; ds.matFlags = s->dMatFlags.ptr;
    ds.textures = s->dTextures.ptr; ds.texBindings = s->dTexBindings.ptr; ds.matTexRange = s->dMatTexRange.ptr;
    ds.primvarBindings = s->dPrimvarBindings.ptr; ds.matPrimvarRange = s->dMatPrimvarRange.ptr; ds.primvarInfos = s->dPrimvarInfos.ptr; ds.primvarData = s->dPrimvarData.ptr;
    ds.sphereLights = s->dSphere.ptr; ds.distantLights = s->dDistant.ptr; ds.rectLights = s->dRect.ptr; ds.diskLights = s->dDisk.ptr;
    if (dome) { ds.domeTexels = dome->dTexels; ds.domeW = dome->width; ds.domeH = dome->height; }

    float4* colorDev = colorRb ? (float4*) colorRb->devMem : nullptr;
    std::vector<TimedSpan> spans;
    g.eventsUsed = 0;
    uint64_t launches = 0;
    auto timed = [&](int kind, auto&& fn) { TimedSpan t; t.kind = kind; t.a = nextEvent(); t.b = nextEvent(); cudaEventRecord(t.a, g.stream); fn(); cudaEventRecord(t.b, g.stream); spans.push_back(t); launches++; };

    CUDA_OK(cudaMemsetAsync(g.counters, 0, sizeof(Counters), g.stream));
    cudaEvent_t evStart = nextEvent(), evEnd = nextEvent();
    cudaEventRecord(evStart, g.stream);
    // sharded render: pixels of other ranks are 0 so that one sum-reduce assembles the image
    if (colorDev && tileShard && s->sampleOffset == 0) CUDA_OK(cudaMemsetAsync(colorDev, 0, (size_t) W * H * 16, g.stream));
    // auxiliary AOVs (everything but Color): device targets + clear values for resolve_aov_kernel
    AovTargets targets; memset(&targets, 0, sizeof(targets));
    bool aovMode = false;
    for (uint32_t i = 0; i < params->aovBindingCount; i++) {
      const GtlcAovBinding& b = params->aovBindings[i];
      if (b.aovId == GTLC_AOV_COLOR || b.aovId < 0 || b.aovId >= GTLC_AOV_COUNT) continue;
      RenderBuffer* rb = as<RenderBuffer>(b.renderBuffer, TAG_RB);
      const bool isInt = b.aovId == GTLC_AOV_OBJECT_ID || b.aovId == GTLC_AOV_FACE_ID || b.aovId == GTLC_AOV_INSTANCE_ID;
      const bool isScalar = b.aovId == GTLC_AOV_DEPTH;
      const int32_t want = isInt ? GTLC_FORMAT_INT32 : (isScalar ? GTLC_FORMAT_FLOAT32 : GTLC_FORMAT_FLOAT32_VEC4);
      if (rb->format != want) { setError("AOV %d needs render buffer format %d (got %d)", b.aovId, want, rb->format); return false; }
      targets.buf[b.aovId] = rb->devMem;
      memcpy(&targets.clear[b.aovId], b.clearValue, 16);
      aovMode = true;
    }
    if (aovMode && !ensureAovSamples(g.poolCapacity)) return false;
    const bool clockMode = targets.buf[GTLC_AOV_CLOCK_CYCLES] != nullptr;          // every kernel then adds its cycles to the sample's counter
    AovSampleBuffers aovBufs = g.aovSamples; aovBufs.clk = clockMode ? g.aovSamples.clk : nullptr;

    const int gridElem = g.smCount * 8;
    const int traceBlocks = g.smCount * (g.tuning.reserved[1] > 0 ? g.tuning.reserved[1] : 8);
    const bool countVisits = g.tuning.countVisits != 0;
    const bool sortByMaterial = g.tuning.sortByMaterial > 0 || (g.tuning.sortByMaterial < 0 && s->flatMaterials.size() > 1);
    int maxClosure = GTLC_CLOSURE_DIFFUSE;
    for (Material* m : s->flatMaterials) maxClosure = std::max(maxClosure, (int) m->closure);
    uint64_t paths = 0;
    for (uint32_t pixBase = 0; pixBase < localPixels; pixBase += pixelsPerWave) {
      const uint32_t wavePixels = std::min(pixelsPerWave, localPixels - pixBase);
      // (a rank whose sample range is empty -- spp < worldSize -- still runs one wave of zero samples: its pixels blend to the previous partial)
      for (uint32_t sBase = sBegin; sBase < sEnd || sBase == sBegin; sBase += samplesPerWave) {
        WaveParams wp; wp.wavePixelBase = pixBase; wp.wavePixels = wavePixels; wp.waveSamples = std::min(samplesPerWave, sEnd - sBase);
        wp.waveSampleBase = sBase; wp.pixelMap = tileShard ? s->dPixelMap.ptr : nullptr;
        wp.firstWave = sBase == sBegin ? 1u : 0u; wp.alpha = (sampleShard && s->shardRank != 0) ? 0.0f : 1.0f;
        paths += (uint64_t) wp.wavePixels * wp.waveSamples;
        PathQueue q0 = g.q[0];
        if (!volumes) { q0.med = nullptr; q0.walk = nullptr; }             // (the pool may exist from an earlier render with media)
        timed(K_RAYGEN, [&] { raygen_kernel<<<gridElem, 256, 0, g.stream>>>(ubo, wp, q0, g.counters, aovBufs.clk); });
        int inIdx = 0;
        if (rs.maxBounces == 0) {
          // The bounce loop never runs (as in rgen:290 the payload's radiance stays 0): the finished-sample slots the
          // resolve kernels read would otherwise hold an earlier render's values.
          CUDA_OK(cudaMemsetAsync(g.sampleRad, 0, (size_t) wp.wavePixels * wp.waveSamples * 16, g.stream));
          if (aovMode) CUDA_OK(cudaMemsetAsync(g.aovSamples.a[0], 0, (size_t) wp.wavePixels * wp.waveSamples * 16, g.stream));
        }
        for (uint32_t bounce = 0; bounce < rs.maxBounces; bounce++) {
          TraceParams P = makeTraceParams(s);
          P.zeroA = &g.counters->numPaths[inIdx ^ 1]; P.zeroB = &g.counters->numShadow; P.rayStat = &g.counters->stats[0];
          P.rayO = g.q[inIdx].rayO; P.rayD = g.q[inIdx].rayD; P.hitTUVP = g.hitTUVP; P.hitInst = g.hitInst;
          P.numRays = &g.counters->numPaths[inIdx]; P.workCounter = &g.counters->workCounter; P.seedRad = g.q[inIdx].rad;
          if (clockMode) { P.clk = aovBufs.clk; P.aovSlotOf = g.q[inIdx].slot; }
          timed(K_TRACE, [&] { launchTrace<TRACE_CLOSEST>(P, countVisits, traceBlocks); });
          if (sortByMaterial) {
            timed(K_SORT, [&] {
              sort_count_kernel<<<g.sortBlocks, kSortBlock, 0, g.stream>>>(g.counters, inIdx, g.hitInst, ds.instInfo, g.sortHist);
              sort_scan_kernel<<<kSortBins, 1024, 0, g.stream>>>(g.sortHist, (uint32_t) g.sortBlocks, g.sortHist + (size_t) kSortBins * g.sortBlocks);
              sort_scatter_kernel<<<g.sortBlocks, kSortBlock, 0, g.stream>>>(g.counters, inIdx, g.hitInst, ds.instInfo, g.sortHist, g.sortHist + (size_t) kSortBins * g.sortBlocks, g.perm);
            });
            launches += 2;
          }
          timed(K_SHADE, [&] {
            const uint32_t* permArg = sortByMaterial ? g.perm : nullptr;
#define GTLC_LAUNCH_SHADE4(MAXC, NEE, AOV, TEX, VOL) shade_kernel<MAXC, NEE, AOV, TEX, VOL><<<gridElem * 256 / shade_block_size(MAXC), shade_block_size(MAXC), 0, g.stream>>>(ubo, ds, g.q[inIdx], g.q[inIdx ^ 1], g.hitTUVP, g.hitInst, permArg, g.shO, g.shD, g.shC, g.shSeed, g.sampleRad, g.counters, inIdx, aovBufs)
// the medium stack only ever changes on transmission events, which only the OpenPBR closure produces
#define GTLC_LAUNCH_SHADE3(MAXC, NEE, AOV, TEX) do { if (volumes && MAXC == GTLC_CLOSURE_OPENPBR) GTLC_LAUNCH_SHADE4(MAXC, NEE, AOV, TEX, (MAXC == GTLC_CLOSURE_OPENPBR)); else GTLC_LAUNCH_SHADE4(MAXC, NEE, AOV, TEX, false); } while (0)
#define GTLC_LAUNCH_SHADE2(MAXC, NEE, AOV) do { if (s->hasTextures) GTLC_LAUNCH_SHADE3(MAXC, NEE, AOV, true); else GTLC_LAUNCH_SHADE3(MAXC, NEE, AOV, false); } while (0)
// (with the ClockCycles AOV bound the AOV instantiation runs on every bounce: its first-bounce captures test the path's own bounce count)
#define GTLC_LAUNCH_SHADE(MAXC, NEE) do { if (aovMode && (bounce == 0 || clockMode)) GTLC_LAUNCH_SHADE2(MAXC, NEE, true); else GTLC_LAUNCH_SHADE2(MAXC, NEE, false); } while (0)
            const bool nee = rs.nextEventEstimation != 0;
            if (maxClosure == GTLC_CLOSURE_DIFFUSE) { if (nee) GTLC_LAUNCH_SHADE(GTLC_CLOSURE_DIFFUSE, true); else GTLC_LAUNCH_SHADE(GTLC_CLOSURE_DIFFUSE, false); }
            else if (maxClosure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) { if (nee) GTLC_LAUNCH_SHADE(GTLC_CLOSURE_USD_PREVIEW_SURFACE, true); else GTLC_LAUNCH_SHADE(GTLC_CLOSURE_USD_PREVIEW_SURFACE, false); }
            else { if (nee) GTLC_LAUNCH_SHADE(GTLC_CLOSURE_OPENPBR, true); else GTLC_LAUNCH_SHADE(GTLC_CLOSURE_OPENPBR, false); }
#undef GTLC_LAUNCH_SHADE
#undef GTLC_LAUNCH_SHADE2
#undef GTLC_LAUNCH_SHADE3
#undef GTLC_LAUNCH_SHADE4
          });
          if (rs.nextEventEstimation) {
            TraceParams S = makeTraceParams(s);
            S.rayStat = &g.counters->stats[3];
            S.rayO = g.shO; S.rayD = g.shD; S.shContrib = g.shC; S.radPath = g.q[inIdx ^ 1].rad; S.radSample = g.sampleRad;
            S.numRays = &g.counters->numShadow; S.workCounter = &g.counters->workCounterShadow; S.seedU32 = g.shSeed;
            if (aovMode && bounce == 0) { S.aovFlags = g.aovSamples.a[0]; S.aovSlotOf = g.q[inIdx ^ 1].slot; }
            if (clockMode) { S.clk = aovBufs.clk; S.aovSlotOf = g.q[inIdx ^ 1].slot; }
            timed(K_SHADOW, [&] { launchTrace<TRACE_SHADOW>(S, countVisits, traceBlocks); });
          }
          inIdx ^= 1;
        }
        const int lastWave = (sBase + samplesPerWave >= sEnd) ? 1 : 0;
        timed(K_RESOLVE, [&] { resolve_kernel<<<gridElem, 256, 0, g.stream>>>(ubo, wp, g.sampleRad, g.accum, colorDev, lastWave); });
        if (aovMode) timed(K_RESOLVE, [&] { resolve_aov_kernel<<<gridElem, 256, 0, g.stream>>>(ubo, wp, g.sampleRad, aovBufs, targets, sBase == sBegin ? 1 : 0, lastWave); });
      }
    }
    // the single collective of the path: every rank's colour AOV holds its own tiles and zeros elsewhere;
    // one sum-reduce to rank 0, on the render stream, before the device->host copy.  The result goes to a buffer of its
    // own on the root so that the shard buffers keep "foreign pixels = 0" for the next progressive call.
    const bool reduceColor = g.nccl.comm && s->shardCount > 1 && colorDev;
    if (reduceColor) {
      if ((uint32_t) g.nccl.world != s->shardCount || (uint32_t) g.nccl.rank != s->shardRank) { setError("gtlc_render: the scene's shard (%u of %u) does not match the communicator (%d of %d)", s->shardRank, s->shardCount, g.nccl.rank, g.nccl.world); return false; }
      const size_t px = (size_t) W * H;
      if (g.nccl.rank == 0 && g.reducedCapacity < px) { cudaFree(g.reduced); g.reduced = nullptr; g.reducedCapacity = 0; CUDA_OK(cudaMalloc(&g.reduced, px * 16)); g.reducedCapacity = px; }
      NCCL_OK(g.nccl.Reduce(colorDev, g.nccl.rank == 0 ? (void*) g.reduced : (void*) colorDev, px * 4, ncclFloat, ncclSum, 0, g.nccl.comm, g.stream));
      launches++;
    }
    cudaEventRecord(evEnd, g.stream);
    CUDA_OK(cudaGetLastError());
    // readback of every bound AOV (Gi.cpp:2439-2489) + blocking wait (Gi.cpp:2492-2502)
    Counters c;
    CUDA_OK(cudaMemcpyAsync(&c, g.counters, sizeof(c), cudaMemcpyDeviceToHost, g.stream));
    if (!g.tuning.skipReadback) {
      for (uint32_t i = 0; i < params->aovBindingCount; i++) {
        RenderBuffer* rb = as<RenderBuffer>(params->aovBindings[i].renderBuffer, TAG_RB);
        if (reduceColor && rb == colorRb) {           // the assembled image exists on rank 0 only
          if (g.nccl.rank == 0) CUDA_OK(cudaMemcpyAsync(rb->hostMem, g.reduced, rb->size, cudaMemcpyDeviceToHost, g.stream));
          continue;
        }
        CUDA_OK(cudaMemcpyAsync(rb->hostMem, rb->devMem, rb->size, cudaMemcpyDeviceToHost, g.stream));
      }
    }
    CUDA_OK(cudaStreamSynchronize(g.stream));
    float keepBuild = g.stats.msBvhBuild, keepDev = g.stats.msBvhDeviceBuild; uint32_t kn = g.stats.bvhNodeCount, kt = g.stats.bvhTriCount, kd = g.stats.bvhMaxDepth;
    memset(&g.stats, 0, sizeof(g.stats));
    g.stats.msBvhBuild = keepBuild; g.stats.msBvhDeviceBuild = keepDev; g.stats.bvhNodeCount = kn; g.stats.bvhTriCount = kt; g.stats.bvhMaxDepth = kd;
    cudaEventElapsedTime(&g.stats.msTotal, evStart, evEnd);
    for (const TimedSpan& t : spans) {
      float ms = 0.0f; cudaEventElapsedTime(&ms, t.a, t.b);
      switch (t.kind) {
        case K_RAYGEN: g.stats.msRaygen += ms; break; case K_TRACE: g.stats.msTraceClosest += ms; break; case K_SHADE: g.stats.msShade += ms; break;
        case K_SHADOW: g.stats.msTraceAny += ms; break; case K_RESOLVE: g.stats.msResolve += ms; break; case K_SORT: g.stats.msSort += ms; break;
      }
    }
    g.stats.closestRays = c.stats[0]; g.stats.shadowRays = c.stats[3]; g.stats.nodeVisits = c.stats[1]; g.stats.triVisits = c.stats[2];
    g.stats.kernelLaunches = launches; g.stats.paths = paths;
    g.stats.maxNodeVisitsPerRay = (uint32_t) c.stats[4]; g.stats.nonFiniteRays = c.stats[5];
    g.stats.closestHits = c.stats[6]; g.stats.shadowNodeVisits = c.stats[7]; g.stats.shadowTriVisits = c.stats[8];
    return true;
  };
  if (!body()) return GTLC_ERROR;
  s->sampleOffset += rs.spp;                                            // Gi.cpp:2515
  return GTLC_OK;
}

} // extern "C"
