// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Scene container, closest-hit / any-hit queries and hit-point reconstruction of the CPU oracle.
// The reference delegates R3/R8 to the Vulkan driver's acceleration structure
// (CgpuVk.cpp:2561-2854); what is pinned by the Vulkan spec and restated here: nearest triangle
// hit with tMin < t < tMax, no face culling (CgpuVk.cpp:2810), barycentrics (u, v) = weights of
// v1, v2.  Driver-defined and therefore fixed by this oracle (ORACLE_DEVIATIONS.md): watertight
// Woop/Benthin/Wald 2013 intersection in world space on pre-transformed vertices, and the tie
// break "equal t -> smaller (instanceIndex, primitiveId)".
#pragma once
#include "orc_bsdf.h"
#include "orc_texture.h"
#include <algorithm>
#include <vector>

namespace orc {

// Flattened instance record handed over by the test harness (mirrors rp_main.h:118-123 BlasPayload
// + CgpuBlasInstance, Gi.cpp:1188-1202).
struct InstanceDesc {
  uint32_t vertexBase;   // first GtlcVertex of the mesh in the global vertex array
  uint32_t faceBase;     // first face of the mesh in the global face array
  uint32_t faceCount;
  uint32_t flags;        // rp_main.h:115-116: bit0 FLIP_FACING, bit1 DOUBLE_SIDED
  uint32_t material;     // index into the material array
  int32_t  instanceId;   // InstanceIds[gl_InstanceID]
  int32_t  objectId;     // preamble.objectId
  uint32_t pad;
  float    objectToWorld[12]; // row-major 3x4 (VkTransformMatrixKHR), Gi.cpp:1191
};

// rp_main.h:58-64
struct FVertex { float f1[4]; float f2[4]; };

// rp_main.h:73-113 (48 bytes each)
struct SphereLight { float pos[3]; uint32_t diffuseSpecularPacked; float baseEmission[3]; float area; float radiusXYZ[3]; float padding; };
struct DistantLight { float direction[3]; float angle; float baseEmission[3]; uint32_t diffuseSpecularPacked; float padding[3]; float invPdf; };
struct RectLight { float origin[3]; float width; float baseEmission[3]; float height; uint32_t tangentFramePacked[2]; uint32_t diffuseSpecularPacked; float padding; };
struct DiskLight { float origin[3]; float radiusX; float baseEmission[3]; float radiusY; uint32_t tangentFramePacked[2]; uint32_t diffuseSpecularPacked; float padding; };

struct Hit { float t, u, v; uint32_t prim, inst; };

struct Tri { V3 v0, v1, v2; uint32_t prim, inst; float opacity; };

struct BNode { V3 lo, hi; uint32_t left, right; uint32_t first, count; }; // count > 0: leaf

// Per-ray shear of the watertight test as three sparse basis vectors (mx = e_kx - Sx e_kz, my = e_ky - Sy e_kz,
// mz = Sz e_kz), Sx = d[kx] * Sz, Sy = d[ky] * Sz, Sz = 1 / d[kz].
struct RayPre { V3 mx, my, mz; };

static inline float comp(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// Woop, Benthin, Wald: "Watertight Ray/Triangle Intersection", JCGT 2(1) 2013, section 3 (dimension choice and
// winding-preserving swap as in the paper; the shear constants use the reciprocal of the dominant component).
static inline RayPre ray_precompute(V3 d)
{
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  if (comp(d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
  const float Sz = 1.0f / comp(d, kz);
  const float Sx = comp(d, kx) * Sz;
  const float Sy = comp(d, ky) * Sz;
  RayPre r;
  r.mx = v3(kx == 0 ? 1.0f : (kz == 0 ? -Sx : 0.0f), kx == 1 ? 1.0f : (kz == 1 ? -Sx : 0.0f), kx == 2 ? 1.0f : (kz == 2 ? -Sx : 0.0f));
  r.my = v3(ky == 0 ? 1.0f : (kz == 0 ? -Sy : 0.0f), ky == 1 ? 1.0f : (kz == 1 ? -Sy : 0.0f), ky == 2 ? 1.0f : (kz == 2 ? -Sy : 0.0f));
  r.mz = v3(kz == 0 ? Sz : 0.0f, kz == 1 ? Sz : 0.0f, kz == 2 ? Sz : 0.0f);
  return r;
}

// Sheared coordinate of a translated vertex: an FMA chain over the sparse basis vector.  It depends only on the
// vertex and the ray, so shared vertices get identical values in every triangle (what watertightness needs).
static inline float shear(V3 a, V3 m) { return fmaf(a.z, m.z, fmaf(a.y, m.y, a.x * m.x)); }

// Returns true and (t, u, v) when the ray hits the triangle's plane inside its edges; the caller
// applies the (tMin, tMax) interval.  The edge functions U, V, W are NOT fused: antisymmetry
// fl(fl(ab) - fl(cd)) = -fl(fl(cd) - fl(ab)) across the two triangles of a shared edge is what makes the test watertight.
static inline bool tri_intersect(const RayPre& rp, V3 o, V3 p0, V3 p1, V3 p2, float& t, float& u, float& v)
{
  V3 A = p0 - o, B = p1 - o, C = p2 - o;
  float Ax = shear(A, rp.mx), Ay = shear(A, rp.my);
  float Bx = shear(B, rp.mx), By = shear(B, rp.my);
  float Cx = shear(C, rp.mx), Cy = shear(C, rp.my);
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    double CxBy = (double) Cx * (double) By, CyBx = (double) Cy * (double) Bx;
    U = (float)(CxBy - CyBx);
    double AxCy = (double) Ax * (double) Cy, AyCx = (double) Ay * (double) Cx;
    V = (float)(AxCy - AyCx);
    double BxAy = (double) Bx * (double) Ay, ByAx = (double) By * (double) Ax;
    W = (float)(BxAy - ByAx);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  float det = (U + V) + W;
  if (det == 0.0f) return false;
  float Az = shear(A, rp.mz), Bz = shear(B, rp.mz), Cz = shear(C, rp.mz);
  float T = fmaf(W, Cz, fmaf(V, Bz, U * Az));
  float rcp = 1.0f / det;
  t = T * rcp;
  u = V * rcp;
  v = W * rcp;
  return true;
}

struct Scene {
  std::vector<GtlcVertex> vertices;   // host vertices as passed to giCreateMesh (Gi.h:109-117)
  std::vector<FVertex> fverts;        // packed like Gi.cpp:845-859
  std::vector<uint32_t> faces;        // 3 per face, mesh-local
  std::vector<int32_t> faceIds;       // optional, one per face (rp_main.chit:232-240)
  std::vector<InstanceDesc> instances;
  std::vector<float> worldToObject;   // 12 per instance
  std::vector<Material> materials;
  std::vector<Texture> textures;                                   // R15
  std::vector<std::vector<GtlcTextureBinding>> materialTextures;   // per material; .texture = 1-based index into `textures`
  std::vector<int> materialOpacityBinding;                         // per material: index of the binding that feeds the cutout input, or -1
  std::vector<int> materialOpacityPrimvar;                         // per material: index of the primvar reader that feeds it (e.g. displayOpacity), or -1
  std::vector<std::vector<GtlcPrimvarBinding>> materialPrimvars;   // per material (names already resolved into primvarInfos)
  std::vector<uint32_t> primvarInfos; std::vector<float> primvarData;
  std::vector<SphereLight> sphereLights;
  std::vector<DistantLight> distantLights;
  std::vector<RectLight> rectLights;
  std::vector<DiskLight> diskLights;
  // dome light (lat-long RGBA float texels; empty = none)
  std::vector<float> domeTexels; uint32_t domeW = 0, domeH = 0;
  float domeRotation[4] = { 0, 0, 0, 1 }; float domeEmission[3] = { 1, 1, 1 };
  // acceleration structure
  std::vector<Tri> tris;
  std::vector<BNode> nodes;
};

static inline V3 xform_point(const float* m, V3 p)
{
  return v3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3],
            ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
            ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
static inline V3 xform_dir(const float* m, V3 p)
{
  return v3((m[0] * p.x + m[1] * p.y) + m[2] * p.z,
            (m[4] * p.x + m[5] * p.y) + m[6] * p.z,
            (m[8] * p.x + m[9] * p.y) + m[10] * p.z);
}
// v * mat3(M) in GLSL = transpose(M3x3) * v
static inline V3 xform_dir_transposed(const float* m, V3 p)
{
  return v3((m[0] * p.x + m[4] * p.y) + m[8] * p.z,
            (m[1] * p.x + m[5] * p.y) + m[9] * p.z,
            (m[2] * p.x + m[6] * p.y) + m[10] * p.z);
}

// Inverse of a row-major affine 3x4, computed in double and rounded once (same recipe as the product
// host code; the Vulkan driver supplies gl_WorldToObjectEXT in the reference).
static inline void invert_affine(const float* m, float* out)
{
  double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
  double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  double det = a * A + b * B + c * C;
  double id = 1.0 / det;
  double r[9] = { A * id, -(b * i - c * h) * id, (b * f - c * e) * id,
                  B * id, (a * i - c * g) * id, -(a * f - c * d) * id,
                  C * id, -(a * h - b * g) * id, (a * e - b * d) * id };
  double tx = m[3], ty = m[7], tz = m[11];
  for (int k = 0; k < 3; k++) {
    out[k * 4 + 0] = (float) r[k * 3 + 0]; out[k * 4 + 1] = (float) r[k * 3 + 1]; out[k * 4 + 2] = (float) r[k * 3 + 2];
    out[k * 4 + 3] = (float) (-(r[k * 3 + 0] * tx + r[k * 3 + 1] * ty + r[k * 3 + 2] * tz));
  }
}

// ---- binary BVH, median split on the largest centroid axis (the oracle's own structure; the
// closest hit is independent of the BVH shape thanks to the deterministic tie break) ------------
static void build_bvh_rec(Scene& s, std::vector<V3>& cen, std::vector<uint32_t>& idx, uint32_t node, uint32_t first, uint32_t count)
{
  V3 lo = v3(FLOAT_MAX), hi = v3(-FLOAT_MAX), clo = v3(FLOAT_MAX), chi = v3(-FLOAT_MAX);
  for (uint32_t i = first; i < first + count; i++) {
    const Tri& t = s.tris[idx[i]];
    const V3* p[3] = { &t.v0, &t.v1, &t.v2 };
    for (int k = 0; k < 3; k++) {
      lo = v3(fminf(lo.x, p[k]->x), fminf(lo.y, p[k]->y), fminf(lo.z, p[k]->z));
      hi = v3(fmaxf(hi.x, p[k]->x), fmaxf(hi.y, p[k]->y), fmaxf(hi.z, p[k]->z));
    }
    V3 c = cen[idx[i]];
    clo = v3(fminf(clo.x, c.x), fminf(clo.y, c.y), fminf(clo.z, c.z));
    chi = v3(fmaxf(chi.x, c.x), fmaxf(chi.y, c.y), fmaxf(chi.z, c.z));
  }
  s.nodes[node].lo = lo; s.nodes[node].hi = hi;
  V3 ext = chi - clo;
  int axis = (ext.x > ext.y) ? ((ext.x > ext.z) ? 0 : 2) : ((ext.y > ext.z) ? 1 : 2);
  if (count <= 4 || comp(ext, axis) <= 0.0f) {
    s.nodes[node].first = first; s.nodes[node].count = count; s.nodes[node].left = s.nodes[node].right = 0;
    return;
  }
  uint32_t mid = first + count / 2;
  std::nth_element(idx.begin() + first, idx.begin() + mid, idx.begin() + first + count,
                   [&](uint32_t a, uint32_t b) { float ca = comp(cen[a], axis), cb = comp(cen[b], axis); return ca < cb || (ca == cb && a < b); });
  uint32_t l = (uint32_t) s.nodes.size(); s.nodes.push_back(BNode()); s.nodes.push_back(BNode());
  s.nodes[node].left = l; s.nodes[node].right = l + 1; s.nodes[node].count = 0; s.nodes[node].first = 0;
  build_bvh_rec(s, cen, idx, l, first, mid - first);
  build_bvh_rec(s, cen, idx, l + 1, mid, first + count - mid);
}

static void scene_build(Scene& s)
{
  // pack FVertex (Gi.cpp:845-859)
  s.fverts.resize(s.vertices.size());
  for (size_t i = 0; i < s.vertices.size(); i++) {
    const GtlcVertex& c = s.vertices[i];
    uint32_t en = host_encode_direction(v3(c.norm[0], c.norm[1], c.norm[2]));
    uint32_t et = host_encode_direction(v3(c.tangent[0], c.tangent[1], c.tangent[2]));
    FVertex f;
    f.f1[0] = c.pos[0]; f.f1[1] = c.pos[1]; f.f1[2] = c.pos[2]; f.f1[3] = c.bitangentSign;
    f.f2[0] = uintBitsToFloat(en); f.f2[1] = uintBitsToFloat(et); f.f2[2] = c.u; f.f2[3] = c.v;
    s.fverts[i] = f;
  }
  s.worldToObject.resize(s.instances.size() * 12);
  s.tris.clear();
  for (size_t ii = 0; ii < s.instances.size(); ii++) {
    const InstanceDesc& in = s.instances[ii];
    invert_affine(in.objectToWorld, &s.worldToObject[ii * 12]);
    for (uint32_t f = 0; f < in.faceCount; f++) {
      const uint32_t* idx = &s.faces[(size_t)(in.faceBase + f) * 3];
      Tri t;
      const GtlcVertex& a = s.vertices[in.vertexBase + idx[0]];
      const GtlcVertex& b = s.vertices[in.vertexBase + idx[1]];
      const GtlcVertex& c = s.vertices[in.vertexBase + idx[2]];
      t.v0 = xform_point(in.objectToWorld, v3(a.pos[0], a.pos[1], a.pos[2]));
      t.v1 = xform_point(in.objectToWorld, v3(b.pos[0], b.pos[1], b.pos[2]));
      t.v2 = xform_point(in.objectToWorld, v3(c.pos[0], c.pos[1], c.pos[2]));
      t.prim = f; t.inst = (uint32_t) ii;
      const bool perHit = (!s.materialOpacityBinding.empty() && s.materialOpacityBinding[in.material] >= 0) || (!s.materialOpacityPrimvar.empty() && s.materialOpacityPrimvar[in.material] >= 0);
      t.opacity = perHit ? -1.0f : mat_cutout_opacity(s.materials[in.material]);   // < 0: per hit
      s.tris.push_back(t);
    }
  }
  s.nodes.clear();
  if (s.tris.empty()) return;
  std::vector<V3> cen(s.tris.size());
  std::vector<uint32_t> idx(s.tris.size());
  for (size_t i = 0; i < s.tris.size(); i++) { cen[i] = ((s.tris[i].v0 + s.tris[i].v1) + s.tris[i].v2) * (1.0f / 3.0f); idx[i] = (uint32_t) i; }
  s.nodes.reserve(s.tris.size());
  s.nodes.push_back(BNode());
  build_bvh_rec(s, cen, idx, 0, 0, (uint32_t) s.tris.size());
  std::vector<Tri> sorted(s.tris.size());
  for (size_t i = 0; i < idx.size(); i++) sorted[i] = s.tris[idx[i]];
  s.tris.swap(sorted);
}

static inline bool box_hit(const BNode& n, V3 o, V3 inv, float tmin, float tmax, float& tnear)
{
  float tx0 = (n.lo.x - o.x) * inv.x, tx1 = (n.hi.x - o.x) * inv.x;
  float ty0 = (n.lo.y - o.y) * inv.y, ty1 = (n.hi.y - o.y) * inv.y;
  float tz0 = (n.lo.z - o.z) * inv.z, tz1 = (n.hi.z - o.z) * inv.z;
  float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
  float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
  tf *= 1.0000004f; // Ize 2013, robust traversal padding
  tnear = tn;
  return tn <= tf;
}

// Stochastic cutout, rp_main.ahit:35-61 (`if (k > opacity) ignoreIntersectionEXT`).  The reference draws k from
// the path RNG once per any-hit invocation in an order Vulkan leaves undefined; the oracle (and the kernel) decide
// with a pure function of (ray seed, primitive, instance) instead -- ORACLE_DEVIATIONS.md #10.
// mdl_cutout_opacity at a candidate hit of a textured cutout (rp_main.ahit:35-42): uv and object-space position as
// setup_mdl_shading_state computes them (mdl_shading_state.glsl:16-23, 61-64), then the material's opacity binding.
static inline float cutout_opacity_at(const struct Scene& s, uint32_t prim, uint32_t inst, float hu, float hv);

static inline bool cutout_ignored(uint32_t seed, float opacity, uint32_t prim, uint32_t inst)
{
  if (opacity >= 1.0f) return false;
  uint32_t st = seed ^ hash_theironborn(prim ^ (inst * 0x9E3779B1u));
  const float xi = uintAsFloat(hash_pcg32(st));
  return xi > opacity;
}

static inline bool hit_better(float t, uint32_t inst, uint32_t prim, const Hit& best)
{
  if (best.prim == 0xFFFFFFFFu) return true;
  if (t < best.t) return true;
  if (t > best.t) return false;
  return (inst < best.inst) || (inst == best.inst && prim < best.prim);
}

// R3: closest hit in (tMin, tMax).  anyHit: R8 (terminate on first hit).
static inline bool scene_trace(const Scene& s, V3 o, V3 d, float tMin, float tMax, bool anyHit, Hit& best, uint32_t seed = 0)
{
  best.t = -1.0f; best.u = best.v = 0.0f; best.prim = best.inst = 0xFFFFFFFFu;
  if (s.nodes.empty() || !(tMax > tMin)) return false;
  RayPre rp = ray_precompute(d);
  V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  uint32_t stack[128]; int sp = 0;
  stack[sp++] = 0;
  float tlimit = tMax;
  while (sp > 0) {
    const BNode& n = s.nodes[stack[--sp]];
    float tn;
    if (!box_hit(n, o, inv, tMin, tlimit, tn)) continue;
    if (n.count > 0) {
      for (uint32_t i = n.first; i < n.first + n.count; i++) {
        const Tri& tr = s.tris[i];
        float t, u, v;
        if (!tri_intersect(rp, o, tr.v0, tr.v1, tr.v2, t, u, v)) continue;
        if (!(t > tMin && t < tMax)) continue;
        float opacity = tr.opacity;
        if (opacity < 0.0f) opacity = cutout_opacity_at(s, tr.prim, tr.inst, u, v);
        if (cutout_ignored(seed, opacity, tr.prim, tr.inst)) continue;
        if (anyHit) { best.t = t; best.u = u; best.v = v; best.prim = tr.prim; best.inst = tr.inst; return true; }
        if (hit_better(t, tr.inst, tr.prim, best)) { best.t = t; best.u = u; best.v = v; best.prim = tr.prim; best.inst = tr.inst; tlimit = t; }
      }
    } else {
      if (sp + 2 > 128) continue;
      stack[sp++] = n.left; stack[sp++] = n.right;
    }
  }
  return best.prim != 0xFFFFFFFFu;
}

// R4: setup_mdl_shading_state, mdl_shading_state.glsl:4-98
static inline void setup_shading_state(const Scene& s, const Hit& h, V3 rayDir, State& st, bool& isFrontFace)
{
  const InstanceDesc& in = s.instances[h.inst];
  const float* o2w = in.objectToWorld;
  const float* w2o = &s.worldToObject[(size_t) h.inst * 12];
  const uint32_t* f = &s.faces[(size_t)(in.faceBase + h.prim) * 3];                // :10
  const FVertex& v0 = s.fverts[in.vertexBase + f[0]];                              // :12-14
  const FVertex& v1 = s.fverts[in.vertexBase + f[1]];
  const FVertex& v2 = s.fverts[in.vertexBase + f[2]];
  V3 bc = v3((1.0f - h.u) - h.v, h.u, h.v);                                        // :16
  V3 p0 = v3(v0.f1[0], v0.f1[1], v0.f1[2]), p1 = v3(v1.f1[0], v1.f1[1], v1.f1[2]), p2 = v3(v2.f1[0], v2.f1[1], v2.f1[2]);
  V3 localPos = (p0 * bc.x + p1 * bc.y) + p2 * bc.z;                               // :23
  V3 pos = xform_point(o2w, localPos);                                             // :24
  V3 geomNormal = normalize(cross(p1 - p0, p2 - p0));                              // :26
  geomNormal = normalize(xform_dir_transposed(w2o, geomNormal));                   // :27
  V3 n0 = decode_direction(floatBitsToUint(v0.f2[0]));                             // :30-32
  V3 n1 = decode_direction(floatBitsToUint(v1.f2[0]));
  V3 n2 = decode_direction(floatBitsToUint(v2.f2[0]));
  V3 localNormal = normalize((n0 * bc.x + n1 * bc.y) + n2 * bc.z);                 // :34
  V3 normal = normalize(xform_dir_transposed(w2o, localNormal));                   // :35
  isFrontFace = dot(geomNormal, -rayDir) >= 0.0f;                                  // :38
  if (!isFrontFace) { geomNormal = -geomNormal; normal = -normal; }                // :40-44
  V3 t0 = decode_direction(floatBitsToUint(v0.f2[1]));                             // :47-49
  V3 t1 = decode_direction(floatBitsToUint(v1.f2[1]));
  V3 t2 = decode_direction(floatBitsToUint(v2.f2[1]));
  V3 localTangent = normalize((t0 * bc.x + t1 * bc.y) + t2 * bc.z);                // :51
  V3 tangent = normalize(xform_dir(o2w, localTangent));                            // :52
  tangent = normalize(tangent - normal * dot(tangent, normal));                    // :55
  float bitangentSign = (bc.x * v0.f1[3] + bc.y * v1.f1[3]) + bc.z * v2.f1[3];     // :57
  V3 bitangent = cross(normal, tangent) * bitangentSign;                           // :58
  float uu = (bc.x * v0.f2[2] + bc.y * v1.f2[2]) + bc.z * v2.f2[2];                // :61-64
  float vv = (bc.x * v0.f2[3] + bc.y * v1.f2[3]) + bc.z * v2.f2[3];
  st.normal = normal; st.geom_normal = geomNormal; st.position = pos;              // :82-90
  st.tangent_u = tangent; st.tangent_v = bitangent; st.text_coords = v3(uu, vv, 0.0f);
  st.local_position = localPos;
}

static inline float cutout_opacity_at(const Scene& s, uint32_t prim, uint32_t inst, float hu, float hv)
{
  const InstanceDesc& in = s.instances[inst];
  const int bi = s.materialOpacityBinding[in.material];
  const uint32_t* f = &s.faces[(size_t) (in.faceBase + prim) * 3];
  if (bi < 0) {                                   // primvar-driven: scene_data_lookup_float / _int (mdl_interface.glsl:281-301, 417-479) at this hit
    const int k = s.materialOpacityPrimvar[in.material];
    const GtlcPrimvarBinding& pb = s.materialPrimvars[in.material][(size_t) k];
    const V3 bc = v3((1.0f - hu) - hv, hu, hv);
    float o = pb.fallback[0];
    if (!s.primvarInfos.empty()) {
      const uint32_t* info = &s.primvarInfos[((size_t) inst * GTLC_MAX_PRIMVAR_BINDINGS + (size_t) k) * 2];
      if (info[1] & 0x100u) {
        const uint32_t stride = info[1] & 15u, interpolation = (info[1] >> 4) & 15u;
        uint32_t idx[3] = { f[0], f[1], f[2] };
        if (interpolation == GTLC_INTERP_UNIFORM) idx[0] = idx[1] = idx[2] = prim;
        else if (interpolation == GTLC_INTERP_INSTANCE) idx[0] = idx[1] = idx[2] = in.instanceId > 0 ? (uint32_t) in.instanceId : 0u;
        else if (interpolation == GTLC_INTERP_CONSTANT) idx[0] = idx[1] = idx[2] = 0u;
        const float* dta = &s.primvarData[info[0]];
        if (info[1] & 0x200u) {
          const uint32_t pick = bc.x > bc.y ? (bc.x > bc.z ? idx[0] : idx[2]) : (bc.y > bc.z ? idx[1] : idx[2]);
          o = (float) (int32_t) floatBitsToUint(dta[(size_t) pick * stride]);
        } else o = (dta[(size_t) idx[0] * stride] * bc.x + dta[(size_t) idx[1] * stride] * bc.y) + dta[(size_t) idx[2] * stride] * bc.z;
      }
    }
    return o < 0.0f ? 0.0f : (o > 1.0f ? 1.0f : o);
  }
  const FVertex& v0 = s.fverts[in.vertexBase + f[0]];
  const FVertex& v1 = s.fverts[in.vertexBase + f[1]];
  const FVertex& v2 = s.fverts[in.vertexBase + f[2]];
  const V3 bc = v3((1.0f - hu) - hv, hu, hv);
  const V3 p0 = v3(v0.f1[0], v0.f1[1], v0.f1[2]), p1 = v3(v1.f1[0], v1.f1[1], v1.f1[2]), p2 = v3(v2.f1[0], v2.f1[1], v2.f1[2]);
  const V3 localPos = (p0 * bc.x + p1 * bc.y) + p2 * bc.z;
  const float uu = (bc.x * v0.f2[2] + bc.y * v1.f2[2]) + bc.z * v2.f2[2];
  const float vv = (bc.x * v0.f2[3] + bc.y * v1.f2[3]) + bc.z * v2.f2[3];
  const GtlcTextureBinding& b = s.materialTextures[in.material][(size_t) bi];
  float ch[4];
  eval_binding(s.textures, b, uu, vv, localPos, ch);
  const float o = ch[b.channels[0]];
  return o < 0.0f ? 0.0f : (o > 1.0f ? 1.0f : o);
}

} // namespace orc
