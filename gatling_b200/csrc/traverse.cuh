// BVH8 traversal kernels: trace_closest (R3) and trace_any / shadow (R8) of SURVEY.md 8a.
//
// Replaces traceRayEXT(sceneAS, ...) of /root/reference/src/gi/shaders/rp_main.rgen:381-393 (closest,
// flags 0, no culling -- CgpuVk.cpp:2810) and :412-424 (shadow, TerminateOnFirstHit|SkipClosestHit);
// the rp_main_shadow.miss "shadowed = false" + "radiance += neeContrib" (rgen:426-429) pair is fused
// into the shadow kernel's epilogue.
//
// Execution model: persistent warps (grid = k x 148 CTAs) pull rays from a global queue with one
// warp-aggregated atomic per refill ("work stealing"); a lane whose ray finished is refilled as soon
// as >= kRefillThreshold lanes of its warp are idle, so divergence in path length does not idle
// the warp (Aila & Laine 2009 "persistent while-while", dynamic fetch).  One ray per lane, per-lane
// traversal stack of 8-byte node-group entries (Ylitie et al. 2017), watertight Woop/Benthin/Wald
// 2013 triangle test in world space.  All node / triangle / ray traffic is 128-bit loads.
//
// Determinism: closest hit = smallest t, ties broken by (instanceIndex, primitiveId); the result does
// not depend on traversal order.
#pragma once
#include "device_math.cuh"
#include "textures.cuh"

namespace gtlc {

#ifndef GTLC_TRACE_MIN_BLOCKS
#define GTLC_TRACE_MIN_BLOCKS 7   // 72 registers: 7 CTAs = 28 warps per SM (measured best of 6/7/8 on C2)
#endif
constexpr int kTraceBlock = 128;         // threads per CTA
constexpr int kStackSize = 48;           // >= wide-tree depth (one deferred node group per level; = kBvh8MaxDepth, checked on the host)
constexpr int kRefillThreshold = 8;      // idle lanes that trigger a refill
constexpr float kBoxPad = 1.0000004f;    // Ize 2013 robust slab padding

enum { TRACE_CLOSEST = 0, TRACE_ANY = 1, TRACE_SHADOW = 2 };

// ---- traversal layout of the BVH8 in HBM --------------------------------------------------------------
// The builders (host SAH, device LBVH / PLOC) and the oracle speak the canonical format of bvh8.h: 80-byte nodes,
// 48-byte triangles.  bvh8_repack_*_kernel re-packs both once per build into what the traversal kernels read:
//
//   node, 96 bytes = 3 x 32-byte sectors, fetched with three 256-bit loads (ld.global.nc.v8 = LDG.E.ENL2.256):
//     words  0..7 : p.x p.y p.z  valid32 | 2^ex 2^ey 2^ez (as floats)  childBase
//     words  8..15: triBase  -  | qlo_x[8] qlo_y[8] qlo_z[8]
//     words 16..23: qhi_x[8] qhi_y[8] qhi_z[8] | -  -
//   valid32 = imask << 24 | triValid24, bit 3 s + k of triValid24 set iff slot s is a leaf with more than k triangles.
//   A hit child sets its whole field (bit 24 + s and bits 3 s .. 3 s + 2) with ONE predicated OR of a compile-time
//   constant; `& valid32` then keeps what exists.  Triangle `b` of the node is triBase + popc(triValid24 below b) --
//   the canonical builders number a node's triangles in slot order, so this is the canonical index.  Against the
//   meta-byte decode of the 80-byte node (a variable shift per child) the node step shrinks from ~264 to ~215 SASS
//   instructions; five 128-bit loads become three 256-bit ones (the L1 data pipe counts wavefronts per load instruction).
//
//   triangle, 64 bytes = 2 sectors, never straddling a 128-byte line (a 48-byte record touches 2.5 sectors on average):
//     (v0.xyz, primitiveId) (v1.xyz, instanceIndex) (v2.xyz, cutout opacity) (rank of (instance, primitive), -, -, -)
//
// The algorithmic bytes of SURVEY.md 8d stay 80 per node visit and 48 per triangle test (the padding is not "useful").
constexpr int kDevNodeWords = 24;
constexpr int kDevTriWords = 16;

// Per-instance record (BlasPayload, rp_main.h:118-123, + CgpuBlasInstance hit-group index)
struct InstanceInfo { uint32_t vertexBase, faceBase, flags, material; };
constexpr int kShadeVertexWords = 16, kShadeFaceWords = 8;   // shading records, see wavefront.cuh (DeviceScene)

// What the any-hit test of a TEXTURED cutout needs (rp_main.ahit:35-61 runs mdl_cutout_opacity on a full shading state;
// the opacity input only ever depends on the texture coordinate / object-space position of the candidate hit).
struct CutoutScene {
  const uint32_t* __restrict__ sverts;              // kShadeVertexWords per vertex
  const uint32_t* __restrict__ sfaces;              // kShadeFaceWords per face
  const InstanceInfo* __restrict__ instInfo;
  const DeviceTexBinding* __restrict__ texBindings;
  const DeviceTexture* __restrict__ textures;
  const uint32_t* __restrict__ matOpacityBinding;   // per material: index into texBindings of the binding that feeds the cutout input, or 0xFFFFFFFF
  const float* __restrict__ matOpacityConst;        // per material: the constant cutout opacity (what a missing binding falls back to)
  // ... or a primvar reader (UsdPreviewSurface.opacity <- displayOpacity is the common case): scene_data_lookup_float / _int at the hit
  const uint2* __restrict__ matOpacityPrimvar;      // per material: (index into primvarFallback, column k of the mesh's scene-data row) or (0xFFFFFFFF, -)
  const float* __restrict__ primvarFallback;        // fallback[0] of every primvar binding
  const uint2* __restrict__ primvarInfos;           // as DeviceScene.primvarInfos
  const float* __restrict__ primvarData;
  const int2* __restrict__ instIds;
};

struct TraceParams {
  const uint32_t* __restrict__ nodes;   // kDevNodeWords per BVH8 node (traversal layout)
  const uint32_t* __restrict__ tris;    // kDevTriWords per triangle (traversal layout)
  const float4* __restrict__ rayO;      // (origin.xyz, tMin)
  const float4* __restrict__ rayD;      // (direction.xyz, tMax)
  float4* __restrict__ hitTUVP;         // (t, u, v, bits primitiveId); miss: t = -1
  uint32_t* __restrict__ hitInst;       // instance index; miss: 0xFFFFFFFF
  const uint32_t* __restrict__ numRays; // device-resident queue length
  uint32_t* workCounter;                // zero at launch (reset by the kernel that ran before: raygen / shade)
  // TRACE_SHADOW epilogue: unoccluded -> target.xyz += contrib
  const float4* __restrict__ shContrib; // (contrib.xyz, bits target): target bit31 = sample slot, else path queue position
  float4* __restrict__ radPath;         // (radiance.xyz, rng) of the continuing paths
  float4* __restrict__ radSample;       // (radiance.xyz, -) of finished samples
  // bookkeeping
  unsigned long long* stats;            // COUNT only: [1] node visits, [2] triangle tests of closest / any rays, [7], [8] the same of shadow rays, [4] max node visits of one ray, [6] closest hits; always: [5] non-finite rays
  uint32_t* visitsPerRay;               // optional 2 x uint32 per ray
  int refillThreshold;                  // idle lanes that trigger a refill (kRefillThreshold by default)
  int triThreshold;                     // unused (was variant 5)
  const float4* __restrict__ seedRad;   // cutout any-hit: per-ray RNG seed = bits of seedRad[i].w (path queue `rad`), or
  const uint32_t* __restrict__ seedU32; //   seedU32[i] (shadow queue); both NULL = seed 0
  uint32_t biasWord;                           // 0x47000000 (see biased_byte)
  float4* aovFlags;                            // NEE debug AOV (first bounce, per-lane kernel only): per-sample capture word a[0]; bit 3 of .w = shadow ray occluded
  const uint32_t* __restrict__ aovSlotOf;      //   sample slot of a continuing path (path queue `slot`)
  // Launch prologue (thread 0 of CTA 0), the wave loop's per-bounce bookkeeping folded into the kernel that needs it --
  // nothing in this kernel reads what it writes:  *zeroA = 0, *zeroB = 0, *rayStat += *numRays.
  uint32_t* zeroA; uint32_t* zeroB; unsigned long long* rayStat;
  float* clk;                                   // ClockCycles AOV (per-lane kernel only): cycles a ray spent here are added to its sample's counter
  const CutoutScene* __restrict__ cut;          // device-resident; read by textured cutouts only (triangle opacity word < 0)
};

GTLC_D void trace_prologue(const TraceParams& P)
{
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (P.zeroA) *P.zeroA = 0u;
    if (P.zeroB) *P.zeroB = 0u;
    if (P.rayStat) *P.rayStat += (unsigned long long) *P.numRays;
  }
}

// 256-bit read-only load (sm_100: one LDG.E.ENL2.256.CONSTANT); `p` must be 32-byte aligned.
GTLC_D void ldg256(const uint32_t* p, uint32_t (&v)[8])
{
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}

// Quantised byte q (0..255) taken from byte `j` of `w`, as the float 32768 + q: one PRMT builds
// 0x4700qq00 (exponent 2^15, q in mantissa bits 8..15), no integer->float conversion pipe involved.
// The bias is folded into the slab origin (o - 32768 s), so a bound costs PRMT + FFMA.
// `k47` = 0x47000000 read from the kernel parameters, NOT a literal: PRMT encodes one immediate only, and with a
// literal ptxas keeps the constant as that immediate and spends a MOV per PRMT on the selector (44 of the ~300
// instructions of a node step); with the constant in a register the selector becomes the immediate.
GTLC_D float biased_byte(uint32_t w, uint32_t k47, uint32_t sel) { return u2f(__byte_perm(w, k47, sel)); }

// A ray with a NaN/Inf origin or direction can never hit a triangle (every comparison of the watertight
// test fails) but would pass every NaN-dropping slab test and walk the whole tree: answer it as a miss.
GTLC_D bool ray_is_finite(float4 ro, float4 rd)
{
  const float m = fmaxf(fmaxf(fabsf(ro.x), fabsf(ro.y)), fmaxf(fabsf(ro.z), fmaxf(fmaxf(fabsf(rd.x), fabsf(rd.y)), fabsf(rd.z))));
  const float sum = ((ro.x + ro.y) + ro.z) + ((rd.x + rd.y) + rd.z);   // NaN propagates through +, fmaxf drops it
  return (m <= kFloatMax) && (sum == sum);
}

struct RaySetup {
  V3 o;
  float tmin;
  V3 idir;
  uint32_t octinv;          // 7 - octant; bit k clear <=> direction component k is negative
  uint32_t pm1, pm2, pm4;   // delta-swap masks (top byte) that move hit bit 24 + s to 24 + (s ^ octinv)
  // Woop shear as three sparse basis vectors: mx = e_kx - Sx e_kz, my = e_ky - Sy e_kz, mz = Sz e_kz, so that
  // the per-triangle code needs no data-dependent component selection (same definition in the CPU oracle).
  V3 mx, my, mz;
};

GTLC_D void ray_setup(RaySetup& r, float4 ro, float4 rd)
{
  const V3 d = v3(rd.x, rd.y, rd.z);
  r.o = v3(ro.x, ro.y, ro.z); r.tmin = ro.w;
  // |d| is clamped to 2^-80 for the slab tests only (Ylitie et al. 2017): with an exactly-zero or denormal
  // component 1/d = inf turns the quantised slab arithmetic q*s + o into inf - inf = NaN, the NaN-dropping
  // min/max then never cull along that axis and the ray walks a whole slice of the tree.
  const float kMinDir = 8.271806125530277e-25f;   // 2^-80
  r.idir = v3(1.0f / (fabsf(d.x) > kMinDir ? d.x : copysignf(kMinDir, d.x)),
              1.0f / (fabsf(d.y) > kMinDir ? d.y : copysignf(kMinDir, d.y)),
              1.0f / (fabsf(d.z) > kMinDir ? d.z : copysignf(kMinDir, d.z)));
  uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
  r.octinv = 7u - oct;
  r.pm1 = (r.octinv & 1u) ? 0x55000000u : 0u;
  r.pm2 = (r.octinv & 2u) ? 0x33000000u : 0u;
  r.pm4 = (r.octinv & 4u) ? 0x0F000000u : 0u;
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  if (comp(d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
  const float Sz = comp(r.idir, kz);          // = 1 / d[kz] exactly (the dominant component is never clamped): no extra division
  const float Sx = comp(d, kx) * Sz;
  const float Sy = comp(d, ky) * Sz;
  r.mx = v3(kx == 0 ? 1.0f : (kz == 0 ? -Sx : 0.0f), kx == 1 ? 1.0f : (kz == 1 ? -Sx : 0.0f), kx == 2 ? 1.0f : (kz == 2 ? -Sx : 0.0f));
  r.my = v3(ky == 0 ? 1.0f : (kz == 0 ? -Sy : 0.0f), ky == 1 ? 1.0f : (kz == 1 ? -Sy : 0.0f), ky == 2 ? 1.0f : (kz == 2 ? -Sy : 0.0f));
  r.mz = v3(kz == 0 ? Sz : 0.0f, kz == 1 ? Sz : 0.0f, kz == 2 ? Sz : 0.0f);
}

// Stochastic cutout (rp_main.ahit:35-61: `if (rand > mdl_cutout_opacity) ignoreIntersectionEXT`).  The reference
// draws from the path's RNG once per any-hit invocation, in an order the Vulkan spec leaves undefined; here the
// decision is a pure function of (ray seed, primitive, instance), so it does not depend on traversal order:
// xi = uintAsFloat(hash_pcg32(seed ^ hash_theironborn(prim ^ inst * 0x9E3779B1))).  `opacity` rides in the spare
// .w of the triangle record (1.0 = opaque, the common case costs one compare).
// mdl_cutout_opacity at a candidate hit of a textured-cutout triangle: the hit's uv / object-space position as
// setup_shading_state computes them (mdl_shading_state.glsl:16-23, 61-64; same expressions as wavefront.cuh), then the
// material's opacity binding, clamped like mat_cutout_opacity.  Out of line: the common any-hit test is one compare.
__device__ __noinline__ float cutout_opacity_at(const CutoutScene* __restrict__ cs, uint32_t prim, uint32_t inst, float hu, float hv)
{
  const CutoutScene c = *cs;
  const InstanceInfo in = c.instInfo[inst];
  const uint32_t bi = c.matOpacityBinding[in.material];
  const uint2 pv = c.matOpacityPrimvar[in.material];
  if (bi == 0xFFFFFFFFu && pv.x == 0xFFFFFFFFu) return c.matOpacityConst[in.material];
  const uint4 fr = __ldg((const uint4*) (c.sfaces + (size_t) (in.faceBase + prim) * kShadeFaceWords));
  const uint32_t* vb = c.sverts + (size_t) in.vertexBase * kShadeVertexWords;
  // words 0..2 position, 7 u, 11 v of each shading vertex
  const uint4 a0 = __ldg((const uint4*) (vb + (size_t) fr.x * kShadeVertexWords)), a1 = __ldg((const uint4*) (vb + (size_t) fr.x * kShadeVertexWords + 4)), a2 = __ldg((const uint4*) (vb + (size_t) fr.x * kShadeVertexWords + 8));
  const uint4 b0 = __ldg((const uint4*) (vb + (size_t) fr.y * kShadeVertexWords)), b1 = __ldg((const uint4*) (vb + (size_t) fr.y * kShadeVertexWords + 4)), b2 = __ldg((const uint4*) (vb + (size_t) fr.y * kShadeVertexWords + 8));
  const uint4 c0 = __ldg((const uint4*) (vb + (size_t) fr.z * kShadeVertexWords)), c1 = __ldg((const uint4*) (vb + (size_t) fr.z * kShadeVertexWords + 4)), c2 = __ldg((const uint4*) (vb + (size_t) fr.z * kShadeVertexWords + 8));
  const V3 bc = v3((1.0f - hu) - hv, hu, hv);
  if (bi == 0xFFFFFFFFu) {                          // primvar-driven (a texture binding on the same input is applied later and wins, as in shading)
    const uint2 info = c.primvarInfos[(size_t) (in.flags >> 8) * GTLC_MAX_PRIMVAR_BINDINGS + pv.y];
    float o;
    if (!(info.y & 0x100u)) o = c.primvarFallback[pv.x];                               // scene_data_isvalid: default_value
    else {
      const uint32_t stride = info.y & 15u, interpolation = (info.y >> 4) & 15u;
      uint32_t i0 = fr.x, i1 = fr.y, i2 = fr.z;                                        // get_scene_data_indices, mdl_interface.glsl:281-301
      if (interpolation == GTLC_INTERP_UNIFORM) i0 = i1 = i2 = prim;
      else if (interpolation == GTLC_INTERP_INSTANCE) { const int id = c.instIds[inst].y; i0 = i1 = i2 = id > 0 ? (uint32_t) id : 0u; }
      else if (interpolation == GTLC_INTERP_CONSTANT) i0 = i1 = i2 = 0u;
      const float* d = c.primvarData + info.x;
      if (info.y & 0x200u) {                                                           // scene_data_lookup_int: nearest vertex
        const uint32_t pick = bc.x > bc.y ? (bc.x > bc.z ? i0 : i2) : (bc.y > bc.z ? i1 : i2);
        o = (float) (int32_t) f2u(__ldg(d + (size_t) pick * stride));
      } else {
        const float v0 = __ldg(d + (size_t) i0 * stride), v1 = __ldg(d + (size_t) i1 * stride), v2 = __ldg(d + (size_t) i2 * stride);
        o = (v0 * bc.x + v1 * bc.y) + v2 * bc.z;
      }
    }
    return o < 0.0f ? 0.0f : (o > 1.0f ? 1.0f : o);
  }
  const V3 p0 = v3(u2f(a0.x), u2f(a0.y), u2f(a0.z)), p1 = v3(u2f(b0.x), u2f(b0.y), u2f(b0.z)), p2 = v3(u2f(c0.x), u2f(c0.y), u2f(c0.z));
  const V3 localPos = (p0 * bc.x + p1 * bc.y) + p2 * bc.z;
  const float stu = (bc.x * u2f(a1.w) + bc.y * u2f(b1.w)) + bc.z * u2f(c1.w);
  const float stv = (bc.x * u2f(a2.w) + bc.y * u2f(b2.w)) + bc.z * u2f(c2.w);
  const DeviceTexBinding& tb = c.texBindings[bi];
  float ch[4];
  eval_tex_binding(tb, c.textures, stu, stv, localPos, ch);
  const float o = ch[(tb.packed >> 4) & 3u];
  return o < 0.0f ? 0.0f : (o > 1.0f ? 1.0f : o);
}

// TEXCUT: the scene has textured cutouts (host dispatch on TraceParams.cut); the kernels of every other scene carry no call.
template <bool TEXCUT>
GTLC_D bool cutout_ignored(const TraceParams& P, uint32_t rayIdx, float opacity, uint32_t prim, uint32_t inst, float hu, float hv)
{
  if (opacity >= 1.0f) return false;
  if (TEXCUT && opacity < 0.0f) {                           // textured cutout: evaluate the material's opacity input at this hit
    opacity = cutout_opacity_at(P.cut, prim, inst, hu, hv);
    if (opacity >= 1.0f) return false;
  }
  uint32_t seed = 0u;
  if (P.seedRad) seed = f2u(__ldg(&P.seedRad[rayIdx]).w);
  else if (P.seedU32) seed = __ldg(&P.seedU32[rayIdx]);
  uint32_t st = seed ^ hash_theironborn(prim ^ (inst * 0x9E3779B1u));
  const float xi = uintAsFloat(hash_pcg32(st));
  return xi > opacity;
}

// Sheared coordinate of a translated vertex: FMA chain over the sparse basis vector (3 instructions).  Per-vertex
// quantities may be fused -- they depend only on (vertex, ray), so shared vertices agree across triangles; the edge
// functions U, V, W below must NOT be fused (their exact antisymmetry across a shared edge is the watertightness).
GTLC_D float shear(V3 a, V3 m) { return fmaf(a.z, m.z, fmaf(a.y, m.y, a.x * m.x)); }

// Rare path of the watertight test: an edge function is exactly zero -> recompute all three in double.
__device__ __noinline__ void tri_edges_double(float Ax, float Ay, float Bx, float By, float Cx, float Cy, float& U, float& V, float& W)
{
  double CxBy = (double) Cx * (double) By, CyBx = (double) Cy * (double) Bx;
  U = (float) (CxBy - CyBx);
  double AxCy = (double) Ax * (double) Cy, AyCx = (double) Ay * (double) Cx;
  V = (float) (AxCy - AyCx);
  double BxAy = (double) Bx * (double) Ay, ByAx = (double) By * (double) Ax;
  W = (float) (BxAy - ByAx);
}

// Woop, Benthin, Wald 2013.  Returns true with (t, u, v) when the ray line pierces the triangle.
GTLC_D bool tri_intersect(const RaySetup& r, V3 p0, V3 p1, V3 p2, float& t, float& u, float& v)
{
  const V3 A = p0 - r.o, B = p1 - r.o, C = p2 - r.o;
  const float Ax = shear(A, r.mx), Ay = shear(A, r.my);
  const float Bx = shear(B, r.mx), By = shear(B, r.my);
  const float Cx = shear(C, r.mx), Cy = shear(C, r.my);
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) tri_edges_double(Ax, Ay, Bx, By, Cx, Cy, U, V, W);
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = (U + V) + W;
  if (det == 0.0f) return false;
  const float Az = shear(A, r.mz), Bz = shear(B, r.mz), Cz = shear(C, r.mz);
  const float T = fmaf(W, Cz, fmaf(V, Bz, U * Az));
  const float rcp = 1.0f / det;
  t = T * rcp;
  u = V * rcp;
  v = W * rcp;
  return true;
}

// One BVH8 node in registers (three 256-bit loads) and its slab test.
struct NodeRegs { uint32_t a[8], b[8], c[8]; };

GTLC_D void node_fetch(const uint32_t* __restrict__ nodes, uint32_t node, NodeRegs& n)
{
  const uint32_t* p = nodes + (size_t) node * kDevNodeWords;
  ldg256(p, n.a); ldg256(p + 8, n.b); ldg256(p + 16, n.c);
}

// Slab-tests the 8 quantised child boxes of one node; returns the 32-bit hit mask
// (bits 31..24: hit internal children in traversal priority order, bit 24 + (slot ^ octinv); bits 23..0: the
// triangle fields 3 slot + k of hit leaf children).
// t = (32768 + q) * s + (o - 32768 s) with s = 2^e / d, o = (p - origin) / d; branch-free.  Same arithmetic, in the
// same order, as the oracle's walk of the canonical node (oracle/orc_bvh8.h: bvh8_node_hitmask).
GTLC_D uint32_t node_hitmask(const NodeRegs& n, const RaySetup& r, float tbest, uint32_t k47)
{
  const float sx = u2f(n.a[4]) * r.idir.x;
  const float sy = u2f(n.a[5]) * r.idir.y;
  const float sz = u2f(n.a[6]) * r.idir.z;
  const float ox = (u2f(n.a[0]) - r.o.x) * r.idir.x - 32768.0f * sx;
  const float oy = (u2f(n.a[1]) - r.o.y) * r.idir.y - 32768.0f * sy;
  const float oz = (u2f(n.a[2]) - r.o.z) * r.idir.z - 32768.0f * sz;
  const bool nx = (r.octinv & 1u) == 0u, ny = (r.octinv & 2u) == 0u, nz = (r.octinv & 4u) == 0u;   // direction component negative
  uint32_t hm = 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const uint32_t qlox = n.b[2 + half], qloy = n.b[4 + half], qloz = n.b[6 + half];
    const uint32_t qhix = n.c[0 + half], qhiy = n.c[2 + half], qhiz = n.c[4 + half];
    const uint32_t nearx = nx ? qhix : qlox, farx = nx ? qlox : qhix;
    const uint32_t neary = ny ? qhiy : qloy, fary = ny ? qloy : qhiy;
    const uint32_t nearz = nz ? qhiz : qloz, farz = nz ? qloz : qhiz;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t sel = 0x7604u | ((uint32_t) j << 4);
      const float tnx = fmaf(biased_byte(nearx, k47, sel), sx, ox);
      const float tny = fmaf(biased_byte(neary, k47, sel), sy, oy);
      const float tnz = fmaf(biased_byte(nearz, k47, sel), sz, oz);
      const float tfx = fmaf(biased_byte(farx, k47, sel), sx, ox);
      const float tfy = fmaf(biased_byte(fary, k47, sel), sy, oy);
      const float tfz = fmaf(biased_byte(farz, k47, sel), sz, oz);
      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
      const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tbest)) * kBoxPad;
      const int s = 4 * half + j;
      if (tn <= tf) hm |= (1u << (24 + s)) | (7u << (3 * s));      // one predicated OR of an immediate
    }
  }
  hm &= n.a[3];
  // internal hits: slot order -> traversal priority order, bit 24 + s -> 24 + (s ^ octinv): three delta swaps whose masks
  // are zero for the octant bits that are clear (the triangle bits below bit 24 never pass a mask)
  uint32_t t;
  t = (hm ^ (hm >> 1)) & r.pm1; hm ^= t ^ (t << 1);
  t = (hm ^ (hm >> 2)) & r.pm2; hm ^= t ^ (t << 2);
  t = (hm ^ (hm >> 4)) & r.pm4; hm ^= t ^ (t << 4);
  return hm;
}

// Index of triangle bit `b` (0..23) of a node: the canonical builders store a node's triangles in slot order.
GTLC_D uint32_t tri_index(uint32_t triBase, uint32_t valid32, uint32_t b) { return triBase + (uint32_t) __popc(valid32 & ~(0xFFFFFFFFu << b)); }

struct TriRegs { uint32_t a[8], b[8]; };
GTLC_D void tri_fetch(const uint32_t* __restrict__ tris, uint32_t tri, TriRegs& t)
{
  const uint32_t* p = tris + (size_t) tri * kDevTriWords;
  ldg256(p, t.a); ldg256(p + 8, t.b);
}

// ---- re-packing kernels: canonical (bvh8.h) -> traversal layout -----------------------------------------
__global__ void __launch_bounds__(256) bvh8_repack_nodes_kernel(const uint4* __restrict__ canon, uint32_t nodeCount, uint4* __restrict__ out, uint32_t* __restrict__ errors)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nodeCount; i += gridDim.x * blockDim.x) {
    const uint4 n0 = canon[(size_t) i * 5 + 0], n1 = canon[(size_t) i * 5 + 1], n2 = canon[(size_t) i * 5 + 2], n3 = canon[(size_t) i * 5 + 3], n4 = canon[(size_t) i * 5 + 4];
    const uint32_t imask = n0.w >> 24;
    uint32_t valid = imask << 24, running = 0;
    bool bad = false;
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const uint32_t m = ((s < 4 ? n1.z : n1.w) >> (8 * (s & 3))) & 0xFFu;
      if (m == 0u) { if (imask & (1u << s)) bad = true; continue; }
      if (imask & (1u << s)) { if (m != (0x20u | (24u + (uint32_t) s))) bad = true; continue; }
      const uint32_t cnt = (uint32_t) __popc(m >> 5), off = m & 31u;
      if (off != running || cnt > 3u || (m >> 5) != ((1u << cnt) - 1u)) bad = true;     // triangles in slot order, unary counts
      valid |= ((1u << cnt) - 1u) << (3 * s);
      running += cnt;
    }
    if (bad) atomicAdd(errors, 1u);
    uint4* o = out + (size_t) i * 6;
    o[0] = make_uint4(n0.x, n0.y, n0.z, valid);
    o[1] = make_uint4((n0.w & 0xFFu) << 23, ((n0.w >> 8) & 0xFFu) << 23, ((n0.w >> 16) & 0xFFu) << 23, n1.x);
    o[2] = make_uint4(n1.y, 0u, n2.x, n2.y);        // triBase, -, qlo_x
    o[3] = make_uint4(n2.z, n2.w, n3.x, n3.y);      // qlo_y, qlo_z
    o[4] = make_uint4(n3.z, n3.w, n4.x, n4.y);      // qhi_x, qhi_y
    o[5] = make_uint4(n4.z, n4.w, 0u, 0u);          // qhi_z
  }
}

__global__ void __launch_bounds__(256) bvh8_repack_tris_kernel(const float4* __restrict__ canon, uint32_t triCount, const uint32_t* __restrict__ instRankBase, float4* __restrict__ out)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < triCount; i += gridDim.x * blockDim.x) {
    const float4 a = canon[(size_t) i * 3 + 0], b = canon[(size_t) i * 3 + 1], c = canon[(size_t) i * 3 + 2];
    float4* o = out + (size_t) i * 4;
    o[0] = a; o[1] = b; o[2] = c;
    o[3] = make_float4(u2f(instRankBase[f2u(b.w)] + f2u(a.w)), 0.0f, 0.0f, 0.0f);   // rank of (instance, primitive): the closest-hit tie-break
  }
}

// The shared body of "one node step": pops the next child of the lane's current node group G (deferring the siblings),
// fetches and slab-tests it.  Returns the new group in G and the node's triangle hits in (triBase, triMask, valid).
#define GTLC_NODE_STEP(G, stack, sp, r, bt, k47, triBase, triMask, triValid)                                  \
  do {                                                                                                         \
    const uint32_t hits_imask = G.y;                                                                           \
    const uint32_t bit = 31u - (uint32_t) __clz(hits_imask);                                                   \
    G.y &= ~(1u << bit);                                                                                       \
    if (G.y & 0xFF000000u) stack[sp++] = G;                                                                    \
    const uint32_t slot = (bit - 24u) ^ r.octinv;                                                              \
    const uint32_t rel = (uint32_t) __popc(hits_imask & ~(0xFFFFFFFFu << slot));                               \
    NodeRegs nr;                                                                                               \
    node_fetch(P.nodes, G.x + rel, nr);                                                                        \
    const uint32_t hm = node_hitmask(nr, r, bt, k47);                                                          \
    G = make_uint2(nr.a[7], (hm & 0xFF000000u) | (nr.a[3] >> 24));                                             \
    triBase = nr.b[0]; triMask = hm & 0x00FFFFFFu; triValid = nr.a[3];                                         \
  } while (0)

template <int MODE, bool COUNT, bool TEXCUT = false>
__global__ void __launch_bounds__(kTraceBlock, GTLC_TRACE_MIN_BLOCKS) trace_kernel(TraceParams P)
{
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t n = *P.numRays;
  const uint32_t k47 = P.biasWord;
  trace_prologue(P);

  uint2 stack[kStackSize];
  int sp = 0;
  bool active = false;
  bool exhausted = false;   // warp-uniform: the global queue has no more rays
  uint32_t rayIdx = 0;
  RaySetup r;
  float bt = 0.0f, bu = 0.0f, bv = 0.0f; uint32_t bprim = 0xFFFFFFFFu, binst = 0xFFFFFFFFu;
  float tmax = 0.0f;
  uint2 G = make_uint2(0u, 0u);
  uint32_t cNodes = 0, cTris = 0;
  uint32_t clk0 = 0;

  while (true) {
    // ---- refill idle lanes from the global queue (one atomic per warp) -----------------------
    if (!exhausted) {
      const unsigned need = __ballot_sync(FULL, !active);
      if (need) {
        const int leader = __ffs(need) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(P.workCounter, (uint32_t) __popc(need));
        base = __shfl_sync(FULL, base, leader);
        if (!active) {
          rayIdx = base + (uint32_t) __popc(need & ((1u << lane) - 1u));
          if (rayIdx < n) {
            const float4 ro = __ldg(&P.rayO[rayIdx]);
            const float4 rd = __ldg(&P.rayD[rayIdx]);
            ray_setup(r, ro, rd);
            tmax = rd.w;
            bt = tmax; bprim = 0xFFFFFFFFu; binst = 0xFFFFFFFFu; bu = 0.0f; bv = 0.0f;
            sp = 0; cNodes = 0; cTris = 0;
            if (P.clk) clk0 = (uint32_t) clock();
            active = tmax > r.tmin;   // empty interval (e.g. disabled shadow ray): immediate miss
            if (!ray_is_finite(ro, rd)) { active = false; atomicAdd(&P.stats[5], 1ull); }
            G = make_uint2(0u, 0x80000000u);
            if (!active) {
              if (MODE == TRACE_CLOSEST || MODE == TRACE_ANY) { P.hitTUVP[rayIdx] = make_float4(-1.0f, 0.0f, 0.0f, u2f(0xFFFFFFFFu)); if (P.hitInst) P.hitInst[rayIdx] = 0xFFFFFFFFu; }
              // TRACE_SHADOW: an empty interval means "no shadow ray" (rgen:407-410): nothing is added
              if (COUNT && P.visitsPerRay) { P.visitsPerRay[2 * (size_t) rayIdx] = 0; P.visitsPerRay[2 * (size_t) rayIdx + 1] = 0; }
            }
          }
        }
        if (base + (uint32_t) __popc(need) >= n) exhausted = true;
      }
    }
    if (!__any_sync(FULL, active)) { if (exhausted) break; else continue; }

    // ---- traversal steps until enough lanes went idle ------------------------------------------
    while (true) {
      if (active) {
        uint32_t triBase = 0, triMask = 0, triValid = 0;
        if (G.y & 0xFF000000u) {
          GTLC_NODE_STEP(G, stack, sp, r, bt, k47, triBase, triMask, triValid);
          if (COUNT) cNodes++;
        }
        bool done = false;
        while (triMask) {
          const uint32_t j = (uint32_t) __ffs(triMask) - 1u;
          triMask &= triMask - 1u;
          TriRegs tr;
          tri_fetch(P.tris, tri_index(triBase, triValid, j), tr);
          if (COUNT) cTris++;
          float t, u, v;
          if (tri_intersect(r, v3(u2f(tr.a[0]), u2f(tr.a[1]), u2f(tr.a[2])), v3(u2f(tr.a[4]), u2f(tr.a[5]), u2f(tr.a[6])), v3(u2f(tr.b[0]), u2f(tr.b[1]), u2f(tr.b[2])), t, u, v)) {
            const uint32_t prim = tr.a[3], inst = tr.a[7];
            if (t > r.tmin && t < tmax && !cutout_ignored<TEXCUT>(P, rayIdx, u2f(tr.b[3]), prim, inst, u, v)) {
              if (MODE != TRACE_CLOSEST) { bt = t; bu = u; bv = v; bprim = prim; binst = inst; done = true; break; }
              const bool better = (bprim == 0xFFFFFFFFu) || (t < bt) || (t == bt && (inst < binst || (inst == binst && prim < bprim)));
              if (better) { bt = t; bu = u; bv = v; bprim = prim; binst = inst; }
            }
          }
        }
        if (!done && !(G.y & 0xFF000000u)) {
          if (sp == 0) done = true;
          else G = stack[--sp];
        }
        if (done) {
          const bool hit = bprim != 0xFFFFFFFFu;
          if (MODE == TRACE_SHADOW) {
            if (!hit) {
              const float4 c = __ldg(&P.shContrib[rayIdx]);
              const uint32_t target = f2u(c.w);
              float4* dst = (target & 0x80000000u) ? (P.radSample + (target & 0x7FFFFFFFu)) : (P.radPath + target);
              float4 cur = *dst;
              cur.x += c.x; cur.y += c.y; cur.z += c.z;
              *dst = cur;
            } else if (P.aovFlags) {                                   // rgen:431-436: NeeAov = shadowed ? red : green
              const uint32_t target = f2u(__ldg(&P.shContrib[rayIdx]).w);
              const uint32_t slot = (target & 0x80000000u) ? (target & 0x7FFFFFFFu) : P.aovSlotOf[target];
              float4 a0 = P.aovFlags[slot];
              a0.w = u2f(f2u(a0.w) | 8u);
              P.aovFlags[slot] = a0;
            }
          } else {
            P.hitTUVP[rayIdx] = make_float4(hit ? bt : -1.0f, bu, bv, u2f(bprim));
            if (P.hitInst) P.hitInst[rayIdx] = binst;
          }
          if (P.clk) {                                                    // rgen:187-189,501-504: the megakernel's per-pixel timer, per ray here
            uint32_t slot;
            if (MODE == TRACE_SHADOW) { const uint32_t target = f2u(__ldg(&P.shContrib[rayIdx]).w); slot = (target & 0x80000000u) ? (target & 0x7FFFFFFFu) : P.aovSlotOf[target]; }
            else slot = P.aovSlotOf[rayIdx];
            P.clk[slot] += (float) ((uint32_t) clock() - clk0);
          }
          if (COUNT) {
            atomicAdd(&P.stats[MODE == TRACE_SHADOW ? 7 : 1], (unsigned long long) cNodes);
            atomicAdd(&P.stats[MODE == TRACE_SHADOW ? 8 : 2], (unsigned long long) cTris);
            atomicMax(&P.stats[4], (unsigned long long) cNodes);
            if (MODE == TRACE_CLOSEST && hit) atomicAdd(&P.stats[6], 1ull);
            if (P.visitsPerRay) { P.visitsPerRay[2 * (size_t) rayIdx] = cNodes; P.visitsPerRay[2 * (size_t) rayIdx + 1] = cTris; }
          }
          active = false;
        }
      }
      const unsigned idle = __ballot_sync(FULL, !active);
      if (idle == FULL) break;
      if (!exhausted && __popc(idle) >= P.refillThreshold) break;
    }
  }
}

// ---- variant 6: ray tiles staged through TMA into shared memory ------------------------------------
// Every warp owns two 32-ray stages of shared memory (2 x 512 B of rayO + 2 x 512 B of rayD).  Lane 0 claims
// the next 32 rays of the SoA queue with one atomic and issues two `cp.async.bulk` (TMA, 1-D) copies that
// complete on an mbarrier; while the warp traverses, the following tile is already in flight, so a refill
// reads its rays from shared memory instead of waiting for an HBM round trip with the whole warp stalled.
// Traversal itself is the per-lane loop above.
GTLC_D uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

GTLC_D void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "LAB_WAIT:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE;\n"
    "bra LAB_WAIT;\n"
    "DONE:\n"
    "}\n" :: "r"(bar), "r"(parity) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kTraceBlock, 6) trace_kernel_tma(TraceParams P)
{
  __shared__ __align__(128) float4 sO[kTraceBlock / 32][2][32];
  __shared__ __align__(128) float4 sD[kTraceBlock / 32][2][32];
  __shared__ __align__(8) unsigned long long sBar[kTraceBlock / 32][2];
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = *P.numRays;
  const uint32_t k47 = P.biasWord;
  trace_prologue(P);

  // claims the next tile and starts its copy into `stage`; returns (base, count) to the whole warp
  auto issue = [&](int stage, uint32_t& base, uint32_t& cnt) {
    base = 0; cnt = 0;
    if (lane == 0) {
      base = atomicAdd(P.workCounter, 32u);
      cnt = base < n ? min(32u, n - base) : 0u;
      if (cnt) {
        const uint32_t bar = smem_u32(&sBar[warp][stage]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(cnt * 32u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(&sO[warp][stage][0])), "l"(P.rayO + base), "r"(cnt * 16u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(&sD[warp][stage][0])), "l"(P.rayD + base), "r"(cnt * 16u), "r"(bar) : "memory");
      }
    }
    base = __shfl_sync(FULL, base, 0); cnt = __shfl_sync(FULL, cnt, 0);
  };

  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sBar[warp][0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sBar[warp][1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  // stage bookkeeping in scalar registers (selected with `cur`), all warp-uniform
  uint32_t base0, cnt0, base1, cnt1, phase0 = 0u, phase1 = 0u;
  issue(0, base0, cnt0);
  issue(1, base1, cnt1);
  int cur = 0; uint32_t cursor = 0;
  if (cnt0) { mbar_wait(smem_u32(&sBar[warp][0]), phase0); phase0 ^= 1u; }
  bool exhausted = cnt0 == 0u;

  uint2 stack[kStackSize];
  int sp = 0;
  bool active = false;
  uint32_t rayIdx = 0;
  RaySetup r;
  float bt = 0.0f, bu = 0.0f, bv = 0.0f; uint32_t bprim = 0xFFFFFFFFu, binst = 0xFFFFFFFFu;
  float tmax = 0.0f;
  uint2 G = make_uint2(0u, 0u);

  while (true) {
    if (!exhausted) {
      const unsigned need = __ballot_sync(FULL, !active);
      if (need) {
        const uint32_t rank = (uint32_t) __popc(need & ((1u << lane) - 1u));
        const uint32_t needCount = (uint32_t) __popc(need);
        const uint32_t curBase = cur ? base1 : base0, curCnt = cur ? cnt1 : cnt0;
        const uint32_t remaining = curCnt - cursor;
        bool got = false; float4 ro, rd;
        if (!active && rank < remaining) { ro = sO[warp][cur][cursor + rank]; rd = sD[warp][cur][cursor + rank]; rayIdx = curBase + cursor + rank; got = true; }
        if (needCount >= remaining) {
          // current stage used up: swap to the prefetched one and start prefetching into the freed stage
          __syncwarp();
          const int old = cur; cur ^= 1;
          const uint32_t nxtBase = cur ? base1 : base0, nxtCnt = cur ? cnt1 : cnt0;
          if (nxtCnt) {
            if (cur) { mbar_wait(smem_u32(&sBar[warp][1]), phase1); phase1 ^= 1u; }
            else { mbar_wait(smem_u32(&sBar[warp][0]), phase0); phase0 ^= 1u; }
          }
          const uint32_t over = rank - remaining;
          if (!active && rank >= remaining && over < nxtCnt) { ro = sO[warp][cur][over]; rd = sD[warp][cur][over]; rayIdx = nxtBase + over; got = true; }
          cursor = min(needCount - remaining, nxtCnt);
          if (nxtCnt == 0u) exhausted = true;
          __syncwarp();
          if (old) issue(1, base1, cnt1); else issue(0, base0, cnt0);
        } else {
          cursor += needCount;
        }
        if (got) {
          ray_setup(r, ro, rd);
          tmax = rd.w;
          bt = tmax; bprim = 0xFFFFFFFFu; binst = 0xFFFFFFFFu; bu = 0.0f; bv = 0.0f;
          sp = 0;
          active = tmax > r.tmin;
          if (!ray_is_finite(ro, rd)) { active = false; atomicAdd(&P.stats[5], 1ull); }
          G = make_uint2(0u, 0x80000000u);
          if (!active && MODE != TRACE_SHADOW) { P.hitTUVP[rayIdx] = make_float4(-1.0f, 0.0f, 0.0f, u2f(0xFFFFFFFFu)); if (P.hitInst) P.hitInst[rayIdx] = 0xFFFFFFFFu; }
        }
      }
    }
    if (!__any_sync(FULL, active)) { if (exhausted) break; else continue; }

    while (true) {
      if (active) {
        uint32_t triBase = 0, triMask = 0, triValid = 0;
        if (G.y & 0xFF000000u) GTLC_NODE_STEP(G, stack, sp, r, bt, k47, triBase, triMask, triValid);
        bool done = false;
        while (triMask) {
          const uint32_t j = (uint32_t) __ffs(triMask) - 1u;
          triMask &= triMask - 1u;
          TriRegs tr;
          tri_fetch(P.tris, tri_index(triBase, triValid, j), tr);
          float t, u, v;
          if (tri_intersect(r, v3(u2f(tr.a[0]), u2f(tr.a[1]), u2f(tr.a[2])), v3(u2f(tr.a[4]), u2f(tr.a[5]), u2f(tr.a[6])), v3(u2f(tr.b[0]), u2f(tr.b[1]), u2f(tr.b[2])), t, u, v)) {
            const uint32_t prim = tr.a[3], inst = tr.a[7];
            if (t > r.tmin && t < tmax && !cutout_ignored<false>(P, rayIdx, u2f(tr.b[3]), prim, inst, u, v)) {
              if (MODE != TRACE_CLOSEST) { bt = t; bu = u; bv = v; bprim = prim; binst = inst; done = true; break; }
              const bool better = (bprim == 0xFFFFFFFFu) || (t < bt) || (t == bt && (inst < binst || (inst == binst && prim < bprim)));
              if (better) { bt = t; bu = u; bv = v; bprim = prim; binst = inst; }
            }
          }
        }
        if (!done && !(G.y & 0xFF000000u)) {
          if (sp == 0) done = true;
          else G = stack[--sp];
        }
        if (done) {
          const bool hit = bprim != 0xFFFFFFFFu;
          if (MODE == TRACE_SHADOW) {
            if (!hit) {
              const float4 c = __ldg(&P.shContrib[rayIdx]);
              const uint32_t target = f2u(c.w);
              float4* dst = (target & 0x80000000u) ? (P.radSample + (target & 0x7FFFFFFFu)) : (P.radPath + target);
              float4 cur4 = *dst;
              cur4.x += c.x; cur4.y += c.y; cur4.z += c.z;
              *dst = cur4;
            }
          } else {
            P.hitTUVP[rayIdx] = make_float4(hit ? bt : -1.0f, bu, bv, u2f(bprim));
            if (P.hitInst) P.hitInst[rayIdx] = binst;
          }
          active = false;
        }
      }
      const unsigned idle = __ballot_sync(FULL, !active);
      if (idle == FULL) break;
      if (!exhausted && __popc(idle) >= P.refillThreshold) break;
    }
  }
}

// ---- variant 7 (default): warp-cooperative batched triangle tests ---------------------------------------
// The per-lane triangle loop above runs with ~3 of 32 lanes active and was 37 % of all issued
// instructions.  Here a lane only APPENDS the triangles its node step found to a per-warp queue in shared
// memory ((owner lane << 27) | triangle index) and goes on traversing; once the queue holds >= 32 entries (or
// nobody has node work left) the whole warp drains it 32 tests at a time: lane L tests entry L against the
// OWNER's ray, whose origin / shear basis / interval live in shared memory, and publishes a hit with one 64-bit
// shared-memory atomicMin on (t bits << 32 | (instance, primitive) rank) -- the same closest-hit order as
// the per-lane loop, so results are bit-identical.  Culling sees a new t a few node steps late (measured on the CPU
// walker: +2.5 % node visits per step of delay), which the 32-wide triangle tests more than pay for.
#ifndef GTLC_ENQUEUE_UNROLL
#define GTLC_ENQUEUE_UNROLL 1   // C3 stand-in: closest trace 65.1 -> 64.3 ms, shadow 19.1 -> 18.9 ms; C2 33.5 -> 33.2 ms (identical queue content)
#endif
constexpr int kBatchQueueCap = 832;   // 31 left-overs + 32 lanes x 24 triangles of one node step, rounded up

template <int MODE, bool TEXCUT = false>
__global__ void __launch_bounds__(kTraceBlock, GTLC_TRACE_MIN_BLOCKS) trace_kernel_batched(TraceParams P)
{
  constexpr int kWarps = kTraceBlock / 32;
  // word-major ([4][32], not [32][4]): a drain reads word k of 32 random owners, which then spread over 8 bank groups
  // instead of 2 (measured on the C3 stand-in: L1 data pipe 77 % busy, mostly these reads)
  __shared__ float4 sRay[kWarps][4][32];             // (o.xyz, tMin) (mx.xyz, tMax) (my.xyz, ray index) (mz.xyz, -)
  __shared__ unsigned long long sKey[kWarps][32];    // best (t bits << 32 | rank) per owner lane
  __shared__ float4 sHit[kWarps][32];                // (u, v, primitive, instance) of the best hit
  __shared__ uint32_t sQueue[kWarps][kBatchQueueCap];
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t n = *P.numRays;
  const uint32_t k47 = P.biasWord;
  trace_prologue(P);

  uint2 stack[kStackSize];
  int sp = 0;
  bool active = false, exhausted = false, pending = false;
  uint32_t rayIdx = 0;
  RaySetup r;
  float bt = 0.0f, tmax = 0.0f;
  unsigned long long keyInit = 0ull;
  uint2 G = make_uint2(0u, 0u);
  uint32_t qCount = 0;    // warp-uniform

  while (true) {
    // ---- refill ----
    if (!exhausted) {
      const unsigned need = __ballot_sync(FULL, !active);
      if (need) {
        const int leader = __ffs(need) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(P.workCounter, (uint32_t) __popc(need));
        base = __shfl_sync(FULL, base, leader);
        if (!active) {
          rayIdx = base + (uint32_t) __popc(need & ((1u << lane) - 1u));
          if (rayIdx < n) {
            const float4 ro = __ldg(&P.rayO[rayIdx]);
            const float4 rd = __ldg(&P.rayD[rayIdx]);
            ray_setup(r, ro, rd);
            tmax = rd.w; bt = tmax; sp = 0; pending = false;
            active = tmax > r.tmin;
            if (!ray_is_finite(ro, rd)) { active = false; atomicAdd(&P.stats[5], 1ull); }
            G = make_uint2(0u, 0x80000000u);
            if (active) {
              sRay[w][0][lane] = make_float4(r.o.x, r.o.y, r.o.z, r.tmin);
              sRay[w][1][lane] = make_float4(r.mx.x, r.mx.y, r.mx.z, tmax);
              sRay[w][2][lane] = make_float4(r.my.x, r.my.y, r.my.z, u2f(rayIdx));
              sRay[w][3][lane] = make_float4(r.mz.x, r.mz.y, r.mz.z, 0.0f);
              keyInit = ((unsigned long long) f2u(tmax) << 32) | 0xFFFFFFFFull;
              sKey[w][lane] = keyInit;
            } else if (MODE != TRACE_SHADOW) {
              P.hitTUVP[rayIdx] = make_float4(-1.0f, 0.0f, 0.0f, u2f(0xFFFFFFFFu)); if (P.hitInst) P.hitInst[rayIdx] = 0xFFFFFFFFu;
            }
          }
        }
        if (base + (uint32_t) __popc(need) >= n) exhausted = true;
        __syncwarp();
      }
    }
    if (!__any_sync(FULL, active)) { if (exhausted) break; else continue; }

    while (true) {
      // ---- A: one node step per lane with node work; its triangles go to the warp queue ----
      uint32_t triBase = 0, triMask = 0, triValid = 0;
      if (active) {
        if (!(G.y & 0xFF000000u) && sp > 0) G = stack[--sp];
        if (G.y & 0xFF000000u) GTLC_NODE_STEP(G, stack, sp, r, bt, k47, triBase, triMask, triValid);
      }
      // enqueue, one triangle per lane per round (owners interleave in the queue)
#if GTLC_ENQUEUE_UNROLL
      {
        // the first three rounds without the loop: their ballots are taken up front from the lanes' triangle counts (a node step
        // rarely finds more than three triangles per lane); queue content and order are those of the loop below
        const uint32_t cnt = (uint32_t) __popc(triMask);
        const unsigned b0 = __ballot_sync(FULL, cnt > 0u);
        if (b0) {
          const unsigned b1 = __ballot_sync(FULL, cnt > 1u), b2 = __ballot_sync(FULL, cnt > 2u);
          const unsigned lt = (1u << lane) - 1u;
          const uint32_t owner = (uint32_t) lane << 27;
          if (cnt > 0u) {
            const uint32_t m = triMask & (0u - triMask);                       // lowest set bit
            sQueue[w][qCount + (uint32_t) __popc(b0 & lt)] = owner | (triBase + (uint32_t) __popc(triValid & (m - 1u)));
            triMask ^= m; pending = true;
          }
          qCount += (uint32_t) __popc(b0);
          if (b1) {
            if (cnt > 1u) {
              const uint32_t m = triMask & (0u - triMask);
              sQueue[w][qCount + (uint32_t) __popc(b1 & lt)] = owner | (triBase + (uint32_t) __popc(triValid & (m - 1u)));
              triMask ^= m;
            }
            qCount += (uint32_t) __popc(b1);
            if (b2) {
              if (cnt > 2u) {
                const uint32_t m = triMask & (0u - triMask);
                sQueue[w][qCount + (uint32_t) __popc(b2 & lt)] = owner | (triBase + (uint32_t) __popc(triValid & (m - 1u)));
                triMask ^= m;
              }
              qCount += (uint32_t) __popc(b2);
            }
          }
        }
      }
#endif
      unsigned have = __ballot_sync(FULL, triMask != 0u);
      while (have) {
        if (triMask) {
          const uint32_t j = (uint32_t) __ffs(triMask) - 1u;
          triMask &= triMask - 1u;
          sQueue[w][qCount + (uint32_t) __popc(have & ((1u << lane) - 1u))] = ((uint32_t) lane << 27) | tri_index(triBase, triValid, j);
          pending = true;
        }
        qCount += (uint32_t) __popc(have);
        have = __ballot_sync(FULL, triMask != 0u);
      }
      // ---- B: drain the queue 32-wide when it is full enough or nothing else can progress ----
      const unsigned nodeWork = __ballot_sync(FULL, active && ((G.y & 0xFF000000u) != 0u || sp > 0));
      if (qCount >= 32u || (qCount > 0u && nodeWork == 0u)) {
        __syncwarp();
        for (uint32_t qb = 0; qb < qCount; qb += 32u) {
          const uint32_t qi = qb + (uint32_t) lane;
          bool won = false; unsigned long long key = 0ull; uint32_t owner = 0; float hu = 0.0f, hv = 0.0f; uint32_t hprim = 0, hinst = 0;
          if (qi < qCount) {
            const uint32_t e = sQueue[w][qi];
            owner = e >> 27;
            TriRegs tr;
            tri_fetch(P.tris, e & 0x07FFFFFFu, tr);
            const float4 r0 = sRay[w][0][owner], r1 = sRay[w][1][owner], r2 = sRay[w][2][owner], r3 = sRay[w][3][owner];
            RaySetup q;
            q.o = v3(r0.x, r0.y, r0.z); q.mx = v3(r1.x, r1.y, r1.z); q.my = v3(r2.x, r2.y, r2.z); q.mz = v3(r3.x, r3.y, r3.z);
            float t, u, v;
            if (tri_intersect(q, v3(u2f(tr.a[0]), u2f(tr.a[1]), u2f(tr.a[2])), v3(u2f(tr.a[4]), u2f(tr.a[5]), u2f(tr.a[6])), v3(u2f(tr.b[0]), u2f(tr.b[1]), u2f(tr.b[2])), t, u, v)) {
              hprim = tr.a[3]; hinst = tr.a[7];
              if (t > r0.w && t < r1.w && !cutout_ignored<TEXCUT>(P, f2u(r2.w), u2f(tr.b[3]), hprim, hinst, u, v)) {
                key = ((unsigned long long) f2u(t) << 32) | (unsigned long long) tr.b[4];
                won = key < atomicMin(&sKey[w][owner], key);
                hu = u; hv = v;
              }
            }
          }
          __syncwarp();
          if (won && sKey[w][owner] == key) sHit[w][owner] = make_float4(hu, hv, u2f(hprim), u2f(hinst));
          __syncwarp();
        }
        qCount = 0;
        pending = false;
        if (active) bt = u2f((uint32_t) (sKey[w][lane] >> 32));
      }
      // ---- C: finished rays ----
      if (active && !pending) {
        const unsigned long long key = sKey[w][lane];
        const bool hit = key != keyInit;
        bool done = !(G.y & 0xFF000000u) && sp == 0;
        if (MODE != TRACE_CLOSEST && hit) done = true;
        if (done) {
          if (MODE == TRACE_SHADOW) {
            if (!hit) {
              const float4 c = __ldg(&P.shContrib[rayIdx]);
              const uint32_t target = f2u(c.w);
              float4* dst = (target & 0x80000000u) ? (P.radSample + (target & 0x7FFFFFFFu)) : (P.radPath + target);
              float4 cur = *dst;
              cur.x += c.x; cur.y += c.y; cur.z += c.z;
              *dst = cur;
            }
          } else {
            const float4 h4 = sHit[w][lane];
            P.hitTUVP[rayIdx] = hit ? make_float4(u2f((uint32_t) (key >> 32)), h4.x, h4.y, h4.z) : make_float4(-1.0f, 0.0f, 0.0f, u2f(0xFFFFFFFFu));
            if (P.hitInst) P.hitInst[rayIdx] = hit ? f2u(h4.w) : 0xFFFFFFFFu;
          }
          active = false;
        }
      }
      const unsigned idle = __ballot_sync(FULL, !active);
      if (idle == FULL) break;
      if (!exhausted && __popc(idle) >= P.refillThreshold) break;
    }
  }
}

} // namespace gtlc
