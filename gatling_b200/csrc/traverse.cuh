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

namespace gtlc {

#ifndef GTLC_TRACE_MIN_BLOCKS
#define GTLC_TRACE_MIN_BLOCKS 7   // 72 registers: 7 CTAs = 28 warps per SM (measured best of 6/7/8 on C2)
#endif
constexpr int kTraceBlock = 128;         // threads per CTA
constexpr int kStackSize = 48;           // >= 2 x wide-tree depth (node groups + postponed triangle groups; checked on the host)
constexpr int kRefillThreshold = 8;      // idle lanes that trigger a refill
constexpr float kBoxPad = 1.0000004f;    // Ize 2013 robust slab padding

enum { TRACE_CLOSEST = 0, TRACE_ANY = 1, TRACE_SHADOW = 2 };

struct TraceParams {
  const uint4* __restrict__ nodes;      // 5 x uint4 per BVH8 node
  const float4* __restrict__ tris;      // 3 x float4 per triangle
  const float4* __restrict__ rayO;      // (origin.xyz, tMin)
  const float4* __restrict__ rayD;      // (direction.xyz, tMax)
  float4* __restrict__ hitTUVP;         // (t, u, v, bits primitiveId); miss: t = -1
  uint32_t* __restrict__ hitInst;       // instance index; miss: 0xFFFFFFFF
  const uint32_t* __restrict__ numRays; // device-resident queue length
  uint32_t* workCounter;                // zeroed before the launch
  // TRACE_SHADOW epilogue: unoccluded -> target.xyz += contrib
  const float4* __restrict__ shContrib; // (contrib.xyz, bits target): target bit31 = sample slot, else path queue position
  float4* __restrict__ radPath;         // (radiance.xyz, rng) of the continuing paths
  float4* __restrict__ radSample;       // (radiance.xyz, -) of finished samples
  // bookkeeping
  unsigned long long* stats;            // [1] node visits, [2] triangle tests, [4] max node visits of one ray, [5] non-finite rays
  uint32_t* visitsPerRay;               // optional 2 x uint32 per ray
  int refillThreshold;                  // idle lanes that trigger a refill (kRefillThreshold by default)
  int triThreshold;                     // variant 5: lanes with pending triangles that trigger a triangle phase
  const float4* __restrict__ seedRad;   // cutout any-hit: per-ray RNG seed = bits of seedRad[i].w (path queue `rad`), or
  const uint32_t* __restrict__ seedU32; //   seedU32[i] (shadow queue); both NULL = seed 0
  uint32_t biasWord;                           // 0x47000000 (see biased_byte)
  float4* aovFlags;                            // NEE debug AOV (first bounce, variant 0 only): per-sample capture word a[0]; bit 3 of .w = shadow ray occluded
  const uint32_t* __restrict__ aovSlotOf;      //   sample slot of a continuing path (path queue `slot`)
  const uint32_t* __restrict__ instRankBase;   // variant 7: first global (instance, primitive)-ordered triangle rank of every instance
};

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
  V3 o, d;
  float tmin;
  V3 idir;
  uint32_t octinv;
  // Woop shear as three sparse basis vectors: mx = e_kx - Sx e_kz, my = e_ky - Sy e_kz, mz = Sz e_kz, so that
  // the per-triangle code needs no data-dependent component selection (same definition in the CPU oracle).
  V3 mx, my, mz;
};

GTLC_D void ray_setup(RaySetup& r, float4 ro, float4 rd)
{
  r.o = v3(ro.x, ro.y, ro.z); r.d = v3(rd.x, rd.y, rd.z); r.tmin = ro.w;
  // |d| is clamped to 2^-80 for the slab tests only (Ylitie et al. 2017): with an exactly-zero or denormal
  // component 1/d = inf turns the quantised slab arithmetic q*s + o into inf - inf = NaN, the NaN-dropping
  // min/max then never cull along that axis and the ray walks a whole slice of the tree.
  const float kMinDir = 8.271806125530277e-25f;   // 2^-80
  r.idir = v3(1.0f / (fabsf(r.d.x) > kMinDir ? r.d.x : copysignf(kMinDir, r.d.x)),
              1.0f / (fabsf(r.d.y) > kMinDir ? r.d.y : copysignf(kMinDir, r.d.y)),
              1.0f / (fabsf(r.d.z) > kMinDir ? r.d.z : copysignf(kMinDir, r.d.z)));
  uint32_t oct = (r.d.x < 0.0f ? 1u : 0u) | (r.d.y < 0.0f ? 2u : 0u) | (r.d.z < 0.0f ? 4u : 0u);
  r.octinv = 7u - oct;
  float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
  int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  if (comp(r.d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
  const float Sz = comp(r.idir, kz);          // = 1 / d[kz] exactly (the dominant component is never clamped): no extra division
  const float Sx = comp(r.d, kx) * Sz;
  const float Sy = comp(r.d, ky) * Sz;
  r.mx = v3(kx == 0 ? 1.0f : (kz == 0 ? -Sx : 0.0f), kx == 1 ? 1.0f : (kz == 1 ? -Sx : 0.0f), kx == 2 ? 1.0f : (kz == 2 ? -Sx : 0.0f));
  r.my = v3(ky == 0 ? 1.0f : (kz == 0 ? -Sy : 0.0f), ky == 1 ? 1.0f : (kz == 1 ? -Sy : 0.0f), ky == 2 ? 1.0f : (kz == 2 ? -Sy : 0.0f));
  r.mz = v3(kz == 0 ? Sz : 0.0f, kz == 1 ? Sz : 0.0f, kz == 2 ? Sz : 0.0f);
}

// Stochastic cutout (rp_main.ahit:35-61: `if (rand > mdl_cutout_opacity) ignoreIntersectionEXT`).  The reference
// draws from the path's RNG once per any-hit invocation, in an order the Vulkan spec leaves undefined; here the
// decision is a pure function of (ray seed, primitive, instance), so it does not depend on traversal order:
// xi = uintAsFloat(hash_pcg32(seed ^ hash_theironborn(prim ^ inst * 0x9E3779B1))).  `opacity` rides in the spare
// .w of the triangle record (1.0 = opaque, the common case costs one compare).
GTLC_D bool cutout_ignored(const TraceParams& P, uint32_t rayIdx, float opacity, uint32_t prim, uint32_t inst)
{
  if (opacity >= 1.0f) return false;
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

// Slab-tests the 8 quantised child boxes of one node; returns the 32-bit hit mask
// (bits 31..24: internal children in traversal priority order, bits 23..0: triangles).
// t = (32768 + q) * s + (o - 32768 s) with s = 2^e / d, o = (p - origin) / d; branch-free.
GTLC_D uint32_t node_hitmask(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4,
                             const RaySetup& r, float tbest, uint32_t k47)
{
  const uint32_t e = n0.w;
  const float sx = u2f((e & 0xFFu) << 23) * r.idir.x;
  const float sy = u2f(((e >> 8) & 0xFFu) << 23) * r.idir.y;
  const float sz = u2f(((e >> 16) & 0xFFu) << 23) * r.idir.z;
  const float ox = (u2f(n0.x) - r.o.x) * r.idir.x - 32768.0f * sx;
  const float oy = (u2f(n0.y) - r.o.y) * r.idir.y - 32768.0f * sy;
  const float oz = (u2f(n0.z) - r.o.z) * r.idir.z - 32768.0f * sz;
  const bool nx = r.d.x < 0.0f, ny = r.d.y < 0.0f, nz = r.d.z < 0.0f;
  const uint32_t octinv4 = r.octinv * 0x01010101u;
  uint32_t hitmask = 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const uint32_t meta4 = half ? n1.w : n1.z;
    // word-level decode of the 4 meta bytes (Ylitie et al. 2017, listing 2)
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;          // bit 4 set iff low5 >= 24
    const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xFFu;                     // 0xFF in every internal-child byte
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
    const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
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
      const uint32_t bits = __byte_perm(child_bits4, 0u, 0x4440u | (uint32_t) j);
      const uint32_t idx = (bit_index4 >> (8 * j)) & 0xFFu;
      hitmask |= (tn <= tf) ? (bits << idx) : 0u;
    }
  }
  return hitmask;
}

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, GTLC_TRACE_MIN_BLOCKS) trace_kernel(TraceParams P)
{
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t n = *P.numRays;
  const uint32_t k47 = P.biasWord;

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
        uint2 T;
        if (G.y & 0xFF000000u) {
          const uint32_t hits_imask = G.y;
          const uint32_t bit = 31u - (uint32_t) __clz(hits_imask);
          G.y &= ~(1u << bit);
          if (G.y & 0xFF000000u) stack[sp++] = G;
          const uint32_t slot = (bit - 24u) ^ r.octinv;
          const uint32_t rel = (uint32_t) __popc(hits_imask & ~(0xFFFFFFFFu << slot));
          const uint32_t node = G.x + rel;
          const uint4* np = P.nodes + (size_t) node * 5;
          const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
          if (COUNT) cNodes++;
          const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, r, bt, k47);
          G = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
          T = make_uint2(n1.y, hm & 0x00FFFFFFu);
        } else {
          T = make_uint2(0u, 0u);
        }
        bool done = false;
        while (T.y) {
          const uint32_t j = (uint32_t) __ffs(T.y) - 1u;
          T.y &= T.y - 1u;
          const float4* tp = P.tris + (size_t) (T.x + j) * 3;
          const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
          if (COUNT) cTris++;
          float t, u, v;
          if (tri_intersect(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v)) {
            if (t > r.tmin && t < tmax && !cutout_ignored(P, rayIdx, c.w, f2u(a.w), f2u(b.w))) {
              const uint32_t prim = f2u(a.w), inst = f2u(b.w);
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
          if (COUNT) {
            atomicAdd(&P.stats[1], (unsigned long long) cNodes);
            atomicAdd(&P.stats[2], (unsigned long long) cTris);
            atomicMax(&P.stats[4], (unsigned long long) cNodes);
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

// ---- variant 5: two-deep pending-triangle FIFO ------------------------------------------------------
// A lane keeps up to two triangle groups (T0, T1) in registers and goes on with node steps while it has
// room; the warp enters a triangle phase only when >= triThreshold lanes have triangles pending or nobody
// can make progress on nodes.  Culling is delayed by at most one node step per lane (unlike variant 1,
// which parks triangle groups on the stack indefinitely).  Hits are identical to variant 0.
template <int MODE>
__global__ void __launch_bounds__(kTraceBlock) trace_kernel_v2(TraceParams P)
{
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t n = *P.numRays;
  const uint32_t k47 = P.biasWord;
  const int triThr = P.triThreshold;

  uint2 stack[kStackSize];
  int sp = 0;
  bool active = false;
  bool exhausted = false;
  uint32_t rayIdx = 0;
  RaySetup r;
  float bt = 0.0f, bu = 0.0f, bv = 0.0f; uint32_t bprim = 0xFFFFFFFFu, binst = 0xFFFFFFFFu;
  float tmax = 0.0f;
  uint2 G = make_uint2(0u, 0u), T0 = make_uint2(0u, 0u), T1 = make_uint2(0u, 0u);

  while (true) {
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
            sp = 0;
            active = tmax > r.tmin;
            if (!ray_is_finite(ro, rd)) { active = false; atomicAdd(&P.stats[5], 1ull); }
            G = make_uint2(0u, 0x80000000u); T0 = make_uint2(0u, 0u); T1 = make_uint2(0u, 0u);
            if (!active && MODE != TRACE_SHADOW) { P.hitTUVP[rayIdx] = make_float4(-1.0f, 0.0f, 0.0f, u2f(0xFFFFFFFFu)); if (P.hitInst) P.hitInst[rayIdx] = 0xFFFFFFFFu; }
          }
        }
        if (base + (uint32_t) __popc(need) >= n) exhausted = true;
      }
    }
    if (!__any_sync(FULL, active)) { if (exhausted) break; else continue; }

    while (true) {
      bool done = false;
      // ---- A: one node step for every lane that has node work and room for its triangles ----
      if (active && T1.y == 0u) {
        if (!(G.y & 0xFF000000u) && sp > 0) G = stack[--sp];
        if (G.y & 0xFF000000u) {
          const uint32_t hits_imask = G.y;
          const uint32_t bit = 31u - (uint32_t) __clz(hits_imask);
          G.y &= ~(1u << bit);
          if (G.y & 0xFF000000u) stack[sp++] = G;
          const uint32_t slot = (bit - 24u) ^ r.octinv;
          const uint32_t rel = (uint32_t) __popc(hits_imask & ~(0xFFFFFFFFu << slot));
          const uint4* np = P.nodes + (size_t) (G.x + rel) * 5;
          const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
          const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, r, bt, k47);
          G = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
          const uint2 Tn = make_uint2(n1.y, hm & 0x00FFFFFFu);
          if (Tn.y) { if (T0.y == 0u) T0 = Tn; else T1 = Tn; }
        }
      }
      // ---- B: triangle phase when enough lanes wait for it, or when nothing else can progress ----
      const bool canNode = active && T1.y == 0u && ((G.y & 0xFF000000u) != 0u || sp > 0);
      unsigned pendMask = __ballot_sync(FULL, active && T0.y != 0u);
      const unsigned nodeMask = __ballot_sync(FULL, canNode);
      if (pendMask && (__popc(pendMask) >= triThr || nodeMask == 0u)) {
        do {
          if (active && !done && T0.y != 0u) {
            const uint32_t j = (uint32_t) __ffs(T0.y) - 1u;
            T0.y &= T0.y - 1u;
            const float4* tp = P.tris + (size_t) (T0.x + j) * 3;
            const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float t, u, v;
            if (tri_intersect(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v)) {
              if (t > r.tmin && t < tmax && !cutout_ignored(P, rayIdx, c.w, f2u(a.w), f2u(b.w))) {
                const uint32_t prim = f2u(a.w), inst = f2u(b.w);
                if (MODE != TRACE_CLOSEST) { bt = t; bu = u; bv = v; bprim = prim; binst = inst; done = true; }
                else {
                  const bool better = (bprim == 0xFFFFFFFFu) || (t < bt) || (t == bt && (inst < binst || (inst == binst && prim < bprim)));
                  if (better) { bt = t; bu = u; bv = v; bprim = prim; binst = inst; }
                }
              }
            }
            if (T0.y == 0u) { T0 = T1; T1 = make_uint2(0u, 0u); }
          }
          pendMask = __ballot_sync(FULL, active && !done && T0.y != 0u);
        } while (pendMask);
      }
      if (active && !done && T0.y == 0u && !(G.y & 0xFF000000u) && sp == 0) done = true;
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
          }
        } else {
          P.hitTUVP[rayIdx] = make_float4(hit ? bt : -1.0f, bu, bv, u2f(bprim));
          if (P.hitInst) P.hitInst[rayIdx] = binst;
        }
        active = false;
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
// Traversal itself is variant 0.
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
        uint2 T;
        if (G.y & 0xFF000000u) {
          const uint32_t hits_imask = G.y;
          const uint32_t bit = 31u - (uint32_t) __clz(hits_imask);
          G.y &= ~(1u << bit);
          if (G.y & 0xFF000000u) stack[sp++] = G;
          const uint32_t slot = (bit - 24u) ^ r.octinv;
          const uint32_t rel = (uint32_t) __popc(hits_imask & ~(0xFFFFFFFFu << slot));
          const uint4* np = P.nodes + (size_t) (G.x + rel) * 5;
          const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
          const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, r, bt, k47);
          G = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
          T = make_uint2(n1.y, hm & 0x00FFFFFFu);
        } else {
          T = make_uint2(0u, 0u);
        }
        bool done = false;
        while (T.y) {
          const uint32_t j = (uint32_t) __ffs(T.y) - 1u;
          T.y &= T.y - 1u;
          const float4* tp = P.tris + (size_t) (T.x + j) * 3;
          const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
          float t, u, v;
          if (tri_intersect(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v)) {
            if (t > r.tmin && t < tmax && !cutout_ignored(P, rayIdx, c.w, f2u(a.w), f2u(b.w))) {
              const uint32_t prim = f2u(a.w), inst = f2u(b.w);
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

// ---- variant 7: warp-cooperative batched triangle tests ----------------------------------------------
// The per-lane triangle loop of variant 0 runs with ~3 of 32 lanes active and was 37 % of all issued
// instructions.  Here a lane only APPENDS the triangles its node step found to a per-warp queue in shared
// memory ((owner lane << 27) | triangle index) and goes on traversing; once the queue holds >= 32 entries (or
// nobody has node work left) the whole warp drains it 32 tests at a time: lane L tests entry L against the
// OWNER's ray, whose origin / shear basis / interval live in shared memory, and publishes a hit with one 64-bit
// shared-memory atomicMin on (t bits << 32 | (instance, primitive) rank) -- the same closest-hit order as
// variant 0, so results are bit-identical.  Culling sees a new t a few node steps late (measured on the CPU
// walker: +2.5 % node visits per step of delay), which the 32-wide triangle tests more than pay for.
constexpr int kBatchQueueCap = 832;   // 31 left-overs + 32 lanes x 24 triangles of one node step, rounded up

template <int MODE>
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
      uint32_t triBase = 0, triMask = 0;
      if (active) {
        if (!(G.y & 0xFF000000u) && sp > 0) G = stack[--sp];
        if (G.y & 0xFF000000u) {
          const uint32_t hits_imask = G.y;
          const uint32_t bit = 31u - (uint32_t) __clz(hits_imask);
          G.y &= ~(1u << bit);
          if (G.y & 0xFF000000u) stack[sp++] = G;
          const uint32_t slot = (bit - 24u) ^ r.octinv;
          const uint32_t rel = (uint32_t) __popc(hits_imask & ~(0xFFFFFFFFu << slot));
          const uint4* np = P.nodes + (size_t) (G.x + rel) * 5;
          const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
          const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, r, bt, k47);
          G = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
          triBase = n1.y; triMask = hm & 0x00FFFFFFu;
        }
      }
      // enqueue, one triangle per lane per round (owners interleave in the queue)
      unsigned have = __ballot_sync(FULL, triMask != 0u);
      while (have) {
        if (triMask) {
          const uint32_t j = (uint32_t) __ffs(triMask) - 1u;
          triMask &= triMask - 1u;
          sQueue[w][qCount + (uint32_t) __popc(have & ((1u << lane) - 1u))] = ((uint32_t) lane << 27) | (triBase + j);
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
            const float4* tp = P.tris + (size_t) (e & 0x07FFFFFFu) * 3;
            const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
            const float4 r0 = sRay[w][0][owner], r1 = sRay[w][1][owner], r2 = sRay[w][2][owner], r3 = sRay[w][3][owner];
            RaySetup q;
            q.o = v3(r0.x, r0.y, r0.z); q.mx = v3(r1.x, r1.y, r1.z); q.my = v3(r2.x, r2.y, r2.z); q.mz = v3(r3.x, r3.y, r3.z);
            float t, u, v;
            if (tri_intersect(q, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, u, v)) {
              hprim = f2u(a.w); hinst = f2u(b.w);
              if (t > r0.w && t < r1.w && !cutout_ignored(P, f2u(r2.w), c.w, hprim, hinst)) {
                key = ((unsigned long long) f2u(t) << 32) | (unsigned long long) (__ldg(&P.instRankBase[hinst]) + hprim);
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
