// Device builder of the compressed BVH8 (replaces cgpuCreateBlas/cgpuCreateTlas,
// /root/reference/src/cgpu/impl/CgpuVk.cpp:2561-2854, fed by _giBuildGeometryStructures, Gi.cpp:784-1204, for
// scenes that are edited interactively: DirtyBvh, Gi.cpp:636,648,663).  The whole build stays on the GPU:
//
//   flatten      instance x face -> world-space triangle record (same arithmetic as the host flatten)
//   bounds       centroid bounds of the scene (block reduce + ordered-int atomics)
//   morton       63-bit Morton key of every centroid (21 bits per axis)
//   radix sort   8 x 8-bit LSD passes over (key, triangle) pairs, stable (sort.cuh's count/scan/scatter scheme)
//   karras       binary radix tree over the sorted keys (Karras 2012); ties broken by position          (bvhBuilder = 1)
//   ploc         or: parallel locally-ordered clustering over the sorted order (Meister & Bittner 2018): every cluster
//                looks `kPlocRadius` neighbours to each side for the partner with the smallest merged surface area,
//                mutual pairs merge; ~SAH-builder quality for a few more milliseconds                          (bvhBuilder = 2)
//   refit + DP   bottom-up with one atomic counter per node: boxes, then the SAH-optimal "present this subtree as
//                <= i children" table of Ylitie, Karras, Laine 2017 section 3.1 (same recurrence as bvh8_build.cpp)
//   collapse     breadth-first, one launch pair per level of the wide tree: count -> scan -> emit, so that node and
//                triangle indices come from prefix sums (no allocation atomics => the result is deterministic)
//
// Everything is a pure function of the input triangles; oracle/orc_lbvh8.h restates it on the CPU and the parity
// test demands byte-identical node and triangle arrays.  Floating point: one IEEE operation per written
// operator (-fmad=false), comparisons exact; the quantisation grid is chosen with exact power-of-two arithmetic
// in double (no log2), unlike the host SAH builder -- the two builders produce different, equally valid trees.
#pragma once
#include "bvh8.h"
#include "wavefront.cuh"

namespace gtlc {

constexpr int kLbvhBlock = 256;
constexpr uint32_t kLbvhMaxLeaf = 3;
constexpr float kLbvhNodeCost = 0.6f;
constexpr float kLbvhInf = 3.402823466e38f;

struct LbvhDp { float cost[7]; uint8_t split[7]; uint8_t rootSplit; uint8_t rootIsLeaf; uint8_t pad[3]; };
static_assert(sizeof(LbvhDp) == 40, "LbvhDp is 40 bytes");

// Binary tree over N sorted triangles: refs 0..N-2 are internal nodes, refs N-1..2N-2 are leaves (ref - (N-1) =
// sorted position).
struct LbvhTree {
  uint32_t n;
  const float4* __restrict__ inTris;     // 3 per triangle, rank order: (v0, primitive) (v1, instance) (v2, opacity)
  unsigned long long* keys; uint32_t* vals;
  uint2* children;      // per internal node
  uint32_t* parent;     // per ref
  uint2* range;         // per internal node: first, last sorted position (Karras tree only)
  uint32_t* total;      // per internal node: triangles in the subtree
  float4* boxLo; float4* boxHi;   // per ref
  uint32_t* visit;      // per internal node
  LbvhDp* dp;           // per internal node
};

GTLC_D float lbvh_area(float4 lo, float4 hi)
{
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return 2.0f * ((dx * dy + dy * dz) + dz * dx);
}

// ---- flatten -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLbvhBlock) lbvh_flatten_kernel(uint32_t n, uint32_t numInst, const uint32_t* __restrict__ rankBase,
    const InstanceInfo* __restrict__ instInfo, const float4* __restrict__ o2w, const float4* __restrict__ fverts,
    const uint32_t* __restrict__ faces, const float* __restrict__ instOpacity, float4* __restrict__ outTris)
{
  const uint32_t rank = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (rank >= n) return;
  uint32_t lo = 0, hi = numInst;            // last instance with rankBase <= rank
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (rankBase[mid] <= rank) lo = mid; else hi = mid; }
  const uint32_t inst = lo, prim = rank - rankBase[inst];
  const InstanceInfo info = instInfo[inst];
  const float4 r0 = o2w[inst * 3 + 0], r1 = o2w[inst * 3 + 1], r2 = o2w[inst * 3 + 2];
  float4 rec[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const uint32_t vi = faces[(size_t) (info.faceBase + prim) * 3 + c];
    const float4 p = fverts[(size_t) (info.vertexBase + vi) * 2];
    rec[c].x = ((r0.x * p.x + r0.y * p.y) + r0.z * p.z) + r0.w;
    rec[c].y = ((r1.x * p.x + r1.y * p.y) + r1.z * p.z) + r1.w;
    rec[c].z = ((r2.x * p.x + r2.y * p.y) + r2.z * p.z) + r2.w;
  }
  rec[0].w = u2f(prim); rec[1].w = u2f(inst); rec[2].w = instOpacity[inst];
  outTris[(size_t) rank * 3 + 0] = rec[0]; outTris[(size_t) rank * 3 + 1] = rec[1]; outTris[(size_t) rank * 3 + 2] = rec[2];
}

// ---- centroid bounds ---------------------------------------------------------------------------------------
GTLC_D uint32_t lbvh_ordered(float f) { const uint32_t u = f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }   // monotone float -> uint
GTLC_D float lbvh_unordered(uint32_t u) { return u2f((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

GTLC_D void lbvh_tri_box(const float4* __restrict__ inTris, uint32_t tri, float4& lo, float4& hi)
{
  const float4 a = inTris[(size_t) tri * 3 + 0], b = inTris[(size_t) tri * 3 + 1], c = inTris[(size_t) tri * 3 + 2];
  lo = make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), 0.0f);
  hi = make_float4(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), 0.0f);
  // A triangle with a NaN / infinite / absurdly large coordinate is inactive (Vulkan's term: the watertight test can never
  // report it) but would poison the centroid bounds, the Morton keys and every box above it: it enters the tree as a point
  // at the origin, exactly as in the host builder (bvh8_build.cpp).  fmaxf / fminf drop NaNs, so the sum carries them.
  const float m = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(lo.y)), fabsf(lo.z)), fmaxf(fmaxf(fabsf(hi.x), fabsf(hi.y)), fabsf(hi.z)));
  const float sum = (((a.x + a.y) + a.z) + ((b.x + b.y) + b.z)) + ((c.x + c.y) + c.z);
  if (!(m <= 1.0e18f) || sum != sum) { lo = make_float4(0.0f, 0.0f, 0.0f, 0.0f); hi = lo; }
}

// bounds[0..2] = min, bounds[3..5] = max of the centroids, as ordered uints (initialised to 0xFFFFFFFF / 0)
__global__ void __launch_bounds__(kLbvhBlock) lbvh_bounds_kernel(uint32_t n, const float4* __restrict__ inTris, uint32_t* __restrict__ bounds)
{
  __shared__ uint32_t sMin[3], sMax[3];
  if (threadIdx.x < 3) { sMin[threadIdx.x] = 0xFFFFFFFFu; sMax[threadIdx.x] = 0u; }
  __syncthreads();
  for (uint32_t i = blockIdx.x * kLbvhBlock + threadIdx.x; i < n; i += gridDim.x * kLbvhBlock) {
    float4 lo, hi; lbvh_tri_box(inTris, i, lo, hi);
    const float c[3] = { 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z) };
#pragma unroll
    for (int k = 0; k < 3; k++) { atomicMin(&sMin[k], lbvh_ordered(c[k])); atomicMax(&sMax[k], lbvh_ordered(c[k])); }
  }
  __syncthreads();
  if (threadIdx.x < 3) { atomicMin(&bounds[threadIdx.x], sMin[threadIdx.x]); atomicMax(&bounds[3 + threadIdx.x], sMax[threadIdx.x]); }
}

// ---- Morton keys -------------------------------------------------------------------------------------------
GTLC_D unsigned long long lbvh_expand21(uint32_t v)
{
  unsigned long long x = v & 0x1FFFFFull;
  x = (x | (x << 32)) & 0x001F00000000FFFFull;
  x = (x | (x << 16)) & 0x001F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}
GTLC_D uint32_t lbvh_quantise(float c, float lo, float hi)
{
  const float ext = hi - lo;
  if (!(ext > 0.0f)) return 0u;
  const float x = (c - lo) / ext;
  const uint32_t q = (uint32_t) (x * 2097152.0f);
  return q > 2097151u ? 2097151u : q;
}
__global__ void __launch_bounds__(kLbvhBlock) lbvh_morton_kernel(uint32_t n, const float4* __restrict__ inTris, const uint32_t* __restrict__ bounds,
                                                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
  const uint32_t i = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (i >= n) return;
  float4 lo, hi; lbvh_tri_box(inTris, i, lo, hi);
  const uint32_t qx = lbvh_quantise(0.5f * (lo.x + hi.x), lbvh_unordered(bounds[0]), lbvh_unordered(bounds[3]));
  const uint32_t qy = lbvh_quantise(0.5f * (lo.y + hi.y), lbvh_unordered(bounds[1]), lbvh_unordered(bounds[4]));
  const uint32_t qz = lbvh_quantise(0.5f * (lo.z + hi.z), lbvh_unordered(bounds[2]), lbvh_unordered(bounds[5]));
  keys[i] = (lbvh_expand21(qx) << 2) | (lbvh_expand21(qy) << 1) | lbvh_expand21(qz);
  vals[i] = i;
}

// ---- stable 8-bit LSD radix pass over (key, value) pairs (same scheme as sort.cuh) ----------------------------
__global__ void __launch_bounds__(kLbvhBlock) lbvh_radix_count_kernel(uint32_t n, const unsigned long long* __restrict__ keys, int shift, uint32_t* __restrict__ hist)
{
  __shared__ uint32_t local[256];
  const uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
  local[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t i = begin + threadIdx.x; i < end; i += kLbvhBlock) atomicAdd(&local[(uint32_t) (keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  hist[threadIdx.x * gridDim.x + blockIdx.x] = local[threadIdx.x];
}

__global__ void __launch_bounds__(kLbvhBlock) lbvh_radix_scatter_kernel(uint32_t n, const unsigned long long* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
    int shift, const uint32_t* __restrict__ offsets, unsigned long long* __restrict__ keysOut, uint32_t* __restrict__ valsOut)
{
  __shared__ uint32_t base[256];
  __shared__ uint32_t warpCnt[kLbvhBlock / 32][256];
  const unsigned FULL = 0xFFFFFFFFu;
  const uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  base[threadIdx.x] = offsets[threadIdx.x * gridDim.x + blockIdx.x];
  for (uint32_t tile = begin; tile < end; tile += kLbvhBlock) {
#pragma unroll
    for (int w = 0; w < kLbvhBlock / 32; w++) warpCnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = tile + threadIdx.x;
    const bool valid = i < end;
    const unsigned long long key = valid ? keysIn[i] : 0ull;
    const uint32_t digit = valid ? ((uint32_t) (key >> shift) & 255u) : 0xFFFFu;
    const unsigned peers = __match_any_sync(FULL, digit);
    const uint32_t rankInWarp = (uint32_t) __popc(peers & ((1u << lane) - 1u));
    if (valid && rankInWarp == 0) warpCnt[warp][digit] = (uint32_t) __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t before = 0;
      for (int w = 0; w < warp; w++) before += warpCnt[w][digit];
      const uint32_t dst = base[digit] + before + rankInWarp;
      keysOut[dst] = key; valsOut[dst] = valsIn[i];
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kLbvhBlock / 32; w++) total += warpCnt[w][threadIdx.x];
    base[threadIdx.x] += total;
    __syncthreads();
  }
}

// exclusive scan of `count` uint32 in place, one CTA of 1024 threads; optionally stores the total
__global__ void __launch_bounds__(1024) lbvh_scan_kernel(uint32_t* __restrict__ a, uint32_t count, uint32_t* __restrict__ total)
{
  __shared__ uint32_t warpSums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < count; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < count ? a[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warpSums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warpSums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, d); if (lane >= d) w += y; }
      warpSums[lane] = w;
    }
    __syncthreads();
    const uint32_t prefix = carry + (warp ? warpSums[warp - 1] : 0u) + (x - v);
    if (i < count) a[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

// ---- Karras 2012 binary radix tree ---------------------------------------------------------------------------
GTLC_D int lbvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long) (a ^ b));
}

__global__ void __launch_bounds__(kLbvhBlock) lbvh_karras_kernel(LbvhTree T)
{
  const int n = (int) T.n;
  const int i = (int) (blockIdx.x * kLbvhBlock + threadIdx.x);
  if (i >= n - 1) return;
  const unsigned long long* keys = T.keys;
  const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2) if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lbvh_delta(keys, n, i, j);
  int s = 0, t = l;
  do { t = (t + 1) >> 1; if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
  const int gamma = i + s * d + (d < 0 ? -1 : 0);
  const int first = i < j ? i : j, last = i < j ? j : i;
  const uint32_t left = (first == gamma) ? (uint32_t) (n - 1 + gamma) : (uint32_t) gamma;
  const uint32_t right = (last == gamma + 1) ? (uint32_t) (n - 1 + gamma + 1) : (uint32_t) (gamma + 1);
  T.children[i] = make_uint2(left, right);
  T.range[i] = make_uint2((uint32_t) first, (uint32_t) last);
  T.total[i] = (uint32_t) (last - first + 1);
  T.parent[left] = (uint32_t) i; T.parent[right] = (uint32_t) i;
  T.visit[i] = 0;
  if (i == 0) T.parent[0] = 0xFFFFFFFFu;
}

GTLC_D uint32_t lbvh_subtree_tris(const LbvhTree& T, uint32_t ref)
{
  if (ref >= T.n - 1) return 1u;
  return T.total[ref];
}

// ---- PLOC (Meister & Bittner 2018): clusters in Morton order merge with their locally best partner ---------------
constexpr int kPlocRadius = 16;

__global__ void __launch_bounds__(kLbvhBlock) lbvh_leaf_box_kernel(LbvhTree T, uint32_t* __restrict__ clusters)
{
  const uint32_t p = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (p >= T.n) return;
  float4 lo, hi; lbvh_tri_box(T.inTris, T.vals[p], lo, hi);
  T.boxLo[T.n - 1 + p] = lo; T.boxHi[T.n - 1 + p] = hi;
  clusters[p] = T.n - 1 + p;
}

GTLC_D float lbvh_union_area(float4 alo, float4 ahi, float4 blo, float4 bhi)
{
  const float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
  const float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
  return lbvh_area(lo, hi);
}

// nearest neighbour within the radius; ties go to the smaller position
__global__ void __launch_bounds__(kLbvhBlock) ploc_nearest_kernel(LbvhTree T, uint32_t count, const uint32_t* __restrict__ clusters, uint32_t* __restrict__ nearest)
{
  const uint32_t i = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (i >= count) return;
  const uint32_t ci = clusters[i];
  const float4 lo = T.boxLo[ci], hi = T.boxHi[ci];
  // Candidates in order of distance along the Morton curve, and at equal distance the side of the pair (2k, 2k + 1) first:
  // with strict comparisons a TIE therefore goes to the closest position, and among coincident or otherwise equal boxes
  // (zero-area clusters, regular grids) even and odd neighbours choose each other.  Scanning the window left to right
  // instead made every such cluster pick the far left end: one mutual pair per round, a chain as deep as the cluster is large.
  float best = kLbvhInf; uint32_t bj = i;
  for (uint32_t d = 1; d <= (uint32_t) kPlocRadius; d++) {
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const bool right = ((i & 1u) == 0u) == (side == 0);
      if (right ? (i + d > count - 1u) : (d > i)) continue;
      const uint32_t j = right ? i + d : i - d;
      const uint32_t cj = clusters[j];
      const float a = lbvh_union_area(lo, hi, T.boxLo[cj], T.boxHi[cj]);
      if (a < best || bj == i) { best = a; bj = j; }
    }
  }
  nearest[i] = bj;
}

// merge[i] = 1 when i and nearest[i] chose each other and i is the left one; keep[i] = 0 for the right one
__global__ void __launch_bounds__(kLbvhBlock) ploc_flags_kernel(uint32_t count, const uint32_t* __restrict__ nearest, uint32_t* __restrict__ merge,
                                                                uint32_t* __restrict__ mergeScan, uint32_t* __restrict__ keepScan)
{
  const uint32_t i = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (i >= count) return;
  const uint32_t j = nearest[i];
  const bool mutual = j != i && nearest[j] == i;
  const uint32_t m = (mutual && i < j) ? 1u : 0u;
  merge[i] = m; mergeScan[i] = m;
  keepScan[i] = (mutual && i > j) ? 0u : 1u;
}

// Internal nodes get descending ids (the last merge, the root, is node 0, as the collapse expects).
__global__ void __launch_bounds__(kLbvhBlock) ploc_apply_kernel(LbvhTree T, uint32_t count, const uint32_t* __restrict__ clusters, const uint32_t* __restrict__ nearest,
    const uint32_t* __restrict__ merge, const uint32_t* __restrict__ mergeRank, const uint32_t* __restrict__ keepPos, uint32_t nextId, uint32_t* __restrict__ out)
{
  const uint32_t i = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (i >= count) return;
  const uint32_t j = nearest[i];
  const bool removed = j != i && nearest[j] == i && i > j;
  if (removed) return;
  uint32_t ref = clusters[i];
  if (merge[i]) {
    const uint32_t a = ref, b = clusters[j];
    const uint32_t id = nextId - mergeRank[i];
    const float4 alo = T.boxLo[a], ahi = T.boxHi[a], blo = T.boxLo[b], bhi = T.boxHi[b];
    T.boxLo[id] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
    T.boxHi[id] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
    T.children[id] = make_uint2(a, b);
    T.total[id] = lbvh_subtree_tris(T, a) + lbvh_subtree_tris(T, b);
    T.parent[a] = id; T.parent[b] = id;
    T.visit[id] = 0;
    if (id == 0) T.parent[0] = 0xFFFFFFFFu;
    ref = id;
  }
  out[keepPos[i]] = ref;
}

// ---- exclusive scan of up to 2^32 elements: tiles of 2048 per CTA, one CTA over the tile sums, add back --------------
constexpr int kScanTile = 2048;
__global__ void __launch_bounds__(1024) lbvh_scan_tiles_kernel(uint32_t* __restrict__ a, uint32_t count, uint32_t* __restrict__ tileSums)
{
  __shared__ uint32_t warpSums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i0 = blockIdx.x * kScanTile + threadIdx.x * 2u;
  const uint32_t v0 = i0 < count ? a[i0] : 0u, v1 = i0 + 1u < count ? a[i0 + 1u] : 0u;
  uint32_t x = v0 + v1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
  if (lane == 31) warpSums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warpSums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, d); if (lane >= d) w += y; }
    warpSums[lane] = w;
  }
  __syncthreads();
  const uint32_t excl = (warp ? warpSums[warp - 1] : 0u) + (x - (v0 + v1));
  if (i0 < count) a[i0] = excl;
  if (i0 + 1u < count) a[i0 + 1u] = excl + v0;
  if (threadIdx.x == 1023) tileSums[blockIdx.x] = warpSums[31];
}
__global__ void __launch_bounds__(kLbvhBlock) lbvh_scan_add_kernel(uint32_t* __restrict__ a, uint32_t count, const uint32_t* __restrict__ tileOffsets)
{
  const uint32_t i = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (i < count) a[i] += tileOffsets[i / kScanTile];
}

// ---- bottom-up: boxes + SAH table --------------------------------------------------------------------------
GTLC_D void lbvh_child_costs(const LbvhTree& T, uint32_t ref, float* cost)
{
  if (ref >= T.n - 1) {                        // single-triangle leaf: the same cost however many slots it is offered
    const float c = lbvh_area(__ldcg(&T.boxLo[ref]), __ldcg(&T.boxHi[ref]));
#pragma unroll
    for (int i = 0; i < 7; i++) cost[i] = c;
  } else {
    const float* src = T.dp[ref].cost;
#pragma unroll
    for (int i = 0; i < 7; i++) cost[i] = __ldcg(&src[i]);
  }
}

__global__ void __launch_bounds__(kLbvhBlock) lbvh_refit_kernel(LbvhTree T)
{
  const uint32_t p = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (p >= T.n) return;
  const uint32_t leafRef = T.n - 1 + p;
  float4 lo, hi; lbvh_tri_box(T.inTris, T.vals[p], lo, hi);
  T.boxLo[leafRef] = lo; T.boxHi[leafRef] = hi;
  if (T.n == 1) return;
  uint32_t cur = T.parent[leafRef];
  while (true) {
    __threadfence();
    if (atomicAdd(&T.visit[cur], 1u) == 0u) return;      // the sibling subtree is not finished yet
    __threadfence();
    const uint2 ch = T.children[cur];
    const float4 llo = __ldcg(&T.boxLo[ch.x]), lhi = __ldcg(&T.boxHi[ch.x]), rlo = __ldcg(&T.boxLo[ch.y]), rhi = __ldcg(&T.boxHi[ch.y]);
    const float4 blo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
    const float4 bhi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
    T.boxLo[cur] = blo; T.boxHi[cur] = bhi;
    const float area = lbvh_area(blo, bhi);
    const uint32_t total = T.total[cur];
    const float leafCost = (total <= kLbvhMaxLeaf) ? area * (float) total : kLbvhInf;
    float L[7], R[7];
    lbvh_child_costs(T, ch.x, L); lbvh_child_costs(T, ch.y, R);
    LbvhDp d;
    float best = kLbvhInf; uint8_t bk = 1;
    for (int k = 1; k <= 7; k++) { const float c = L[k - 1] + R[8 - k - 1]; if (c < best) { best = c; bk = (uint8_t) k; } }
    const float internalCost = best + area * kLbvhNodeCost;
    d.rootSplit = bk; d.rootIsLeaf = leafCost <= internalCost ? 1 : 0;
    d.cost[0] = d.rootIsLeaf ? leafCost : internalCost; d.split[0] = 0;
    for (int i = 2; i <= 7; i++) {
      float bd = kLbvhInf; uint8_t bs = 0;
      for (int k = 1; k < i; k++) { const float c = L[k - 1] + R[i - k - 1]; if (c < bd) { bd = c; bs = (uint8_t) k; } }
      if (bd < d.cost[i - 2]) { d.cost[i - 1] = bd; d.split[i - 1] = bs; }
      else { d.cost[i - 1] = d.cost[i - 2]; d.split[i - 1] = 0; }
    }
    d.pad[0] = d.pad[1] = d.pad[2] = 0;
    T.dp[cur] = d;
    if (cur == 0) return;
    cur = T.parent[cur];
  }
}

// ---- breadth-first collapse ----------------------------------------------------------------------------------
struct LbvhLevel {
  uint32_t count;                        // wide nodes on this level
  const uint2* __restrict__ queue;       // (binary ref, wide node index)
  uint2* __restrict__ nextQueue;
  uint32_t* __restrict__ childList;      // 8 per entry: ref | leaf << 31
  uint32_t* __restrict__ childCount;     // per entry
  uint32_t* __restrict__ cntNodes;       // per entry -> exclusive scan
  uint32_t* __restrict__ cntTris;        // per entry -> exclusive scan
  uint32_t nodeBase, triBase;            // first free wide node / triangle index before this level's children
};


// expands "subtree `ref` as at most `i` children" (bvh8_build.cpp's gather) with an explicit stack, left first
GTLC_D void lbvh_gather(const LbvhTree& T, uint32_t ref, int i, uint32_t* list, int& n)
{
  uint32_t stRef[16]; int stI[16]; int sp = 0;
  stRef[sp] = ref; stI[sp] = i; sp++;
  while (sp > 0) {
    sp--;
    const uint32_t r = stRef[sp]; int k = stI[sp];
    if (r >= T.n - 1) { list[n++] = r | 0x80000000u; continue; }
    const LbvhDp& d = T.dp[r];
    while (k > 1 && d.split[k - 1] == 0) k--;
    if (k == 1) { list[n++] = r | (d.rootIsLeaf ? 0x80000000u : 0u); continue; }
    const int s = d.split[k - 1];
    const uint2 ch = T.children[r];
    stRef[sp] = ch.y; stI[sp] = k - s; sp++;
    stRef[sp] = ch.x; stI[sp] = s; sp++;
  }
}

__global__ void __launch_bounds__(kLbvhBlock) lbvh_collapse_count_kernel(LbvhTree T, LbvhLevel Lv)
{
  const uint32_t e = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (e >= Lv.count) return;
  const uint32_t ref = Lv.queue[e].x;
  uint32_t list[8]; int n = 0;
  if (ref >= T.n - 1) list[n++] = ref | 0x80000000u;                     // one-triangle scene
  else if (T.dp[ref].rootIsLeaf) list[n++] = ref | 0x80000000u;          // <= kLbvhMaxLeaf triangles in total
  else {
    const uint2 ch = T.children[ref];
    const int k = T.dp[ref].rootSplit;
    lbvh_gather(T, ch.x, k, list, n);
    lbvh_gather(T, ch.y, 8 - k, list, n);
  }
  uint32_t nodes = 0, tris = 0;
  for (int i = 0; i < 8; i++) {
    Lv.childList[(size_t) e * 8 + i] = i < n ? list[i] : 0xFFFFFFFFu;
    if (i < n) { if (list[i] & 0x80000000u) tris += lbvh_subtree_tris(T, list[i] & 0x7FFFFFFFu); else nodes++; }
  }
  Lv.childCount[e] = (uint32_t) n;
  Lv.cntNodes[e] = nodes; Lv.cntTris[e] = tris;
}

GTLC_D double lbvh_pow2(int e) { return __longlong_as_double((long long) (e + 1023) << 52); }
GTLC_D float lbvh_next_below(float f)
{
  const uint32_t u = f2u(f);
  if ((u & 0x7FFFFFFFu) == 0u) return u2f(0x80000001u);
  return (u & 0x80000000u) ? u2f(u + 1u) : u2f(u - 1u);
}

__global__ void __launch_bounds__(kLbvhBlock) lbvh_collapse_emit_kernel(LbvhTree T, LbvhLevel Lv, uint4* __restrict__ outNodes, float4* __restrict__ outTris)
{
  const uint32_t e = blockIdx.x * kLbvhBlock + threadIdx.x;
  if (e >= Lv.count) return;
  const uint2 entry = Lv.queue[e];
  const int n = (int) Lv.childCount[e];
  uint32_t ch[8];
  for (int i = 0; i < 8; i++) ch[i] = Lv.childList[(size_t) e * 8 + i];
  const float4 plo = T.boxLo[entry.x], phi = T.boxHi[entry.x];
  // ---- octant-aware slot assignment (greedy; same rule as bvh8_build.cpp) ----
  const float pc[3] = { 0.5f * (plo.x + phi.x), 0.5f * (plo.y + phi.y), 0.5f * (plo.z + phi.z) };
  float cost[8][8];
  for (int i = 0; i < n; i++) {
    const uint32_t r = ch[i] & 0x7FFFFFFFu;
    const float4 lo = T.boxLo[r], hi = T.boxHi[r];
    const float d0 = 0.5f * (lo.x + hi.x) - pc[0], d1 = 0.5f * (lo.y + hi.y) - pc[1], d2 = 0.5f * (lo.z + hi.z) - pc[2];
    for (int s = 0; s < 8; s++) cost[i][s] = (((s & 1) ? d0 : -d0) + ((s & 2) ? d1 : -d1)) + ((s & 4) ? d2 : -d2);
  }
  int childAtSlot[8]; for (int s = 0; s < 8; s++) childAtSlot[s] = -1;
  uint32_t childDone = 0, slotUsed = 0;
  for (int k = 0; k < n; k++) {
    int bi = -1, bs = -1; float bc = -kLbvhInf;
    for (int i = 0; i < n; i++) {
      if (childDone & (1u << i)) continue;
      for (int s = 0; s < 8; s++) {
        if (slotUsed & (1u << s)) continue;
        if (cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; }
      }
    }
    if (bi < 0) {                                  // only with NaN bounds: first free pair
      for (int i = 0; i < n && bi < 0; i++) if (!(childDone & (1u << i))) bi = i;
      for (int s = 0; s < 8 && bs < 0; s++) if (!(slotUsed & (1u << s))) bs = s;
    }
    childAtSlot[bs] = bi; childDone |= 1u << bi; slotUsed |= 1u << bs;
  }
  // ---- quantisation grid: origin one quantum below the lower corner, at most 250 quanta of extent ----
  const float ploA[3] = { plo.x, plo.y, plo.z }, phiA[3] = { phi.x, phi.y, phi.z };
  uint32_t ebits[3]; double scale[3]; float porigin[3];
  for (int k = 0; k < 3; k++) {
    const double ext = (double) phiA[k] - (double) ploA[k];
    int ex = -100;
    if (ext > 0.0) {
      ex = (int) ((__double_as_longlong(ext) >> 52) & 0x7FF) - 1023 - 8;
      if (ex < -100) ex = -100;
      while (ex < 100 && ext > 250.0 * lbvh_pow2(ex)) ex++;
    }
    ebits[k] = (uint32_t) (ex + 127);
    scale[k] = lbvh_pow2(ex);
    float po = (float) ((double) ploA[k] - scale[k]);
    while ((double) po > (double) ploA[k] - 0.75 * scale[k]) po = lbvh_next_below(po);
    porigin[k] = po;
  }
  uint32_t q[6][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };   // 8 bytes each, slot s = byte s
  uint32_t meta[2] = { 0, 0 };
  uint32_t imask = 0, nInternal = 0, triOffset = 0;
  const uint32_t childBase = Lv.nodeBase + Lv.cntNodes[e];
  const uint32_t triBase = Lv.triBase + Lv.cntTris[e];
  for (int s = 0; s < 8; s++) {
    const int i = childAtSlot[s];
    if (i < 0) continue;
    const uint32_t r = ch[i] & 0x7FFFFFFFu;
    const float4 lo4 = T.boxLo[r], hi4 = T.boxHi[r];
    const float cl[3] = { lo4.x, lo4.y, lo4.z }, chh[3] = { hi4.x, hi4.y, hi4.z };
    for (int k = 0; k < 3; k++) {
      const double lo = ((double) cl[k] - (double) porigin[k]) / scale[k];
      const double hi = ((double) chh[k] - (double) porigin[k]) / scale[k];
      double ql = floor(lo), qh = ceil(hi);
      if (lo - ql < 1.0 / 64.0 && ql > 0.0) ql -= 1.0;
      if (qh - hi < 1.0 / 64.0 && qh < 255.0) qh += 1.0;
      if (!(ql >= 0.0)) ql = 0.0;
      if (!(qh <= 255.0)) qh = 255.0;
      q[k][s >> 2] |= (uint32_t) ql << (8 * (s & 3));
      q[3 + k][s >> 2] |= (uint32_t) qh << (8 * (s & 3));
    }
    if (ch[i] & 0x80000000u) {
      const uint32_t cnt = lbvh_subtree_tris(T, r);
      meta[s >> 2] |= (((((1u << cnt) - 1u) << 5) | triOffset) & 0xFFu) << (8 * (s & 3));
      uint32_t walk[4]; int wp = 0; uint32_t t = 0;                   // the <= kLbvhMaxLeaf leaves of the subtree, left to right
      walk[wp++] = r;
      while (wp > 0) {
        const uint32_t x = walk[--wp];
        if (x >= T.n - 1) {
          const uint32_t tri = T.vals[x - (T.n - 1)];
          const size_t dst = (size_t) (triBase + triOffset + t) * 3;
          outTris[dst + 0] = T.inTris[(size_t) tri * 3 + 0]; outTris[dst + 1] = T.inTris[(size_t) tri * 3 + 1]; outTris[dst + 2] = T.inTris[(size_t) tri * 3 + 2];
          t++;
        } else { const uint2 c2 = T.children[x]; walk[wp++] = c2.y; walk[wp++] = c2.x; }
      }
      triOffset += cnt;
    } else {
      meta[s >> 2] |= (0x20u | (24u + (uint32_t) s)) << (8 * (s & 3));
      imask |= 1u << s;
      Lv.nextQueue[Lv.cntNodes[e] + nInternal] = make_uint2(r, childBase + nInternal);
      nInternal++;
    }
  }
  uint4* dst = outNodes + (size_t) entry.y * 5;
  dst[0] = make_uint4(f2u(porigin[0]), f2u(porigin[1]), f2u(porigin[2]), ebits[0] | (ebits[1] << 8) | (ebits[2] << 16) | (imask << 24));
  dst[1] = make_uint4(childBase, triBase, meta[0], meta[1]);
  dst[2] = make_uint4(q[0][0], q[0][1], q[1][0], q[1][1]);
  dst[3] = make_uint4(q[2][0], q[2][1], q[3][0], q[3][1]);
  dst[4] = make_uint4(q[4][0], q[4][1], q[5][0], q[5][1]);
}

} // namespace gtlc
