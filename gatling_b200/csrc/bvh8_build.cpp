// Host builder of the compressed BVH8 (see bvh8.h).  Binned-SAH binary BVH (OpenMP tasks) ->
// greedy surface-area collapse to 8-wide nodes -> octant-aware slot assignment -> quantisation.
// Stands in for the Vulkan driver's PREFER_FAST_TRACE build the reference requests
// (/root/reference/src/cgpu/impl/CgpuVk.cpp:2575); a GPU LBVH/PLOC builder is the "next" row 1 of
// SURVEY.md 8f.
#include "bvh8.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <emmintrin.h>

namespace gtlc {
namespace {

struct Box {
  float lo[3], hi[3];
  void reset() { for (int k = 0; k < 3; k++) { lo[k] = 3.402823466e38f; hi[k] = -3.402823466e38f; } }
  void grow(const float* p) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
  void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
  float area() const
  {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0.0f) return 0.0f;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
  }
};

struct BinNode { Box box; uint32_t left, right, first, count, total; };   // count > 0: binary leaf; total = triangles in the subtree

// SAH-optimal collapse to 8-wide nodes (Ylitie, Karras, Laine 2017, section 3.1) as a dynamic programme over the binary
// tree: cost[i-1] = cheapest way to present this subtree to its parent as at most i children (i = 1..7), each child
// either a leaf of <= kMaxLeaf triangles or a wide node; split[i-1] = how many of them the left subtree gets (0 = same
// as with one child fewer); rootSplit / rootIsLeaf describe the subtree as ONE child.
struct Dp { float cost[7]; uint8_t split[7]; uint8_t rootSplit; uint8_t rootIsLeaf; };

constexpr int kBins = 16;
constexpr uint32_t kSmallNode = 4;           // nodes up to this size pick their split without bin arrays
constexpr uint32_t kParallelRange = 262144;   // ranges at least this long are scanned by several tasks
constexpr uint32_t kTaskRange = 4096;          // subtrees larger than this become OpenMP tasks (a few hundred tasks per million triangles)
// SAH parameters (overridable through GTLC_BVH_MAX_LEAF / GTLC_BVH_NODE_COST for tuning experiments)
uint32_t kMaxLeaf = 3;
float kNodeCost = 0.6f;  // relative to one triangle test
float kBigTriFraction = 0.4f;   // GTLC_BVH_BIG_TRIS: a triangle is "big" when its extent sum exceeds this fraction of the node's (0 = off)
constexpr float kTriCost = 1.0f;
constexpr float kMaxCoordinate = 1.0e18f;   // beyond this, surface areas leave the float range

// One record per triangle, moved (not indexed) by the partition passes so that every pass of the builder streams
// through memory: bounds in the first three lanes of two SSE registers, the triangle's index in the fourth lane of `lo`.
struct alignas(16) Prim { float lo[3]; uint32_t id; float hi[3]; uint32_t pad; };

// lower corner with the index lane cleared (an integer reinterpreted as float is a denormal: slow in arithmetic)
inline __m128 loadLo(const Prim& p) { return _mm_and_ps(_mm_load_ps(p.lo), _mm_castsi128_ps(_mm_setr_epi32(-1, -1, -1, 0))); }

// Box::area() on registers: the same products and the same order of additions
inline float area4(__m128 lo, __m128 hi)
{
  const __m128 d = _mm_sub_ps(hi, lo);
  if (_mm_cvtss_f32(d) < 0.0f) return 0.0f;
  const __m128 p = _mm_mul_ps(d, _mm_shuffle_ps(d, d, _MM_SHUFFLE(3, 0, 2, 1)));   // (dx dy, dy dz, dz dx, .)
  const float p0 = _mm_cvtss_f32(p), p1 = _mm_cvtss_f32(_mm_shuffle_ps(p, p, 1)), p2 = _mm_cvtss_f32(_mm_shuffle_ps(p, p, 2));
  return 2.0f * ((p0 + p1) + p2);
}

struct Bounds4 {           // box of the triangles and box of their centroids, as SSE registers (lane 3 unused)
  __m128 blo, bhi, clo, chi;
  void reset()
  {
    blo = clo = _mm_set1_ps(3.402823466e38f);
    bhi = chi = _mm_set1_ps(-3.402823466e38f);
  }
  void grow(const Prim& p)
  {
    const __m128 lo = loadLo(p), hi = _mm_load_ps(p.hi);
    const __m128 c = _mm_mul_ps(_mm_set1_ps(0.5f), _mm_add_ps(lo, hi));
    blo = _mm_min_ps(blo, lo); bhi = _mm_max_ps(bhi, hi);
    clo = _mm_min_ps(clo, c);  chi = _mm_max_ps(chi, c);
  }
  void grow(const Bounds4& o)
  {
    blo = _mm_min_ps(blo, o.blo); bhi = _mm_max_ps(bhi, o.bhi);
    clo = _mm_min_ps(clo, o.clo); chi = _mm_max_ps(chi, o.chi);
  }
  static Box toBox(__m128 lo, __m128 hi)
  {
    alignas(16) float a[4], b[4];
    _mm_store_ps(a, lo); _mm_store_ps(b, hi);
    Box r;
    for (int k = 0; k < 3; k++) { r.lo[k] = a[k]; r.hi[k] = b[k]; }
    return r;
  }
  Box box() const { return toBox(blo, bhi); }
  Box centroidBox() const { return toBox(clo, chi); }
};

struct Builder {
  RawVector<Prim> prims;
  RawVector<BinNode> nodes;
  RawVector<Dp> dp;                      // filled bottom-up as the recursion unwinds: a node's row right after its children's
  uint32_t depthCap = 0;                 // 0 = plain SAH; else no binary leaf below this depth (median splits where needed)
  static uint32_t ceilLog2(uint32_t c) { return c <= 1u ? 0u : 32u - (uint32_t) __builtin_clz(c - 1u); }

  // The top of the tree is a handful of huge ranges: their passes are cut into chunks that run as tasks of the team
  // (min / max / counts merge exactly, so the tree does not depend on the chunking).
  static uint32_t chunksOf(uint32_t count) { return count >= kParallelRange ? 64u : 1u; }

  Bounds4 boundsOf(uint32_t begin, uint32_t end)
  {
    const uint32_t count = end - begin, chunks = chunksOf(count), chunkLen = (count + chunks - 1) / chunks;
    Bounds4 r; r.reset();
    if (chunks == 1) {
      for (uint32_t i = begin; i < end; i++) r.grow(prims[i]);
      return r;
    }
    std::vector<Bounds4> part(chunks);
#pragma omp taskloop default(shared) grainsize(1)
    for (uint32_t c = 0; c < chunks; c++) {
      Bounds4 a; a.reset();
      const uint32_t e0 = begin + c * chunkLen, e1 = std::min(end, e0 + chunkLen);
      for (uint32_t i = e0; i < e1; i++) a.grow(prims[i]);
      part[c] = a;
    }
    for (uint32_t c = 0; c < chunks; c++) r.grow(part[c]);
    return r;
  }

  struct Bins {
    __m128 lo[3][kBins], hi[3][kBins];
    uint32_t count[3][kBins];
    uint32_t big;
    void reset()
    {
      for (int a = 0; a < 3; a++) for (int b = 0; b < kBins; b++) { lo[a][b] = _mm_set1_ps(3.402823466e38f); hi[a][b] = _mm_set1_ps(-3.402823466e38f); count[a][b] = 0; }
      big = 0;
    }
    void merge(const Bins& o)
    {
      for (int a = 0; a < 3; a++) for (int b = 0; b < kBins; b++) {
        lo[a][b] = _mm_min_ps(lo[a][b], o.lo[a][b]); hi[a][b] = _mm_max_ps(hi[a][b], o.hi[a][b]); count[a][b] += o.count[a][b];
      }
      big += o.big;
    }
  };

  // one pass over [e0, e1): all three axes binned at once, big triangles counted on the way
  void binRange(uint32_t e0, uint32_t e1, __m128 cLo, __m128 scale, float bigLimit, Bins& bins) const
  {
    for (uint32_t i = e0; i < e1; i++) {
      const Prim& p = prims[i];
      const __m128 lo = loadLo(p), hi = _mm_load_ps(p.hi);
      const __m128 c = _mm_mul_ps(_mm_set1_ps(0.5f), _mm_add_ps(lo, hi));
      alignas(16) int b[4]; alignas(16) float d[4];
      _mm_store_si128((__m128i*) b, _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(c, cLo), scale)));
      _mm_store_ps(d, _mm_sub_ps(hi, lo));
      for (int a = 0; a < 3; a++) {
        const int k = b[a] < 0 ? 0 : (b[a] >= kBins ? kBins - 1 : b[a]);
        bins.lo[a][k] = _mm_min_ps(bins.lo[a][k], lo); bins.hi[a][k] = _mm_max_ps(bins.hi[a][k], hi); bins.count[a][k]++;
      }
      bins.big += ((d[0] + d[1]) + d[2]) > bigLimit ? 1u : 0u;
    }
  }

  // std::partition's result for a huge range, computed by the whole team.  The two-pointer scheme swaps the k-th misplaced
  // record from the left (predicate false, position before the final boundary) with the k-th misplaced record from the
  // right (predicate true, position at or after it) and touches nothing else; so: flag everything, count, list the two
  // kinds of misplaced positions with prefix sums, swap the pairs.  Also returns the bounds of the two sides.
  template <class Pred>
  uint32_t partitionChunked(uint32_t begin, uint32_t end, const Pred& goesLeft, Bounds4& lb, Bounds4& rb)
  {
    const uint32_t count = end - begin, chunks = chunksOf(count), chunkLen = (count + chunks - 1) / chunks;
    struct Part { Bounds4 side[2]; uint32_t trues, misplacedLeft, misplacedRight; };
    std::vector<Part> part(chunks);
    std::vector<uint8_t> flag(count);
#pragma omp taskloop default(shared) grainsize(1)
    for (uint32_t c = 0; c < chunks; c++) {
      Part& pt = part[c];
      pt.side[0].reset(); pt.side[1].reset(); pt.trues = 0;
      const uint32_t e0 = std::min(end, begin + c * chunkLen), e1 = std::min(end, e0 + chunkLen);
      for (uint32_t i = e0; i < e1; i++) {
        const bool f = goesLeft(prims[i]);
        flag[i - begin] = f ? 1 : 0;
        pt.side[f ? 0 : 1].grow(prims[i]);
        pt.trues += f ? 1u : 0u;
      }
    }
    uint32_t trues = 0;
    for (uint32_t c = 0; c < chunks; c++) { trues += part[c].trues; lb.grow(part[c].side[0]); rb.grow(part[c].side[1]); }
    const uint32_t mid = begin + trues;
#pragma omp taskloop default(shared) grainsize(1)
    for (uint32_t c = 0; c < chunks; c++) {
      const uint32_t e0 = std::min(end, begin + c * chunkLen), e1 = std::min(end, e0 + chunkLen);
      uint32_t nl = 0, nr = 0;
      for (uint32_t i = e0; i < std::min(e1, mid); i++) nl += 1u - flag[i - begin];
      for (uint32_t i = std::max(e0, mid); i < e1; i++) nr += flag[i - begin];
      part[c].misplacedLeft = nl; part[c].misplacedRight = nr;
    }
    uint32_t swaps = 0, swapsRight = 0;
    for (uint32_t c = 0; c < chunks; c++) {
      const uint32_t nl = part[c].misplacedLeft, nr = part[c].misplacedRight;
      part[c].misplacedLeft = swaps; part[c].misplacedRight = swapsRight;
      swaps += nl; swapsRight += nr;
    }
    std::vector<uint32_t> fromLeft(swaps), fromRight(swaps);     // both ascending; swapsRight == swaps
#pragma omp taskloop default(shared) grainsize(1)
    for (uint32_t c = 0; c < chunks; c++) {
      const uint32_t e0 = std::min(end, begin + c * chunkLen), e1 = std::min(end, e0 + chunkLen);
      uint32_t kl = part[c].misplacedLeft, kr = part[c].misplacedRight;
      for (uint32_t i = e0; i < std::min(e1, mid); i++) if (!flag[i - begin]) fromLeft[kl++] = i;
      for (uint32_t i = std::max(e0, mid); i < e1; i++) if (flag[i - begin]) fromRight[kr++] = i;
    }
    const uint32_t swapChunk = (swaps + chunks - 1) / chunks;
#pragma omp taskloop default(shared) grainsize(1)
    for (uint32_t c = 0; c < chunks; c++) {
      const uint32_t k0 = std::min(swaps, c * swapChunk), k1 = std::min(swaps, k0 + swapChunk);
      for (uint32_t k = k0; k < k1; k++) std::swap(prims[fromLeft[k]], prims[fromRight[swaps - 1 - k]]);
    }
    return mid;
  }

  // The same permutation by one thread, without a data-dependent branch per record (the predicate is a coin flip for
  // the branch predictor): pass 1 flags every record and grows the bounds of its side through an indexed accumulator,
  // pass 2 lists the misplaced positions of both sides with unconditional stores, then the pairs are swapped.
  template <class Pred>
  uint32_t partitionSerial(uint32_t begin, uint32_t end, const Pred& goesLeft, Bounds4& lb, Bounds4& rb)
  {
    const uint32_t count = end - begin;
    static thread_local std::vector<uint8_t> flagScratch;
    static thread_local std::vector<uint32_t> listScratch;
    uint8_t flagSmall[256]; uint32_t listSmall[512];
    uint8_t* flag = flagSmall; uint32_t* fromLeft = listSmall;
    if (count > 256) {
      if (flagScratch.size() < count) { flagScratch.resize(count); listScratch.resize((size_t) 2 * count); }
      flag = flagScratch.data(); fromLeft = listScratch.data();
    }
    uint32_t* fromRight = fromLeft + count;
    Bounds4 side[2]; side[0].reset(); side[1].reset();
    uint32_t trues = 0;
    Prim* base = prims.data() + begin;
    for (uint32_t i = 0; i < count; i++) {
      const uint32_t f = goesLeft(base[i]) ? 1u : 0u;
      flag[i] = (uint8_t) f; trues += f;
      side[1u - f].grow(base[i]);
    }
    lb = side[0]; rb = side[1];
    uint32_t nl = 0, nr = 0;
    for (uint32_t i = 0; i < trues; i++) { fromLeft[nl] = i; nl += 1u - flag[i]; }
    for (uint32_t i = count; i-- > trues;) { fromRight[nr] = i; nr += flag[i]; }
    for (uint32_t k = 0; k < nl; k++) std::swap(base[fromLeft[k]], base[fromRight[k]]);
    return begin + trues;
  }

  // the collapse table row of `ni`; both children's rows are complete (post-order)
  void solveNode(uint32_t ni)
  {
    const float INF = 3.402823466e38f;
    const BinNode& n = nodes[ni];
    Dp& d = dp[ni];
    const float area = n.box.area();
    const float leafCost = (n.total <= kMaxLeaf) ? area * (float) n.total * kTriCost : INF;
    if (n.count > 0) {                       // binary leaf (single triangle)
      for (int i = 0; i < 7; i++) { d.cost[i] = leafCost; d.split[i] = 0; }
      d.rootSplit = 0; d.rootIsLeaf = 1;
      return;
    }
    const Dp& L = dp[n.left]; const Dp& R = dp[n.right];
    // as one wide node: distribute 8 child slots between the two binary children
    float best = INF; uint8_t bk = 1;
    for (int k = 1; k <= 7; k++) { const float c = L.cost[k - 1] + R.cost[8 - k - 1]; if (c < best) { best = c; bk = (uint8_t) k; } }
    const float internalCost = best + area * kNodeCost;
    d.rootSplit = bk; d.rootIsLeaf = n.total <= kMaxLeaf && leafCost <= internalCost;
    d.cost[0] = d.rootIsLeaf ? leafCost : internalCost; d.split[0] = 0;
    for (int i = 2; i <= 7; i++) {            // as at most i children of the parent
      float bd = INF; uint8_t bs = 0;
      for (int k = 1; k < i; k++) { const float c = L.cost[k - 1] + R.cost[i - k - 1]; if (c < bd) { bd = c; bs = (uint8_t) k; } }
      if (bd < d.cost[i - 2]) { d.cost[i - 1] = bd; d.split[i - 1] = bs; }
      else { d.cost[i - 1] = d.cost[i - 2]; d.split[i - 1] = 0; }   // 0 = same as with one child fewer
    }
  }

  // Node ids are the pre-order of the binary tree: every leaf holds one triangle, so the subtree over `c` triangles has
  // exactly 2c - 1 nodes and the right child of `node` sits 2 (mid - begin) after it.  No shared counter, and a subtree is
  // one contiguous block for the collapse and emission passes that walk it afterwards.
  void split(uint32_t node, uint32_t begin, uint32_t mid, uint32_t end, const Bounds4& lb, const Bounds4& rb, bool leftInline, uint32_t depth)
  {
    BinNode& n = nodes[node];
    const uint32_t l = node + 1, r = node + 2 * (mid - begin);
    n.left = l; n.right = r; n.count = 0; n.first = begin; n.total = end - begin;
    if (end - begin > kTaskRange) {
      if (leftInline) build(l, begin, mid, lb, depth + 1);
      else {
#pragma omp task default(shared) firstprivate(l, begin, mid, lb, depth)
        build(l, begin, mid, lb, depth + 1);
      }
#pragma omp task default(shared) firstprivate(r, mid, end, rb, depth)
      build(r, mid, end, rb, depth + 1);
#pragma omp taskwait
    } else {
      build(l, begin, mid, lb, depth + 1);
      build(r, mid, end, rb, depth + 1);
    }
    solveNode(node);
  }

  // `bounds` = box of the triangles in [begin, end) and of their centroids, handed down by whoever partitioned them
  void build(uint32_t node, uint32_t begin, uint32_t end, const Bounds4& bounds, uint32_t depth)
  {
    const uint32_t count = end - begin;
    const uint32_t chunks = chunksOf(count), chunkLen = (count + chunks - 1) / chunks;
    const Box box = bounds.box(), cbox = bounds.centroidBox();
    BinNode& n = nodes[node];
    n.box = box;
    auto makeLeaf = [&]() { n.first = begin; n.count = count; n.total = count; n.left = n.right = 0; };
    if (count == 1) { makeLeaf(); solveNode(node); return; }
    // Depth-capped rebuild (bvh8_build's second attempt, only after a first tree came out deeper than the traversal stack):
    // once the levels left under the cap are just enough for a balanced subtree, split by index.
    if (depthCap && depth + ceilLog2(count) >= depthCap) {
      const uint32_t half = begin + count / 2;
      split(node, begin, half, end, boundsOf(begin, half), boundsOf(half, end), false, depth);
      return;
    }

    float scaleA[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    bool axisOk[3];
    for (int axis = 0; axis < 3; axis++) {
      const float ext = cbox.hi[axis] - cbox.lo[axis];
      axisOk[axis] = ext > 0.0f;
      if (axisOk[axis]) scaleA[axis] = (float) kBins / ext;
    }
    const __m128 cLo = _mm_setr_ps(cbox.lo[0], cbox.lo[1], cbox.lo[2], 0.0f), scale = _mm_loadu_ps(scaleA);
    float bestCost = 3.402823466e38f; int bestAxis = -1; int bestBin = -1;
    // Zero-area clusters (point triangles, axis-aligned collinear segments) make EVERY candidate cost 0; the strict
    // comparison alone would then keep the lowest occupied bin and peel 1/16 of the range per level (depth in the
    // hundreds).  Among zero-cost candidates the most balanced one wins; non-zero ties keep the first candidate.
    uint32_t bestSkew = 0xFFFFFFFFu;
    auto better = [&](float cost, uint32_t leftCount) {
      if (cost < bestCost) { bestSkew = leftCount > count - leftCount ? 2 * leftCount - count : count - 2 * leftCount; return true; }
      if (cost == 0.0f && bestCost == 0.0f) {
        const uint32_t skew = leftCount > count - leftCount ? 2 * leftCount - count : count - 2 * leftCount;
        if (skew < bestSkew) { bestSkew = skew; return true; }
      }
      return false;
    };

    if (count <= kSmallNode) {
      // ---- the bottom of the tree (most nodes hold a handful of triangles): no bin arrays.  A split after an empty bin
      // separates the same two sets as the split after the occupied bin below it and loses the strict comparison, so only
      // the occupied bins are candidates; their two boxes are gathered directly (min / max are exact in any order). ----
      uint8_t binOf[3][kSmallNode];
      uint32_t occupied[3] = { 0u, 0u, 0u };
      for (uint32_t i = 0; i < count; i++) {
        const Prim& p = prims[begin + i];
        const __m128 c = _mm_mul_ps(_mm_set1_ps(0.5f), _mm_add_ps(loadLo(p), _mm_load_ps(p.hi)));
        alignas(16) int b[4];
        _mm_store_si128((__m128i*) b, _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(c, cLo), scale)));
        for (int a = 0; a < 3; a++) {
          const int k = b[a] < 0 ? 0 : (b[a] >= kBins ? kBins - 1 : b[a]);
          binOf[a][i] = (uint8_t) k; occupied[a] |= 1u << k;
        }
      }
      for (int axis = 0; axis < 3; axis++) {
        if (!axisOk[axis]) continue;
        uint32_t m = occupied[axis];
        while (m & (m - 1)) {                         // every occupied bin but the highest
          const int b = __builtin_ctz(m); m &= m - 1;
          __m128 llo = _mm_set1_ps(3.402823466e38f), lhi = _mm_set1_ps(-3.402823466e38f), rlo = llo, rhi = lhi;
          uint32_t cl = 0;
          for (uint32_t i = 0; i < count; i++) {
            const Prim& p = prims[begin + i];
            const __m128 lo = loadLo(p), hi = _mm_load_ps(p.hi);
            if (binOf[axis][i] <= b) { llo = _mm_min_ps(llo, lo); lhi = _mm_max_ps(lhi, hi); cl++; }
            else { rlo = _mm_min_ps(rlo, lo); rhi = _mm_max_ps(rhi, hi); }
          }
          const float cost = area4(llo, lhi) * (float) cl + area4(rlo, rhi) * (float) (count - cl);
          if (better(cost, cl)) { bestCost = cost; bestAxis = axis; bestBin = b; }
        }
      }
    } else {
      // ---- binning pass (three axes at once) ----
      const float sd = (box.hi[0] - box.lo[0]) + (box.hi[1] - box.lo[1]) + (box.hi[2] - box.lo[2]);
      const bool wantBig = count >= 64 && kBigTriFraction > 0.0f;
      const float bigLimit = wantBig ? kBigTriFraction * sd : 3.402823466e38f;
      Bins bins; bins.reset();
      if (chunks == 1) binRange(begin, end, cLo, scale, bigLimit, bins);
      else {
        std::vector<Bins> part(chunks);
#pragma omp taskloop default(shared) grainsize(1)
        for (uint32_t c = 0; c < chunks; c++) {
          part[c].reset();
          const uint32_t e0 = begin + c * chunkLen, e1 = std::min(end, e0 + chunkLen);
          binRange(e0, e1, cLo, scale, bigLimit, part[c]);
        }
        for (uint32_t c = 0; c < chunks; c++) bins.merge(part[c]);
      }

      // Triangles that span most of this node (room walls around furniture, a table top over what stands on it) go into a
      // subtree of their own: sorted by centroid they would sit deep inside one child and inflate every box on the way.
      // Measured in BVH8 node visits per ray (oracle walker): kitchen stand-in 6.59 -> 5.03, nothing lost elsewhere.
      if (wantBig && bins.big > 0 && bins.big <= count / 16) {
        auto isBig = [&](const Prim& x) { return ((x.hi[0] - x.lo[0]) + (x.hi[1] - x.lo[1]) + (x.hi[2] - x.lo[2])) > bigLimit; };
        const uint32_t mid = (uint32_t) (std::stable_partition(prims.begin() + begin, prims.begin() + end, isBig) - prims.begin());
        const Bounds4 lb = boundsOf(begin, mid), rb = boundsOf(mid, end);
        split(node, begin, mid, end, lb, rb, true, depth);      // the few big triangles are built on the spot
        return;
      }

      // ---- SAH sweep over the OCCUPIED bins of each axis (a split after an empty bin repeats the previous candidate) ----
      for (int axis = 0; axis < 3; axis++) {
        if (!axisOk[axis]) continue;
        const uint32_t* bc = bins.count[axis];
        uint32_t occupied = 0;
        for (int b = 0; b < kBins; b++) occupied |= (bc[b] != 0u ? 1u : 0u) << b;
        float rightArea[kBins]; uint32_t rightCount[kBins];
        __m128 alo = _mm_set1_ps(3.402823466e38f), ahi = _mm_set1_ps(-3.402823466e38f); uint32_t c = 0;
        for (uint32_t m = occupied; m;) {             // descending: rightArea[b] = area of everything in bins >= b
          const int b = 31 - __builtin_clz(m); m ^= 1u << b;
          alo = _mm_min_ps(alo, bins.lo[axis][b]); ahi = _mm_max_ps(ahi, bins.hi[axis][b]); c += bc[b];
          rightArea[b] = area4(alo, ahi); rightCount[b] = c;
        }
        alo = _mm_set1_ps(3.402823466e38f); ahi = _mm_set1_ps(-3.402823466e38f); c = 0;
        for (uint32_t m = occupied; m & (m - 1);) {   // ascending, every occupied bin but the highest
          const int b = __builtin_ctz(m); m &= m - 1;
          const int nb = __builtin_ctz(m);
          alo = _mm_min_ps(alo, bins.lo[axis][b]); ahi = _mm_max_ps(ahi, bins.hi[axis][b]); c += bc[b];
          const float cost = area4(alo, ahi) * (float) c + rightArea[nb] * (float) rightCount[nb];
          if (better(cost, c)) { bestCost = cost; bestAxis = axis; bestBin = b; }
        }
      }
    }

    // ---- partition; the children's bounds are gathered while the records move ----
    uint32_t mid;
    Bounds4 lb, rb;
    bool haveBounds = false;
    if (bestAxis < 0) {
      mid = begin + count / 2;  // identical centroids: split by index
    } else {
      const float lo = cbox.lo[bestAxis], sc = scaleA[bestAxis];
      lb.reset(); rb.reset();
      auto goesLeft = [&](const Prim& p) {
        int b = (int) ((0.5f * (p.lo[bestAxis] + p.hi[bestAxis]) - lo) * sc);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        return b <= bestBin;
      };
      if (chunks > 1) mid = partitionChunked(begin, end, goesLeft, lb, rb);
      else mid = partitionSerial(begin, end, goesLeft, lb, rb);
      haveBounds = true;
      if (mid == begin || mid == end) { mid = begin + count / 2; haveBounds = false; }
    }
    if (!haveBounds) { lb = boundsOf(begin, mid); rb = boundsOf(mid, end); }
    split(node, begin, mid, end, lb, rb, false, depth);
  }
};

struct WideChild { uint32_t bin; bool leaf; };

// floor / ceil without the library call for the values that occur here (a few hundred at most)
inline double floorSmall(double x)
{
  if (!(x > -1e9 && x < 1e9)) return std::floor(x);
  double f = (double) (int) x;
  if (f > x) f -= 1.0;
  return f;
}
inline double ceilSmall(double x)
{
  if (!(x > -1e9 && x < 1e9)) return std::ceil(x);
  double f = (double) (int) x;
  if (f < x) f += 1.0;
  return f;
}

inline uint32_t floatExponentBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return (u >> 23) & 0xFFu; }

} // namespace

void bvh8_warm_up_threads()
{
#pragma omp parallel
  {
    const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(300);
    while (std::chrono::steady_clock::now() < until) {}
  }
}

static void bvh8_build_once(const BuildTri* tris, size_t count, Bvh8& out, uint32_t depthCap)
{
  auto t0 = std::chrono::steady_clock::now();
  out.nodes.clear(); out.tris.clear(); out.maxDepth = 0;
  const uint32_t N = (uint32_t) count;
  if (N == 0) {
    Bvh8Node empty; std::memset(&empty, 0, sizeof(empty));
    out.nodes.push_back(empty);
    out.buildSeconds = 0.0;
    return;
  }
  if (const char* e = getenv("GTLC_BVH_MAX_LEAF")) { int v = atoi(e); if (v >= 1 && v <= 3) kMaxLeaf = (uint32_t) v; }
  if (const char* e = getenv("GTLC_BVH_NODE_COST")) { float v = (float) atof(e); if (v > 0.0f) kNodeCost = v; }
  if (const char* e = getenv("GTLC_BVH_BIG_TRIS")) kBigTriFraction = (float) atof(e);
  Builder b;
  b.depthCap = depthCap;
  b.prims.resize(N);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t) N; i++) {
    Box bx; bx.reset();
    for (int k = 0; k < 3; k++) bx.grow(tris[i].v[k]);
    // A triangle with a NaN / infinite / absurdly large coordinate can never be hit (every comparison of the watertight
    // test fails; Vulkan calls such triangles inactive) but would poison the boxes and the SAH arithmetic above it:
    // it enters the tree as a point at the origin.
    bool finite = true;
    for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) finite = finite && (std::fabs(tris[i].v[v][k]) <= kMaxCoordinate);
    if (!finite) for (int k = 0; k < 3; k++) bx.lo[k] = bx.hi[k] = 0.0f;
    Prim& p = b.prims[(size_t) i];
    for (int k = 0; k < 3; k++) { p.lo[k] = bx.lo[k]; p.hi[k] = bx.hi[k]; }
    p.id = (uint32_t) i; p.pad = 0;
  }
  b.nodes.resize((size_t) 2 * N - 1);
  b.dp.resize((size_t) 2 * N - 1);
#pragma omp parallel
#pragma omp single
  {
    const Bounds4 all = b.boundsOf(0, N);
    b.build(0, 0, N, all, 0);
  }

  const auto tBinary = std::chrono::steady_clock::now();
  const RawVector<BinNode>& bn = b.nodes;
  const RawVector<Dp>& dp = b.dp;
  // expands "subtree `node` as at most `i` children" into the child list
  struct WideChildList { WideChild ch[8]; int n; };
  auto gather = [&](auto&& self, uint32_t node, int i, WideChildList& out) -> void {
    const Dp& d = dp[node];
    while (i > 1 && d.split[i - 1] == 0) i--;
    if (i == 1) { out.ch[out.n++] = { node, d.rootIsLeaf != 0 }; return; }
    const int k = d.split[i - 1];
    self(self, bn[node].left, k, out);
    self(self, bn[node].right, i - k, out);
  };

  // ---- emission, one level of the wide tree at a time (the same scheme as the device builder): child lists and their
  // sizes in parallel, a prefix sum fixes every node's childBase / triBase, then the nodes are written in parallel.
  // The layout is the breadth-first order a serial queue would produce.
  struct Pending { uint32_t bin, wide, depth; };
  RawVector<Pending> level, next;
  out.nodes.reserve((size_t) N + 1);                  // a wide node has at least two children (virtual memory until used)
  out.nodes.resize(1);
  out.tris.clear();
  out.tris.resize((size_t) N * 12);                   // every triangle lands in exactly one leaf
  level.push_back({ 0u, 0u, 1u });
  RawVector<WideChildList> lists;                     // per-level scratch, reused so its pages are touched once
  RawVector<uint32_t> nodeOff, triOff;
  uint32_t nodeBase = 1, triBaseLevel = 0;
  while (!level.empty()) {
    const int64_t M = (int64_t) level.size();
    out.maxDepth = std::max(out.maxDepth, level[0].depth);
    lists.resize((size_t) M); nodeOff.resize((size_t) M); triOff.resize((size_t) M);
#pragma omp parallel for schedule(static) if (M > 256)
    for (int64_t e = 0; e < M; e++) {
      const Pending cur = level[(size_t) e];
      WideChildList& list = lists[(size_t) e]; list.n = 0;
      if (dp[cur.bin].rootIsLeaf) { list.ch[list.n++] = { cur.bin, true }; }       // only possible for the global root
      else {
        gather(gather, bn[cur.bin].left, dp[cur.bin].rootSplit, list);
        gather(gather, bn[cur.bin].right, 8 - dp[cur.bin].rootSplit, list);
      }
      uint32_t nn = 0, nt = 0;
      for (int i = 0; i < list.n; i++) { if (list.ch[i].leaf) nt += bn[list.ch[i].bin].total; else nn++; }
      nodeOff[(size_t) e] = nn; triOff[(size_t) e] = nt;
    }
    uint32_t totalNodes = 0, totalTris = 0;
    for (int64_t e = 0; e < M; e++) { const uint32_t a = nodeOff[(size_t) e], t = triOff[(size_t) e]; nodeOff[(size_t) e] = totalNodes; triOff[(size_t) e] = totalTris; totalNodes += a; totalTris += t; }
    out.nodes.resize((size_t) nodeBase + totalNodes);
    next.resize(totalNodes);                          // one entry per internal child, all written below
#pragma omp parallel for schedule(static) if (M > 256)
    for (int64_t ei = 0; ei < M; ei++) {
    const size_t e = (size_t) ei;
    const Pending cur = level[e];
    const BinNode& root = bn[cur.bin];
    WideChild* ch = lists[e].ch; const int n = lists[e].n;
    // ---- octant-aware slot assignment (greedy): slot s sits towards the corner whose sign bits are s
    const Box& pb = root.box;
    float pc[3] = { 0.5f * (pb.lo[0] + pb.hi[0]), 0.5f * (pb.lo[1] + pb.hi[1]), 0.5f * (pb.lo[2] + pb.hi[2]) };
    int slotOf[8]; bool slotUsed[8] = { false, false, false, false, false, false, false, false };
    bool childDone[8] = { false, false, false, false, false, false, false, false };
    float cost[8][8];
    for (int i = 0; i < n; i++) {
      const Box& cb = bn[ch[i].bin].box;
      float d[3] = { 0.5f * (cb.lo[0] + cb.hi[0]) - pc[0], 0.5f * (cb.lo[1] + cb.hi[1]) - pc[1], 0.5f * (cb.lo[2] + cb.hi[2]) - pc[2] };
      for (int s = 0; s < 8; s++)
        cost[i][s] = ((s & 1) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 4) ? d[2] : -d[2]);
    }
    // n rounds of "take the largest entry of the matrix whose row and column are still free" (ties: first in row-major
    // order).  Every row remembers its best free column; a round only rescans the rows that just lost theirs.
    float rowBest[8]; int rowSlot[8];
    auto rescan = [&](int i) {
      float bc = -3.402823466e38f; int bs = -1;
      for (int s = 0; s < 8; s++) if (!slotUsed[s] && cost[i][s] > bc) { bc = cost[i][s]; bs = s; }
      rowBest[i] = bc; rowSlot[i] = bs;
    };
    for (int i = 0; i < n; i++) rescan(i);
    for (int k = 0; k < n; k++) {
      int bi = -1; float bc = -3.402823466e38f;
      for (int i = 0; i < n; i++) if (!childDone[i] && rowBest[i] > bc) { bc = rowBest[i]; bi = i; }
      if (bi < 0) { for (int i = 0; i < n && bi < 0; i++) if (!childDone[i]) bi = i; }          // no finite entry left: any order will do
      int bs = rowSlot[bi];
      if (bs < 0 || slotUsed[bs]) { bs = 0; while (slotUsed[bs]) bs++; }
      slotOf[bi] = bs; childDone[bi] = true; slotUsed[bs] = true;
      for (int i = 0; i < n; i++) if (!childDone[i] && rowSlot[i] == bs) rescan(i);
    }
    int childAtSlot[8]; for (int s = 0; s < 8; s++) childAtSlot[s] = -1;
    for (int i = 0; i < n; i++) childAtSlot[slotOf[i]] = i;

    // ---- quantisation grid: origin one quantum below the node's lower corner and 251 quanta of extent,
    // so that every child bound can be padded by a safety quantum on BOTH sides (the slab test's
    // 2^15-bias arithmetic is accurate to ~1/512 quantum; see traverse.cuh) ----------------------
    uint32_t ebits[3]; double scale[3], invScale[3]; float porigin[3];
    for (int k = 0; k < 3; k++) {
      // smallest e with ext <= 250 * 2^e, kept inside [-100, 100]: ext = m * 2^x with m in [0.5, 1), 250 = 0.9765625 * 2^8
      const double ext = (double) pb.hi[k] - (double) pb.lo[k];
      int e = -100;
      if (ext > 0.0) { int x; const double m = std::frexp(ext, &x); e = (m <= 0.9765625) ? x - 8 : x - 7; }
      if (e < -100) e = -100;
      if (e > 100) e = 100;
      ebits[k] = (uint32_t) (e + 127);
      scale[k] = std::ldexp(1.0, e); invScale[k] = std::ldexp(1.0, -e);
      float po = (float) ((double) pb.lo[k] - scale[k]);
      while ((double) po > (double) pb.lo[k] - 0.75 * scale[k]) po = std::nextafterf(po, -3.402823466e38f);
      porigin[k] = po;
    }
    Bvh8Node node; std::memset(&node, 0, sizeof(node));
    std::memcpy(&node.w[0], &porigin[0], 4); std::memcpy(&node.w[1], &porigin[1], 4); std::memcpy(&node.w[2], &porigin[2], 4);
    uint8_t meta[8] = { 0 }; uint8_t q[6][8]; std::memset(q, 0, sizeof(q));
    uint32_t imask = 0;
    const uint32_t childBase = nodeBase + nodeOff[e];
    const uint32_t triBase = triBaseLevel + triOff[e];
    uint32_t nInternal = 0, triOffset = 0;
    for (int s = 0; s < 8; s++) {
      int i = childAtSlot[s];
      if (i < 0) continue;
      const BinNode& c = bn[ch[i].bin];
      for (int k = 0; k < 3; k++) {
        double lo = ((double) c.box.lo[k] - (double) porigin[k]) * invScale[k];   // power of two: exact
        double hi = ((double) c.box.hi[k] - (double) porigin[k]) * invScale[k];
        double ql = floorSmall(lo), qh = ceilSmall(hi);
        if (lo - ql < 1.0 / 64.0 && ql > 0.0) ql -= 1.0;   // safety quantum when the bound sits on a grid line
        if (qh - hi < 1.0 / 64.0 && qh < 255.0) qh += 1.0;
        if (ql < 0.0) ql = 0.0;
        if (qh > 255.0) qh = 255.0;
        q[k][s] = (uint8_t) ql; q[3 + k][s] = (uint8_t) qh;
      }
      if (ch[i].leaf) {
        uint32_t cnt = c.total;  // 1..kMaxLeaf triangles: the whole binary subtree becomes one leaf
        meta[s] = (uint8_t) ((((1u << cnt) - 1u) << 5) | triOffset);
        for (uint32_t t = 0; t < cnt; t++) {
          const BuildTri& tr = tris[b.prims[c.first + t].id];
          float rec[12];
          for (int v = 0; v < 3; v++) { rec[v * 4 + 0] = tr.v[v][0]; rec[v * 4 + 1] = tr.v[v][1]; rec[v * 4 + 2] = tr.v[v][2]; }
          std::memcpy(&rec[3], &tr.prim, 4); std::memcpy(&rec[7], &tr.inst, 4); rec[11] = tr.opacity;
          std::memcpy(&out.tris[(size_t) (triBase + triOffset + t) * 12], rec, sizeof(rec));
        }
        triOffset += cnt;
      } else {
        meta[s] = (uint8_t) (0x20u | (24u + (uint32_t) s));
        imask |= 1u << s;
        nInternal++;
      }
    }
    node.w[3] = ebits[0] | (ebits[1] << 8) | (ebits[2] << 16) | (imask << 24);
    node.w[4] = childBase; node.w[5] = triBase;
    std::memcpy(&node.w[6], meta, 8);
    std::memcpy(&node.w[8], q[0], 8);  std::memcpy(&node.w[10], q[1], 8);   // qlo_x, qlo_y
    std::memcpy(&node.w[12], q[2], 8); std::memcpy(&node.w[14], q[3], 8);   // qlo_z, qhi_x
    std::memcpy(&node.w[16], q[4], 8); std::memcpy(&node.w[18], q[5], 8);   // qhi_y, qhi_z
    out.nodes[cur.wide] = node;
    // internal children get consecutive indices in slot order
    uint32_t k = 0;
    for (int sl = 0; sl < 8; sl++) {
      const int i = childAtSlot[sl];
      if (i < 0 || ch[i].leaf) continue;
      next[nodeOff[e] + k] = { ch[i].bin, childBase + k, cur.depth + 1 };
      k++;
    }
    }
    nodeBase += totalNodes; triBaseLevel += totalTris;
    level.swap(next);
  }
  out.tris.resize((size_t) triBaseLevel * 12);
  out.buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("GTLC_BUILD_TIMING")) {
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[gtlc] host BVH8: binary SAH + collapse table %.0f ms, wide-node emission %.0f ms\n", ms(t0, tBinary), ms(tBinary, std::chrono::steady_clock::now()));
  }
}

void bvh8_build(const BuildTri* tris, size_t count, Bvh8& out)
{
  uint32_t cap = 0u;
  if (const char* e = getenv("GTLC_BVH_DEPTH_CAP")) cap = (uint32_t) std::max(0, atoi(e));    // tests: exercise the capped build directly
  bvh8_build_once(tris, count, out, cap);
  if (out.maxDepth > kBvh8MaxDepth) {
    // SAH can legitimately produce chains (geometric progressions of triangle sizes peel one triangle per level).  The
    // traversal stack is fixed, so such a scene is built again with a depth cap: SAH splits until the levels left are just
    // enough for a balanced subtree, median splits below.  The wide tree is never deeper than the binary one.
    const double first = out.buildSeconds;
    if (getenv("GTLC_BUILD_TIMING")) fprintf(stderr, "[gtlc] host BVH8: wide depth %u > %u, rebuilding with a depth cap\n", out.maxDepth, kBvh8MaxDepth);
    bvh8_build_once(tris, count, out, kBvh8MaxDepth - 4u);
    out.buildSeconds += first;
  }
}

} // namespace gtlc
