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

constexpr int kBins = 16;
constexpr uint32_t kParallelRange = 262144;   // ranges at least this long are scanned by several tasks
// SAH parameters (overridable through GTLC_BVH_MAX_LEAF / GTLC_BVH_NODE_COST for tuning experiments)
uint32_t kMaxLeaf = 3;
float kNodeCost = 0.6f;  // relative to one triangle test
float kBigTriFraction = 0.4f;   // GTLC_BVH_BIG_TRIS: a triangle is "big" when its extent sum exceeds this fraction of the node's (0 = off)
constexpr float kTriCost = 1.0f;

struct Builder {
  const std::vector<BuildTri>& tris;
  std::vector<Box> triBox;
  std::vector<float> cen;  // 3 per tri
  std::vector<uint32_t> idx;
  std::vector<BinNode> nodes;
  std::atomic<uint32_t> nodeCount{ 0 };

  explicit Builder(const std::vector<BuildTri>& t) : tris(t) {}

  void build(uint32_t node, uint32_t begin, uint32_t end)
  {
    const uint32_t count = end - begin;
    // The top of the tree is a handful of huge ranges: their passes are cut into chunks that run as tasks of the team
    // (min / max / counts merge exactly, so the tree does not depend on the chunking).
    const uint32_t chunks = count >= kParallelRange ? 64u : 1u;
    const uint32_t chunkLen = (count + chunks - 1) / chunks;
    Box box, cbox;
    box.reset(); cbox.reset();
    if (chunks == 1) {
      for (uint32_t i = begin; i < end; i++) {
        box.grow(triBox[idx[i]]);
        cbox.grow(&cen[(size_t) idx[i] * 3]);
      }
    } else {
      std::vector<Box> pb(chunks), pc(chunks);
#pragma omp taskloop default(shared) grainsize(1)
      for (uint32_t c = 0; c < chunks; c++) {
        Box a, d; a.reset(); d.reset();
        const uint32_t e0 = begin + c * chunkLen, e1 = std::min(end, e0 + chunkLen);
        for (uint32_t i = e0; i < e1; i++) { a.grow(triBox[idx[i]]); d.grow(&cen[(size_t) idx[i] * 3]); }
        pb[c] = a; pc[c] = d;
      }
      for (uint32_t c = 0; c < chunks; c++) { if (pb[c].lo[0] <= pb[c].hi[0]) box.grow(pb[c]); if (pc[c].lo[0] <= pc[c].hi[0]) cbox.grow(pc[c]); }
    }
    BinNode& n = nodes[node];
    n.box = box;
    auto makeLeaf = [&]() { n.first = begin; n.count = count; n.total = count; n.left = n.right = 0; };
    if (count == 1) { makeLeaf(); return; }

    // Triangles that span most of this node (room walls around furniture, a table top over what stands on it) go into a
    // subtree of their own: sorted by centroid they would sit deep inside one child and inflate every box on the way.
    // Measured in BVH8 node visits per ray (oracle walker): kitchen stand-in 6.59 -> 5.03, nothing lost elsewhere.
    if (count >= 64 && kBigTriFraction > 0.0f) {
      const float sd = (box.hi[0] - box.lo[0]) + (box.hi[1] - box.lo[1]) + (box.hi[2] - box.lo[2]);
      auto isBig = [&](uint32_t t) { const Box& x = triBox[t]; return ((x.hi[0] - x.lo[0]) + (x.hi[1] - x.lo[1]) + (x.hi[2] - x.lo[2])) > kBigTriFraction * sd; };
      uint32_t big = 0;
      for (uint32_t i = begin; i < end && big <= count / 16; i++) big += isBig(idx[i]) ? 1u : 0u;
      if (big > 0 && big <= count / 16) {
        const uint32_t mid = (uint32_t) (std::stable_partition(idx.begin() + begin, idx.begin() + end, isBig) - idx.begin());
        const uint32_t l = nodeCount.fetch_add(2);
        n.left = l; n.right = l + 1; n.count = 0; n.first = begin; n.total = count;
        build(l, begin, mid);
        if (count > 16384) {
#pragma omp task default(shared)
          build(l + 1, mid, end);
#pragma omp taskwait
        } else build(l + 1, mid, end);
        return;
      }
    }

    // binned SAH over the three axes
    float bestCost = 3.402823466e38f; int bestAxis = -1; int bestBin = -1;
    for (int axis = 0; axis < 3; axis++) {
      float lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
      if (!(ext > 0.0f)) continue;
      float scale = (float) kBins / ext;
      Box bb[kBins]; uint32_t bc[kBins];
      for (int b = 0; b < kBins; b++) { bb[b].reset(); bc[b] = 0; }
      if (chunks == 1) {
        for (uint32_t i = begin; i < end; i++) {
          int b = (int) ((cen[(size_t) idx[i] * 3 + axis] - lo) * scale);
          b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
          bb[b].grow(triBox[idx[i]]); bc[b]++;
        }
      } else {
        struct Partial { Box bb[kBins]; uint32_t bc[kBins]; };
        std::vector<Partial> part(chunks);
#pragma omp taskloop default(shared) grainsize(1)
        for (uint32_t c = 0; c < chunks; c++) {
          Partial& pt = part[c];
          for (int b = 0; b < kBins; b++) { pt.bb[b].reset(); pt.bc[b] = 0; }
          const uint32_t e0 = begin + c * chunkLen, e1 = std::min(end, e0 + chunkLen);
          for (uint32_t i = e0; i < e1; i++) {
            int b = (int) ((cen[(size_t) idx[i] * 3 + axis] - lo) * scale);
            b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
            pt.bb[b].grow(triBox[idx[i]]); pt.bc[b]++;
          }
        }
        for (uint32_t c = 0; c < chunks; c++) for (int b = 0; b < kBins; b++) if (part[c].bc[b]) { bb[b].grow(part[c].bb[b]); bc[b] += part[c].bc[b]; }
      }
      float rightArea[kBins]; uint32_t rightCount[kBins];
      Box acc; acc.reset(); uint32_t c = 0;
      for (int b = kBins - 1; b > 0; b--) { acc.grow(bb[b]); c += bc[b]; rightArea[b] = acc.area(); rightCount[b] = c; }
      acc.reset(); c = 0;
      for (int b = 0; b < kBins - 1; b++) {
        acc.grow(bb[b]); c += bc[b];
        if (c == 0 || rightCount[b + 1] == 0) continue;
        float cost = acc.area() * (float) c + rightArea[b + 1] * (float) rightCount[b + 1];
        if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
      }
    }
    uint32_t mid;
    if (bestAxis < 0) {
      mid = begin + count / 2;  // identical centroids: split by index
    } else {
      float lo = cbox.lo[bestAxis], scale = (float) kBins / (cbox.hi[bestAxis] - cbox.lo[bestAxis]);
      auto it = std::partition(idx.begin() + begin, idx.begin() + end, [&](uint32_t t) {
        int b = (int) ((cen[(size_t) t * 3 + bestAxis] - lo) * scale);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        return b <= bestBin;
      });
      mid = (uint32_t) (it - idx.begin());
      if (mid == begin || mid == end) mid = begin + count / 2;
    }
    uint32_t l = nodeCount.fetch_add(2);
    n.left = l; n.right = l + 1; n.count = 0; n.first = begin; n.total = count;
    if (count > 16384) {
#pragma omp task default(shared)
      build(l, begin, mid);
#pragma omp task default(shared)
      build(l + 1, mid, end);
#pragma omp taskwait
    } else {
      build(l, begin, mid);
      build(l + 1, mid, end);
    }
  }
};

struct WideChild { uint32_t bin; bool leaf; };

inline uint32_t floatExponentBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return (u >> 23) & 0xFFu; }

} // namespace

void bvh8_build(const std::vector<BuildTri>& tris, Bvh8& out)
{
  auto t0 = std::chrono::steady_clock::now();
  out.nodes.clear(); out.tris.clear(); out.maxDepth = 0;
  const uint32_t N = (uint32_t) tris.size();
  if (N == 0) {
    Bvh8Node empty; std::memset(&empty, 0, sizeof(empty));
    out.nodes.push_back(empty);
    out.buildSeconds = 0.0;
    return;
  }
  if (const char* e = getenv("GTLC_BVH_MAX_LEAF")) { int v = atoi(e); if (v >= 1 && v <= 3) kMaxLeaf = (uint32_t) v; }
  if (const char* e = getenv("GTLC_BVH_NODE_COST")) { float v = (float) atof(e); if (v > 0.0f) kNodeCost = v; }
  if (const char* e = getenv("GTLC_BVH_BIG_TRIS")) kBigTriFraction = (float) atof(e);
  Builder b(tris);
  b.triBox.resize(N); b.cen.resize((size_t) N * 3); b.idx.resize(N);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t) N; i++) {
    Box bx; bx.reset();
    for (int k = 0; k < 3; k++) bx.grow(tris[i].v[k]);
    b.triBox[i] = bx;
    for (int k = 0; k < 3; k++) b.cen[(size_t) i * 3 + k] = 0.5f * (bx.lo[k] + bx.hi[k]);
    b.idx[i] = (uint32_t) i;
  }
  b.nodes.resize((size_t) 2 * N);
  b.nodeCount = 1;
#pragma omp parallel
#pragma omp single
  b.build(0, 0, N);

  const auto tBinary = std::chrono::steady_clock::now();
  // ---- SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3.1), dynamic programme bottom-up over the
  // binary tree: cost[n][i-1] = cheapest way to present subtree n to its parent as at most i children
  // (i = 1..7), each child either a leaf of <= kMaxLeaf triangles or a wide node.  Bottom-up = post-order; the two
  // subtrees of a big node are independent OpenMP tasks.
  const std::vector<BinNode>& bn = b.nodes;
  const uint32_t binCount = b.nodeCount.load();
  struct Dp { float cost[7]; uint8_t split[7]; uint8_t rootSplit; uint8_t rootIsLeaf; };
  std::vector<Dp> dp(binCount);
  const float INF = 3.402823466e38f;
  auto solveNode = [&](uint32_t ni) {
    const BinNode& n = bn[ni];
    Dp& d = dp[ni];
    const float area = n.box.area();
    const float leafCost = (n.total <= kMaxLeaf) ? area * (float) n.total * kTriCost : INF;
    if (n.count > 0) {                       // binary leaf (single triangle or coincident centroids)
      for (int i = 0; i < 7; i++) { d.cost[i] = leafCost; d.split[i] = 0; }
      d.rootSplit = 0; d.rootIsLeaf = 1;
      return;
    }
    const Dp& L = dp[n.left]; const Dp& R = dp[n.right];
    // as one wide node: distribute 8 child slots between the two binary children
    float best = INF; uint8_t bk = 1;
    for (int k = 1; k <= 7; k++) { const float c = L.cost[k - 1] + R.cost[8 - k - 1]; if (c < best) { best = c; bk = (uint8_t) k; } }
    const float internalCost = best + area * kNodeCost;
    d.rootSplit = bk; d.rootIsLeaf = leafCost <= internalCost;
    d.cost[0] = d.rootIsLeaf ? leafCost : internalCost; d.split[0] = 0;
    for (int i = 2; i <= 7; i++) {            // as at most i children of the parent
      float bd = INF; uint8_t bs = 0;
      for (int k = 1; k < i; k++) { const float c = L.cost[k - 1] + R.cost[i - k - 1]; if (c < bd) { bd = c; bs = (uint8_t) k; } }
      if (bd < d.cost[i - 2]) { d.cost[i - 1] = bd; d.split[i - 1] = bs; }
      else { d.cost[i - 1] = d.cost[i - 2]; d.split[i - 1] = 0; }   // 0 = same as with one child fewer
    }
  };
  // post-order; big subtrees become OpenMP tasks, small ones run as an explicit-stack walk inside their task
  auto solveSubtree = [&](auto&& self, uint32_t root) -> void {
    if (bn[root].count == 0 && bn[root].total > 32768) {
#pragma omp task default(shared)
      self(self, bn[root].left);
#pragma omp task default(shared)
      self(self, bn[root].right);
#pragma omp taskwait
      solveNode(root);
      return;
    }
    std::vector<uint32_t> order; order.reserve(2 * (size_t) bn[root].total);
    order.push_back(root);
    for (size_t h = 0; h < order.size(); h++) { const BinNode& n = bn[order[h]]; if (n.count == 0) { order.push_back(n.left); order.push_back(n.right); } }
    for (size_t h = order.size(); h-- > 0;) solveNode(order[h]);      // children come after their parent in `order`
  };
#pragma omp parallel
#pragma omp single
  solveSubtree(solveSubtree, 0u);
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

  const auto tDp = std::chrono::steady_clock::now();
  // ---- emission, one level of the wide tree at a time (the same scheme as the device builder): child lists and their
  // sizes in parallel, a prefix sum fixes every node's childBase / triBase, then the nodes are written in parallel.
  // The layout is the breadth-first order a serial queue would produce.
  struct Pending { uint32_t bin, wide, depth; };
  std::vector<Pending> level, next;
  out.nodes.assign(1, Bvh8Node());
  out.tris.clear();
  level.push_back({ 0u, 0u, 1u });
  uint32_t nodeBase = 1, triBaseLevel = 0;
  while (!level.empty()) {
    const int64_t M = (int64_t) level.size();
    out.maxDepth = std::max(out.maxDepth, level[0].depth);
    std::vector<WideChildList> lists((size_t) M);
    std::vector<uint32_t> nodeOff((size_t) M), triOff((size_t) M);
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
    out.tris.resize((size_t) (triBaseLevel + totalTris) * 12);
    next.assign(totalNodes, Pending{ 0u, 0u, 0u });
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
    for (int k = 0; k < n; k++) {
      int bi = -1, bs = -1; float bc = -3.402823466e38f;
      for (int i = 0; i < n; i++) {
        if (childDone[i]) continue;
        for (int s = 0; s < 8; s++) {
          if (slotUsed[s]) continue;
          if (cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; }
        }
      }
      slotOf[bi] = bs; childDone[bi] = true; slotUsed[bs] = true;
    }
    int childAtSlot[8]; for (int s = 0; s < 8; s++) childAtSlot[s] = -1;
    for (int i = 0; i < n; i++) childAtSlot[slotOf[i]] = i;

    // ---- quantisation grid: origin one quantum below the node's lower corner and 251 quanta of extent,
    // so that every child bound can be padded by a safety quantum on BOTH sides (the slab test's
    // 2^15-bias arithmetic is accurate to ~1/512 quantum; see traverse.cuh) ----------------------
    uint32_t ebits[3]; double scale[3]; float porigin[3];
    for (int k = 0; k < 3; k++) {
      double ext = (double) pb.hi[k] - (double) pb.lo[k];
      int e = (ext > 0.0) ? (int) std::ceil(std::log2(ext / 250.0)) : -100;
      if (e < -100) e = -100;
      if (e > 100) e = 100;
      while (ext / std::ldexp(1.0, e) > 250.0 && e < 100) e++;
      ebits[k] = (uint32_t) (e + 127);
      scale[k] = std::ldexp(1.0, e);
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
        double lo = ((double) c.box.lo[k] - (double) porigin[k]) / scale[k];
        double hi = ((double) c.box.hi[k] - (double) porigin[k]) / scale[k];
        double ql = std::floor(lo), qh = std::ceil(hi);
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
          const BuildTri& tr = tris[b.idx[c.first + t]];
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
  out.buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("GTLC_BUILD_TIMING")) {
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[gtlc] host BVH8: binary SAH %.0f ms, collapse table %.0f ms, wide-node emission %.0f ms\n", ms(t0, tBinary), ms(tBinary, tDp), ms(tDp, std::chrono::steady_clock::now()));
  }
}

} // namespace gtlc
