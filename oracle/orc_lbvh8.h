// TEST INFRASTRUCTURE ONLY -- CPU restatement of the device BVH8 builder (gatling_b200/csrc/bvh8_gpu.cuh), which
// takes the place of the driver's acceleration-structure build the reference requests through
// cgpuCreateBlas / cgpuCreateTlas (/root/reference/src/cgpu/impl/CgpuVk.cpp:2561-2854).  The reference has no
// algorithm of its own here (the Vulkan driver owns it), so this oracle pins the product's builder, not the
// reference's: parity unpinned against the reference, byte-exact between this file and the CUDA kernels.
//
// Sequential and deliberately written in a different style from the kernels (recursion instead of explicit
// stacks, std::stable_sort instead of a radix sort, a post-order walk instead of atomic counters).  Every
// floating-point expression keeps the kernels' operator order (compiled with -ffp-contract=off).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc_lbvh8 {

struct Box3 { float lo[3], hi[3]; };
struct Table { float cost[7]; uint8_t split[7]; uint8_t rootSplit, rootIsLeaf; };

static const uint32_t MAX_LEAF = 3;
static const float NODE_COST = 0.6f;
static const float BIG = 3.402823466e38f;

inline float box_area(const Box3& b)
{
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return 2.0f * ((dx * dy + dy * dz) + dz * dx);
}

inline uint64_t spread21(uint32_t v)   // bit i of v -> bit 3i
{
  uint64_t r = 0;
  for (int i = 0; i < 21; i++) r |= (uint64_t) ((v >> i) & 1u) << (3 * i);
  return r;
}

inline uint32_t grid21(float c, float lo, float hi)
{
  const float ext = hi - lo;
  if (!(ext > 0.0f)) return 0;
  const float x = (c - lo) / ext;
  const uint32_t q = (uint32_t) (x * 2097152.0f);
  return q > 2097151u ? 2097151u : q;
}

inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }

struct Builder {
  uint32_t n;
  const float* in;                         // 12 floats per triangle, rank order
  std::vector<uint64_t> key; std::vector<uint32_t> val;
  std::vector<uint32_t> left, right, total;           // internal nodes 0..n-2: children, triangles in the subtree
  std::vector<Box3> box;                   // refs: internal 0..n-2, leaves n-1..2n-2
  std::vector<Table> tab;                  // internal nodes
  std::vector<uint32_t> outNodes;          // 20 words per wide node
  std::vector<float> outTris;
  uint32_t depth = 0;

  bool is_leaf(uint32_t ref) const { return ref >= n - 1; }
  uint32_t tri_count(uint32_t ref) const { return is_leaf(ref) ? 1u : total[ref]; }
  void leaves_of(uint32_t ref, std::vector<uint32_t>& positions) const     // sorted positions of the subtree's triangles, left to right
  {
    if (is_leaf(ref)) { positions.push_back(ref - (n - 1)); return; }
    leaves_of(left[ref], positions); leaves_of(right[ref], positions);
  }

  Box3 tri_box(uint32_t t) const
  {
    Box3 b;
    for (int k = 0; k < 3; k++) {
      const float a = in[(size_t) t * 12 + k], bb = in[(size_t) t * 12 + 4 + k], c = in[(size_t) t * 12 + 8 + k];
      b.lo[k] = std::fmin(std::fmin(a, bb), c);
      b.hi[k] = std::fmax(std::fmax(a, bb), c);
    }
    // inactive triangles (NaN / infinite / > 1e18 coordinates) enter the tree as a point at the origin (bvh8_gpu.cuh: lbvh_tri_box)
    bool finite = true;
    for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) finite = finite && (std::fabs(in[(size_t) t * 12 + 4 * v + k]) <= 1.0e18f);
    if (!finite) for (int k = 0; k < 3; k++) b.lo[k] = b.hi[k] = 0.0f;
    return b;
  }

  int common_prefix(int i, int j) const
  {
    if (j < 0 || j >= (int) n) return -1;
    if (key[i] == key[j]) return 64 + clz32((uint32_t) (i ^ j));
    return clz64(key[i] ^ key[j]);
  }

  void sort_keys()
  {
    float cmin[3] = { BIG, BIG, BIG }, cmax[3] = { -BIG, -BIG, -BIG };
    std::vector<float> cen((size_t) n * 3);
    for (uint32_t t = 0; t < n; t++) {
      const Box3 b = tri_box(t);
      for (int k = 0; k < 3; k++) {
        const float c = 0.5f * (b.lo[k] + b.hi[k]);
        cen[(size_t) t * 3 + k] = c;
        cmin[k] = std::fmin(cmin[k], c); cmax[k] = std::fmax(cmax[k], c);
      }
    }
    std::vector<std::pair<uint64_t, uint32_t>> kv(n);
    for (uint32_t t = 0; t < n; t++) {
      const uint64_t code = (spread21(grid21(cen[(size_t) t * 3 + 0], cmin[0], cmax[0])) << 2) |
                            (spread21(grid21(cen[(size_t) t * 3 + 1], cmin[1], cmax[1])) << 1) |
                            spread21(grid21(cen[(size_t) t * 3 + 2], cmin[2], cmax[2]));
      kv[t] = { code, t };
    }
    std::stable_sort(kv.begin(), kv.end(), [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first < b.first; });
    key.resize(n); val.resize(n);
    for (uint32_t i = 0; i < n; i++) { key[i] = kv[i].first; val[i] = kv[i].second; }
  }

  // Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", section 4
  void radix_tree()
  {
    left.assign(n, 0); right.assign(n, 0); total.assign(n, 0);
    for (int i = 0; i + 1 < (int) n; i++) {
      const int d = common_prefix(i, i + 1) - common_prefix(i, i - 1) >= 0 ? 1 : -1;
      const int floorPrefix = common_prefix(i, i - d);
      int reach = 2;
      while (common_prefix(i, i + reach * d) > floorPrefix) reach *= 2;
      int len = 0;
      for (int step = reach / 2; step >= 1; step /= 2) if (common_prefix(i, i + (len + step) * d) > floorPrefix) len += step;
      const int j = i + len * d;
      const int nodePrefix = common_prefix(i, j);
      int s = 0, step = len;
      do { step = (step + 1) / 2; if (common_prefix(i, i + (s + step) * d) > nodePrefix) s += step; } while (step > 1);
      const int cut = i + s * d + std::min(d, 0);
      const int a = std::min(i, j), b = std::max(i, j);
      left[i] = (a == cut) ? (uint32_t) (n - 1 + cut) : (uint32_t) cut;
      right[i] = (b == cut + 1) ? (uint32_t) (n - 1 + cut + 1) : (uint32_t) (cut + 1);
      total[i] = (uint32_t) (b - a + 1);
    }
  }

  // Meister & Bittner 2018, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction": clusters in
  // Morton order, each looks PLOC_RADIUS positions to either side for the partner with the smallest merged area (candidates by
  // distance, the side of the pair (2k, 2k + 1) first, strictly smaller wins: ties go to the closest position); mutual pairs merge into the left one's place;
  // the order of the surviving clusters is kept.  Node ids count down so that the root, the last merge, is node 0.
  static const int PLOC_RADIUS = 16;
  static float merged_area(const Box3& a, const Box3& b)
  {
    Box3 u;
    for (int k = 0; k < 3; k++) { u.lo[k] = std::fmin(a.lo[k], b.lo[k]); u.hi[k] = std::fmax(a.hi[k], b.hi[k]); }
    return box_area(u);
  }
  void ploc_tree()
  {
    left.assign(n, 0); right.assign(n, 0); total.assign(n, 0);
    box.resize((size_t) 2 * n);
    std::vector<uint32_t> cl(n), next;
    for (uint32_t p = 0; p < n; p++) { box[n - 1 + p] = tri_box(val[p]); cl[p] = n - 1 + p; }
    uint32_t nextId = n - 2;
    while (cl.size() > 1) {
      const uint32_t count = (uint32_t) cl.size();
      std::vector<uint32_t> nn(count);
      for (uint32_t i = 0; i < count; i++) {
        float best = BIG; uint32_t bj = i;
        for (uint32_t d = 1; d <= (uint32_t) PLOC_RADIUS; d++) {
          for (int side = 0; side < 2; side++) {
            const bool right = ((i & 1u) == 0u) == (side == 0);            // the pair (2k, 2k + 1) looks at each other first
            if (right ? (i + d > count - 1u) : (d > i)) continue;
            const uint32_t j = right ? i + d : i - d;
            const float a = merged_area(box[cl[i]], box[cl[j]]);
            if (a < best || bj == i) { best = a; bj = j; }
          }
        }
        nn[i] = bj;
      }
      next.clear();
      uint32_t merges = 0;
      for (uint32_t i = 0; i < count; i++) {
        const uint32_t j = nn[i];
        const bool mutual = j != i && nn[j] == i;
        if (mutual && i > j) continue;                         // absorbed by its partner
        if (mutual) {
          const uint32_t id = nextId - merges++;
          const uint32_t a = cl[i], b = cl[j];
          left[id] = a; right[id] = b; total[id] = tri_count(a) + tri_count(b);
          for (int k = 0; k < 3; k++) { box[id].lo[k] = std::fmin(box[a].lo[k], box[b].lo[k]); box[id].hi[k] = std::fmax(box[a].hi[k], box[b].hi[k]); }
          next.push_back(id);
        } else next.push_back(cl[i]);
      }
      nextId -= merges;
      cl.swap(next);
    }
  }

  void slot_costs(uint32_t ref, float* c) const
  {
    if (is_leaf(ref)) { const float a = box_area(box[ref]); for (int i = 0; i < 7; i++) c[i] = a; }
    else for (int i = 0; i < 7; i++) c[i] = tab[ref].cost[i];
  }

  // post-order without recursion depth problems: an explicit list in "children after parents" order, walked backwards
  void fit_and_cost()
  {
    box.resize((size_t) 2 * n);
    for (uint32_t p = 0; p < n; p++) box[n - 1 + p] = tri_box(val[p]);
    if (n < 2) return;
    tab.resize(n - 1);
    std::vector<uint32_t> order; order.reserve(n - 1);
    order.push_back(0);
    for (size_t h = 0; h < order.size(); h++) {
      const uint32_t r = order[h];
      if (!is_leaf(left[r])) order.push_back(left[r]);
      if (!is_leaf(right[r])) order.push_back(right[r]);
    }
    for (size_t h = order.size(); h-- > 0;) {
      const uint32_t r = order[h];
      const Box3 &a = box[left[r]], &b = box[right[r]];
      Box3 u;
      for (int k = 0; k < 3; k++) { u.lo[k] = std::fmin(a.lo[k], b.lo[k]); u.hi[k] = std::fmax(a.hi[k], b.hi[k]); }
      box[r] = u;
      const float area = box_area(u);
      const uint32_t tris = total[r];
      const float asLeaf = tris <= MAX_LEAF ? area * (float) tris : BIG;
      float L[7], R[7];
      slot_costs(left[r], L); slot_costs(right[r], R);
      Table t;
      float best = BIG; uint8_t bestK = 1;
      for (int k = 1; k <= 7; k++) { const float c = L[k - 1] + R[7 - k]; if (c < best) { best = c; bestK = (uint8_t) k; } }
      const float asNode = best + area * NODE_COST;
      t.rootSplit = bestK; t.rootIsLeaf = asLeaf <= asNode ? 1 : 0;
      t.cost[0] = t.rootIsLeaf ? asLeaf : asNode; t.split[0] = 0;
      for (int i = 2; i <= 7; i++) {
        float bd = BIG; uint8_t bs = 0;
        for (int k = 1; k < i; k++) { const float c = L[k - 1] + R[i - k - 1]; if (c < bd) { bd = c; bs = (uint8_t) k; } }
        if (bd < t.cost[i - 2]) { t.cost[i - 1] = bd; t.split[i - 1] = bs; }
        else { t.cost[i - 1] = t.cost[i - 2]; t.split[i - 1] = 0; }
      }
      tab[r] = t;
    }
  }

  struct Child { uint32_t ref; bool leaf; };

  void expand(uint32_t ref, int slots, std::vector<Child>& out) const
  {
    if (is_leaf(ref)) { out.push_back({ ref, true }); return; }
    const Table& t = tab[ref];
    while (slots > 1 && t.split[slots - 1] == 0) slots--;
    if (slots == 1) { out.push_back({ ref, t.rootIsLeaf != 0 }); return; }
    const int k = t.split[slots - 1];
    expand(left[ref], k, out);
    expand(right[ref], slots - k, out);
  }

  static double two_to(int e) { return std::ldexp(1.0, e); }
  static float below(float f) { return std::nextafterf(f, -BIG); }

  void emit_node(uint32_t ref, uint32_t wide, const std::vector<Child>& ch, uint32_t childBase, uint32_t triBase, std::vector<std::pair<uint32_t, uint32_t>>& next)
  {
    const int nc = (int) ch.size();
    const Box3& pb = box[ref];
    const float pc[3] = { 0.5f * (pb.lo[0] + pb.hi[0]), 0.5f * (pb.lo[1] + pb.hi[1]), 0.5f * (pb.lo[2] + pb.hi[2]) };
    float score[8][8];
    for (int i = 0; i < nc; i++) {
      const Box3& cb = box[ch[i].ref];
      const float d0 = 0.5f * (cb.lo[0] + cb.hi[0]) - pc[0], d1 = 0.5f * (cb.lo[1] + cb.hi[1]) - pc[1], d2 = 0.5f * (cb.lo[2] + cb.hi[2]) - pc[2];
      for (int s = 0; s < 8; s++) score[i][s] = (((s & 1) ? d0 : -d0) + ((s & 2) ? d1 : -d1)) + ((s & 4) ? d2 : -d2);
    }
    int atSlot[8] = { -1, -1, -1, -1, -1, -1, -1, -1 };
    bool placed[8] = { false }, taken[8] = { false };
    for (int round = 0; round < nc; round++) {
      int bi = -1, bs = -1; float bc = -BIG;
      for (int i = 0; i < nc; i++) if (!placed[i]) for (int s = 0; s < 8; s++) if (!taken[s] && score[i][s] > bc) { bc = score[i][s]; bi = i; bs = s; }
      if (bi < 0) {
        for (int i = 0; i < nc && bi < 0; i++) if (!placed[i]) bi = i;
        for (int s = 0; s < 8 && bs < 0; s++) if (!taken[s]) bs = s;
      }
      atSlot[bs] = bi; placed[bi] = true; taken[bs] = true;
    }
    uint32_t ebits[3]; double scale[3]; float origin[3];
    for (int k = 0; k < 3; k++) {
      const double ext = (double) pb.hi[k] - (double) pb.lo[k];
      int e = -100;
      if (ext > 0.0) {
        e = std::ilogb(ext) - 8;
        if (e < -100) e = -100;
        while (e < 100 && ext > 250.0 * two_to(e)) e++;
      }
      ebits[k] = (uint32_t) (e + 127);
      scale[k] = two_to(e);
      float po = (float) ((double) pb.lo[k] - scale[k]);
      while ((double) po > (double) pb.lo[k] - 0.75 * scale[k]) po = below(po);
      origin[k] = po;
    }
    uint8_t meta[8] = { 0 }, q[6][8]; std::memset(q, 0, sizeof(q));
    uint32_t imask = 0, internal = 0, triOffset = 0;
    for (int s = 0; s < 8; s++) {
      const int i = atSlot[s];
      if (i < 0) continue;
      const Box3& cb = box[ch[i].ref];
      for (int k = 0; k < 3; k++) {
        const double lo = ((double) cb.lo[k] - (double) origin[k]) / scale[k];
        const double hi = ((double) cb.hi[k] - (double) origin[k]) / scale[k];
        double ql = std::floor(lo), qh = std::ceil(hi);
        if (lo - ql < 1.0 / 64.0 && ql > 0.0) ql -= 1.0;
        if (qh - hi < 1.0 / 64.0 && qh < 255.0) qh += 1.0;
        if (!(ql >= 0.0)) ql = 0.0;
        if (!(qh <= 255.0)) qh = 255.0;
        q[k][s] = (uint8_t) ql; q[3 + k][s] = (uint8_t) qh;
      }
      if (ch[i].leaf) {
        const uint32_t cnt = tri_count(ch[i].ref);
        std::vector<uint32_t> pos; leaves_of(ch[i].ref, pos);
        meta[s] = (uint8_t) ((((1u << cnt) - 1u) << 5) | triOffset);
        for (uint32_t t = 0; t < cnt; t++)
          std::memcpy(&outTris[(size_t) (triBase + triOffset + t) * 12], &in[(size_t) val[pos[t]] * 12], 48);
        triOffset += cnt;
      } else {
        meta[s] = (uint8_t) (0x20u | (24u + (uint32_t) s));
        imask |= 1u << s;
        next.push_back({ ch[i].ref, childBase + internal });
        internal++;
      }
    }
    uint32_t* w = &outNodes[(size_t) wide * 20];
    std::memcpy(&w[0], origin, 12);
    w[3] = ebits[0] | (ebits[1] << 8) | (ebits[2] << 16) | (imask << 24);
    w[4] = childBase; w[5] = triBase;
    std::memcpy(&w[6], meta, 8);
    std::memcpy(&w[8], q[0], 8); std::memcpy(&w[10], q[1], 8); std::memcpy(&w[12], q[2], 8);
    std::memcpy(&w[14], q[3], 8); std::memcpy(&w[16], q[4], 8); std::memcpy(&w[18], q[5], 8);
  }

  void collapse()
  {
    outTris.assign((size_t) n * 12, 0.0f);
    outNodes.assign(20, 0u);
    std::vector<std::pair<uint32_t, uint32_t>> level = { { 0u, 0u } }, next;
    uint32_t nodeCount = 1, triCount = 0;
    while (!level.empty()) {
      depth++;
      // children of every node of the level first (their counts fix the index ranges), then the nodes
      std::vector<std::vector<Child>> lists(level.size());
      uint32_t nodesHere = 0, trisHere = 0;
      std::vector<uint32_t> nodeOff(level.size()), triOff(level.size());
      for (size_t e = 0; e < level.size(); e++) {
        const uint32_t ref = level[e].first;
        std::vector<Child>& l = lists[e];
        if (is_leaf(ref) || tab[ref].rootIsLeaf) l.push_back({ ref, true });
        else { expand(left[ref], tab[ref].rootSplit, l); expand(right[ref], 8 - tab[ref].rootSplit, l); }
        nodeOff[e] = nodesHere; triOff[e] = trisHere;
        for (const Child& c : l) { if (c.leaf) trisHere += tri_count(c.ref); else nodesHere++; }
      }
      outNodes.resize((size_t) (nodeCount + nodesHere) * 20, 0u);
      next.clear();
      for (size_t e = 0; e < level.size(); e++)
        emit_node(level[e].first, level[e].second, lists[e], nodeCount + nodeOff[e], triCount + triOff[e], next);
      nodeCount += nodesHere; triCount += trisHere;
      level.swap(next);
    }
  }
};

// tris: n x 12 floats in (instance, primitive) order.  Returns the node count; nodes/outTris sized by the caller
// for the worst case (n nodes x 20 words, n x 12 floats).
// mode 1: Karras radix tree (GtlcTuning.bvhBuilder = 1); mode 2: PLOC (bvhBuilder = 2)
inline uint32_t build(const float* tris, uint32_t n, uint32_t* nodes, float* outTris, uint32_t* depth, int mode = 1)
{
  if (n == 0) { std::memset(nodes, 0, 80); if (depth) *depth = 0; return 1; }
  Builder b; b.n = n; b.in = tris;
  b.sort_keys();
  if (n > 1) { if (mode == 2) b.ploc_tree(); else b.radix_tree(); }
  b.fit_and_cost();
  b.collapse();
  std::memcpy(nodes, b.outNodes.data(), b.outNodes.size() * 4);
  std::memcpy(outTris, b.outTris.data(), b.outTris.size() * 4);
  if (depth) *depth = b.depth;
  return (uint32_t) (b.outNodes.size() / 20);
}

} // namespace orc_lbvh8
