// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Scalar CPU walker over the PRODUCT's compressed BVH8 bytes (layout: gatling_b200/csrc/bvh8.h).
// Purpose: (1) check on the CPU that the host-built BVH8 returns the same closest hits as the oracle's
// own binary BVH; (2) produce the exact per-ray node / triangle visit counts that define the
// algorithmic bytes of the traversal roofline (32 + 20 + 80*n_nodes + 48*n_tris), by
// restating the kernel's traversal order: children in octant order (slot ^ octinv, highest first),
// siblings deferred as one stack entry, a node's triangles tested right after its slab tests in
// ascending slot order, no re-test of a deferred child's own box.
#pragma once

namespace orc {

static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline float ldf(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

static inline uint32_t bvh8_node_hitmask(const uint8_t* nd, V3 o, V3 d, V3 idir, float tmin, float tbest, uint32_t octinv)
{
  const uint32_t e = ld32(nd + 12);
  const float sx = uintBitsToFloat((e & 0xFFu) << 23) * idir.x;
  const float sy = uintBitsToFloat(((e >> 8) & 0xFFu) << 23) * idir.y;
  const float sz = uintBitsToFloat(((e >> 16) & 0xFFu) << 23) * idir.z;
  // slab origin with the kernel's 2^15 byte bias folded in: t = (32768 + q) * s + (o - 32768 s)
  const float ox = (ldf(nd + 0) - o.x) * idir.x - 32768.0f * sx;
  const float oy = (ldf(nd + 4) - o.y) * idir.y - 32768.0f * sy;
  const float oz = (ldf(nd + 8) - o.z) * idir.z - 32768.0f * sz;
  const uint8_t* meta = nd + 24;
  const uint8_t* qlox = nd + 32; const uint8_t* qloy = nd + 40; const uint8_t* qloz = nd + 48;
  const uint8_t* qhix = nd + 56; const uint8_t* qhiy = nd + 64; const uint8_t* qhiz = nd + 72;
  uint32_t hitmask = 0;
  for (int i = 0; i < 8; i++) {
    const float nx = 32768.0f + (float) (d.x < 0.0f ? qhix[i] : qlox[i]), fx = 32768.0f + (float) (d.x < 0.0f ? qlox[i] : qhix[i]);
    const float ny = 32768.0f + (float) (d.y < 0.0f ? qhiy[i] : qloy[i]), fy = 32768.0f + (float) (d.y < 0.0f ? qloy[i] : qhiy[i]);
    const float nz = 32768.0f + (float) (d.z < 0.0f ? qhiz[i] : qloz[i]), fz = 32768.0f + (float) (d.z < 0.0f ? qloz[i] : qhiz[i]);
    const float tnx = fmaf(nx, sx, ox), tny = fmaf(ny, sy, oy), tnz = fmaf(nz, sz, oz);
    const float tfx = fmaf(fx, sx, ox), tfy = fmaf(fy, sy, oy), tfz = fmaf(fz, sz, oz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tbest)) * 1.0000004f;
    const uint32_t m = meta[i];
    if (tn <= tf && m != 0u) {
      const uint32_t bits = m >> 5, lo5 = m & 31u;
      const uint32_t idx = (lo5 >= 24u) ? (24u + ((lo5 & 7u) ^ octinv)) : lo5;
      hitmask |= bits << idx;
    }
  }
  return hitmask;
}

static inline void bvh8_trace_one(const uint8_t* nodes, const float* tris, V3 o, V3 d, float tmin, float tmax, bool anyHit,
                                  Hit& best, uint32_t& nNodes, uint32_t& nTris, uint32_t seed = 0)
{
  best.t = -1.0f; best.u = best.v = 0.0f; best.prim = best.inst = 0xFFFFFFFFu;
  nNodes = nTris = 0;
  if (!(tmax > tmin)) return;
  {  // the kernel answers non-finite rays as misses without traversal (they cannot hit a triangle)
    const float c[6] = { o.x, o.y, o.z, d.x, d.y, d.z };
    for (int k = 0; k < 6; k++) if (!(fabsf(c[k]) <= FLOAT_MAX)) return;
  }
  RayPre rp = ray_precompute(d);
  const float kMinDir = 8.271806125530277e-25f;   // 2^-80, slab tests only (mirrors the kernel)
  V3 idir = v3(1.0f / (fabsf(d.x) > kMinDir ? d.x : copysignf(kMinDir, d.x)),
               1.0f / (fabsf(d.y) > kMinDir ? d.y : copysignf(kMinDir, d.y)),
               1.0f / (fabsf(d.z) > kMinDir ? d.z : copysignf(kMinDir, d.z)));
  const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
  const uint32_t octinv = 7u - oct;
  float bt = tmax;
  uint32_t stackX[64], stackY[64]; int sp = 0;
  uint32_t Gx = 0, Gy = 0x80000000u;
  while (true) {
    uint32_t Tx = 0, Ty = 0;
    if (Gy & 0xFF000000u) {
      const uint32_t hits_imask = Gy;
      uint32_t bit = 31; while (!(hits_imask & (1u << bit))) bit--;
      Gy &= ~(1u << bit);
      if (Gy & 0xFF000000u) { stackX[sp] = Gx; stackY[sp] = Gy; sp++; }
      const uint32_t slot = (bit - 24u) ^ octinv;
      const uint32_t below = (slot == 0) ? 0u : (hits_imask & ((1u << slot) - 1u));
      const uint32_t rel = (uint32_t) __builtin_popcount(below);
      const uint8_t* nd = nodes + (size_t) (Gx + rel) * 80;
      nNodes++;
      const uint32_t hm = bvh8_node_hitmask(nd, o, d, idir, tmin, bt, octinv);
      Gx = ld32(nd + 16); Gy = (hm & 0xFF000000u) | (ld32(nd + 12) >> 24);
      Tx = ld32(nd + 20); Ty = hm & 0x00FFFFFFu;
    }
    bool done = false;
    while (Ty) {
      const uint32_t j = (uint32_t) __builtin_ctz(Ty);
      Ty &= Ty - 1u;
      const float* tp = tris + (size_t) (Tx + j) * 12;
      nTris++;
      float t, u, v;
      if (tri_intersect(rp, o, v3(tp[0], tp[1], tp[2]), v3(tp[4], tp[5], tp[6]), v3(tp[8], tp[9], tp[10]), t, u, v)) {
        if (t > tmin && t < tmax && !cutout_ignored(seed, tp[11], floatBitsToUint(tp[3]), floatBitsToUint(tp[7]))) {
          const uint32_t prim = floatBitsToUint(tp[3]), inst = floatBitsToUint(tp[7]);
          if (anyHit) { best.t = t; best.u = u; best.v = v; best.prim = prim; best.inst = inst; done = true; break; }
          if (hit_better(t, inst, prim, best)) { best.t = t; best.u = u; best.v = v; best.prim = prim; best.inst = inst; bt = t; }
        }
      }
    }
    if (done) return;
    if (!(Gy & 0xFF000000u)) {
      if (sp == 0) return;
      sp--; Gx = stackX[sp]; Gy = stackY[sp];
    }
  }
}

} // namespace orc

extern "C" void orc_bvh8_trace(const void* nodes, uint32_t nodeCount, const void* tris, uint32_t triCount, const float* rays, uint32_t n,
                               int anyHit, void* hitsOut, void* visitsOut)
{
  (void) nodeCount; (void) triCount;
  orc::Hit* hits = (orc::Hit*) hitsOut;
  uint32_t* visits = (uint32_t*) visitsOut;
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < (int) n; i++) {
    const float* r = &rays[(size_t) i * 8];
    orc::Hit h; uint32_t nn, nt;
    orc::bvh8_trace_one((const uint8_t*) nodes, (const float*) tris, orc::v3(r[0], r[1], r[2]), orc::v3(r[4], r[5], r[6]), r[3], r[7], anyHit != 0, h, nn, nt);
    hits[i] = h;
    if (visits) { visits[2 * (size_t) i] = nn; visits[2 * (size_t) i + 1] = nt; }
  }
}

// Analysis helper (not used by any test): the same walk, but a node's triangle group is only tested after `delay`
// further node steps of that ray -- how many extra node visits does late culling cost?  Guides kernel design.
extern "C" void orc_bvh8_trace_delayed(const void* nodesV, const void* trisV, const float* rays, uint32_t n, int delay, uint64_t* outCounts)
{
  using namespace orc;
  const uint8_t* nodes = (const uint8_t*) nodesV; const float* tris = (const float*) trisV;
  uint64_t totN = 0, totT = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : totN, totT)
  for (int i = 0; i < (int) n; i++) {
    const float* rr = &rays[(size_t) i * 8];
    V3 o = v3(rr[0], rr[1], rr[2]), d = v3(rr[4], rr[5], rr[6]); float tmin = rr[3], tmax = rr[7];
    RayPre rp = ray_precompute(d);
    const float kMinDir = 8.271806125530277e-25f;
    V3 idir = v3(1.0f / (fabsf(d.x) > kMinDir ? d.x : copysignf(kMinDir, d.x)), 1.0f / (fabsf(d.y) > kMinDir ? d.y : copysignf(kMinDir, d.y)), 1.0f / (fabsf(d.z) > kMinDir ? d.z : copysignf(kMinDir, d.z)));
    const uint32_t octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
    float bt = tmax; Hit best; best.prim = best.inst = 0xFFFFFFFFu; best.t = -1;
    uint32_t stackX[64], stackY[64]; int sp = 0;
    uint32_t Gx = 0, Gy = 0x80000000u;
    struct Pend { uint32_t x, y; uint64_t due; } fifo[256]; int fh = 0, ft = 0; uint64_t step = 0;
    auto testGroup = [&](uint32_t Tx, uint32_t Ty) {
      while (Ty) {
        const uint32_t j = (uint32_t) __builtin_ctz(Ty); Ty &= Ty - 1u;
        const float* tp = tris + (size_t) (Tx + j) * 12; totT++;
        float t, u, v;
        if (tri_intersect(rp, o, v3(tp[0], tp[1], tp[2]), v3(tp[4], tp[5], tp[6]), v3(tp[8], tp[9], tp[10]), t, u, v) && t > tmin && t < tmax) {
          const uint32_t prim = floatBitsToUint(tp[3]), inst = floatBitsToUint(tp[7]);
          if (hit_better(t, inst, prim, best)) { best.t = t; best.prim = prim; best.inst = inst; bt = t; }
        }
      }
    };
    while (true) {
      while (fh != ft && fifo[fh & 255].due <= step) { testGroup(fifo[fh & 255].x, fifo[fh & 255].y); fh++; }
      if (Gy & 0xFF000000u) {
        const uint32_t hits_imask = Gy;
        uint32_t bit = 31; while (!(hits_imask & (1u << bit))) bit--;
        Gy &= ~(1u << bit);
        if (Gy & 0xFF000000u) { stackX[sp] = Gx; stackY[sp] = Gy; sp++; }
        const uint32_t slot = (bit - 24u) ^ octinv;
        const uint32_t below = (slot == 0) ? 0u : (hits_imask & ((1u << slot) - 1u));
        const uint8_t* nd = nodes + (size_t) (Gx + (uint32_t) __builtin_popcount(below)) * 80;
        totN++; step++;
        const uint32_t hm = bvh8_node_hitmask(nd, o, d, idir, tmin, bt, octinv);
        Gx = ld32(nd + 16); Gy = (hm & 0xFF000000u) | (ld32(nd + 12) >> 24);
        if (hm & 0x00FFFFFFu) { fifo[ft & 255] = { ld32(nd + 20), hm & 0x00FFFFFFu, step + (uint64_t) delay }; ft++; }
      }
      if (!(Gy & 0xFF000000u)) {
        if (sp == 0) { while (fh != ft) { testGroup(fifo[fh & 255].x, fifo[fh & 255].y); fh++; } break; }
        sp--; Gx = stackX[sp]; Gy = stackY[sp];
      }
    }
  }
  outCounts[0] = totN; outCounts[1] = totT;
}

// Analysis helper (not a restatement of any kernel): the same BVH8 walked in exact near-to-far order -- every hit child of
// a node, internal or leaf, is pushed with its entry distance and the closest pending entry is processed next.  Gives the
// node-visit / triangle-test counts an ideally ordered traversal would need; outCounts = {nodes, triangles}.
// mode 0: global priority queue (ideal order); mode 1: a stack, the hit children of each node pushed far-to-near (what a
// kernel that sorts the <= 8 hits of a node would do); mode 2: a stack, children pushed in the kernel's octant order
// (slot ^ octinv, highest first) but leaf children deferred like internal ones; mode 3: octant order for internal children,
// the node's leaf triangles tested at once (= the production order, as a cross-check of this helper).  Triangles are tested
// when their leaf entry is popped.
extern "C" void orc_bvh8_trace_ordered(const void* nodesV, const void* trisV, const float* rays, uint32_t n, uint64_t* outCounts, int mode)
{
  using namespace orc;
  const uint8_t* nodes = (const uint8_t*) nodesV; const float* tris = (const float*) trisV;
  uint64_t totalNodes = 0, totalTris = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : totalNodes, totalTris)
  for (int64_t r = 0; r < (int64_t) n; r++) {
    const float* ry = rays + r * 8;
    const V3 o = v3(ry[0], ry[1], ry[2]), d = v3(ry[4], ry[5], ry[6]);
    const float tmin = ry[3], tmax = ry[7];
    if (!(tmax > tmin)) continue;
    RayPre rp = ray_precompute(d);
    const float kMinDir = 8.271806125530277e-25f;
    const V3 idir = v3(1.0f / (fabsf(d.x) > kMinDir ? d.x : copysignf(kMinDir, d.x)), 1.0f / (fabsf(d.y) > kMinDir ? d.y : copysignf(kMinDir, d.y)),
                       1.0f / (fabsf(d.z) > kMinDir ? d.z : copysignf(kMinDir, d.z)));
    struct Entry { float tn; uint32_t ref; };      // ref: bit 31 = triangle run (base | count << 24), else node index
    Entry heap[512]; int hn = 0;
    const uint32_t octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
    auto push = [&](float tn, uint32_t ref) { int i = hn++; heap[i] = { tn, ref }; while (i > 0 && heap[(i - 1) / 2].tn > heap[i].tn) { std::swap(heap[i], heap[(i - 1) / 2]); i = (i - 1) / 2; } };
    auto pop = [&]() { Entry top = heap[0]; heap[0] = heap[--hn]; int i = 0; while (true) { int l = 2 * i + 1, rr = l + 1, m = i; if (l < hn && heap[l].tn < heap[m].tn) m = l; if (rr < hn && heap[rr].tn < heap[m].tn) m = rr; if (m == i) break; std::swap(heap[i], heap[m]); i = m; } return top; };
    float bt = tmax;
    push(tmin, 0u);
    while (hn > 0) {
      Entry cur;
      if (mode == 0) { cur = pop(); if (cur.tn > bt) break; }
      else { cur = heap[--hn]; if (cur.tn > bt) continue; }
      if (cur.ref & 0x80000000u) {
        const uint32_t base = cur.ref & 0x00FFFFFFu, cnt = (cur.ref >> 24) & 0x7Fu;
        for (uint32_t k = 0; k < cnt; k++) {
          const float* tp = tris + (size_t) (base + k) * 12;
          totalTris++;
          float t, u, v;
          if (tri_intersect(rp, o, v3(tp[0], tp[1], tp[2]), v3(tp[4], tp[5], tp[6]), v3(tp[8], tp[9], tp[10]), t, u, v) && t > tmin && t < bt) bt = t;
        }
        continue;
      }
      const uint8_t* nd = nodes + (size_t) cur.ref * 80;
      totalNodes++;
      const uint32_t e = ld32(nd + 12);
      const float sc[3] = { uintBitsToFloat((e & 0xFFu) << 23), uintBitsToFloat(((e >> 8) & 0xFFu) << 23), uintBitsToFloat(((e >> 16) & 0xFFu) << 23) };
      const float org[3] = { ldf(nd + 0), ldf(nd + 4), ldf(nd + 8) }, oo[3] = { o.x, o.y, o.z }, id[3] = { idir.x, idir.y, idir.z };
      const uint32_t imask = e >> 24, childBase = ld32(nd + 16), triBase = ld32(nd + 20);
      Entry local[8]; float key[8]; int nl = 0;
      for (int i = 0; i < 8; i++) {
        const uint32_t m = nd[24 + i];
        if (!m) continue;
        float tn = tmin, tf = bt;
        for (int a = 0; a < 3; a++) {
          const float lo = (org[a] + (float) nd[32 + 8 * a + i] * sc[a] - oo[a]) * id[a], hi = (org[a] + (float) nd[56 + 8 * a + i] * sc[a] - oo[a]) * id[a];
          tn = fmaxf(tn, fminf(lo, hi)); tf = fminf(tf, fmaxf(lo, hi));
        }
        if (tn > tf * 1.0000004f) continue;
        const uint32_t ref = ((m & 31u) >= 24u) ? childBase + (uint32_t) __builtin_popcount(imask & ((1u << i) - 1u))
                                                : (0x80000000u | ((uint32_t) __builtin_popcount(m >> 5) << 24) | (triBase + (m & 31u)));
        if (mode == 0) push(tn, ref);
        else if (mode == 3 && (ref & 0x80000000u)) {           // immediate triangle tests
          const uint32_t base = ref & 0x00FFFFFFu, cnt = (ref >> 24) & 0x7Fu;
          for (uint32_t k = 0; k < cnt; k++) {
            const float* tp = tris + (size_t) (base + k) * 12;
            totalTris++;
            float t, u, v;
            if (tri_intersect(rp, o, v3(tp[0], tp[1], tp[2]), v3(tp[4], tp[5], tp[6]), v3(tp[8], tp[9], tp[10]), t, u, v) && t > tmin && t < bt) bt = t;
          }
        }
        else { local[nl].tn = (mode == 1) ? tn : -(float) ((uint32_t) i ^ octinv); local[nl].ref = ref; key[nl] = tn; nl++; }
      }
      if (mode != 0) {                       // lowest priority first onto the stack
        int ord[8]; for (int q = 0; q < nl; q++) ord[q] = q;
        for (int a = 1; a < nl; a++) { const int v = ord[a]; int b = a - 1; while (b >= 0 && local[ord[b]].tn < local[v].tn) { ord[b + 1] = ord[b]; b--; } ord[b + 1] = v; }   // descending priority key
        for (int q = 0; q < nl; q++) { heap[hn].tn = key[ord[q]]; heap[hn].ref = local[ord[q]].ref; hn++; }
      }
    }
    if (tmax > tmin) { totalNodes += 0; }
  }
  outCounts[0] = totalNodes; outCounts[1] = totalTris;
}

// Analysis helper: the production walk (octant order, node groups on the stack, a node's triangles tested at once) with ONE
// extra float per stack entry -- the smallest entry distance among the group's remaining children.  A popped group whose
// bound lies beyond the current best hit is dropped without fetching any of its nodes.  outCounts = {nodes, triangles}.
// mode 0: bound = smallest entry distance of the children still in the group; mode 1: bound = smallest entry distance of all
// internal children the node step found (computed once, never updated -- cheaper in a kernel, weaker).
extern "C" void orc_bvh8_trace_groupcull(const void* nodesV, const void* trisV, const float* rays, uint32_t n, uint64_t* outCounts, int mode)
{
  using namespace orc;
  const uint8_t* nodes = (const uint8_t*) nodesV; const float* tris = (const float*) trisV;
  uint64_t totalNodes = 0, totalTris = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : totalNodes, totalTris)
  for (int64_t r = 0; r < (int64_t) n; r++) {
    const float* ry = rays + r * 8;
    const V3 o = v3(ry[0], ry[1], ry[2]), d = v3(ry[4], ry[5], ry[6]);
    const float tmin = ry[3], tmax = ry[7];
    if (!(tmax > tmin)) continue;
    RayPre rp = ray_precompute(d);
    const float kMinDir = 8.271806125530277e-25f;
    const V3 idir = v3(1.0f / (fabsf(d.x) > kMinDir ? d.x : copysignf(kMinDir, d.x)), 1.0f / (fabsf(d.y) > kMinDir ? d.y : copysignf(kMinDir, d.y)),
                       1.0f / (fabsf(d.z) > kMinDir ? d.z : copysignf(kMinDir, d.z)));
    const uint32_t octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
    struct Group { uint32_t base, imask, hits; float tn[8]; float minAll, minRest; };       // hits: bit (s ^ octinv) per hit internal child in slot s; tn by slot
    Group stack[64]; int sp = 0;
    float bt = tmax;
    Group G; G.base = 0; G.imask = 1; G.hits = 1u << (0 ^ octinv); for (int k = 0; k < 8; k++) G.tn[k] = tmin;
    G.minAll = tmin; G.minRest = tmin;
    bool haveG = true;
    while (true) {
      if (!haveG) {
        if (sp == 0) break;
        G = stack[--sp]; haveG = true;
        float mn = 3.4e38f;                                          // the one float a real stack entry would carry
        for (int k = 0; k < 8; k++) if (G.hits & (1u << (k ^ octinv))) mn = fminf(mn, G.tn[k]);
        if (mode == 1) mn = G.minAll;
        if (mode == 2) mn = G.minRest;
        if (mn > bt) { haveG = false; continue; }
      }
      uint32_t bit = 31; while (!(G.hits & (1u << bit))) bit--;
      G.hits &= ~(1u << bit);
      const uint32_t slot = bit ^ octinv;
      const uint32_t node = G.base + (uint32_t) __builtin_popcount(G.imask & ((1u << slot) - 1u));
      if (G.hits) {
        if (G.minRest < 0.0f) {   // first push of this group: one float = min over the children that stay behind (mode 2)
          float m2 = 3.4e38f; for (int k = 0; k < 8; k++) if (G.hits & (1u << (k ^ octinv))) m2 = fminf(m2, G.tn[k]);
          G.minRest = m2;
        }
        stack[sp++] = G;
      }
      haveG = false;
      const uint8_t* nd = nodes + (size_t) node * 80;
      totalNodes++;
      const uint32_t e = ld32(nd + 12);
      const float sc[3] = { uintBitsToFloat((e & 0xFFu) << 23), uintBitsToFloat(((e >> 8) & 0xFFu) << 23), uintBitsToFloat(((e >> 16) & 0xFFu) << 23) };
      const float org[3] = { ldf(nd + 0), ldf(nd + 4), ldf(nd + 8) }, oo[3] = { o.x, o.y, o.z }, id[3] = { idir.x, idir.y, idir.z };
      Group C; C.base = ld32(nd + 16); C.imask = e >> 24; C.hits = 0; C.minAll = 3.4e38f; C.minRest = -1.0f;
      const uint32_t triBase = ld32(nd + 20);
      for (int i = 0; i < 8; i++) {
        const uint32_t m = nd[24 + i];
        C.tn[i] = 3.4e38f;
        if (!m) continue;
        float tn = tmin, tf = bt;
        for (int a = 0; a < 3; a++) {
          const float lo = (org[a] + (float) nd[32 + 8 * a + i] * sc[a] - oo[a]) * id[a], hi = (org[a] + (float) nd[56 + 8 * a + i] * sc[a] - oo[a]) * id[a];
          tn = fmaxf(tn, fminf(lo, hi)); tf = fminf(tf, fmaxf(lo, hi));
        }
        if (tn > tf * 1.0000004f) continue;
        if ((m & 31u) >= 24u) { C.hits |= 1u << ((uint32_t) i ^ octinv); C.tn[i] = tn; C.minAll = fminf(C.minAll, tn); }
        else {
          const uint32_t base = triBase + (m & 31u), cnt = (uint32_t) __builtin_popcount(m >> 5);
          for (uint32_t k = 0; k < cnt; k++) {
            const float* tp = tris + (size_t) (base + k) * 12;
            totalTris++;
            float t, u, v;
            if (tri_intersect(rp, o, v3(tp[0], tp[1], tp[2]), v3(tp[4], tp[5], tp[6]), v3(tp[8], tp[9], tp[10]), t, u, v) && t > tmin && t < bt) bt = t;
          }
        }
      }
      if (C.hits) { G = C; haveG = true; }
    }
  }
  outCounts[0] = totalNodes; outCounts[1] = totalTris;
}

// Analysis helper: of the nodes the production walk fetches, how many could be rejected by a slab test against the node's
// own quantisation grid [p, p + 255 * 2^e] (available from the first 16 bytes of the node) before any child is looked at?
// outCounts = {node visits, of which the grid box is missed or starts beyond the current best hit, of which no child is hit,
// of which even the union of the children's boxes is missed or beyond the best hit, of which that union is missed
// whatever the best hit (quantisation slack of the parent's box for this node)}.
extern "C" void orc_bvh8_grid_earlyout(const void* nodesV, const void* trisV, const float* rays, uint32_t n, uint64_t* outCounts)
{
  using namespace orc;
  const uint8_t* nodes = (const uint8_t*) nodesV; const float* tris = (const float*) trisV;
  uint64_t visits = 0, gridCulled = 0, emptyHit = 0, unionMissed = 0, unionMissedGeom = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : visits, gridCulled, emptyHit, unionMissed, unionMissedGeom)
  for (int64_t r = 0; r < (int64_t) n; r++) {
    const float* ry = rays + r * 8;
    const V3 o = v3(ry[0], ry[1], ry[2]), d = v3(ry[4], ry[5], ry[6]);
    const float tmin = ry[3], tmax = ry[7];
    if (!(tmax > tmin)) continue;
    RayPre rp = ray_precompute(d);
    const float kMinDir = 8.271806125530277e-25f;
    const V3 idir = v3(1.0f / (fabsf(d.x) > kMinDir ? d.x : copysignf(kMinDir, d.x)), 1.0f / (fabsf(d.y) > kMinDir ? d.y : copysignf(kMinDir, d.y)),
                       1.0f / (fabsf(d.z) > kMinDir ? d.z : copysignf(kMinDir, d.z)));
    const uint32_t octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
    float bt = tmax;
    uint32_t stackX[64], stackY[64]; int sp = 0;
    uint32_t Gx = 0, Gy = 0x80000000u;
    while (true) {
      uint32_t Tx = 0, Ty = 0;
      if (Gy & 0xFF000000u) {
        const uint32_t hits_imask = Gy;
        uint32_t bit = 31; while (!(hits_imask & (1u << bit))) bit--;
        Gy &= ~(1u << bit);
        if (Gy & 0xFF000000u) { stackX[sp] = Gx; stackY[sp] = Gy; sp++; }
        const uint32_t slot = (bit - 24u) ^ octinv;
        const uint32_t below = (slot == 0) ? 0u : (hits_imask & ((1u << slot) - 1u));
        const uint8_t* nd = nodes + (size_t) (Gx + (uint32_t) __builtin_popcount(below)) * 80;
        visits++;
        {  // the early-out candidate
          const uint32_t e = ld32(nd + 12);
          const float sc[3] = { uintBitsToFloat((e & 0xFFu) << 23), uintBitsToFloat(((e >> 8) & 0xFFu) << 23), uintBitsToFloat(((e >> 16) & 0xFFu) << 23) };
          const float org[3] = { ldf(nd + 0), ldf(nd + 4), ldf(nd + 8) }, oo[3] = { o.x, o.y, o.z }, id[3] = { idir.x, idir.y, idir.z };
          float tn = tmin, tf = bt;
          for (int a = 0; a < 3; a++) {
            const float lo = (org[a] - oo[a]) * id[a], hi = (org[a] + 255.0f * sc[a] - oo[a]) * id[a];
            tn = fmaxf(tn, fminf(lo, hi)); tf = fminf(tf, fmaxf(lo, hi));
          }
          if (tn > tf * 1.0000004f) gridCulled++;
        }
        const uint32_t hm = bvh8_node_hitmask(nd, o, d, idir, tmin, bt, octinv);
        if (hm == 0) emptyHit++;
        {  // does the ray at least cross the union of the children's boxes (the node's real extent)?
          const uint32_t e = ld32(nd + 12);
          const float sc[3] = { uintBitsToFloat((e & 0xFFu) << 23), uintBitsToFloat(((e >> 8) & 0xFFu) << 23), uintBitsToFloat(((e >> 16) & 0xFFu) << 23) };
          const float org[3] = { ldf(nd + 0), ldf(nd + 4), ldf(nd + 8) }, oo[3] = { o.x, o.y, o.z }, id[3] = { idir.x, idir.y, idir.z };
          float tn = tmin, tf = bt, tf0 = 3e38f;
          for (int a = 0; a < 3; a++) {
            uint32_t qlo = 255, qhi = 0;
            for (int i = 0; i < 8; i++) if (nd[24 + i]) { qlo = std::min<uint32_t>(qlo, nd[32 + 8 * a + i]); qhi = std::max<uint32_t>(qhi, nd[56 + 8 * a + i]); }
            const float lo = (org[a] + (float) qlo * sc[a] - oo[a]) * id[a], hi = (org[a] + (float) qhi * sc[a] - oo[a]) * id[a];
            tn = fmaxf(tn, fminf(lo, hi)); tf = fminf(tf, fmaxf(lo, hi)); tf0 = fminf(tf0, fmaxf(lo, hi));
          }
          if (tn > tf * 1.0000004f) unionMissed++;
          if (tn > fminf(tf0, 3e38f) * 1.0000004f) unionMissedGeom++;
        }
        Gx = ld32(nd + 16); Gy = (hm & 0xFF000000u) | (ld32(nd + 12) >> 24);
        Tx = ld32(nd + 20); Ty = hm & 0x00FFFFFFu;
      }
      while (Ty) {
        const uint32_t j = (uint32_t) __builtin_ctz(Ty);
        Ty &= Ty - 1u;
        const float* tp = tris + (size_t) (Tx + j) * 12;
        float t, u, v;
        if (tri_intersect(rp, o, v3(tp[0], tp[1], tp[2]), v3(tp[4], tp[5], tp[6]), v3(tp[8], tp[9], tp[10]), t, u, v) && t > tmin && t < bt) bt = t;
      }
      if (!(Gy & 0xFF000000u)) {
        if (sp == 0) break;
        sp--; Gx = stackX[sp]; Gy = stackY[sp];
      }
    }
  }
  outCounts[0] = visits; outCounts[1] = gridCulled; outCounts[2] = emptyHit; outCounts[3] = unionMissed; outCounts[4] = unionMissedGeom;
}
