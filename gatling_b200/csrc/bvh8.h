// Compressed 8-wide BVH ("BVH8"): node layout and host builder interface.
//
// Replaces the driver-internal acceleration structure the reference builds through
// cgpuCreateBlas/cgpuCreateTlas (/root/reference/src/cgpu/impl/CgpuVk.cpp:2561-2854; inputs assembled
// in /root/reference/src/gi/impl/Gi.cpp:784-1315).  Instances are flattened: every instance's
// triangles are pre-transformed to world space and live in ONE BVH8 (180 GB of HBM makes the
// two-level indirection unnecessary), so the traversal kernel never transforms rays.
//
// Node = 80 bytes = 5 x 16-byte words (layout after Ylitie, Karras, Laine 2017, "Efficient Incoherent
// Ray Traversal on GPUs Through Compressed Wide BVHs"):
//   w0: p.x p.y p.z  (float32 origin of the quantisation grid)
//       u32 { ex, ey, ez, imask }      ex/ey/ez = biased float exponents of the grid scale 2^e,
//                                      imask bit s = child slot s is an internal node
//   w1: u32 childBase  (index of the first internal child; internal children are consecutive in slot order)
//       u32 triBase    (index of the first triangle referenced by this node's leaf children)
//       u8 meta[8]     internal child: 0b001_11000 + slot  (= 0x20 | (24 + slot))
//                      leaf child:     (unary triangle count << 5) | offset from triBase   (count 1..3)
//                      empty slot:     0
//   w2: u8 qlo_x[8], u8 qlo_y[8]
//   w3: u8 qlo_z[8], u8 qhi_x[8]
//   w4: u8 qhi_y[8], u8 qhi_z[8]
// child box = p + q * 2^e, quantised conservatively (floor/ceil, plus a safety quantum when a bound
// falls on a grid line).
//
// Triangle = 48 bytes = 3 x float4: (v0.xyz, bits primitiveId) (v1.xyz, bits instanceIndex) (v2.xyz, cutout opacity; 1 = opaque).
#pragma once
#include <stdint.h>
#include <memory>
#include <new>
#include <utility>
#include <vector>

namespace gtlc {

// std::vector whose resize() leaves new trivially-constructible elements untouched: the arrays below run to hundreds of
// megabytes, every element is written by whoever sized them, and the pages should be first touched by the threads that
// fill them rather than zeroed by the one that called resize().
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
  DefaultInitAllocator() = default;
  template <class U> DefaultInitAllocator(const DefaultInitAllocator<U>&) {}
  template <class U> void construct(U* p) { ::new ((void*) p) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*) p) U(std::forward<A>(a)...); }
};
template <class T> using RawVector = std::vector<T, DefaultInitAllocator<T>>;

constexpr uint32_t kBvh8MaxDepth = 48;   // wide-tree depth the traversal kernels' stacks hold (traverse.cuh: kStackSize)

struct Bvh8Node { uint32_t w[20]; };
static_assert(sizeof(Bvh8Node) == 80, "BVH8 node must be 80 bytes");

struct Bvh8 {
  RawVector<Bvh8Node> nodes;
  RawVector<float> tris;        // 12 floats per triangle, leaf order
  uint32_t maxDepth = 0;        // depth of the wide tree (bounds the traversal stack)
  double buildSeconds = 0.0;
};

struct BuildTri { float v[3][3]; uint32_t prim, inst; float opacity; };   // opacity < 1: stochastic cutout (rp_main.ahit)

// Starts the OpenMP team the builder will use and keeps it busy for a fraction of a millisecond.  Some container
// schedulers leave freshly created threads on their creator's core for hundreds of milliseconds; a team that was
// created when the first scene was (long before its first build) has been spread over the cores by then.
void bvh8_warm_up_threads();

// Builds the BVH8 over `count` world-space triangles.  Deterministic for a given input, whatever the thread count.
void bvh8_build(const BuildTri* tris, size_t count, Bvh8& out);

} // namespace gtlc
