// Material-id sort of the hit queue: an 8-bit counting (radix) sort written by hand -- per-CTA histograms in
// shared memory, a per-bin scan of the per-CTA counts, stable scatter with warp-level __match_any ranking.  It produces a
// gather permutation `perm` that the shade kernel walks, so that the lanes of a warp evaluate the same
// closure with the same parameter block.  Takes the place of the reference's optional
// GL_NV_shader_invocation_reorder hint (/root/reference/src/gi/shaders/rp_main.rgen:351-379, reorderThreadNV
// on the closest-hit shader index = material index).
// Key: material index of the hit instance (0..254), 255 for misses and anything beyond 254 materials.
// Stable, so paths keep their queue order inside a material bucket (memory coherence of the state gathers).
#pragma once
#include "wavefront.cuh"

namespace gtlc {

constexpr int kSortBins = 256;
constexpr int kSortBlock = 256;
static_assert(kSortBlock == kSortBins, "the scatter kernel scans one bin total per thread");

GTLC_D uint32_t sort_key(const uint32_t* __restrict__ hitInst, const InstanceInfo* __restrict__ instInfo, uint32_t i)
{
  const uint32_t inst = hitInst[i];
  if (inst == 0xFFFFFFFFu) return 255u;
  const uint32_t m = instInfo[inst].material;
  return m < 255u ? m : 255u;
}

// Pass 1: hist[bin * gridDim.x + block] = number of keys == bin in this CTA's contiguous chunk.
__global__ void __launch_bounds__(kSortBlock) sort_count_kernel(const Counters* ctr, int inIdx, const uint32_t* __restrict__ hitInst,
                                                                const InstanceInfo* __restrict__ instInfo, uint32_t* __restrict__ hist)
{
  __shared__ uint32_t local[kSortBins];
  const uint32_t n = ctr->numPaths[inIdx];
  const uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
  local[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t i = begin + threadIdx.x; i < end; i += kSortBlock) atomicAdd(&local[sort_key(hitInst, instInfo, i)], 1u);
  __syncthreads();
  hist[threadIdx.x * gridDim.x + blockIdx.x] = local[threadIdx.x];
}

// Pass 2: one CTA per bin scans that bin's row of per-CTA counts (exclusive, in place) and leaves the bin's total in
// binTotals[bin]; the 256-entry scan over the totals is done by every scatter CTA for itself.  (A single CTA scanning the
// whole 256 x 592 table took ~100 us per sort: 3 % of a C3 frame, 15 % of an eighth of a C2 frame.)
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t* __restrict__ hist, uint32_t blocks, uint32_t* __restrict__ binTotals)
{
  __shared__ uint32_t warpSums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  uint32_t* row = hist + (size_t) blockIdx.x * blocks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < blocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < blocks ? row[i] : 0u;
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
    if (i < blocks) row[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) binTotals[blockIdx.x] = carry;
}

// Pass 3: stable scatter.  perm[destination] = source queue index.
__global__ void __launch_bounds__(kSortBlock) sort_scatter_kernel(const Counters* ctr, int inIdx, const uint32_t* __restrict__ hitInst,
                                                                  const InstanceInfo* __restrict__ instInfo, const uint32_t* __restrict__ offsets,
                                                                  const uint32_t* __restrict__ binTotals, uint32_t* __restrict__ perm)
{
  __shared__ uint32_t base[kSortBins];
  __shared__ uint32_t binScan[kSortBins / 32];
  __shared__ uint32_t warpCnt[kSortBlock / 32][kSortBins];
  const unsigned FULL = 0xFFFFFFFFu;
  const uint32_t n = ctr->numPaths[inIdx];
  const uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {                                                    // first destination of bin `threadIdx.x` = exclusive scan of the 256 bin totals (kSortBlock == kSortBins)
    const uint32_t v = binTotals[threadIdx.x];
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL, x, d); if (lane >= d) x += y; }
    if (lane == 31) binScan[warp] = x;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; w++) before += binScan[w];
    base[threadIdx.x] = (before + (x - v)) + offsets[threadIdx.x * gridDim.x + blockIdx.x];
  }
  for (uint32_t tile = begin; tile < end; tile += kSortBlock) {
#pragma unroll
    for (int w = 0; w < kSortBlock / 32; w++) warpCnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = tile + threadIdx.x;
    const bool valid = i < end;
    const uint32_t key = valid ? sort_key(hitInst, instInfo, i) : 0xFFFFu;
    const unsigned peers = __match_any_sync(FULL, key);
    const uint32_t rankInWarp = (uint32_t) __popc(peers & ((1u << lane) - 1u));
    if (valid && rankInWarp == 0) warpCnt[warp][key] = (uint32_t) __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t before = 0;
      for (int w = 0; w < warp; w++) before += warpCnt[w][key];
      perm[base[key] + before + rankInWarp] = i;
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kSortBlock / 32; w++) total += warpCnt[w][threadIdx.x];
    base[threadIdx.x] += total;
    __syncthreads();
  }
}

} // namespace gtlc
