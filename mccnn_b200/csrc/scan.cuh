// scan.cuh -- two-level exclusive prefix sum over int32 with a device-resident element count.
//
// Replaces the reference's Hillis-Steele smem scans + atomic second level
// (propagate_offsets tf_ops/sort_gpu.cu:87-146, computeOffsets tf_ops/find_neighbors.cu:122-182).
// Level 1: one CTA per tile of kScanTile items writes the tile-local exclusive prefix in place and the tile
// total.  Level 2: one CTA scans the tile totals.  Consumers fold the two levels with scan_prefix(), the way the
// reference's determine_new_index / findNeighbors add their block offsets, so no third pass over the data exists.
#pragma once
#include "common.cuh"

namespace mc {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

__device__ __forceinline__ int scan_prefix(const int* __restrict__ local, const int* __restrict__ tileOffs, int i) {
    return local[i] + tileOffs[i / kScanTile];
}

// CTA-wide exclusive scan of one int per thread; returns the exclusive prefix, total through `total`.
template <int THREADS>
__device__ __forceinline__ int block_exclusive_scan(int v, int& total, int* smemWarp /* THREADS/32 + 1 ints */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) smemWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < THREADS / 32 ? smemWarp[lane] : 0;
        int winc = warp_inclusive_scan(w);
        if (lane < THREADS / 32) smemWarp[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smemWarp[THREADS / 32] = winc;       // grand total
    }
    __syncthreads();
    int res = inc - v + smemWarp[warp];
    total = smemWarp[THREADS / 32];
    __syncthreads();  // smemWarp reusable by the caller afterwards
    return res;
}

// level 1: data[i] <- exclusive prefix inside its tile; tileSums[tile] <- tile total.
static __global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(int* __restrict__ data,
                                                                   const int* __restrict__ dCount, int capacity,
                                                                   int* __restrict__ tileSums) {
    __shared__ int sWarp[kScanThreads / 32 + 1];
    const int n = live_count(dCount, capacity);
    const int tileBase = blockIdx.x * kScanTile;
    if (tileBase >= n) return;
    const int base = tileBase + threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        sum += v[k];
    }
    int total;
    int excl = block_exclusive_scan<kScanThreads>(sum, total, sWarp);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

// level 2: tileSums <- exclusive scan of the live tile totals (single CTA); *dTotal <- grand total.
static __global__ void __launch_bounds__(1024) scan_sums_kernel(int* __restrict__ tileSums, const int* __restrict__ dCount,
                                                         int capacity, int* __restrict__ dTotal) {
    __shared__ int sWarp[1024 / 32 + 1];
    const int n = live_count(dCount, capacity);
    const int numTiles = (n + kScanTile - 1) / kScanTile;
    int carry = 0;
    for (int base = 0; base < numTiles; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < numTiles ? tileSums[i] : 0;
        int total;
        int excl = block_exclusive_scan<1024>(v, total, sWarp);
        if (i < numTiles) tileSums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0 && dTotal != nullptr) *dTotal = carry;
}

inline int scan_num_tiles(int capacity) { return capacity > 0 ? cdiv(capacity, kScanTile) : 1; }

// host helper: in-place two-level scan of data[0..count); tileSums needs scan_num_tiles(capacity) ints.
inline void launch_scan(int* data, const int* dCount, int capacity, int* tileSums, int* dTotal,
                        cudaStream_t stream) {
    int tiles = scan_num_tiles(capacity);
    scan_tiles_kernel<<<tiles, kScanThreads, 0, stream>>>(data, dCount, capacity, tileSums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(tileSums, dCount, capacity, dTotal);
}

}  // namespace mc
