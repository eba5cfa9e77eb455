// poisson.cu -- grid-based greedy Poisson-disk sampling of every cloud of a packed, grid-sorted batch.
//
// Replaces samplePointCloud / selectSamples / copyPoints (tf_ops/poisson_sampling.cu:51-124,175-244).  The
// reference launches one tiny kernel per (cloud, phase) -- B*27 sequential launches (864 at B = 32), one thread per
// cell walking 27 cells serially -- then blocks the host on the sample count.  Here:
//   select  : ONE launch for the whole batch, one CTA per cloud, the 27 colour phases run inside the kernel
//             separated by __syncthreads(); a warp owns a cell of the phase, its lanes own the 27 neighbour cells,
//             the accepted flags of the cloud live in shared memory.  The accepted SET is identical to the
//             reference's (it is deterministic there too: same-phase cells are >= 3 cells apart and a thread only
//             flags points of its own cell).
//   compact : the reference's output order inside a (cloud, phase) launch is atomic-arrival order; the canonical
//             order (ascending position in the sorted array == ascending cell, then point) is produced by a
//             counting sort on the 27*B (cloud, phase) buckets.  The sample count stays on the device.
#include "common.cuh"
#include "scan.cuh"

namespace mc {
namespace {

constexpr int kSelThreads = 512;
constexpr int kSelFlagBytes = 64 * 1024;  // per-cloud accepted flags kept in shared memory up to this many points
constexpr int kPtThreads = 256;

__global__ void __launch_bounds__(kSelThreads)
    poisson_select_kernel(bool scaleInv, float radius, int n, const float* __restrict__ mn,
                          const float* __restrict__ mx, const float* __restrict__ pts,
                          const int2* __restrict__ cellIdx, unsigned char* sel) {
    extern __shared__ unsigned char sFlags[];
    __shared__ int sLo, sHi;
    const int b = blockIdx.x;
    const int lane = lane_id(), warp = threadIdx.x >> 5, numWarps = kSelThreads / 32;
    const int cellsPerCloud = n * n * n;
    const int2* cells = cellIdx + (size_t)b * cellsPerCloud;

    // extent of this cloud inside the sorted array (the sort is batch-major)
    if (threadIdx.x == 0) {
        sLo = INT_MAX;
        sHi = 0;
    }
    __syncthreads();
    {
        int lo = INT_MAX, hi = 0;
        for (int c = threadIdx.x; c < cellsPerCloud; c += kSelThreads) {
            int2 ce = cells[c];
            if (ce.y > ce.x) {
                lo = min(lo, ce.x);
                hi = max(hi, ce.y);
            }
        }
        lo = __reduce_min_sync(MC_FULL_MASK, lo);
        hi = __reduce_max_sync(MC_FULL_MASK, hi);
        if (lane == 0) {
            atomicMin(&sLo, lo);
            atomicMax(&sHi, hi);
        }
    }
    __syncthreads();
    const int cLo = sLo, cHi = sHi;
    const int nCloud = cHi > cLo ? cHi - cLo : 0;
    if (nCloud == 0) return;
    const bool inSmem = nCloud <= kSelFlagBytes;
    if (inSmem)
        for (int t = threadIdx.x; t < nCloud; t += kSelThreads) sFlags[t] = 0;
    __syncthreads();

    const float ext = max_extent(mn, mx, b);
    const float r = scaleInv ? __fmul_rn(radius, ext) : radius;

    for (int ph = 0; ph < 27; ++ph) {
        const int ox = kOffsPool[ph][0], oy = kOffsPool[ph][1], oz = kOffsPool[ph][2];
        // cells of this phase: 3t+1+o < n  (only the upper bound exists, poisson_sampling.cu:70-74)
        const int tX = (n + 1 - ox) / 3, tY = (n + 1 - oy) / 3, tZ = (n + 1 - oz) / 3;
        const int total = tX * tY * tZ;
        for (int c = warp; c < total; c += numWarps) {
            const int tx = c / (tY * tZ), rem = c - tx * (tY * tZ);
            const int ty = rem / tZ, tz = rem - ty * tZ;
            const int x = tx * 3 + 1 + ox, y = ty * 3 + 1 + oy, z = tz * 3 + 1 + oz;
            const int2 own = cells[(x * n + y) * n + z];
            if (own.y <= own.x) continue;
            int nlo = 0, nhi = 0;
            if (lane < 27) {
                int qx = x + kOffsPool[lane][0], qy = y + kOffsPool[lane][1], qz = z + kOffsPool[lane][2];
                if (qx >= 0 && qx < n && qy >= 0 && qy < n && qz >= 0 && qz < n) {
                    int2 ce = cells[(qx * n + qy) * n + qz];
                    nlo = ce.x;
                    nhi = ce.y;
                }
            }
            for (int i = own.x; i < own.y; ++i) {
                const float px = pts[(size_t)i * 3], py = pts[(size_t)i * 3 + 1], pz = pts[(size_t)i * 3 + 2];
                bool coll = false;
                for (int j = nlo; j < nhi; ++j) {
                    unsigned char fl = inSmem ? sFlags[j - cLo] : __ldcg(&sel[j]);
                    if (fl && within_radius(px, py, pz, pts[(size_t)j * 3], pts[(size_t)j * 3 + 1],
                                            pts[(size_t)j * 3 + 2], r)) {
                        coll = true;
                        break;
                    }
                }
                if (!__any_sync(MC_FULL_MASK, coll) && lane == 0) {
                    if (inSmem)
                        sFlags[i - cLo] = (unsigned char)(ph + 1);
                    else
                        __stcg(&sel[i], (unsigned char)(ph + 1));
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    if (inSmem)
        for (int t = threadIdx.x; t < nCloud; t += kSelThreads) sel[cLo + t] = sFlags[t];
}

// (cloud, phase) histogram; arrival[i] = arrival rank inside the bucket
__global__ void __launch_bounds__(kPtThreads) poisson_count_kernel(const int* __restrict__ dN, int cap, int B,
                                                                   const unsigned char* __restrict__ sel,
                                                                   const int* __restrict__ bids,
                                                                   int* __restrict__ counters,
                                                                   int* __restrict__ arrival) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int ph = sel[i];
    if (ph == 0) return;
    int b = bids[i];
    b = (unsigned)b < (unsigned)B ? b : 0;
    arrival[i] = atomicAdd(&counters[b * 27 + ph - 1], 1);
}

__global__ void __launch_bounds__(kPtThreads) poisson_bucket_kernel(const int* __restrict__ dN, int cap, int B,
                                                                    const unsigned char* __restrict__ sel,
                                                                    const int* __restrict__ bids,
                                                                    const int* __restrict__ local,
                                                                    const int* __restrict__ tileOffs,
                                                                    const int* __restrict__ arrival,
                                                                    int* __restrict__ bucket) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int ph = sel[i];
    if (ph == 0) return;
    int b = bids[i];
    b = (unsigned)b < (unsigned)B ? b : 0;
    bucket[scan_prefix(local, tileOffs, b * 27 + ph - 1) + arrival[i]] = i;
}

__global__ void __launch_bounds__(kPtThreads) poisson_emit_kernel(const int* __restrict__ dN, int cap, int B,
                                                                  const unsigned char* __restrict__ sel,
                                                                  const int* __restrict__ bids,
                                                                  const float* __restrict__ pts,
                                                                  const int* __restrict__ local,
                                                                  const int* __restrict__ tileOffs,
                                                                  const int* __restrict__ dTotal,
                                                                  const int* __restrict__ bucket, int capOut,
                                                                  int* __restrict__ dOverflow,
                                                                  float* __restrict__ outPts,
                                                                  int* __restrict__ outBids, int* __restrict__ outIdx) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int ph = sel[i];
    if (ph == 0) return;
    int b = bids[i];
    b = (unsigned)b < (unsigned)B ? b : 0;
    int k = b * 27 + ph - 1;
    int beg = scan_prefix(local, tileOffs, k);
    int end = (k + 1 < B * 27) ? scan_prefix(local, tileOffs, k + 1) : *dTotal;
    int rnk = 0;
    for (int p = beg; p < end; ++p) rnk += (bucket[p] < i);
    int pos = beg + rnk;
    if (pos >= capOut) {
        if (dOverflow != nullptr) *dOverflow = 1;
        return;
    }
    outPts[(size_t)pos * 3] = pts[(size_t)i * 3];
    outPts[(size_t)pos * 3 + 1] = pts[(size_t)i * 3 + 1];
    outPts[(size_t)pos * 3 + 2] = pts[(size_t)i * 3 + 2];
    outBids[pos] = b;
    outIdx[pos] = i;
}

}  // namespace
}  // namespace mc

using namespace mc;

// workspace ints: sel bytes (N/4+1) | counters[27B] | tileSums | arrival[N] | bucket[N]
extern "C" long long mc_poisson_workspace_ints(int numPoints, int batchSize) {
    long long N = numPoints > 0 ? numPoints : 1;
    return (N / 4 + 1) + 27LL * batchSize + scan_num_tiles(27 * batchSize) + 2 * N + 8;
}

// samplePointCloud + copyPoints, tf_ops/poisson_sampling.cu:175-244.  Outputs hold capacityOut rows (numPoints
// always suffices); the live number of samples is written to *dNumSamples (device); samples past the capacity are
// dropped and *dOverflow (optional) is set.
extern "C" int mc_poisson_sampling(int scaleInv, float radius, int numPoints, const int* dNumPoints, int batchSize,
                                   int numCells, const float* aabbMin, const float* aabbMax, const float* pts,
                                   const int* batchIds, const int* cellIdx, int* workspace, int capacityOut,
                                   float* outPts, int* outBatchIds, int* outIdx, int* dNumSamples, int* dOverflow,
                                   void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_poisson_sampling: radius must be positive");
    MC_REQUIRE(batchSize > 0 && numCells > 0, "mc_poisson_sampling: batch size and cell count must be positive");
    MC_REQUIRE(dNumSamples != nullptr, "mc_poisson_sampling: dNumSamples required");
    cudaStream_t s = (cudaStream_t)stream;
    static bool attrSet = false;
    if (!attrSet) {
        cudaFuncSetAttribute(poisson_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelFlagBytes);
        attrSet = true;
    }
    const long long N = numPoints > 0 ? numPoints : 1;
    const int nb = 27 * batchSize;
    unsigned char* sel = reinterpret_cast<unsigned char*>(workspace);
    int* counters = workspace + (N / 4 + 1);
    int* tileSums = counters + nb;
    int* arrival = tileSums + scan_num_tiles(nb);
    int* bucket = arrival + N;
    cudaMemsetAsync(workspace, 0, sizeof(int) * (size_t)((N / 4 + 1) + nb), s);  // sel + counters
    if (numPoints > 0) {
        poisson_select_kernel<<<batchSize, kSelThreads, kSelFlagBytes, s>>>(
            scaleInv != 0, radius, numCells, aabbMin, aabbMax, pts, reinterpret_cast<const int2*>(cellIdx), sel);
        int grid = cdiv(numPoints, kPtThreads);
        poisson_count_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, batchSize, sel, batchIds, counters,
                                                         arrival);
        launch_scan(counters, nullptr, nb, tileSums, dNumSamples, s);
        poisson_bucket_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, batchSize, sel, batchIds, counters,
                                                          tileSums, arrival, bucket);
        poisson_emit_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, batchSize, sel, batchIds, pts,
                                                        counters, tileSums, dNumSamples, bucket, capacityOut,
                                                        dOverflow, outPts, outBatchIds, outIdx);
    } else {
        cudaMemsetAsync(dNumSamples, 0, sizeof(int), s);
    }
    return check_launch("mc_poisson_sampling");
}
