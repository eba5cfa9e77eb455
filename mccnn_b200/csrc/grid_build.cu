// grid_build.cu -- the whole grid sort (SortPointsStep1 + SortPointsStep2 of the reference, tf_ops/sort_gpu.cu:436-497)
// as ONE kernel launch for packed batches whose clouds are contiguous (utils/DataSet.py:816-818; every Poisson level
// produced by poisson.cu is contiguous by construction).
//
// The reference needs 3 memsets + 7 launches per grid and global atomics over the whole cell range.  Because the key
// is batch-major (sort_gpu.cu:59) a cloud's points only ever land in that cloud's n^3 cells and the sorted array keeps
// clouds in place, so one CTA per cloud can do the complete counting sort on its own:
//   histogram in shared memory -> CTA-wide exclusive scan -> cell table rows of the cloud -> arrival-order scatter ->
//   stable rank inside the (small) cell segment -> move points / batch ids / keys / feature rows.
// The result is bit-identical to the general multi-kernel path in grid_sort.cu (same keys, same canonical stable
// order).  Contiguity is VERIFIED on the device: every point of a CTA's range must carry the CTA's batch id and the
// ranges must cover all live points; otherwise *dViolation is set and the caller falls back to the general path.
#include "common.cuh"
#include "scan.cuh"

namespace mc {
namespace {

constexpr int kGbThreads = 512;
constexpr int kGbMaxCells = 8192;  // n^3 limit of the fused path (32 KB histogram)

struct GbShared {
    int warpScan[kGbThreads / 32 + 1];
    int carry;
};

// first index in [0, N) with bids[i] >= target (N if none), bids non-decreasing: CTA-wide 512-ary search.
// Invariant: everything before lo is < target, bids[hi] (if any) is >= target.  All threads return the same value.
__device__ int cta_lower_bound(const int* __restrict__ bids, int N, int target) {
    int lo = 0, hi = N;
    while (hi > lo) {
        const int chunk = (hi - lo + kGbThreads - 1) / kGbThreads;
        const int p = lo + threadIdx.x * chunk;
        const int less = (p < hi) && (bids[p] < target);
        const int cnt = __syncthreads_count(less);  // probes are monotone: the first cnt of them are < target
        if (cnt == 0) {
            hi = lo;
        } else {
            const int nhi = min(lo + cnt * chunk, hi);
            lo = lo + (cnt - 1) * chunk + 1;
            hi = nhi;
        }
    }
    return lo;
}

__global__ void __launch_bounds__(kGbThreads)
    grid_build_kernel(const int* __restrict__ dN, int capN, int B, int n, int C, const float* __restrict__ mn,
                      const float* __restrict__ mx, const float* __restrict__ pts, const int* __restrict__ bids,
                      const float* __restrict__ feats, int* __restrict__ arrival, int* __restrict__ perm,
                      unsigned long long* __restrict__ coverage, int* __restrict__ dViolation, int* __restrict__ keys,
                      int* __restrict__ newIdx, float* __restrict__ outPts, int* __restrict__ outBids,
                      float* __restrict__ outFeats, int2* __restrict__ cellIdx) {
    extern __shared__ int sHist[];  // n^3 ints
    __shared__ GbShared sh;
    const int N = live_count(dN, capN);
    const int b = blockIdx.x;
    const int cells = n * n * n;

    const int s = cta_lower_bound(bids, N, b);
    const int e = cta_lower_bound(bids, N, b + 1);

    for (int c = threadIdx.x; c < cells; c += kGbThreads) sHist[c] = 0;
    __syncthreads();

    // keys + histogram (arrival order inside the cell)
    const CloudFrame f = cloud_frame(mn, mx, b, n, 0.f, false);
    bool bad = false;
    for (int i = s + threadIdx.x; i < e; i += kGbThreads) {
        bad |= bids[i] != b;
        int cx, cy, cz;
        cell_of(f, pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2], n, cx, cy, cz);
        const int kl = (cx * n + cy) * n + cz;
        keys[i] = b * cells + kl;
        arrival[i] = atomicAdd(&sHist[kl], 1);
    }
    if (bad) *dViolation = 1;
    if (threadIdx.x == 0) {  // coverage: the ranges of all clouds must add up to the live point count
        const unsigned long long prev = atomicAdd(coverage, (1ull << 32) | (unsigned)(e - s));
        const int arrived = (int)(prev >> 32) + 1;
        const long long covered = (long long)(prev & 0xffffffffull) + (e - s);
        if (arrived == B && covered != N) *dViolation = 1;
        sh.carry = 0;
    }
    __syncthreads();

    // exclusive scan of the histogram (in place) + this cloud's rows of the cell table
    for (int c0 = 0; c0 < cells; c0 += kGbThreads) {
        const int c = c0 + threadIdx.x;
        const int cnt = c < cells ? sHist[c] : 0;
        int total;
        const int excl = block_exclusive_scan<kGbThreads>(cnt, total, sh.warpScan);
        const int start = sh.carry + excl;
        if (c < cells) {
            sHist[c] = start;
            cellIdx[(size_t)b * cells + c] = cnt > 0 ? make_int2(s + start, s + start + cnt) : make_int2(0, 0);
        }
        __syncthreads();
        if (threadIdx.x == 0) sh.carry += total;
        __syncthreads();
    }

    // arrival-order scatter into the cell segments
    for (int i = s + threadIdx.x; i < e; i += kGbThreads) perm[s + sHist[keys[i] - b * cells] + arrival[i]] = i;
    __syncthreads();

    // stable rank inside the segment, then move the point
    for (int i = s + threadIdx.x; i < e; i += kGbThreads) {
        const int key = keys[i];
        const int kl = key - b * cells;
        const int beg = s + sHist[kl];
        const int end = (kl + 1 < cells) ? s + sHist[kl + 1] : e;
        int r = 0;
        for (int p = beg; p < end; ++p) r += (perm[p] < i);
        const int d = beg + r;
        newIdx[i] = d;
        outPts[(size_t)d * 3] = pts[(size_t)i * 3];
        outPts[(size_t)d * 3 + 1] = pts[(size_t)i * 3 + 1];
        outPts[(size_t)d * 3 + 2] = pts[(size_t)i * 3 + 2];
        outBids[d] = b;
    }
    if (C <= 0) return;
    __syncthreads();
    // feature rows of the cloud (coalesced over the row)
    const long long total = (long long)(e - s) * C;
    for (long long t = threadIdx.x; t < total; t += kGbThreads) {
        const int i = s + (int)(t / C), c = (int)(t % C);
        outFeats[(size_t)newIdx[i] * C + c] = feats[(size_t)i * C + c];
    }
}

}  // namespace
}  // namespace mc

using namespace mc;

// 1 when (batchSize, numCells) fits the fused one-CTA-per-cloud path
extern "C" int mc_build_grid_supported(int batchSize, int numCells) {
    long long cells = (long long)numCells * numCells * numCells;
    return (cells <= kGbMaxCells && batchSize > 0) ? 1 : 0;
}

// workspace ints: coverage (one 64-bit word) | arrival[N] | perm[N]
extern "C" long long mc_build_grid_workspace_ints(int numPoints) {
    return 2LL * (numPoints > 0 ? numPoints : 1) + 2 + 8;
}

// Fused SortPointsStep1 + SortPointsStep2 for cloud-contiguous batches.  Outputs exactly those of
// mc_sort_points_step1 + mc_sort_points_step2.  *dViolation (required) is set when the batch is NOT contiguous (the
// outputs are then unspecified and the caller must use the general path).
extern "C" int mc_build_grid(int numPoints, const int* dNumPoints, int batchSize, int numFeatures, int numCells,
                             const float* aabbMin, const float* aabbMax, const float* pts, const int* batchIds,
                             const float* features, int* workspace, int* dViolation, int* keys, int* newIdx,
                             float* outPts, int* outBatchIds, float* outFeatures, int* outCellIdx, void* stream) {
    MC_REQUIRE(mc_build_grid_supported(batchSize, numCells), "mc_build_grid: grid too fine for the fused path");
    MC_REQUIRE(dViolation != nullptr, "mc_build_grid: dViolation required");
    cudaStream_t s = (cudaStream_t)stream;
    const int N = numPoints > 0 ? numPoints : 1;
    MC_REQUIRE(((uintptr_t)workspace) % 8 == 0, "mc_build_grid: workspace must be 8-byte aligned");
    unsigned long long* coverage = reinterpret_cast<unsigned long long*>(workspace);
    int* arrival = workspace + 2;
    int* perm = arrival + N;
    cudaMemsetAsync(coverage, 0, sizeof(unsigned long long), s);
    const int cells = numCells * numCells * numCells;
    grid_build_kernel<<<batchSize, kGbThreads, sizeof(int) * (size_t)cells, s>>>(
        dNumPoints, numPoints, batchSize, numCells, numFeatures, aabbMin, aabbMax, pts, batchIds, features, arrival,
        perm, coverage, dViolation, keys, newIdx, outPts, outBatchIds, outFeatures, reinterpret_cast<int2*>(outCellIdx));
    return check_launch("mc_build_grid");
}
