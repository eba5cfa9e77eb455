// neighbors.cu -- fixed-radius neighbour lists (CSR: start[M], pairs[K,2] = (point j, sample i)) over the grid.
//
// Replaces countNeighbors / computeOffsets / findNeighbors (tf_ops/find_neighbors.cu:40-264).  The reference runs
// one THREAD per sample (a serial walk over up to 27 cells x population, uncoalesced), a cudaMalloc + symbol
// upload + blocking D2H for the total, and Hillis-Steele scans.  Here one WARP owns a sample: lanes 0..26 fetch the
// 27 cell ranges at once, then the warp sweeps the non-empty cells with one candidate per lane (coalesced 12-byte
// rows), and __ballot_sync/__popc give both the count and the ORDERED write slot, so the output order is exactly
// the reference's (cells in table order find_neighbors.cu:282-291, ascending j inside a cell).  The predicate is
// the float-exact one in common.cuh.  The total stays on the device (dNumNeighbors); the caller decides whether to
// read it (exact-shape API) or to keep going with a capacity-sized pair buffer (sync-free pipeline).
#include <cstdlib>

#include "common.cuh"
#include "lookback.cuh"
#include "scan.cuh"

namespace mc {
namespace {

constexpr int kWarpsPerCta = 8;

template <bool FILL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
    neighbors_kernel(bool scaleInv, const int* __restrict__ dM, int capM, int n, float radius,
                     const float* __restrict__ samples, const int* __restrict__ sbids,
                     const float* __restrict__ pts2, const int2* __restrict__ cellIdx, const float* __restrict__ mn,
                     const float* __restrict__ mx, int* __restrict__ start, const int* __restrict__ tileOffs,
                     int2* __restrict__ pairs, int capK, int* __restrict__ dOverflow) {
    const int M = live_count(dM, capM);
    const int i = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (i >= M) return;
    const int lane = lane_id();
    const int b = sbids[i];
    const CloudFrame f = cloud_frame(mn, mx, b, n, radius, scaleInv);
    const float sx = samples[(size_t)i * 3], sy = samples[(size_t)i * 3 + 1], sz = samples[(size_t)i * 3 + 2];
    int cx, cy, cz;
    cell_of(f, sx, sy, sz, n, cx, cy, cz);
    const float thr = exact_radius_threshold(f.radius);  // d2 <= thr  <=>  sqrt(d2) < radius, exactly

    // lanes 0..26: (lo,hi) of neighbour cell `lane`
    int lo = 0, hi = 0;
    if (lane < 27) {
        // kOffsNeigh[lane] in closed form (x cycles +1,0,-1 fastest, find_neighbors.cu:282-291): a lane-indexed
        // constant-memory load would serialise over the 27 distinct addresses
        int x = cx + 1 - lane % 3, y = cy + 1 - (lane / 3) % 3, z = cz + 1 - lane / 9;
        if (x >= 0 && x < n && y >= 0 && y < n && z >= 0 && z < n) {
            int2 c = cellIdx[(((size_t)b * n + x) * n + y) * n + z];
            lo = c.x;
            hi = c.y;
        }
    }
    unsigned nonEmpty = __ballot_sync(MC_FULL_MASK, hi > lo);

    int base = 0;
    if (FILL) {
        base = start[i] + tileOffs[i / kScanTile];
        if (lane == 0) start[i] = base;  // finalise in place, as find_neighbors.cu:223-224
    }
    const unsigned ltMask = (1u << lane) - 1u;
    int count = 0;
    while (nonEmpty) {
        int o = __ffs(nonEmpty) - 1;
        nonEmpty &= nonEmpty - 1;
        const int clo = __shfl_sync(MC_FULL_MASK, lo, o), chi = __shfl_sync(MC_FULL_MASK, hi, o);
        for (int j0 = clo; j0 < chi; j0 += 32) {
            int j = j0 + lane;
            bool in = false;
            if (j < chi)
                in = dist2_ref(sx, sy, sz, pts2[(size_t)j * 3], pts2[(size_t)j * 3 + 1], pts2[(size_t)j * 3 + 2]) <= thr;
            unsigned hit = __ballot_sync(MC_FULL_MASK, in);
            if (FILL && in) {
                int slot = base + count + __popc(hit & ltMask);
                if (slot < capK)
                    pairs[slot] = make_int2(j, i);
                else if (dOverflow != nullptr)
                    *dOverflow = 1;
            }
            count += __popc(hit);
        }
    }
    if (!FILL && lane == 0) start[i] = count;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused count + scan + fill for the sync-free pipeline (the pair capacity is known up front): ONE launch.
// A CTA owns 32 consecutive samples (4 per warp).  Sweep 1 evaluates every candidate once and caches the hit ballots
// of the rounds that had hits in shared memory; the 32 counts are scanned in the CTA, the CTA total goes through the
// decoupled look-back (lookback.cuh) to obtain the CTA's offset in the global pair list, and the cached ballots are
// replayed into ordered (j, i) writes -- distances are computed once, no second kernel, no separate scan kernels.
constexpr int kFusedPerWarp = 4;
constexpr int kFusedPerCta = kWarpsPerCta * kFusedPerWarp;  // 32 samples per CTA
constexpr int kFusedRounds = 64;                            // cached hit rounds per sample (else: re-sweep)
static_assert(kFusedPerCta == 32, "the CTA scan of the per-sample counts uses one warp");

constexpr int kPdfList = 256;  // neighbour coordinates a warp keeps in shared memory for the fused density estimate

struct FusedShared {
    int2 rounds[kFusedPerCta][kFusedRounds];  // (first candidate of the round, hit mask)
    int count[kFusedPerCta];
    int nRounds[kFusedPerCta];
    int offset[kFusedPerCta];
    int cta, base;
};

// one sweep over the 27 cells of sample i; WRITE: emit pairs at `base` instead of caching
template <bool WRITE>
__device__ __forceinline__ void sweep_sample(int i, bool scaleInv, int n, float radius,
                                             const float* __restrict__ samples, const int* __restrict__ sbids,
                                             const float* __restrict__ pts2, const int2* __restrict__ cellIdx,
                                             const float* __restrict__ mn, const float* __restrict__ mx, int lane,
                                             int2* roundsOut, int& outCount, int& outRounds, int base,
                                             int2* __restrict__ pairs, int capK, int* __restrict__ dOverflow) {
    const int b = sbids[i];
    const CloudFrame f = cloud_frame(mn, mx, b, n, radius, scaleInv);
    const float sx = samples[(size_t)i * 3], sy = samples[(size_t)i * 3 + 1], sz = samples[(size_t)i * 3 + 2];
    int cx, cy, cz;
    cell_of(f, sx, sy, sz, n, cx, cy, cz);
    const float thr = exact_radius_threshold(f.radius);
    int lo = 0, hi = 0;
    if (lane < 27) {
        // kOffsNeigh[lane] in closed form (x cycles +1,0,-1 fastest, find_neighbors.cu:282-291): a lane-indexed
        // constant-memory load would serialise over the 27 distinct addresses
        int x = cx + 1 - lane % 3, y = cy + 1 - (lane / 3) % 3, z = cz + 1 - lane / 9;
        if (x >= 0 && x < n && y >= 0 && y < n && z >= 0 && z < n) {
            int2 c = cellIdx[(((size_t)b * n + x) * n + y) * n + z];
            lo = c.x;
            hi = c.y;
        }
    }
    unsigned nonEmpty = __ballot_sync(MC_FULL_MASK, hi > lo);
    const unsigned ltMask = (1u << lane) - 1u;
    int count = 0, nr = 0;
    if (!WRITE) {
        // Sparse neighbourhoods (fine grids: a fraction of a point per cell, e.g. MCSeg r = 0.03, ScanNet rooms): the
        // candidate-per-lane sweep below would spend one round per non-empty cell on one or two candidates.  Instead
        // LANE = CELL: every lane walks its own cell's few candidates and records the hits as a bit mask; the cells
        // are the lanes in table order, so a warp prefix sum of the per-lane counts gives the same ordered slots.
        int maxPop = hi - lo;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) maxPop = max(maxPop, __shfl_xor_sync(MC_FULL_MASK, maxPop, d));
        if (maxPop <= 32 && maxPop < __popc(nonEmpty)) {
            unsigned mask = 0;
            for (int c = 0; c < maxPop; ++c) {
                const int j = lo + c;
                if (j < hi &&
                    dist2_ref(sx, sy, sz, pts2[(size_t)j * 3], pts2[(size_t)j * 3 + 1], pts2[(size_t)j * 3 + 2]) <= thr)
                    mask |= 1u << c;
            }
            roundsOut[lane] = make_int2(lo, (int)mask);
            int c = __popc(mask);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(MC_FULL_MASK, c, d);
            outCount = c;
            outRounds = -1;  // cached in the lane = cell format
            return;
        }
    }
    while (nonEmpty) {
        const int o = __ffs(nonEmpty) - 1;
        nonEmpty &= nonEmpty - 1;
        const int clo = __shfl_sync(MC_FULL_MASK, lo, o), chi = __shfl_sync(MC_FULL_MASK, hi, o);
        for (int j0 = clo; j0 < chi; j0 += 32) {
            const int j = j0 + lane;
            bool in = false;
            if (j < chi)
                in = dist2_ref(sx, sy, sz, pts2[(size_t)j * 3], pts2[(size_t)j * 3 + 1], pts2[(size_t)j * 3 + 2]) <= thr;
            const unsigned hit = __ballot_sync(MC_FULL_MASK, in);
            if (hit) {
                if (WRITE) {
                    if (in) {
                        const int slot = base + count + __popc(hit & ltMask);
                        if (slot < capK)
                            pairs[slot] = make_int2(j, i);
                        else if (dOverflow != nullptr)
                            *dOverflow = 1;
                    }
                } else if (nr < kFusedRounds && lane == 0) {
                    roundsOut[nr] = make_int2(j0, (int)hit);
                }
                ++nr;
                count += __popc(hit);
            }
        }
    }
    outCount = count;
    outRounds = nr;
}

// PER_WARP samples per warp: 1 keeps the per-warp serial chain short when there are few samples (latency bound),
// 4 keeps the look-back chain short when there are many.
// PDF: also evaluate compute_pdf for the pairs of every sample while its neighbour list is at hand (window > 0): the
// same terms in the same order as pdf_kernel (common.cuh::pdf_term), so the values are bit-identical to the separate op.
template <int PER_WARP, bool PDF>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
    neighbors_fused_kernel(bool scaleInv, const int* __restrict__ dM, int capM, int n, float radius,
                           const float* __restrict__ samples, const int* __restrict__ sbids,
                           const float* __restrict__ pts2, const int2* __restrict__ cellIdx,
                           const float* __restrict__ mn, const float* __restrict__ mx, unsigned long long* desc,
                           int numCtas, int* __restrict__ start, int2* __restrict__ pairs, int capK,
                           int* __restrict__ dNumNeighbors, int* __restrict__ dOverflow, float window,
                           float* __restrict__ pdfs) {
    constexpr int kFusedPerWarp = PER_WARP;
    constexpr int kPerCta = kWarpsPerCta * PER_WARP;
    __shared__ FusedShared sh;
    __shared__ float sPx[PDF ? kWarpsPerCta : 1][PDF ? kPdfList : 1], sPy[PDF ? kWarpsPerCta : 1][PDF ? kPdfList : 1],
        sPz[PDF ? kWarpsPerCta : 1][PDF ? kPdfList : 1];
    const int M = live_count(dM, capM);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sh.cta = lookback_ticket(desc);
    __syncthreads();
    const int cta = sh.cta;
    const int first = cta * kPerCta + warp * kFusedPerWarp;

    // sweep 1: count + cache the hit ballots
    for (int s = 0; s < kFusedPerWarp; ++s) {
        const int slot = warp * kFusedPerWarp + s;
        const int i = first + s;
        int cnt = 0, nr = 0;
        if (i < M)
            sweep_sample<false>(i, scaleInv, n, radius, samples, sbids, pts2, cellIdx, mn, mx, lane, sh.rounds[slot], cnt,
                                nr, 0, nullptr, 0, nullptr);
        if (lane == 0) {
            sh.count[slot] = cnt;
            sh.nRounds[slot] = nr;
        }
    }
    __syncthreads();

    // CTA scan of the 32 counts + look-back across CTAs
    if (warp == 0) {
        const int c = lane < kPerCta ? sh.count[lane] : 0;
        const int inc = warp_inclusive_scan(c);
        if (lane < kPerCta) sh.offset[lane] = inc - c;
        const int total = __shfl_sync(MC_FULL_MASK, inc, 31);
        const int base = lookback_exclusive_prefix(desc, cta, total);
        if (lane == 0) {
            sh.base = base;
            if (cta == numCtas - 1) *dNumNeighbors = base + total;
        }
    }
    __syncthreads();

    // replay: ordered writes from the cached ballots
    const unsigned ltMask = (1u << lane) - 1u;
    for (int s = 0; s < kFusedPerWarp; ++s) {
        const int slot = warp * kFusedPerWarp + s;
        const int i = first + s;
        if (i >= M) break;
        const int base = sh.base + sh.offset[slot];
        if (lane == 0) start[i] = base;
        const int nr = sh.nRounds[slot];
        if (nr < 0) {  // lane = cell format: (first candidate, hit mask) per lane, cells in lane order
            const int2 rd = sh.rounds[slot][lane];
            unsigned mask = (unsigned)rd.y;
            const int c = __popc(mask);
            int dst = base + warp_inclusive_scan(c) - c;
            while (mask) {
                const int bit = __ffs(mask) - 1;
                mask &= mask - 1;
                if (dst < capK)
                    pairs[dst] = make_int2(rd.x + bit, i);
                else if (dOverflow != nullptr)
                    *dOverflow = 1;
                ++dst;
            }
        } else if (nr <= kFusedRounds) {
            int written = 0;
            for (int r = 0; r < nr; ++r) {
                const int2 rd = sh.rounds[slot][r];
                const unsigned hit = (unsigned)rd.y;
                if (hit & (1u << lane)) {
                    const int dst = base + written + __popc(hit & ltMask);
                    if (dst < capK)
                        pairs[dst] = make_int2(rd.x + lane, i);
                    else if (dOverflow != nullptr)
                        *dOverflow = 1;
                }
                written += __popc(hit);
            }
        } else {  // more hit rounds than the cache holds: sweep again, writing
            int cnt, nr2;
            sweep_sample<true>(i, scaleInv, n, radius, samples, sbids, pts2, cellIdx, mn, mx, lane, nullptr, cnt, nr2, base,
                               pairs, capK, dOverflow);
        }
        if (PDF) {
            // density estimate of this sample's pairs: pdf_p = (1/k) sum_q g(x_q - x_p) over ALL k neighbours, q in list
            // order.  The warp re-reads the neighbour indices it has just written, keeps up to kPdfList coordinates in
            // shared memory (longer lists: the tail through L1 / L2) and every lane takes the pairs lane, lane + 32, ...
            __syncwarp();
            const int k = sh.count[slot];
            const int kw = max(0, min(k, capK - base));  // pairs actually written
            const float ext = max_extent(mn, mx, sbids[i]);
            const float e2 = pdf_exponent_scale(scaleInv ? radius * ext : radius, window);
            const int staged = min(kw, kPdfList);
            for (int t = lane; t < staged; t += 32) {
                const int j = pairs[base + t].x;
                sPx[warp][t] = pts2[(size_t)j * 3];
                sPy[warp][t] = pts2[(size_t)j * 3 + 1];
                sPz[warp][t] = pts2[(size_t)j * 3 + 2];
            }
            __syncwarp();
            for (int t = lane; t < kw; t += 32) {
                float xj, yj, zj;
                if (t < staged) {
                    xj = sPx[warp][t];
                    yj = sPy[warp][t];
                    zj = sPz[warp][t];
                } else {
                    const int j = pairs[base + t].x;
                    xj = pts2[(size_t)j * 3];
                    yj = pts2[(size_t)j * 3 + 1];
                    zj = pts2[(size_t)j * 3 + 2];
                }
                float acc = 0.0f;
                for (int q = 0; q < staged; ++q) acc += pdf_term(sPx[warp][q] - xj, sPy[warp][q] - yj, sPz[warp][q] - zj, e2);
                for (int q = staged; q < kw; ++q) {
                    const int jj = pairs[base + q].x;
                    acc += pdf_term(pts2[(size_t)jj * 3] - xj, pts2[(size_t)jj * 3 + 1] - yj, pts2[(size_t)jj * 3 + 2] - zj,
                                    e2);
                }
                pdfs[base + t] = pdf_finish(acc, window, k);
            }
            __syncwarp();
        }
    }
}

}  // namespace
}  // namespace mc

using namespace mc;

extern "C" long long mc_find_neighbors_workspace_ints(int numSamples) {
    long long tiles = scan_num_tiles(numSamples);
    long long fused = 2LL * (cdiv(numSamples > 0 ? numSamples : 1, kWarpsPerCta) + 2);  // look-back descriptors
    return (tiles > fused ? tiles : fused) + 8;
}

// countNeighborsCPU + packNeighborsCPU (tf_ops/find_neighbors.cu:268-371) in ONE launch, for callers that provide a
// pair capacity up front (sync-free pipeline).  startIdx receives the final start indices, *dNumNeighbors the total;
// pairs past capacityPairs are dropped and *dOverflow is set.  workspace: mc_find_neighbors_workspace_ints, 8-byte
// aligned.
constexpr int kPdfFuseMaxSamples = 8192;

// MCCNN_PDF_FUSE_MAX=<samples> overrides the threshold (0: always two launches) -- measurement knob.
static int pdf_fuse_max_samples() {
    static const int v = [] {
        const char* e = getenv("MCCNN_PDF_FUSE_MAX");
        return e ? atoi(e) : kPdfFuseMaxSamples;
    }();
    return v;
}

static int launch_fused(int scaleInv, int numSamples, const int* dNumSamples, int numCells, float radius,
                        const float* samples, const int* sampleBatchIds, const float* pts2, const int* cellIdx,
                        const float* aabbMin, const float* aabbMax, int* startIdx, int* workspace, int* pairs,
                        int capacityPairs, int* dNumNeighbors, int* dOverflow, float window, float* pdfs, void* stream,
                        const char* what) {
    cudaStream_t s = (cudaStream_t)stream;
    if (numSamples <= 0) {
        cudaMemsetAsync(dNumNeighbors, 0, sizeof(int), s);
        return check_launch(what);
    }
    const bool few = numSamples <= 16384;
    const int ctas = cdiv(numSamples, few ? kWarpsPerCta : kFusedPerCta);
    unsigned long long* desc = reinterpret_cast<unsigned long long*>(workspace);
    cudaMemsetAsync(desc, 0, sizeof(unsigned long long) * (size_t)(ctas + 1), s);
    const int2* cells = reinterpret_cast<const int2*>(cellIdx);
    int2* prs = reinterpret_cast<int2*>(pairs);
#define MC_LAUNCH_FUSED(PW, PDF)                                                                                      \
    neighbors_fused_kernel<PW, PDF><<<ctas, kWarpsPerCta * 32, 0, s>>>(                                               \
        scaleInv != 0, dNumSamples, numSamples, numCells, radius, samples, sampleBatchIds, pts2, cells, aabbMin, aabbMax, \
        desc, ctas, startIdx, prs, capacityPairs, dNumNeighbors, dOverflow, window, pdfs)
    if (pdfs != nullptr) {
        if (few) MC_LAUNCH_FUSED(1, true);
        else MC_LAUNCH_FUSED(kFusedPerWarp, true);
    } else {
        if (few) MC_LAUNCH_FUSED(1, false);
        else MC_LAUNCH_FUSED(kFusedPerWarp, false);
    }
#undef MC_LAUNCH_FUSED
    return check_launch(what);
}

// countNeighborsCPU + packNeighborsCPU (tf_ops/find_neighbors.cu:268-371) in ONE launch, for callers that provide a
// pair capacity up front (sync-free pipeline).  startIdx receives the final start indices, *dNumNeighbors the total;
// pairs past capacityPairs are dropped and *dOverflow is set.  workspace: mc_find_neighbors_workspace_ints, 8-byte
// aligned.
extern "C" int mc_find_neighbors_fused(int scaleInv, int numSamples, const int* dNumSamples, int numCells,
                                       float radius, const float* samples, const int* sampleBatchIds,
                                       const float* pts2, const int* cellIdx, const float* aabbMin,
                                       const float* aabbMax, int* startIdx, int* workspace, int* pairs,
                                       int capacityPairs, int* dNumNeighbors, int* dOverflow, void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_find_neighbors: radius must be positive");
    MC_REQUIRE(numCells > 0, "mc_find_neighbors: cell count must be positive");
    MC_REQUIRE(dNumNeighbors != nullptr, "mc_find_neighbors_fused: dNumNeighbors required");
    MC_REQUIRE(((uintptr_t)workspace) % 8 == 0, "mc_find_neighbors_fused: workspace must be 8-byte aligned");
    return launch_fused(scaleInv, numSamples, dNumSamples, numCells, radius, samples, sampleBatchIds, pts2, cellIdx,
                        aabbMin, aabbMax, startIdx, workspace, pairs, capacityPairs, dNumNeighbors, dOverflow, 0.f,
                        nullptr, stream, "mc_find_neighbors_fused");
}

// The same launch also evaluating computeDPFsCPU (tf_ops/compute_pdf.cu:98-120) for every pair while the sample's
// neighbour list is at hand: pdfs[capacityPairs] receives exactly the values mc_compute_pdf would produce (same terms,
// same order).  The samples' batch ids stand in for the neighbours' (a neighbour lies in its sample's cloud).
extern "C" int mc_find_neighbors_pdf_fused(int scaleInv, float window, int numSamples, const int* dNumSamples,
                                           int numCells, float radius, const float* samples,
                                           const int* sampleBatchIds, const float* pts2, const int* cellIdx,
                                           const float* aabbMin, const float* aabbMax, int* startIdx, int* workspace,
                                           int* pairs, float* pdfs, int capacityPairs, int* dNumNeighbors,
                                           int* dOverflow, void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_find_neighbors: radius must be positive");
    MC_REQUIRE(window > 0.0f, "mc_find_neighbors_pdf_fused: window must be positive");
    MC_REQUIRE(numCells > 0, "mc_find_neighbors: cell count must be positive");
    MC_REQUIRE(dNumNeighbors != nullptr && pdfs != nullptr, "mc_find_neighbors_pdf_fused: outputs required");
    MC_REQUIRE(((uintptr_t)workspace) % 8 == 0, "mc_find_neighbors_pdf_fused: workspace must be 8-byte aligned");
    // One launch while the call is latency-bound; past kPdfFuseMaxSamples the per-sample density loop (one lane per
    // pair of ONE sample: ~6 of 32 lanes busy on the up-sampling lists) costs more than the launch it saves
    // (profiles/r2_neighbors_pdf_fused.txt), so the densities go to the pair-per-thread kernel of pdf.cu.
    if (numSamples <= pdf_fuse_max_samples())
        return launch_fused(scaleInv, numSamples, dNumSamples, numCells, radius, samples, sampleBatchIds, pts2,
                            cellIdx, aabbMin, aabbMax, startIdx, workspace, pairs, capacityPairs, dNumNeighbors,
                            dOverflow, window, pdfs, stream, "mc_find_neighbors_pdf_fused");
    int rc = launch_fused(scaleInv, numSamples, dNumSamples, numCells, radius, samples, sampleBatchIds, pts2, cellIdx,
                          aabbMin, aabbMax, startIdx, workspace, pairs, capacityPairs, dNumNeighbors, dOverflow, 0.f,
                          nullptr, stream, "mc_find_neighbors_pdf_fused");
    if (rc != 0) return rc;
    return launch_pdf_by_sample(scaleInv != 0, window, radius, dNumSamples, numSamples, dNumNeighbors, capacityPairs,
                                pts2, sampleBatchIds, aabbMin, aabbMax, startIdx, pairs, pdfs, (cudaStream_t)stream);
}

// countNeighborsCPU, tf_ops/find_neighbors.cu:268-317 -- without the cudaMalloc / D2H: the total lands in
// *dNumNeighbors (device).  startIdx receives the tile-local exclusive prefix; mc_find_neighbors_pack finalises it.
extern "C" int mc_find_neighbors_count(int scaleInv, int numSamples, const int* dNumSamples, int numCells,
                                       float radius, const float* samples, const int* sampleBatchIds,
                                       const float* pts2, const int* cellIdx, const float* aabbMin,
                                       const float* aabbMax, int* startIdx, int* workspace, int* dNumNeighbors,
                                       void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_find_neighbors: radius must be positive");
    MC_REQUIRE(numCells > 0, "mc_find_neighbors: cell count must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    if (numSamples > 0)
        neighbors_kernel<false><<<cdiv(numSamples, kWarpsPerCta), kWarpsPerCta * 32, 0, s>>>(
            scaleInv != 0, dNumSamples, numSamples, numCells, radius, samples, sampleBatchIds, pts2,
            reinterpret_cast<const int2*>(cellIdx), aabbMin, aabbMax, startIdx, nullptr, nullptr, 0, nullptr);
    launch_scan(startIdx, dNumSamples, numSamples, workspace, dNumNeighbors, s);
    count_launches(2);
    return check_launch("mc_find_neighbors_count");
}

// packNeighborsCPU, tf_ops/find_neighbors.cu:330-371.  pairs holds capacityPairs (j,i) rows; rows past the
// capacity are dropped and *dOverflow is set (caller zeroes it beforehand).
extern "C" int mc_find_neighbors_pack(int scaleInv, int numSamples, const int* dNumSamples, int numCells,
                                      float radius, const float* samples, const int* sampleBatchIds,
                                      const float* pts2, const int* cellIdx, const float* aabbMin,
                                      const float* aabbMax, int* startIdx, const int* workspace, int* pairs,
                                      int capacityPairs, int* dOverflow, void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_find_neighbors: radius must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    if (numSamples > 0)
        neighbors_kernel<true><<<cdiv(numSamples, kWarpsPerCta), kWarpsPerCta * 32, 0, s>>>(
            scaleInv != 0, dNumSamples, numSamples, numCells, radius, samples, sampleBatchIds, pts2,
            reinterpret_cast<const int2*>(cellIdx), aabbMin, aabbMax, startIdx, workspace,
            reinterpret_cast<int2*>(pairs), capacityPairs, dOverflow);
    return check_launch("mc_find_neighbors_pack");
}
