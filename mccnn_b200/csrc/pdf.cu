// pdf.cu -- kernel density estimate per (sample, neighbour) pair.
//
// Replaces computeDPFsCPU / computePDFs (tf_ops/compute_pdf.cu:40-120):
//     pdf_p = (1/k_i) * sum_{j' in N(i)} prod_a [ (1/h) * 0.39894228 * exp(-0.5 * ((x_j'a - x_ja) / (r_b h))^2) ]
// The reference evaluates three double-precision exp() per term and makes every thread re-gather all k_i neighbour
// coordinates through two dependent global loads.  Here the product of the three Gaussians is folded into ONE
// exponential of the summed exponent, evaluated in fp32 on the SFU (ex2.approx with the -0.5*log2(e)/(r h)^2 factor
// pre-multiplied; relative error of a term <= ~1e-6 and of the sum far below the 1e-5 parity bound), and the CTA
// stages the coordinates of every neighbour its 256 pairs can touch -- a contiguous window of the pair list, because
// the list is CSR-ordered by sample -- in shared memory once (SoA, conflict-free broadcast reads).
#include "common.cuh"

namespace mc {
namespace {

constexpr int kPdfThreads = 256;
constexpr int kPdfStage = 2048;  // staged neighbour coordinates per CTA (24 KB)

// BY_SAMPLE: bids holds the SAMPLES' batch ids (a neighbour lies in its sample's cloud), else the points'.
template <bool BY_SAMPLE>
__global__ void __launch_bounds__(kPdfThreads)
    pdf_kernel(bool scaleInv, float window, float radius, const int* __restrict__ dM, int capM,
               const int* __restrict__ dK, int capK, const float* __restrict__ pts, const int* __restrict__ bids,
               const float* __restrict__ mn, const float* __restrict__ mx, const int* __restrict__ start,
               const int2* __restrict__ pairs, float* __restrict__ pdf) {
    __shared__ float sx[kPdfStage], sy[kPdfStage], sz[kPdfStage];
    const int M = live_count(dM, capM);
    const int K = live_count(dK, capK);
    const int p0 = blockIdx.x * kPdfThreads;
    if (p0 >= K) return;
    const int pLast = min(p0 + kPdfThreads, K) - 1;
    // window of the pair list covering every neighbour list this CTA reads
    const int iFirst = pairs[p0].y, iLast = pairs[pLast].y;
    const int wLo = start[iFirst];
    const int wHi = min((iLast < M - 1) ? start[iLast + 1] : K, K);  // clamp: K may be capacity-truncated
    const int staged = min(wHi - wLo, kPdfStage);
    for (int t = threadIdx.x; t < staged; t += kPdfThreads) {
        int j = pairs[wLo + t].x;
        sx[t] = pts[(size_t)j * 3];
        sy[t] = pts[(size_t)j * 3 + 1];
        sz[t] = pts[(size_t)j * 3 + 2];
    }
    __syncthreads();

    const int p = p0 + threadIdx.x;
    if (p >= K) return;
    const int2 pr = pairs[p];
    const int j = pr.x, i = pr.y;
    const float ext = max_extent(mn, mx, bids[BY_SAMPLE ? i : j]);
    const float r = scaleInv ? radius * ext : radius;
    const int lo = start[i];
    const int hi = min((i < M - 1) ? start[i + 1] : K, K);
    // exp(-0.5 * |d|^2 / (r h)^2) = 2^(|d|^2 * e2)
    const float e2 = pdf_exponent_scale(r, window);
    const float xj = pts[(size_t)j * 3], yj = pts[(size_t)j * 3 + 1], zj = pts[(size_t)j * 3 + 2];
    float acc = 0.0f;
    int q = lo;
    const int qStagedEnd = min(hi, wLo + staged);
    for (; q < qStagedEnd; ++q) {
        int t = q - wLo;
        acc += pdf_term(sx[t] - xj, sy[t] - yj, sz[t] - zj, e2);
    }
    for (; q < hi; ++q) {  // neighbour list longer than the stage: read through L1/L2
        int jj = pairs[q].x;
        acc += pdf_term(pts[(size_t)jj * 3] - xj, pts[(size_t)jj * 3 + 1] - yj, pts[(size_t)jj * 3 + 2] - zj, e2);
    }
    pdf[p] = pdf_finish(acc, window, hi - lo);
}

}  // namespace

// mc_compute_pdf for a pair list just produced by the fused neighbour kernel (neighbors.cu), which has the samples'
// batch ids at hand, not the points'.
int launch_pdf_by_sample(bool scaleInv, float window, float radius, const int* dM, int capM, const int* dK, int capK,
                         const float* pts, const int* sampleBids, const float* mn, const float* mx, const int* start,
                         const int* pairs, float* pdf, cudaStream_t s) {
    if (capK <= 0) return 0;
    pdf_kernel<true><<<cdiv(capK, kPdfThreads), kPdfThreads, 0, s>>>(scaleInv, window, radius, dM, capM, dK, capK, pts,
                                                                      sampleBids, mn, mx, start,
                                                                      reinterpret_cast<const int2*>(pairs), pdf);
    return check_launch("mc_find_neighbors_pdf_fused (densities)");
}

}  // namespace mc

using namespace mc;

extern "C" int mc_compute_pdf(int scaleInv, float window, int numSamples, const int* dNumSamples, int numNeighbors,
                              const int* dNumNeighbors, float radius, const float* pts, const int* batchIds,
                              const float* aabbMin, const float* aabbMax, const int* startIdx, const int* pairs,
                              float* pdfs, void* stream) {
    MC_REQUIRE(window > 0.0f, "mc_compute_pdf: window must be positive");
    MC_REQUIRE(radius > 0.0f, "mc_compute_pdf: radius must be positive");
    if (numNeighbors <= 0) return 0;
    pdf_kernel<false><<<cdiv(numNeighbors, kPdfThreads), kPdfThreads, 0, (cudaStream_t)stream>>>(
        scaleInv != 0, window, radius, dNumSamples, numSamples, dNumNeighbors, numNeighbors, pts, batchIds, aabbMin,
        aabbMax, startIdx, reinterpret_cast<const int2*>(pairs), pdfs);
    return check_launch("mc_compute_pdf");
}
