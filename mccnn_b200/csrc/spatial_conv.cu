// spatial_conv.cu -- Monte-Carlo convolution forward:
//     out[i,o] = sum_{p=(j,i)} sum_c  f[j,c] * kappa_{o*C+c}(delta_p) / (pdf_p * cnt_i),   delta_p = (x_j - s_i) / r_b
// with kappa the block-structured MLP 3 -> 8 -> 8 -> 8 per block of 8 output neurons.
//
// Replaces spatialConvCPU / evaluateMLPKernel / evaluateMLPNoCombinKernel (tf_ops/spatial_conv.cu:24-325,796-871).
// The reference gives every (pair, MLP block) an 8-thread group that re-reads the block's 176 parameters from
// global memory, exchanges the hidden activations through shared memory with two CTA-wide barriers, and issues one
// global float atomicAdd per (pair, neuron).  Here the neighbour list is consumed in its CSR order:
//   * a WARP owns one sample i and a chunk of kChunkBlocks MLP blocks; a LANE owns a pair (32 pairs per sweep) and
//     evaluates the whole block MLP in registers, reading parameters as 128-bit shared-memory broadcasts
//     (first layer packed as (w_x, w_y, w_z, b) per neuron);
//   * the Monte-Carlo sum over the sample's pairs is accumulated per lane and closed with a 31-shuffle
//     reduce-scatter, after which lane l holds neuron chunk*32 + l: one coalesced 128-byte store, NO atomics, and a
//     summation order that is fixed (deterministic output, unlike the reference's atomics).
// FP32 FFMA throughout: parity with the reference is 1e-5 relative, which TF32 tensor-core inputs (10-bit
// mantissa) cannot meet, and at 8x8 per block the layers are not GEMM-shaped ("why no tensor cores").
#include "common.cuh"

namespace mc {
namespace {

constexpr int kConvWarps = 4;
constexpr int kChunkBlocks = 4;                   // MLP blocks per warp -> 32 neurons, one per lane after the reduce
constexpr int kChunkNeurons = kChunkBlocks * 8;   // 32
constexpr int kBlockParams = 32 + 8 + 64 + 8 + 64;  // W1p[8][4] (w,b1) | b2[8] | W2[8][8] | bo[8] | Wo[8][8] = 176

enum ConvMode { kGeneric = 0, kDepthwise = 1, kSingleIn = 2 };

struct ConvParams {
    const float* W1;
    const float* b1;
    const float* W2;
    const float* b2;
    const float* Wo;
    const float* bo;
};

// one butterfly step of the reduce-scatter: lanes with bit OFF set keep the upper half of the live values
template <int OFF>
__device__ __forceinline__ void reduce_scatter_step(float (&acc)[kChunkNeurons], int lane) {
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int k = 0; k < OFF; ++k) {
        float send = upper ? acc[k] : acc[k + OFF];
        float keep = upper ? acc[k + OFF] : acc[k];
        acc[k] = keep + __shfl_xor_sync(MC_FULL_MASK, send, OFF);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kConvWarps * 32)
    spatial_conv_kernel(bool avg, bool scaleInv, float radius, const int* __restrict__ dM, int capM,
                        const int* __restrict__ dK, int capK, int C, int nOut, int numBlocks, int outStride,
                        const float* __restrict__ pts, const int* __restrict__ bids,
                        const float* __restrict__ feats, const float* __restrict__ pdfs,
                        const float* __restrict__ samples, const int* __restrict__ start,
                        const int2* __restrict__ pairs, const float* __restrict__ mn, const float* __restrict__ mx,
                        ConvParams prm, float* __restrict__ out) {
    __shared__ __align__(16) float sW[kChunkBlocks * kBlockParams];
    const int chunk = blockIdx.y;
    const int lane = lane_id();

    // stage this chunk's parameters (zero for blocks past the end)
    for (int t = threadIdx.x; t < kChunkBlocks * kBlockParams; t += blockDim.x) {
        int q = t / kBlockParams, e = t - q * kBlockParams;
        int blk = chunk * kChunkBlocks + q;
        float v = 0.0f;
        if (blk < numBlocks) {
            int nr0 = blk * 8;
            if (e < 32) {
                int tt = e >> 2, a = e & 3;
                v = a < 3 ? prm.W1[(nr0 + tt) * 3 + a] : prm.b1[nr0 + tt];
            } else if (e < 40) {
                v = prm.b2[nr0 + e - 32];
            } else if (e < 104) {
                v = prm.W2[blk * 64 + e - 40];
            } else if (e < 112) {
                v = prm.bo[nr0 + e - 104];
            } else {
                v = prm.Wo[blk * 64 + e - 112];
            }
        }
        sW[t] = v;
    }
    __syncthreads();

    const int M = live_count(dM, capM);
    const int K = live_count(dK, capK);
    const int i = blockIdx.x * kConvWarps + (threadIdx.x >> 5);
    if (i >= capM) return;
    const int n = chunk * kChunkNeurons + lane;  // the neuron this lane stores
    if (i >= M) {                                // capacity padding rows: defined (zero) output
        if (n < nOut) out[(size_t)i * outStride + n] = 0.0f;
        return;
    }
    const int lo = start[i];
    const int hi = min((i < M - 1) ? start[i + 1] : K, K);  // clamp: K may be capacity-truncated
    const float cnt = avg ? (float)(hi - lo) : 1.0f;
    const float six = samples[(size_t)i * 3], siy = samples[(size_t)i * 3 + 1], siz = samples[(size_t)i * 3 + 2];

    float acc[kChunkNeurons];
#pragma unroll
    for (int a = 0; a < kChunkNeurons; ++a) acc[a] = 0.0f;

    for (int p = lo + lane; p < hi; p += 32) {
        const int j = pairs[p].x;
        const float ext = max_extent(mn, mx, bids[j]);
        const float r = scaleInv ? radius * ext : radius;
        const float dx = (pts[(size_t)j * 3] - six) / r;
        const float dy = (pts[(size_t)j * 3 + 1] - siy) / r;
        const float dz = (pts[(size_t)j * 3 + 2] - siz) / r;
        const float invw = 1.0f / (pdfs[p] * cnt);
        const float* frow = feats + (size_t)j * C;
        float fSingle = 0.0f;
        if (MODE == kSingleIn) fSingle = frow[0] * invw;

#pragma unroll
        for (int q = 0; q < kChunkBlocks; ++q) {
            const int blk = chunk * kChunkBlocks + q;
            if (blk >= numBlocks) continue;  // warp-uniform; parameters of such blocks are zero anyway
            // parameters are re-read from shared memory (128-bit broadcasts) for every pair; the empty asm keeps
            // the compiler from hoisting the 704 loop-invariant loads into registers (it spills otherwise)
            int wOff = q * kBlockParams;
            asm volatile("" : "+r"(wOff));
            const float4* w = reinterpret_cast<const float4*>(sW + wOff);
            float h1[8], h2[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float4 a = w[t];
                h1[t] = fmaxf(fmaf(dx, a.x, fmaf(dy, a.y, fmaf(dz, a.z, a.w))), 0.0f);
            }
            const float4 b2lo = w[8], b2hi = w[9];
            const float b2[8] = {b2lo.x, b2lo.y, b2lo.z, b2lo.w, b2hi.x, b2hi.y, b2hi.z, b2hi.w};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float4 a = w[10 + 2 * t], c = w[11 + 2 * t];
                float s = b2[t];
                s = fmaf(h1[0], a.x, s);
                s = fmaf(h1[1], a.y, s);
                s = fmaf(h1[2], a.z, s);
                s = fmaf(h1[3], a.w, s);
                s = fmaf(h1[4], c.x, s);
                s = fmaf(h1[5], c.y, s);
                s = fmaf(h1[6], c.z, s);
                s = fmaf(h1[7], c.w, s);
                h2[t] = fmaxf(s, 0.0f);
            }
            const float4 bolo = w[26], bohi = w[27];
            const float bo[8] = {bolo.x, bolo.y, bolo.z, bolo.w, bohi.x, bohi.y, bohi.z, bohi.w};
            float f[8];
            if (MODE == kDepthwise) {
                const float4 f0 = *reinterpret_cast<const float4*>(frow + blk * 8);
                const float4 f1 = *reinterpret_cast<const float4*>(frow + blk * 8 + 4);
                f[0] = f0.x * invw; f[1] = f0.y * invw; f[2] = f0.z * invw; f[3] = f0.w * invw;
                f[4] = f1.x * invw; f[5] = f1.y * invw; f[6] = f1.z * invw; f[7] = f1.w * invw;
            } else if (MODE == kGeneric) {
#pragma unroll
                for (int t = 0; t < 8; ++t) f[t] = frow[(blk * 8 + t) % C] * invw;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float4 a = w[28 + 2 * t], c = w[29 + 2 * t];
                float s = bo[t];
                s = fmaf(h2[0], a.x, s);
                s = fmaf(h2[1], a.y, s);
                s = fmaf(h2[2], a.z, s);
                s = fmaf(h2[3], a.w, s);
                s = fmaf(h2[4], c.x, s);
                s = fmaf(h2[5], c.y, s);
                s = fmaf(h2[6], c.z, s);
                s = fmaf(h2[7], c.w, s);
                acc[q * 8 + t] = fmaf(MODE == kSingleIn ? fSingle : f[t], s, acc[q * 8 + t]);
            }
        }
    }

    // reduce-scatter over the 32 lanes: afterwards lane l holds the warp total of acc[l]
    reduce_scatter_step<16>(acc, lane);
    reduce_scatter_step<8>(acc, lane);
    reduce_scatter_step<4>(acc, lane);
    reduce_scatter_step<2>(acc, lane);
    reduce_scatter_step<1>(acc, lane);
    if (n < nOut) out[(size_t)i * outStride + n] = acc[0];
}

// generic multi-feature case (C > 1): out[i,o] = sum_c S[i, o*C + c]
__global__ void __launch_bounds__(256) combine_kernel(const int* __restrict__ dM, int capM, int C, int Co, int sStride,
                                                      const float* __restrict__ S, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)capM * Co) return;
    int i = (int)(t / Co), o = (int)(t - (long long)i * Co);
    const float* s = S + (size_t)i * sStride + (size_t)o * C;
    float a = 0.0f;
    for (int c = 0; c < C; ++c) a += s[c];
    out[t] = a;
}

}  // namespace
}  // namespace mc

using namespace mc;

// floats of workspace spatial_conv needs (only the multi-feature case with more than one input feature)
extern "C" long long mc_spatial_conv_workspace_floats(int numSamples, int numInFeatures, int numOutFeatures,
                                                      int combin) {
    if (!combin || numInFeatures == 1) return 0;
    long long nOut = (long long)numInFeatures * numOutFeatures;
    long long nb = (nOut + 7) / 8;
    return (long long)(numSamples > 0 ? numSamples : 1) * nb * 8;
}

// spatialConvCPU, tf_ops/spatial_conv.cu:796-871.  Parameter buffers are the flat layouts the reference kernels
// index (spatial_conv.cu:49-73,172): W1[(8q+t)*3+a], b1[8q+t], W2[q*64+t*8+u], b2[8q+t], Wo[(8q+t)*8+u], bo[8q+t].
extern "C" int mc_spatial_conv(int avg, int scaleInv, int numNeighbors, const int* dNumNeighbors, int numInFeatures,
                               int numOutFeatures, int numSamples, const int* dNumSamples, int combin, float radius,
                               const float* pts, const int* batchIds, const float* features, const float* pdfs,
                               const float* samples, const int* startIdx, const int* pairs, const float* aabbMin,
                               const float* aabbMax, const float* W1, const float* b1, const float* W2,
                               const float* b2, const float* Wout, const float* bout, float* workspace,
                               float* outFeatures, void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_spatial_conv: radius must be positive");
    MC_REQUIRE(numInFeatures > 0 && numOutFeatures > 0, "mc_spatial_conv: feature counts must be positive");
    MC_REQUIRE(combin || numInFeatures == numOutFeatures,
               "mc_spatial_conv: depth-wise convolution needs equal input and output feature counts");
    MC_REQUIRE(combin || numInFeatures % 8 == 0,
               "mc_spatial_conv: depth-wise convolution needs a feature count multiple of the block size 8");
    if (numSamples <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int C = numInFeatures;
    const int nOut = combin ? C * numOutFeatures : C;
    const int nb = (nOut + 7) / 8;
    const int chunks = cdiv(nb, kChunkBlocks);
    dim3 grid(cdiv(numSamples, kConvWarps), chunks);
    ConvParams prm{W1, b1, W2, b2, Wout, bout};
    const int2* prs = reinterpret_cast<const int2*>(pairs);
    if (!combin) {
        MC_REQUIRE(((uintptr_t)features) % 16 == 0, "mc_spatial_conv: feature buffer must be 16-byte aligned");
        spatial_conv_kernel<kDepthwise><<<grid, kConvWarps * 32, 0, s>>>(
            avg != 0, scaleInv != 0, radius, dNumSamples, numSamples, dNumNeighbors, numNeighbors, C, nOut, nb, C, pts,
            batchIds, features, pdfs, samples, startIdx, prs, aabbMin, aabbMax, prm, outFeatures);
    } else if (C == 1) {
        spatial_conv_kernel<kSingleIn><<<grid, kConvWarps * 32, 0, s>>>(
            avg != 0, scaleInv != 0, radius, dNumSamples, numSamples, dNumNeighbors, numNeighbors, C, nOut, nb,
            numOutFeatures, pts, batchIds, features, pdfs, samples, startIdx, prs, aabbMin, aabbMax, prm,
            outFeatures);
    } else {
        MC_REQUIRE(workspace != nullptr, "mc_spatial_conv: workspace required for multi-feature convolutions");
        spatial_conv_kernel<kGeneric><<<grid, kConvWarps * 32, 0, s>>>(
            avg != 0, scaleInv != 0, radius, dNumSamples, numSamples, dNumNeighbors, numNeighbors, C, nb * 8, nb,
            nb * 8, pts, batchIds, features, pdfs, samples, startIdx, prs, aabbMin, aabbMax, prm, workspace);
        long long total = (long long)numSamples * numOutFeatures;
        combine_kernel<<<cdiv(total, 256), 256, 0, s>>>(dNumSamples, numSamples, C, numOutFeatures, nb * 8, workspace,
                                                        outFeatures);
        count_launches(1);
    }
    return check_launch("mc_spatial_conv");
}
