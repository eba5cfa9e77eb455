// spatial_conv.cu -- Monte-Carlo convolution forward:
//     out[i,o] = sum_{p=(j,i)} sum_c  f[j,c] * kappa_{o*C+c}(delta_p) / (pdf_p * cnt_i),   delta_p = (x_j - s_i) / r_b
// with kappa the block-structured MLP 3 -> 8 -> 8 -> 8 per block of 8 output neurons.
//
// Replaces spatialConvCPU / evaluateMLPKernel / evaluateMLPNoCombinKernel (tf_ops/spatial_conv.cu:24-325,796-871).
// The reference gives every (pair, MLP block) an 8-thread group that re-reads the block's 176 parameters from
// global memory, exchanges the hidden activations through shared memory with two CTA-wide barriers, and issues one
// global float atomicAdd per (pair, neuron).  Here the neighbour list is consumed in its CSR order:
//   * a WARP owns a run of G consecutive samples and a chunk of 4 MLP blocks (32 neurons).  It walks the run's pair
//     range in tiles of 64 pairs, TWO pairs per lane, so every lane is busy regardless of how many neighbours a
//     sample has, and each 128-bit shared-memory broadcast of parameters feeds two pairs;
//   * the whole block MLP of a pair lives in registers (first layer packed as (w_x, w_y, w_z, b) per neuron);
//   * the Monte-Carlo sum is a SEGMENTED reduction over the lanes of a tile: for each sample present in the tile the
//     lanes' contributions are masked and closed with a 31-shuffle reduce-scatter over the chunk's 32 neurons, after
//     which lane l holds neuron l (one coalesced 128-byte store); partial sums of a sample that continues into the
//     next tile are carried in a register.  NO atomics, and a summation order that is a pure function of the
//     neighbour list: bit-reproducible, unlike the reference.
// FP32 FFMA throughout: parity with the reference is 1e-5 relative, which TF32 tensor-core inputs (10-bit
// mantissa) cannot meet, and at 8x8 per block the layers are not GEMM-shaped ("why no tensor cores").
#include "common.cuh"

namespace mc {
namespace {

constexpr int kConvWarps = 4;
constexpr int kChunkBlocks = 4;                     // MLP blocks per warp
constexpr int kChunkNeurons = kChunkBlocks * 8;     // 32
constexpr int kBlockParams = 32 + 8 + 64 + 8 + 64;  // W1p[8][4] (w,b1) | b2[8] | W2[8][8] | bo[8] | Wo[8][8] = 176
constexpr int kTile = 64;                           // pairs per tile: two per lane
constexpr int kRadiusTable = 1024;                  // per-cloud 1/r kept in shared memory up to this batch size

enum ConvMode { kGeneric = 0, kDepthwise = 1, kSingleIn = 2 };

struct ConvArgs {
    bool avg, scaleInv;
    float radius;
    const int* dM;
    int capM;
    const int* dK;
    int capK;
    int B, C, nOut, numBlocks, outStride, G;
    const float* pts;
    const int* bids;
    const float* feats;
    const float* pdfs;
    const float* samples;
    const int* start;
    const int2* pairs;
    const float* mn;
    const float* mx;
    const float *W1, *b1, *W2, *b2, *Wo, *bo;
    float* out;
};

struct PairState {  // one (neighbour, sample) pair held by a lane
    float dx, dy, dz;  // (x_j - s_i) / r
    float invw;        // 1 / (pdf * cnt); 0 for an empty slot
    int i;             // sample, -1 for an empty slot
    int j;             // neighbour point
};

// 3 -> 8 -> 8 -> 8 block MLP for two pairs at once; parameters from shared memory (broadcast reads)
__device__ __forceinline__ void block_mlp2(const float4* __restrict__ w, const PairState& a, const PairState& b,
                                           float (&ka)[8], float (&kb)[8]) {
    float h1a[8], h1b[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float4 p = w[t];
        h1a[t] = fmaxf(fmaf(a.dx, p.x, fmaf(a.dy, p.y, fmaf(a.dz, p.z, p.w))), 0.0f);
        h1b[t] = fmaxf(fmaf(b.dx, p.x, fmaf(b.dy, p.y, fmaf(b.dz, p.z, p.w))), 0.0f);
    }
    const float4 b2lo = w[8], b2hi = w[9];
    const float b2[8] = {b2lo.x, b2lo.y, b2lo.z, b2lo.w, b2hi.x, b2hi.y, b2hi.z, b2hi.w};
    float h2a[8], h2b[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float4 p = w[10 + 2 * t], q = w[11 + 2 * t];
        float sa = b2[t], sb = b2[t];
        sa = fmaf(h1a[0], p.x, sa); sb = fmaf(h1b[0], p.x, sb);
        sa = fmaf(h1a[1], p.y, sa); sb = fmaf(h1b[1], p.y, sb);
        sa = fmaf(h1a[2], p.z, sa); sb = fmaf(h1b[2], p.z, sb);
        sa = fmaf(h1a[3], p.w, sa); sb = fmaf(h1b[3], p.w, sb);
        sa = fmaf(h1a[4], q.x, sa); sb = fmaf(h1b[4], q.x, sb);
        sa = fmaf(h1a[5], q.y, sa); sb = fmaf(h1b[5], q.y, sb);
        sa = fmaf(h1a[6], q.z, sa); sb = fmaf(h1b[6], q.z, sb);
        sa = fmaf(h1a[7], q.w, sa); sb = fmaf(h1b[7], q.w, sb);
        h2a[t] = fmaxf(sa, 0.0f);
        h2b[t] = fmaxf(sb, 0.0f);
    }
    const float4 bolo = w[26], bohi = w[27];
    const float bo[8] = {bolo.x, bolo.y, bolo.z, bolo.w, bohi.x, bohi.y, bohi.z, bohi.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float4 p = w[28 + 2 * t], q = w[29 + 2 * t];
        float sa = bo[t], sb = bo[t];
        sa = fmaf(h2a[0], p.x, sa); sb = fmaf(h2b[0], p.x, sb);
        sa = fmaf(h2a[1], p.y, sa); sb = fmaf(h2b[1], p.y, sb);
        sa = fmaf(h2a[2], p.z, sa); sb = fmaf(h2b[2], p.z, sb);
        sa = fmaf(h2a[3], p.w, sa); sb = fmaf(h2b[3], p.w, sb);
        sa = fmaf(h2a[4], q.x, sa); sb = fmaf(h2b[4], q.x, sb);
        sa = fmaf(h2a[5], q.y, sa); sb = fmaf(h2b[5], q.y, sb);
        sa = fmaf(h2a[6], q.z, sa); sb = fmaf(h2b[6], q.z, sb);
        sa = fmaf(h2a[7], q.w, sa); sb = fmaf(h2b[7], q.w, sb);
        ka[t] = sa;
        kb[t] = sb;
    }
}

// one butterfly level of the reduce-scatter over the 32 lanes: lanes with bit OFF set keep the upper half of the
// live values.  After levels 16, 8, 4, 2, 1 lane l holds the warp total of value index l.
template <int OFF>
__device__ __forceinline__ void reduce_scatter_step(float (&v)[kChunkNeurons], int lane) {
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int k = 0; k < OFF; ++k) {
        const float send = upper ? v[k] : v[k + OFF];
        const float keep = upper ? v[k + OFF] : v[k];
        v[k] = keep + __shfl_xor_sync(MC_FULL_MASK, send, OFF);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kConvWarps * 32) spatial_conv_kernel(const ConvArgs a) {
    __shared__ __align__(16) float sW[kChunkBlocks * kBlockParams];
    __shared__ float sInvR[kRadiusTable];
    const int chunk = blockIdx.y;
    const int lane = lane_id();

    // stage this chunk's parameters (zero for blocks past the end) and the per-cloud 1/r
    for (int t = threadIdx.x; t < kChunkBlocks * kBlockParams; t += blockDim.x) {
        const int q = t / kBlockParams, e = t - q * kBlockParams;
        const int blk = chunk * kChunkBlocks + q;
        float v = 0.0f;
        if (blk < a.numBlocks) {
            const int nr0 = blk * 8;
            if (e < 32) {
                const int tt = e >> 2, c = e & 3;
                v = c < 3 ? a.W1[(nr0 + tt) * 3 + c] : a.b1[nr0 + tt];
            } else if (e < 40) {
                v = a.b2[nr0 + e - 32];
            } else if (e < 104) {
                v = a.W2[blk * 64 + e - 40];
            } else if (e < 112) {
                v = a.bo[nr0 + e - 104];
            } else {
                v = a.Wo[blk * 64 + e - 112];
            }
        }
        sW[t] = v;
    }
    const bool radiusTable = a.B <= kRadiusTable;
    if (radiusTable)
        for (int b = threadIdx.x; b < a.B; b += blockDim.x) {
            const float r = a.scaleInv ? a.radius * max_extent(a.mn, a.mx, b) : a.radius;
            sInvR[b] = 1.0f / r;
        }
    __syncthreads();

    const int M = live_count(a.dM, a.capM);
    const int K = live_count(a.dK, a.capK);
    const int s0 = (blockIdx.x * kConvWarps + (threadIdx.x >> 5)) * a.G;  // first sample of this warp's run
    if (s0 >= a.capM) return;
    const int n = chunk * kChunkNeurons + lane;  // neuron this lane zero-fills

    // samples without neighbours (and capacity padding rows) get a defined zero output
    const int sEndCap = min(s0 + a.G, a.capM);
    for (int s = s0; s < sEndCap; ++s) {
        bool empty = s >= M;
        if (!empty) {
            const int lo = a.start[s];
            const int hi = min((s < M - 1) ? a.start[s + 1] : K, K);
            empty = hi <= lo;
        }
        if (empty && n < a.nOut) a.out[(size_t)s * a.outStride + n] = 0.0f;
    }
    if (s0 >= M) return;
    const int s1 = min(s0 + a.G, M);
    const int P0 = min(a.start[s0], K);
    const int P1 = min((s1 < M) ? a.start[s1] : K, K);

    float carry = 0.0f;   // partial sum (neuron n) of the sample that continues from the previous tile
    int openSample = -1;  // that sample

    for (int tile = P0; tile < P1; tile += kTile) {
        PairState ps[2];
        int hiOf[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = tile + u * 32 + lane;
            ps[u].i = -1;
            ps[u].j = 0;
            ps[u].dx = ps[u].dy = ps[u].dz = ps[u].invw = 0.0f;
            hiOf[u] = 0;
            if (p < P1) {
                const int2 pr = a.pairs[p];
                const int j = pr.x, i = pr.y;
                float invR;
                if (radiusTable) {
                    invR = sInvR[a.bids[j]];
                } else {
                    const float ext = max_extent(a.mn, a.mx, a.bids[j]);
                    invR = 1.0f / (a.scaleInv ? a.radius * ext : a.radius);
                }
                const int lo = a.start[i];
                const int hi = min((i < M - 1) ? a.start[i + 1] : K, K);
                const float cnt = a.avg ? (float)(hi - lo) : 1.0f;
                ps[u].i = i;
                ps[u].j = j;
                ps[u].dx = (a.pts[(size_t)j * 3] - a.samples[(size_t)i * 3]) * invR;
                ps[u].dy = (a.pts[(size_t)j * 3 + 1] - a.samples[(size_t)i * 3 + 1]) * invR;
                ps[u].dz = (a.pts[(size_t)j * 3 + 2] - a.samples[(size_t)i * 3 + 2]) * invR;
                ps[u].invw = 1.0f / (a.pdfs[p] * cnt);
                hiOf[u] = hi;
            }
        }
        // samples present in the tile: consecutive ids iFirst..iLast (pairs are CSR-ordered by sample)
        const int iFirst = __shfl_sync(MC_FULL_MASK, ps[0].i, 0);
        const int lastSlot = min(tile + kTile, P1) - 1 - tile;
        const int iLast = __shfl_sync(MC_FULL_MASK, lastSlot >= 32 ? ps[1].i : ps[0].i, lastSlot & 31);
        const int hiLast = __shfl_sync(MC_FULL_MASK, lastSlot >= 32 ? hiOf[1] : hiOf[0], lastSlot & 31);
        const bool lastComplete = hiLast <= tile + kTile;  // does sample iLast end inside this tile?
        float fs[2] = {0.f, 0.f};
        if (MODE == kSingleIn) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (ps[u].i >= 0) fs[u] = a.feats[ps[u].j] * ps[u].invw;
        }

        float va[kChunkNeurons], vb[kChunkNeurons];  // f * kappa / (pdf * cnt) of this lane's two pairs
#pragma unroll
        for (int q = 0; q < kChunkBlocks; ++q) {
            const int blk = chunk * kChunkBlocks + q;
            if (blk >= a.numBlocks) {  // warp-uniform
#pragma unroll
                for (int t = 0; t < 8; ++t) va[q * 8 + t] = vb[q * 8 + t] = 0.0f;
                continue;
            }
            // parameters are re-read from shared memory (128-bit broadcasts) for every tile; the empty asm keeps the
            // compiler from hoisting the loop-invariant loads into registers (it spills otherwise)
            int wOff = q * kBlockParams;
            asm volatile("" : "+r"(wOff));
            float ka[8], kb[8];
            block_mlp2(reinterpret_cast<const float4*>(sW + wOff), ps[0], ps[1], ka, kb);
            // weight each kernel value by the pair's feature / (pdf * cnt)
            if (MODE == kSingleIn) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    va[q * 8 + t] = ka[t] * fs[0];
                    vb[q * 8 + t] = kb[t] * fs[1];
                }
            } else {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (ps[u].i >= 0) {
                        const float* frow = a.feats + (size_t)ps[u].j * a.C;
                        if (MODE == kDepthwise) {
                            const float4 f0 = *reinterpret_cast<const float4*>(frow + blk * 8);
                            const float4 f1 = *reinterpret_cast<const float4*>(frow + blk * 8 + 4);
                            f[0] = f0.x; f[1] = f0.y; f[2] = f0.z; f[3] = f0.w;
                            f[4] = f1.x; f[5] = f1.y; f[6] = f1.z; f[7] = f1.w;
                        } else {
#pragma unroll
                            for (int t = 0; t < 8; ++t) f[t] = frow[(blk * 8 + t) % a.C];
                        }
                    }
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (u == 0)
                            va[q * 8 + t] = ka[t] * (f[t] * ps[0].invw);
                        else
                            vb[q * 8 + t] = kb[t] * (f[t] * ps[1].invw);
                    }
                }
            }
        }
        // segmented sum over the samples of the tile: mask, then one 31-shuffle reduce-scatter over the 32 neurons of
        // the chunk (independent shuffles per butterfly level), after which lane l holds neuron chunk*32 + l
        for (int cur = iFirst; cur <= iLast; ++cur) {
            const bool m0 = ps[0].i == cur, m1 = ps[1].i == cur;
            if (__ballot_sync(MC_FULL_MASK, m0 || m1) == 0u) continue;  // sample without pairs
            float w[kChunkNeurons];
#pragma unroll
            for (int t = 0; t < kChunkNeurons; ++t) w[t] = (m0 ? va[t] : 0.0f) + (m1 ? vb[t] : 0.0f);
            reduce_scatter_step<16>(w, lane);
            reduce_scatter_step<8>(w, lane);
            reduce_scatter_step<4>(w, lane);
            reduce_scatter_step<2>(w, lane);
            reduce_scatter_step<1>(w, lane);
            float tot = w[0];
            if (cur == openSample) tot += carry;
            if (cur < iLast || lastComplete) {
                if (n < a.nOut) a.out[(size_t)cur * a.outStride + n] = tot;  // one coalesced 128-byte store per warp
            } else {
                carry = tot;  // the sample continues in the next tile
            }
        }
        openSample = lastComplete ? -1 : iLast;
    }
}

// generic multi-feature case (C > 1): out[i,o] = sum_c S[i, o*C + c]
__global__ void __launch_bounds__(256) combine_kernel(const int* __restrict__ dM, int capM, int C, int Co, int sStride,
                                                      const float* __restrict__ S, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)capM * Co) return;
    int i = (int)(t / Co), o = (int)(t - (long long)i * Co);
    const float* s = S + (size_t)i * sStride + (size_t)o * C;
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc += s[c];
    out[t] = acc;
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
                               int batchSize, const float* pts, const int* batchIds, const float* features,
                               const float* pdfs, const float* samples, const int* startIdx, const int* pairs,
                               const float* aabbMin, const float* aabbMax, const float* W1, const float* b1,
                               const float* W2, const float* b2, const float* Wout, const float* bout,
                               float* workspace, float* outFeatures, void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_spatial_conv: radius must be positive");
    MC_REQUIRE(batchSize > 0, "mc_spatial_conv: batch size must be positive");
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
    // run length per warp: ~3 tiles of pairs per warp (the partially filled last tile stays cheap) -- but the op is
    // latency bound, so never at the price of leaving the GPU short of warps
    long long avgK = numNeighbors / (long long)numSamples;
    int G = (int)(192 / (avgK > 0 ? avgK : 1));
    G = G < 1 ? 1 : (G > 16 ? 16 : G);
    while (G > 1 && (long long)cdiv(numSamples, G) * chunks < 2048) --G;
    ConvArgs a;
    a.avg = avg != 0;
    a.scaleInv = scaleInv != 0;
    a.radius = radius;
    a.dM = dNumSamples;
    a.capM = numSamples;
    a.dK = dNumNeighbors;
    a.capK = numNeighbors;
    a.B = batchSize;
    a.C = C;
    a.numBlocks = nb;
    a.G = G;
    a.pts = pts;
    a.bids = batchIds;
    a.feats = features;
    a.pdfs = pdfs;
    a.samples = samples;
    a.start = startIdx;
    a.pairs = reinterpret_cast<const int2*>(pairs);
    a.mn = aabbMin;
    a.mx = aabbMax;
    a.W1 = W1;
    a.b1 = b1;
    a.W2 = W2;
    a.b2 = b2;
    a.Wo = Wout;
    a.bo = bout;
    dim3 grid(cdiv(cdiv(numSamples, G), kConvWarps), chunks);
    if (!combin) {
        MC_REQUIRE(((uintptr_t)features) % 16 == 0, "mc_spatial_conv: feature buffer must be 16-byte aligned");
        a.nOut = nOut;
        a.outStride = C;
        a.out = outFeatures;
        spatial_conv_kernel<kDepthwise><<<grid, kConvWarps * 32, 0, s>>>(a);
    } else if (C == 1) {
        a.nOut = nOut;
        a.outStride = numOutFeatures;
        a.out = outFeatures;
        spatial_conv_kernel<kSingleIn><<<grid, kConvWarps * 32, 0, s>>>(a);
    } else {
        MC_REQUIRE(workspace != nullptr, "mc_spatial_conv: workspace required for multi-feature convolutions");
        a.nOut = nb * 8;
        a.outStride = nb * 8;
        a.out = workspace;
        spatial_conv_kernel<kGeneric><<<grid, kConvWarps * 32, 0, s>>>(a);
        long long total = (long long)numSamples * numOutFeatures;
        combine_kernel<<<cdiv(total, 256), 256, 0, s>>>(dNumSamples, numSamples, C, numOutFeatures, nb * 8, workspace,
                                                        outFeatures);
        count_launches(1);
    }
    return check_launch("mc_spatial_conv");
}
