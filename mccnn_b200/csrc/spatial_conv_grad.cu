// spatial_conv_grad.cu -- backward of the Monte-Carlo convolution and of the two permutation ops.
//
// Replaces spatialConvGradsCPU / computedconvj_dKernel / computedconvj_dNoCombinKernel
// (tf_ops/spatial_conv.cu:327-700,873-966), sortPointsStep2GradGPUKernel (tf_ops/sort_gpu.cu:260-277,499-513) and
// getFeaturesSampledPointsGradients (tf_ops/poisson_sampling.cu:158-171,260-274).
//
// The reference gives every (pair, MLP block) an 8-thread group (one thread per neuron) that exchanges the forward
// and backward activations through shared memory with four CTA-wide barriers and issues ~180 global float atomics per
// (pair, block) -- all onto the SAME 176 parameter-gradient addresses of the block.  Here:
//   * the decomposition is kept -- lane t of an 8-lane group is neuron t of the block -- but a group stays on ONE
//     block for a whole range of pairs: the neuron's 38 parameters (its rows AND columns of the two 8x8 layers) live
//     in registers, the activations are exchanged with width-8 shuffles (32 per pair), and the 22 parameter gradients
//     of the neuron are accumulated in registers over the whole range;
//   * a CTA = 4 warps x 4 groups on consecutive pairs of the same block; its partial sums are combined in a fixed
//     order (shuffles across groups, shared memory across warps) and written to a workspace row; a second small
//     kernel adds the rows in order.  Parameter gradients are therefore bit-reproducible; NO parameter atomics;
//   * the feature gradient dF[j,c] is scattered by neighbour j, which the pair list is not sorted by: one float
//     atomicAdd per (pair, neuron) as in the reference (pre-reduced over the group when there is one input feature).
// Arithmetic: FP32 FFMA, forward recomputed per pair exactly as the forward op defines it; the ReLU derivative is taken
// with >= 0 like the reference (spatial_conv.cu:405,431).
#include "common.cuh"

namespace mc {
namespace {

constexpr int kGradWarps = 4;
constexpr int kGradPairsPerCta = 1024;
constexpr int kGradRadiusTable = 1024;  // per-cloud radius kept in shared memory up to this batch size
constexpr int kGradVals = 22;  // per neuron: dW1[3], db1, dW2[8], db2, dWo[8], dbo

struct GradArgs {
    bool avg, scaleInv, combin;
    float radius;
    int K, M, B, C, outC, nOut, numBlocks;
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
    const float* outGrad;
    float* partial;  // [gridDim.x][numBlocks][8][22]
    float* dFeat;
};

__global__ void __launch_bounds__(kGradWarps * 32, 4) spatial_conv_grad_kernel(const GradArgs a) {
    __shared__ float sAcc[kGradWarps][8][kGradVals];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int t = lane & 7, grp = lane >> 3, gbase = lane & 24;
    const int q = blockIdx.y;
    const int n = q * 8 + t;  // this lane's neuron
    const bool neuronLive = n < a.nOut;
    const int c = neuronLive ? (a.combin ? n % a.C : n) : 0;
    const int o = neuronLive ? (a.combin ? n / a.C : n) : 0;

    // parameters of neuron t of block q: flat reference layouts W1[(8q+t)*3+a], W2[q*64+t*8+u], Wo[(8q+t)*8+u]
    const float w1x = a.W1[n * 3], w1y = a.W1[n * 3 + 1], w1z = a.W1[n * 3 + 2], b1v = a.b1[n];
    const float b2v = a.b2[n], bov = a.bo[n];
    float w2row[8], w2col[8], worow[8], wocol[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        w2row[u] = a.W2[q * 64 + t * 8 + u];
        w2col[u] = a.W2[q * 64 + u * 8 + t];
        worow[u] = a.Wo[(size_t)n * 8 + u];
        wocol[u] = (q * 8 + u < a.nOut) ? a.Wo[(size_t)(q * 8 + u) * 8 + t] : 0.0f;
    }
    float aW1x = 0.f, aW1y = 0.f, aW1z = 0.f, ab1 = 0.f, ab2 = 0.f, abo = 0.f;
    float aW2[8], aWo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) aW2[u] = aWo[u] = 0.f;

    // per-cloud radius in shared memory: removes the batch id -> box -> radius chain from every pair
    __shared__ float sRad[kGradRadiusTable];
    const bool radiusTable = a.B <= kGradRadiusTable;
    if (radiusTable) {
        for (int bb = threadIdx.x; bb < a.B; bb += blockDim.x)
            sRad[bb] = a.scaleInv ? a.radius * max_extent(a.mn, a.mx, bb) : a.radius;
        __syncthreads();
    }

    const int p0 = blockIdx.x * kGradPairsPerCta;
    const int p1 = min(a.K, p0 + kGradPairsPerCta);
    // two-stage software pipeline over the warp's steps: (A) pair + pdf of step s+2, (B) everything addressed through
    // the pair (coordinates, batch id, neighbour range, feature, incoming gradient) of step s+1, compute step s --
    // the dependent global loads of a step are in flight while the two previous steps are computed
    struct StageA {
        int2 pr;
        float pdf;
        bool live;
    };
    struct StageB {
        float px, py, pz, sx, sy, sz, pdf, f, g;
        int b, lo, hi, j;
        bool live;
    };
    auto load_a = [&](int base) {
        StageA A;
        const int p = base + grp;
        A.live = p < p1;
        A.pr = A.live ? a.pairs[p] : make_int2(0, 0);
        A.pdf = A.live ? a.pdfs[p] : 1.0f;
        return A;
    };
    auto load_b = [&](const StageA& A) {
        StageB Bv;
        const int j = A.pr.x, i = A.pr.y;
        Bv.j = j;
        Bv.live = A.live;
        Bv.pdf = A.pdf;
        Bv.px = a.pts[(size_t)j * 3];
        Bv.py = a.pts[(size_t)j * 3 + 1];
        Bv.pz = a.pts[(size_t)j * 3 + 2];
        Bv.sx = a.samples[(size_t)i * 3];
        Bv.sy = a.samples[(size_t)i * 3 + 1];
        Bv.sz = a.samples[(size_t)i * 3 + 2];
        Bv.b = a.bids[j];
        Bv.lo = a.start[i];
        Bv.hi = (i < a.M - 1) ? a.start[i + 1] : a.K;
        const bool use = A.live && neuronLive;
        Bv.f = use ? a.feats[(size_t)j * a.C + c] : 0.0f;
        Bv.g = use ? a.outGrad[(size_t)i * a.outC + o] : 0.0f;
        return Bv;
    };
    const int step = kGradWarps * 4;
    const int base0 = p0 + warp * 4;
    StageA nextA = load_a(base0);
    StageB nextB = load_b(nextA);
    nextA = load_a(base0 + step);
    for (int base = base0; base < p1; base += step) {  // warp-uniform trip count
        const StageB cur = nextB;
        if (base + step < p1) {
            nextB = load_b(nextA);
            nextA = load_a(base + 2 * step);
        }
        const bool live = cur.live;
        const int j = cur.j;
        const float r = radiusTable ? sRad[cur.b]
                                    : (a.scaleInv ? a.radius * max_extent(a.mn, a.mx, cur.b) : a.radius);
        const float dx = (cur.px - cur.sx) / r;
        const float dy = (cur.py - cur.sy) / r;
        const float dz = (cur.pz - cur.sz) / r;
        const float w = live ? cur.pdf * (a.avg ? (float)(cur.hi - cur.lo) : 1.0f) : 1.0f;  // dead slots: exact zeros
        const bool use = live && neuronLive;
        const float f = cur.f, g = cur.g;

        // forward, recomputed
        const float z1 = fmaf(dx, w1x, fmaf(dy, w1y, fmaf(dz, w1z, b1v)));
        const float h1 = fmaxf(z1, 0.f);
        float h1u[8], z2 = b2v;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            h1u[u] = __shfl_sync(MC_FULL_MASK, h1, gbase + u);
            z2 = fmaf(h1u[u], w2row[u], z2);
        }
        const float h2 = fmaxf(z2, 0.f);
        float h2u[8], kappa = bov;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            h2u[u] = __shfl_sync(MC_FULL_MASK, h2, gbase + u);
            kappa = fmaf(h2u[u], worow[u], kappa);
        }
        // output layer: dWo[n,u] += (f g / w) relu(z2[u]); dbo[n] += f g / w; dF[j,c] += g kappa / w
        const float common = (f * g) / w;
#pragma unroll
        for (int u = 0; u < 8; ++u) aWo[u] = fmaf(common, h2u[u], aWo[u]);
        abo += common;
        float df = (g * kappa) / w;
        if (a.combin && a.C == 1) {  // all eight neurons feed the same feature: one atomic per pair
            df += __shfl_xor_sync(MC_FULL_MASK, df, 1);
            df += __shfl_xor_sync(MC_FULL_MASK, df, 2);
            df += __shfl_xor_sync(MC_FULL_MASK, df, 4);
            if (t == 0 && live) atomicAdd(&a.dFeat[j], df);
        } else if (use) {
            atomicAdd(&a.dFeat[(size_t)j * a.C + c], df);
        }
        // second hidden layer: dz2[t] = (z2 >= 0) * sum_i g_i f_i Wo[i,t] / w
        const float gf = g * f;
        float dh2 = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) dh2 = fmaf(__shfl_sync(MC_FULL_MASK, gf, gbase + u), wocol[u], dh2);
        const float dz2 = ((z2 >= 0.f ? 1.0f : 0.0f) * dh2) / w;
#pragma unroll
        for (int u = 0; u < 8; ++u) aW2[u] = fmaf(dz2, h1u[u], aW2[u]);
        ab2 += dz2;
        // first hidden layer: dz1[t] = (z1 >= 0) * sum_i dz2[i] W2[i,t]
        float dh1 = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) dh1 = fmaf(__shfl_sync(MC_FULL_MASK, dz2, gbase + u), w2col[u], dh1);
        const float dz1 = (z1 >= 0.f ? 1.0f : 0.0f) * dh1;
        aW1x = fmaf(dz1, dx, aW1x);
        aW1y = fmaf(dz1, dy, aW1y);
        aW1z = fmaf(dz1, dz, aW1z);
        ab1 += dz1;
    }

    // ---- fixed-order reduction: groups of the warp (shuffles), warps of the CTA (shared memory), then one row
    float v[kGradVals] = {aW1x, aW1y, aW1z, ab1, aW2[0], aW2[1], aW2[2], aW2[3], aW2[4], aW2[5], aW2[6],
                          aW2[7], ab2,  aWo[0], aWo[1], aWo[2], aWo[3], aWo[4], aWo[5], aWo[6], aWo[7], abo};
#pragma unroll
    for (int e = 0; e < kGradVals; ++e) {
        v[e] += __shfl_xor_sync(MC_FULL_MASK, v[e], 8);
        v[e] += __shfl_xor_sync(MC_FULL_MASK, v[e], 16);
    }
    if (grp == 0) {
#pragma unroll
        for (int e = 0; e < kGradVals; ++e) sAcc[warp][t][e] = v[e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 8 * kGradVals; e += blockDim.x) {
        const float* s0 = &sAcc[0][0][0];
        float acc = s0[e];
#pragma unroll
        for (int wv = 1; wv < kGradWarps; ++wv) acc += s0[wv * 8 * kGradVals + e];
        a.partial[((size_t)blockIdx.x * a.numBlocks + q) * (8 * kGradVals) + e] = acc;
    }
}

// sums the per-CTA rows in order and scatters the 22 values of every neuron to the six gradient tensors
__global__ void __launch_bounds__(256) spatial_conv_grad_reduce_kernel(int rows, int numBlocks, int nOut,
                                                                       const float* __restrict__ partial,
                                                                       float* __restrict__ dW1, float* __restrict__ db1,
                                                                       float* __restrict__ dW2, float* __restrict__ db2,
                                                                       float* __restrict__ dWo,
                                                                       float* __restrict__ dbo) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int perRow = numBlocks * 8 * kGradVals;
    if (e >= perRow) return;
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += partial[(size_t)r * perRow + e];
    const int n = e / kGradVals, k = e - n * kGradVals;  // neuron (8q + t), value
    if (k < 3)
        dW1[n * 3 + k] = acc;
    else if (k == 3)
        db1[n] = acc;
    else if (k < 12)
        dW2[n * 8 + (k - 4)] = acc;
    else if (k == 12)
        db2[n] = acc;
    else if (k < 21)
        dWo[(size_t)n * 8 + (k - 13)] = n < nOut ? acc : 0.f;
    else
        dbo[n] = n < nOut ? acc : 0.f;
}

__global__ void __launch_bounds__(256) scatter_rows_kernel(int rows, int rowVecs, const int* __restrict__ idx,
                                                           const float* __restrict__ in, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long r = t / rowVecs;
    if (r >= rows) return;
    const int cc = (int)(t - r * rowVecs);
    out[(size_t)idx[r] * rowVecs + cc] = in[(size_t)r * rowVecs + cc];
}

__global__ void __launch_bounds__(256) gather_rows_kernel(int rows, int rowVecs, const int* __restrict__ idx,
                                                          const float* __restrict__ in, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long r = t / rowVecs;
    if (r >= rows) return;
    const int cc = (int)(t - r * rowVecs);
    out[(size_t)r * rowVecs + cc] = in[(size_t)idx[r] * rowVecs + cc];
}

}  // namespace
}  // namespace mc

using namespace mc;

// floats of workspace mc_spatial_conv_grad needs: one row of 176 * numBlocks partial sums per 1024 pairs
extern "C" long long mc_spatial_conv_grad_workspace_floats(int numNeighbors, int numInFeatures, int numOutFeatures,
                                                           int combin) {
    long long nOut = combin ? (long long)numInFeatures * numOutFeatures : numInFeatures;
    long long nb = (nOut + 7) / 8;
    long long rows = (numNeighbors + kGradPairsPerCta - 1) / kGradPairsPerCta;
    return (rows > 0 ? rows : 1) * nb * 8 * kGradVals;
}

// spatialConvGradsCPU, tf_ops/spatial_conv.cu:873-966.  Gradient buffers have the shapes / flat layouts of the
// parameters they belong to ([3*8nb], [8nb], [8nb*8], [8nb], [8nb*8], [8nb]) and are fully written (the padding
// neurons of the output layer get 0); featureGrads is [numPoints, numInFeatures].
extern "C" int mc_spatial_conv_grad(int avg, int scaleInv, int numNeighbors, int numInFeatures, int numOutFeatures,
                                    int numSamples, int numPoints, int combin, float radius, int batchSize,
                                    const float* pts, const int* batchIds, const float* features, const float* pdfs,
                                    const float* samples, const int* startIdx, const int* pairs,
                                    const float* aabbMin, const float* aabbMax, const float* W1, const float* b1,
                                    const float* W2, const float* b2, const float* Wout, const float* bout,
                                    const float* outFeatureGrads, float* workspace, float* featureGrads,
                                    float* W1Grads, float* b1Grads, float* W2Grads, float* b2Grads, float* WoutGrads,
                                    float* boutGrads, void* stream) {
    MC_REQUIRE(radius > 0.0f, "mc_spatial_conv_grad: radius must be positive");
    MC_REQUIRE(batchSize > 0, "mc_spatial_conv_grad: batch size must be positive");
    MC_REQUIRE(numInFeatures > 0 && numOutFeatures > 0, "mc_spatial_conv_grad: feature counts must be positive");
    MC_REQUIRE(combin || numInFeatures == numOutFeatures,
               "mc_spatial_conv_grad: depth-wise convolution needs equal input and output feature counts");
    MC_REQUIRE(combin || numInFeatures % 8 == 0,
               "mc_spatial_conv_grad: depth-wise convolution needs a feature count multiple of the block size 8");
    MC_REQUIRE(workspace != nullptr, "mc_spatial_conv_grad: workspace required");
    cudaStream_t s = (cudaStream_t)stream;
    const int C = numInFeatures;
    const int nOut = combin ? C * numOutFeatures : C;
    const int nb = (nOut + 7) / 8;
    if (numPoints > 0) cudaMemsetAsync(featureGrads, 0, sizeof(float) * (size_t)numPoints * C, s);
    const int rows = numNeighbors > 0 ? cdiv(numNeighbors, kGradPairsPerCta) : 0;
    if (rows > 0) {
        GradArgs a;
        a.avg = avg != 0;
        a.scaleInv = scaleInv != 0;
        a.combin = combin != 0;
        a.radius = radius;
        a.K = numNeighbors;
        a.M = numSamples;
        a.B = batchSize;
        a.C = C;
        a.outC = combin ? numOutFeatures : C;
        a.nOut = nOut;
        a.numBlocks = nb;
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
        a.outGrad = outFeatureGrads;
        a.partial = workspace;
        a.dFeat = featureGrads;
        MC_REQUIRE(nb <= 65535, "mc_spatial_conv_grad: too many MLP blocks");
        spatial_conv_grad_kernel<<<dim3(rows, nb), kGradWarps * 32, 0, s>>>(a);
        count_launches(1);
    }
    spatial_conv_grad_reduce_kernel<<<cdiv((long long)nb * 8 * kGradVals, 256), 256, 0, s>>>(
        rows, nb, nOut, workspace, W1Grads, b1Grads, W2Grads, b2Grads, WoutGrads, boutGrads);
    return check_launch("mc_spatial_conv_grad");
}

// sortPointsStep2GradGPUKernel, tf_ops/sort_gpu.cu:499-513: inGrad[i,:] = outGrad[newIdx[i],:] for points and features
extern "C" int mc_sort_points_step2_grad(int numPoints, int numFeatures, const float* outPtsGrads,
                                         const float* outFeatureGrads, const int* newIdx, float* inPtsGrads,
                                         float* inFeatureGrads, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (numPoints > 0) {
        gather_rows_kernel<<<cdiv((long long)numPoints * 3, 256), 256, 0, s>>>(numPoints, 3, newIdx, outPtsGrads,
                                                                             inPtsGrads);
        if (numFeatures > 0) {
            gather_rows_kernel<<<cdiv((long long)numPoints * numFeatures, 256), 256, 0, s>>>(
                numPoints, numFeatures, newIdx, outFeatureGrads, inFeatureGrads);
            count_launches(1);
        }
    }
    return check_launch("mc_sort_points_step2_grad");
}

// getFeaturesSampledPointsGradients, tf_ops/poisson_sampling.cu:260-274: zeros, then grad[idx[s],:] = outGrad[s,:]
extern "C" int mc_get_sampled_features_grad(int numPoints, int numFeatures, int numSamples, const int* idx,
                                            const float* outFeatureGrads, float* featureGrads, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (numPoints > 0 && numFeatures > 0)
        cudaMemsetAsync(featureGrads, 0, sizeof(float) * (size_t)numPoints * numFeatures, s);
    if (numSamples > 0 && numFeatures > 0)
        scatter_rows_kernel<<<cdiv((long long)numSamples * numFeatures, 256), 256, 0, s>>>(
            numSamples, numFeatures, idx, outFeatureGrads, featureGrads);
    return check_launch("mc_get_sampled_features_grad");
}
