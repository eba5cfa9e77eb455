// spatial_conv.cu -- Monte-Carlo convolution forward:
//     out[i,o] = sum_{p=(j,i)} sum_c  f[j,c] * kappa_{o*C+c}(delta_p) / (pdf_p * cnt_i),   delta_p = (x_j - s_i) / r_b
// with kappa the block-structured MLP 3 -> 8 -> 8 -> 8 per block of 8 output neurons.
//
// Replaces spatialConvCPU / evaluateMLPKernel / evaluateMLPNoCombinKernel (tf_ops/spatial_conv.cu:24-325,796-871).
// The reference gives every (pair, MLP block) an 8-thread group that re-reads the block's 176 parameters from
// global memory, exchanges the hidden activations through shared memory with two CTA-wide barriers, and issues one
// global float atomicAdd per (pair, neuron).  Here the neighbour list is consumed in its CSR order and the two 8x8
// layers AND the Monte-Carlo sum run on the tensor cores:
//
//   * a WARP owns a run of G <= 16 consecutive samples and a chunk of NB MLP blocks.  It walks the run's pair range
//     in tiles of 16 pairs (the M of an m16n8k8 MMA); pair records (delta, weight, local sample, neighbour) are
//     staged 32 at a time -- one pair per lane, coalesced, two loads deep in flight -- through a per-warp
//     shared-memory slab;
//   * layer 1 (K = 3) is FFMA and is computed directly in the A-fragment layout of layer 2;
//   * layer 2:  H2[16 pairs x 8] = H1[16 x 8] . W2^T          one MMA shape, the bias is the C operand;
//   * layer 3 is evaluated transposed, K3^T[8 neurons x 8 pairs] = Wo . H2^T, so that the C fragment of layer 2 IS the
//     B fragment of layer 3 (with the k index permuted identically on both operands) and the C fragment of layer 3
//     IS the B fragment of the reduction -- no shuffles between the layers;
//   * the Monte-Carlo sum  out[16 samples x 8 neurons] += S[16 x 8 pairs] . V[8 pairs x 8 neurons]  is an MMA with
//     the 0/1 sample-membership matrix S; the accumulator fragment lives in registers for the whole run and is
//     stored once.  NO atomics, and a summation order that is a pure function of the neighbour list.
//
// Precision: parity with the reference is 1e-5 relative, which single-pass TF32 (10-bit mantissa) cannot meet.  Every
// product is therefore the 3xTF32 split  a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo  (a_hi = tf32(a), a_lo = a - a_hi,
// FP32 accumulate): ~2^-21 relative per product.  For layer 3 the hi and lo halves of Wo are stacked in the 16 rows
// of one A operand, so two MMAs per 8 pairs give all four partial products.  S is exact in TF32.
// Per 16 pairs and block: 11 MMAs (3 + 4 + 4), or 8 in the single-input mode where the linear last layer is applied
// after the sum, and ~70-90 other instructions, against ~370 for the FFMA formulation.
// (tcgen05 was considered and rejected, DESIGN.md "spatial_conv": the operands are 8x8 per block and change layout
// between every layer through a ReLU, which would round-trip TMEM <-> registers <-> shared memory four times per
// tile; the legacy mma.sync path keeps every intermediate in the issuing warp's registers.)
#include "common.cuh"

#include <cstdlib>

namespace mc {
namespace {

constexpr int kConvWarps = 4;
constexpr int kRadiusTable = 1024;  // per-cloud 1/r kept in shared memory up to this batch size
constexpr int kMaxRun = 16;         // samples per warp: rows of the accumulator fragment

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

// D = A(16x8, row) . B(8x8, col) + C, TF32 inputs, FP32 accumulate.  lane = 4g + t:
//   A: a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4];  B: b0 = B[t][g], b1 = B[t+4][g];
//   C/D: c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1, float c0, float c1, float c2, float c3) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
__device__ __forceinline__ void mma_tf32_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// x = hi + lo (+ 2^-21 |x|): hi is x rounded to TF32 (nearest, ties away); lo is the exact remainder, of which the
// tensor core reads the upper 19 bits
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

struct PairLoads {  // second-stage loads of one pair, in flight while the previous 32 pairs are computed
    float px, py, pz, sx, sy, sz, pdf, f;
    int lo, hi, bid, i, j;
};

__device__ __forceinline__ float layer1(const float4& d, const float4& w) {
    return fmaxf(fmaf(d.x, w.x, fmaf(d.y, w.y, fmaf(d.z, w.z, w.w))), 0.f);
}

// MODE kGeneric / kDepthwise: every neuron has its own feature weight, all three layers are evaluated per pair.
// MODE kSingleIn (one input feature, the first convolution of every network): the weight f_p / (pdf_p cnt) does not
// depend on the neuron and the last layer is linear, so it is applied AFTER the Monte-Carlo sum:
//     out[i,n] = sum_u Wo[n,u] . Z[i,u] + bo[n] . F[i],   Z[i,u] = sum_p w_p H2[p,u],   F[i] = sum_p w_p
// which removes layer 3 from the per-pair work (8 instead of 11 MMAs per 16 pairs and block).
template <int MODE, int NB>
__global__ void __launch_bounds__(kConvWarps * 32) spatial_conv_kernel(const ConvArgs a) {
    __shared__ float sInvR[kRadiusTable];
    __shared__ __align__(16) float4 sDelta[kConvWarps][2][32];  // (dx, dy, dz, -) of the staged pairs
    __shared__ __align__(8) float sWgt[kConvWarps][2][32];      // 1 / (pdf * cnt)  [x feature, single-input mode]
    __shared__ __align__(8) int sLocal[kConvWarps][2][32];      // sample - first sample of the run; -1 = empty slot
    __shared__ __align__(8) int sNeigh[kConvWarps][2][32];      // neighbour point (row of the feature matrix)

    // the chunk's parameters in fragment order, pre-split for the 3xTF32 products: five 128-bit words per
    // (block, lane), lane-contiguous so that every read in the tile loop is one conflict-free LDS.128
    __shared__ __align__(16) float4 sFrag[NB][5][32];

    const bool radiusTable = a.B <= kRadiusTable;
    if (radiusTable) {
        for (int b = threadIdx.x; b < a.B; b += blockDim.x) {
            const float r = a.scaleInv ? a.radius * max_extent(a.mn, a.mx, b) : a.radius;
            sInvR[b] = 1.0f / r;
        }
    }
    const int blk0 = blockIdx.y * NB;
    // flat reference layouts: W1[(8q+u)*3+axis], b1[8q+u], W2[q*64 + out*8 + in], b2[8q+out], Wo[(8q+out)*8 + in],
    // bo[8q+out]
    for (int idx = threadIdx.x; idx < NB * 32; idx += blockDim.x) {
        const int q = idx >> 5, ln = idx & 31, fg = ln >> 2, ft = ln & 3;
        const int blk = blk0 + q;
        float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, f2 = f0, f3 = f0, f4 = f0;
        if (blk < a.numBlocks) {
            const int u0 = blk * 8 + ft, u1 = u0 + 4;
            // layer 1, hidden units t and t+4: (w_x, w_y, w_z, bias)
            f0 = make_float4(a.W1[u0 * 3], a.W1[u0 * 3 + 1], a.W1[u0 * 3 + 2], a.b1[u0]);
            f1 = make_float4(a.W1[u1 * 3], a.W1[u1 * 3 + 1], a.W1[u1 * 3 + 2], a.b1[u1]);
            uint32_t h0, l0, h1, l1;
            split_tf32(a.W2[blk * 64 + fg * 8 + ft], h0, l0);
            split_tf32(a.W2[blk * 64 + fg * 8 + ft + 4], h1, l1);
            if (MODE == kSingleIn)  // layer 2 as A fragment (transposed layer): rows 0-7 hi(W2), rows 8-15 lo(W2)
                f2 = make_float4(__uint_as_float(h0), __uint_as_float(l0), __uint_as_float(h1), __uint_as_float(l1));
            else  // layer 2 as B fragment: W2[out g][in t], W2[out g][in t+4] as (hi0, hi1, lo0, lo1)
                f2 = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
            // layer 3: Wo[out g][in 2t], Wo[out g][in 2t+1] as (hi0, lo0, hi1, lo1).  A fragment of the transposed
            // layer (rows 0-7 hi, rows 8-15 lo; k' = t <-> in 2t, k' = t+4 <-> in 2t+1: the order the layer-2 C
            // fragment holds its columns in), or B fragment of the post-sum projection of the single-input mode
            split_tf32(a.Wo[(blk * 8 + fg) * 8 + 2 * ft], h0, l0);
            split_tf32(a.Wo[(blk * 8 + fg) * 8 + 2 * ft + 1], h1, l1);
            f3 = make_float4(__uint_as_float(h0), __uint_as_float(l0), __uint_as_float(h1), __uint_as_float(l1));
            if (MODE == kSingleIn) {  // layer 2 bias of unit g | layer 3 bias of neurons 2t, 2t+1
                f4 = make_float4(a.b2[blk * 8 + fg], a.bo[blk * 8 + 2 * ft], a.bo[blk * 8 + 2 * ft + 1], 0.f);
            } else {  // layer 2 bias of units 2t, 2t+1 | layer 3 bias of neuron g | feature column of neuron g
                const int fc = (MODE == kGeneric) ? (blk * 8 + fg) % a.C : min(blk * 8 + fg, a.C - 1);
                f4 = make_float4(a.b2[blk * 8 + 2 * ft], a.b2[blk * 8 + 2 * ft + 1], a.bo[blk * 8 + fg],
                                 __int_as_float(fc));
            }
        }
        sFrag[q][0][ln] = f0;
        sFrag[q][1][ln] = f1;
        sFrag[q][2][ln] = f2;
        sFrag[q][3][ln] = f3;
        sFrag[q][4][ln] = f4;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int g = lane >> 2, t = lane & 3;

    const int M = live_count(a.dM, a.capM);
    const int K = live_count(a.dK, a.capK);
    const int s0 = (blockIdx.x * kConvWarps + warp) * a.G;  // first sample of this warp's run
    if (s0 >= a.capM) return;
    const int rowsOut = min(a.G, a.capM - s0);  // rows written: live samples and capacity padding (zeros)
    const int s1 = min(s0 + a.G, M);
    int P0 = 0, P1 = 0;
    if (s0 < M) {
        P0 = min(a.start[s0], K);
        P1 = min((s1 < M) ? a.start[s1] : K, K);
    }

    // accumulator fragments, rows = samples s0 + {g, g+8}: out[.][8 blk + {2t, 2t+1}], or (single-input mode)
    // Z[.][unit {2t, 2t+1}] and the weight sums F[.]
    float acc[NB][4];
    float accF[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < NB; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;

    // ---- pair staging, two loads deep: (A) pair + pdf of batch k+2, (B) geometry of batch k+1, compute batch k
    int2 prA = make_int2(0, -1);
    float pdfA = 1.0f;
    auto load_a = [&](int p) {
        const bool ok = p < P1;
        prA = ok ? a.pairs[p] : make_int2(0, -1);
        pdfA = ok ? a.pdfs[p] : 1.0f;
    };
    auto load_b = [&]() {
        PairLoads L;
        L.j = prA.x;
        L.i = prA.y;
        L.pdf = pdfA;
        const int j = L.j, i = max(L.i, 0);
        L.px = a.pts[(size_t)j * 3];
        L.py = a.pts[(size_t)j * 3 + 1];
        L.pz = a.pts[(size_t)j * 3 + 2];
        L.sx = a.samples[(size_t)i * 3];
        L.sy = a.samples[(size_t)i * 3 + 1];
        L.sz = a.samples[(size_t)i * 3 + 2];
        L.bid = a.bids[j];
        L.lo = a.start[i];
        L.hi = (i < M - 1) ? a.start[i + 1] : K;
        L.f = (MODE == kSingleIn) ? a.feats[j] : 1.0f;
        return L;
    };
    auto stage = [&](const PairLoads& L, int buf) {
        float invR;
        if (radiusTable) {
            invR = sInvR[L.bid];
        } else {
            const float ext = max_extent(a.mn, a.mx, L.bid);
            invR = 1.0f / (a.scaleInv ? a.radius * ext : a.radius);
        }
        const bool ok = L.i >= 0;
        const float cnt = a.avg ? (float)(min(L.hi, K) - L.lo) : 1.0f;
        const float invw = ok ? L.f / (L.pdf * cnt) : 0.0f;
        sDelta[warp][buf][lane] = ok ? make_float4((L.px - L.sx) * invR, (L.py - L.sy) * invR, (L.pz - L.sz) * invR, 0.f)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        sWgt[warp][buf][lane] = invw;
        sLocal[warp][buf][lane] = ok ? L.i - s0 : -1;
        if (MODE != kSingleIn) sNeigh[warp][buf][lane] = ok ? L.j : 0;
    };

    const int nBatches = (P1 - P0 + 31) >> 5;
    if (nBatches > 0) {
        load_a(P0 + lane);
        PairLoads L = load_b();
        load_a(P0 + 32 + lane);
        stage(L, 0);
        __syncwarp();
    }
    for (int k = 0; k < nBatches; ++k) {
        const int buf = k & 1;
        PairLoads L;
        const bool more = k + 1 < nBatches;
        if (more) {
            L = load_b();
            load_a(P0 + (k + 2) * 32 + lane);
        }
        const int left = P1 - (P0 + k * 32);  // pairs in this batch (may exceed 32)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h * 16 >= left) break;  // warp-uniform
            const int r0 = h * 16;
            const float4 da = sDelta[warp][buf][r0 + g], db = sDelta[warp][buf][r0 + g + 8];
            // pairs this lane holds as COLUMNS of the transposed fragments: rows 2t, 2t+1 (first 8) and 8+2t, 8+2t+1
            const float2 wA = *reinterpret_cast<const float2*>(&sWgt[warp][buf][r0 + 2 * t]);
            const float2 wB = *reinterpret_cast<const float2*>(&sWgt[warp][buf][r0 + 8 + 2 * t]);
            const int2 lA = *reinterpret_cast<const int2*>(&sLocal[warp][buf][r0 + 2 * t]);
            const int2 lB = *reinterpret_cast<const int2*>(&sLocal[warp][buf][r0 + 8 + 2 * t]);
            // membership matrix S[sample row][pair column] as A fragments (exact in TF32)
            const uint32_t one = 0x3f800000u;
            const uint32_t sA0 = lA.x == g ? one : 0u, sA1 = lA.x == g + 8 ? one : 0u;
            const uint32_t sA2 = lA.y == g ? one : 0u, sA3 = lA.y == g + 8 ? one : 0u;
            const uint32_t sB0 = lB.x == g ? one : 0u, sB1 = lB.x == g + 8 ? one : 0u;
            const uint32_t sB2 = lB.y == g ? one : 0u, sB3 = lB.y == g + 8 ? one : 0u;

            if constexpr (MODE == kSingleIn) {
                {  // F[i] += sum of the weights of sample i's pairs (every column of the fragment holds the same sum)
                    uint32_t fh[4], fl[4];
                    split_tf32(wA.x, fh[0], fl[0]);
                    split_tf32(wA.y, fh[1], fl[1]);
                    split_tf32(wB.x, fh[2], fl[2]);
                    split_tf32(wB.y, fh[3], fl[3]);
                    mma_tf32_acc(accF, sA0, sA1, sA2, sA3, fl[0], fl[1]);
                    mma_tf32_acc(accF, sB0, sB1, sB2, sB3, fl[2], fl[3]);
                    mma_tf32_acc(accF, sA0, sA1, sA2, sA3, fh[0], fh[1]);
                    mma_tf32_acc(accF, sB0, sB1, sB2, sB3, fh[2], fh[3]);
                }
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    // parameters are re-read from shared memory for every tile; the empty asm keeps the compiler from
                    // hoisting the loop-invariant loads into (far too many) registers
                    int fo = q * 5 * 32 + lane;
                    asm volatile("" : "+r"(fo));
                    const float4* frag = &sFrag[0][0][0] + fo;
                    const float4 wa = frag[0], wb = frag[32], p2 = frag[64];
                    const float b2g = frag[128].x;
                    const uint32_t a20 = __float_as_uint(p2.x), a21 = __float_as_uint(p2.y);
                    const uint32_t a22 = __float_as_uint(p2.z), a23 = __float_as_uint(p2.w);
                    // layer 1 as B fragments of the transposed layer 2: units t / t+4 of pair g (first / second 8)
                    uint32_t xh[4], xl[4];
                    split_tf32(layer1(da, wa), xh[0], xl[0]);
                    split_tf32(layer1(da, wb), xh[1], xl[1]);
                    split_tf32(layer1(db, wa), xh[2], xl[2]);
                    split_tf32(layer1(db, wb), xh[3], xl[3]);
                    // layer 2, transposed: rows = unit g (hi part of W2) / unit g (lo part), columns = 8 pairs
                    float d2a[4], d2b[4];
                    mma_tf32(d2a, a20, a21, a22, a23, xl[0], xl[1], b2g, b2g, 0.f, 0.f);
                    mma_tf32(d2b, a20, a21, a22, a23, xl[2], xl[3], b2g, b2g, 0.f, 0.f);
                    mma_tf32_acc(d2a, a20, a21, a22, a23, xh[0], xh[1]);
                    mma_tf32_acc(d2b, a20, a21, a22, a23, xh[2], xh[3]);
                    // w_p . H2[p][unit g] of pairs 2t, 2t+1 | 8+2t, 8+2t+1
                    uint32_t vh[4], vl[4];
                    split_tf32(fmaxf(d2a[0] + d2a[2], 0.f) * wA.x, vh[0], vl[0]);
                    split_tf32(fmaxf(d2a[1] + d2a[3], 0.f) * wA.y, vh[1], vl[1]);
                    split_tf32(fmaxf(d2b[0] + d2b[2], 0.f) * wB.x, vh[2], vl[2]);
                    split_tf32(fmaxf(d2b[1] + d2b[3], 0.f) * wB.y, vh[3], vl[3]);
                    // Z[i][u] += sum over the pairs of sample i
                    mma_tf32_acc(acc[q], sA0, sA1, sA2, sA3, vl[0], vl[1]);
                    mma_tf32_acc(acc[q], sB0, sB1, sB2, sB3, vl[2], vl[3]);
                    mma_tf32_acc(acc[q], sA0, sA1, sA2, sA3, vh[0], vh[1]);
                    mma_tf32_acc(acc[q], sB0, sB1, sB2, sB3, vh[2], vh[3]);
                }
            } else {
                const int2 jA = *reinterpret_cast<const int2*>(&sNeigh[warp][buf][r0 + 2 * t]);
                const int2 jB = *reinterpret_cast<const int2*>(&sNeigh[warp][buf][r0 + 8 + 2 * t]);
                float fA0[NB], fA1[NB], fB0[NB], fB1[NB];  // feature x weight of the four column pairs, per block
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int fcol = __float_as_int(sFrag[q][4][lane].w);
                    fA0[q] = a.feats[(size_t)jA.x * a.C + fcol] * wA.x;
                    fA1[q] = a.feats[(size_t)jA.y * a.C + fcol] * wA.y;
                    fB0[q] = a.feats[(size_t)jB.x * a.C + fcol] * wB.x;
                    fB1[q] = a.feats[(size_t)jB.y * a.C + fcol] * wB.y;
                }
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    int fo = q * 5 * 32 + lane;
                    asm volatile("" : "+r"(fo));
                    const float4* frag = &sFrag[0][0][0] + fo;
                    const float4 wa = frag[0], wb = frag[32], p2 = frag[64], p3 = frag[96], p4 = frag[128];
                    const uint32_t w2h0 = __float_as_uint(p2.x), w2h1 = __float_as_uint(p2.y);
                    const uint32_t w2l0 = __float_as_uint(p2.z), w2l1 = __float_as_uint(p2.w);
                    const uint32_t w30 = __float_as_uint(p3.x), w31 = __float_as_uint(p3.y);
                    const uint32_t w32 = __float_as_uint(p3.z), w33 = __float_as_uint(p3.w);
                    // layer 1 in the A-fragment layout of layer 2: hidden t / t+4 of pairs g / g+8
                    uint32_t h1h[4], h1l[4];
                    split_tf32(layer1(da, wa), h1h[0], h1l[0]);
                    split_tf32(layer1(db, wa), h1h[1], h1l[1]);
                    split_tf32(layer1(da, wb), h1h[2], h1l[2]);
                    split_tf32(layer1(db, wb), h1h[3], h1l[3]);
                    // layer 2: small terms first
                    float d2[4];
                    mma_tf32(d2, h1l[0], h1l[1], h1l[2], h1l[3], w2h0, w2h1, p4.x, p4.y, p4.x, p4.y);
                    mma_tf32_acc(d2, h1h[0], h1h[1], h1h[2], h1h[3], w2l0, w2l1);
                    mma_tf32_acc(d2, h1h[0], h1h[1], h1h[2], h1h[3], w2h0, w2h1);
                    uint32_t h2h[4], h2l[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_tf32(fmaxf(d2[e], 0.f), h2h[e], h2l[e]);
                    // layer 3, transposed: rows = neuron g (hi part of Wo) / neuron g (lo part), columns = 8 pairs
                    float d3a[4], d3b[4];
                    mma_tf32(d3a, w30, w31, w32, w33, h2l[0], h2l[1], p4.z, p4.z, 0.f, 0.f);
                    mma_tf32(d3b, w30, w31, w32, w33, h2l[2], h2l[3], p4.z, p4.z, 0.f, 0.f);
                    mma_tf32_acc(d3a, w30, w31, w32, w33, h2h[0], h2h[1]);
                    mma_tf32_acc(d3b, w30, w31, w32, w33, h2h[2], h2h[3]);
                    // kernel value x feature / (pdf * cnt): neuron g of pairs 2t, 2t+1 | 8+2t, 8+2t+1
                    uint32_t vh[4], vl[4];
                    split_tf32((d3a[0] + d3a[2]) * fA0[q], vh[0], vl[0]);
                    split_tf32((d3a[1] + d3a[3]) * fA1[q], vh[1], vl[1]);
                    split_tf32((d3b[0] + d3b[2]) * fB0[q], vh[2], vl[2]);
                    split_tf32((d3b[1] + d3b[3]) * fB1[q], vh[3], vl[3]);
                    // Monte-Carlo sum over the pairs of each sample of the run
                    mma_tf32_acc(acc[q], sA0, sA1, sA2, sA3, vl[0], vl[1]);
                    mma_tf32_acc(acc[q], sB0, sB1, sB2, sB3, vl[2], vl[3]);
                    mma_tf32_acc(acc[q], sA0, sA1, sA2, sA3, vh[0], vh[1]);
                    mma_tf32_acc(acc[q], sB0, sB1, sB2, sB3, vh[2], vh[3]);
                }
            }
        }
        if (more) stage(L, buf ^ 1);
        __syncwarp();
    }

    // ---- store: lane (g, t) holds samples s0+g, s0+g+8 x neurons 8 blk + {2t, 2t+1}.  Samples without neighbours
    // and the capacity-padding rows past the live count get their defined zeros here as well.
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int n = (blk0 + q) * 8 + 2 * t;
        if (blk0 + q >= a.numBlocks) break;
        float o4[4] = {acc[q][0], acc[q][1], acc[q][2], acc[q][3]};
        if constexpr (MODE == kSingleIn) {
            // out[i][n] = sum_u Z[i][u] Wo[n][u] + bo[n] F[i]: the accumulator fragment is the A fragment (its columns
            // 2t, 2t+1 are k' = t, t+4), the pre-split Wo slice the B fragment
            const float4 p3 = sFrag[q][3][lane], p4 = sFrag[q][4][lane];
            uint32_t zh[4], zl[4];
            split_tf32(acc[q][0], zh[0], zl[0]);
            split_tf32(acc[q][2], zh[1], zl[1]);
            split_tf32(acc[q][1], zh[2], zl[2]);
            split_tf32(acc[q][3], zh[3], zl[3]);
            mma_tf32(o4, zl[0], zl[1], zl[2], zl[3], __float_as_uint(p3.x), __float_as_uint(p3.z), p4.y * accF[0],
                     p4.z * accF[0], p4.y * accF[2], p4.z * accF[2]);
            mma_tf32_acc(o4, zh[0], zh[1], zh[2], zh[3], __float_as_uint(p3.y), __float_as_uint(p3.w));
            mma_tf32_acc(o4, zh[0], zh[1], zh[2], zh[3], __float_as_uint(p3.x), __float_as_uint(p3.z));
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = g + 8 * half;
            if (row < rowsOut) {
                float* o = a.out + (size_t)(s0 + row) * a.outStride + n;
                if (n < a.nOut) o[0] = o4[2 * half];
                if (n + 1 < a.nOut) o[1] = o4[2 * half + 1];
            }
        }
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

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

constexpr int kChunkBlocks = 4;  // MLP blocks per warp (fewer when the convolution has fewer)

// blocks past the end of the parameter arrays are evaluated with zero parameters (no branch in the tile loop, so the
// compiler interleaves the blocks' MMA chains); the chunk width is matched to small convolutions instead
template <int MODE>
void launch_conv(const ConvArgs& a, int nb, int G, cudaStream_t s) {
    const int runs = cdiv(a.capM, G);
    static const int forcedNB = env_int("MCCNN_CONV_NB", 0);  // tuning knob: MLP blocks per warp (2 or 4), 0 = default
    if (nb >= 3 && forcedNB == 2) {
        dim3 grid(cdiv(runs, kConvWarps), cdiv(nb, 2));
        spatial_conv_kernel<MODE, 2><<<grid, kConvWarps * 32, 0, s>>>(a);
    } else if (nb >= 3) {
        dim3 grid(cdiv(runs, kConvWarps), cdiv(nb, 4));
        spatial_conv_kernel<MODE, 4><<<grid, kConvWarps * 32, 0, s>>>(a);
    } else if (nb == 2) {
        dim3 grid(cdiv(runs, kConvWarps), 1);
        spatial_conv_kernel<MODE, 2><<<grid, kConvWarps * 32, 0, s>>>(a);
    } else {
        dim3 grid(cdiv(runs, kConvWarps), 1);
        spatial_conv_kernel<MODE, 1><<<grid, kConvWarps * 32, 0, s>>>(a);
    }
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
    // run length per warp: ~8 tiles of 16 pairs (the partially filled last tile stays cheap), at most the 16 rows of
    // the accumulator fragment -- but the op is latency bound, so never at the price of leaving the GPU short of warps
    static const int forcedG = env_int("MCCNN_CONV_RUN", 0);  // tuning knob (samples per warp), 0 = heuristic
    long long avgK = numNeighbors / (long long)numSamples;
    int G = (int)(128 / (avgK > 0 ? avgK : 1));
    G = G < 1 ? 1 : (G > kMaxRun ? kMaxRun : G);
    while (G > 1 && (long long)cdiv(numSamples, G) * chunks < 1776) --G;
    if (forcedG > 0) G = forcedG > kMaxRun ? kMaxRun : forcedG;
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
    if (!combin) {
        a.nOut = nOut;
        a.outStride = C;
        a.out = outFeatures;
        launch_conv<kDepthwise>(a, nb, G, s);
    } else if (C == 1) {
        a.nOut = nOut;
        a.outStride = numOutFeatures;
        a.out = outFeatures;
        launch_conv<kSingleIn>(a, nb, G, s);
    } else {
        MC_REQUIRE(workspace != nullptr, "mc_spatial_conv: workspace required for multi-feature convolutions");
        a.nOut = nb * 8;
        a.outStride = nb * 8;
        a.out = workspace;
        launch_conv<kGeneric>(a, nb, G, s);
        long long total = (long long)numSamples * numOutFeatures;
        combine_kernel<<<cdiv(total, 256), 256, 0, s>>>(dNumSamples, numSamples, C, numOutFeatures, nb * 8, workspace,
                                                        outFeatures);
        count_launches(1);
    }
    return check_launch("mc_spatial_conv");
}
