// spatial_conv.cu -- Monte-Carlo convolution forward:
//     out[i,o] = sum_{p=(j,i)} sum_c  f[j,c] * kappa_{o*C+c}(delta_p) / (pdf_p * cnt_i),   delta_p = (x_j - s_i) / r_b
// with kappa the block-structured MLP 3 -> 8 -> 8 -> 8 per block of 8 output neurons.
//
// Replaces spatialConvCPU / evaluateMLPKernel / evaluateMLPNoCombinKernel (tf_ops/spatial_conv.cu:24-325,796-871).
// The reference gives every (pair, MLP block) an 8-thread group that re-reads the block's 176 parameters from
// global memory, exchanges the hidden activations through shared memory with two CTA-wide barriers, and issues one
// global float atomicAdd per (pair, neuron).  Here the neighbour list is consumed in its CSR order, all three MLP
// layers AND the Monte-Carlo sum run on the tensor cores, and nothing is accumulated with atomics.
//
// Structure (persistent, warp-specialised CTAs of 5 warps; one CTA = a contiguous range of sample RUNS x a chunk of
// 4 x NB MLP blocks; a run = G <= 16 consecutive samples = the rows of one accumulator fragment):
//   * 2 PRODUCER warps (alternate batches): stream the pair list of the CTA's runs in batches of <= 32 pairs, two
//     global loads deep (pair + pdf of the batch after next, geometry of the next one in flight while a batch is
//     written): delta, 1 / (pdf * cnt), the sample's row in the run; for depth-wise convolutions the 32 NB .. 128
//     features of every neighbour's row which this CTA's blocks need are staged with 16-byte asynchronous copies
//     (cp.async, a whole row piece per warp instruction: coalesced 128-bit reads), their completion tied to the
//     stage's mbarrier.  Stages form a 4-deep full / empty mbarrier ring.  (One cp.async.bulk per row -- TMA -- was
//     built and measured first: a per-lane bulk copy is a uniform-datapath instruction, the warp serialises 32 of
//     them per batch and the producer became the bottleneck: profiles/r2_spatial_conv_experiments.md.)
//   * 4 CONSUMER warps, NB blocks each: per tile of 16 pairs
//       layer 1   [16 pairs x 4] . [4 x 8]  (delta | 1) x (W1 | b1): m16n8k4; the hidden units are PERMUTED over the
//                 output columns (column 2t <- unit t, 2t+1 <- unit t+4), so that the C fragment IS the A fragment of
//       layer 2   H2[16 x 8] = H1[16 x 8] . W2^T, m16n8k8, bias as the C operand;
//       layer 3   transposed, K3^T[8 neurons x 8 pairs] = Wo . H2^T: the C fragment of layer 2 IS its B fragment (k
//                 permuted identically on both operands), and its C fragment IS the B fragment of
//       the sum   out[16 samples x 8 neurons] += S[16 x 8 pairs] . V[8 pairs x 8 neurons] with the 0/1 membership
//                 matrix S; the accumulator fragment lives in registers for the whole run and is stored once.
//     The single-input mode (C = 1, the first convolution of every network) applies the linear last layer after the sum.
// Precision: parity is 1e-5 relative, single-pass TF32 cannot meet it: every product is the 3xTF32 split
//     a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo,  FP32 accumulate, ~2^-20 relative per product.
// The tensor core TRUNCATES an fp32 operand to TF32 (measured, tools/probes/mma_probe.cu), so "a_hi" is the raw
// register and only a_lo = a - trunc(a) is computed (2 instructions per value; parameters are split once per CTA).
// Per 16 pairs and block: 14 MMAs and ~55 other instructions (the round-1 kernel: 11 + ~145, issue bound at 33 % tensor
// pipe, profiles/r2_ncu_full_spatial_conv_c4_before.txt).
// (tcgen05 does not fit this op: the operands are 8 x 8 per block and change layout between the layers through a
// ReLU; a tcgen05.mma tile is >= 64 x 8 x 8 with operands in shared memory and the result in TMEM.  The dense layers
// BETWEEN the convolutions are plain GEMMs and do run on tcgen05: dense_tc.cu.)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mc {
namespace {

constexpr int kProducers = 2;
constexpr int kStages = 4;
constexpr int kRadiusTable = 1024;  // per-cloud 1/r kept in shared memory up to this batch size
constexpr int kMaxRun = 16;         // samples per run: rows of the accumulator fragment
constexpr int kMaxRunsPerCta = 2040;
constexpr int kGenericStageC = 64;  // multi-feature convolutions stage up to this many input features per pair

enum ConvMode { kGeneric = 0, kDepthwise = 1, kSingleIn = 2 };

struct ConvArgs {
    bool avg, scaleInv;
    float radius;
    const int* dM;
    int capM;
    const int* dK;
    int capK;
    int B, C, nOut, numBlocks, outStride, G, runsPerCta;
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

// ---- mbarrier / bulk copy -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(sm_u32(bar)), "r"(parity)
            : "memory");
        if (++spins > (1u << 26)) __trap();  // a wait this long is a protocol bug: fail loudly instead of hanging
    }
}
// 16-byte asynchronous global -> shared copy (LDGSTS); srcBytes = 0 writes zeros without reading
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t srcBytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(srcBytes) : "memory");
}
// this thread's arrival on `bar`, delivered once all its earlier cp.async copies have landed
__device__ __forceinline__ void bar_arrive_after_copies(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sm_u32(bar)) : "memory");
}

// ---- tensor-core primitives -----------------------------------------------------------------------------------------
// D = A(16x8, row) . B(8x8, col) + C, TF32 inputs, FP32 accumulate.  lane = 4g + t:
//   A: a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4];  B: b0 = B[t][g], b1 = B[t+4][g];
//   C/D: c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]
__device__ __forceinline__ void mma8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                     uint32_t b1, float c0, float c1, float c2, float c3) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
__device__ __forceinline__ void mma8_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// m16n8k4: D = A(16x4) . B(4x8) + C with a0 = A[g][t], a1 = A[g+8][t]; b0 = B[t][g]; C/D as above.
// the same instructions as volatile asm: issued in program order (the phase-interleaved tile loop relies on it)
__device__ __forceinline__ void mma8v(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                      uint32_t b1, float c0, float c1, float c2, float c3) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
__device__ __forceinline__ void mma8v_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma4v(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
__device__ __forceinline__ void mma4v_acc(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
// word `w` (0..3) of block q's parameter fragments; the volatile offset pins the load to its phase
__device__ __forceinline__ float4 frag_ld(const float4* myFrag, int q, int w) {
    int fo = (q * 4 + w) * 32;
    asm volatile("" : "+r"(fo));
    return myFrag[fo];
}
__device__ __forceinline__ float2 frag_ld2(const float4* myFrag, int q, int w) {
    int fo = (q * 4 + w) * 32;
    asm volatile("" : "+r"(fo));
    return *reinterpret_cast<const float2*>(myFrag + fo);
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// x = hi + lo with hi = x truncated to TF32 -- what the tensor core reads from the raw register -- and lo the exact
// remainder (of which the tensor core again reads the upper 19 bits): only lo has to be computed
__device__ __forceinline__ uint32_t lo_tf32(float x) {
    return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}
__device__ __forceinline__ uint32_t raw(float x) { return __float_as_uint(x); }

// One tile of 16 pairs x NB blocks.  myFrag: this lane's parameter fragments (four words per block); da / db: component t
// of delta of pairs g / g+8 (component 3 = 1); wA / wB, lA / lB: weights and run-local sample rows of the pairs this
// lane holds as columns (2t, 2t+1 | 8+2t, 8+2t+1); f: feature x weight of those pairs for neuron g of every block,
// bo: the block's layer-3 bias of neuron g (both unused in the single-input mode).
//
// The blocks of a tile are independent chains of dependent MMAs (latency ~20 cycles each).  The phases below issue
// the same step of ALL NB blocks back to back (volatile asm keeps that order), so that a single warp keeps its
// tensor pipe busy.  Parameters are re-read from shared memory per phase (asm-volatile offsets keep the compiler from
// hoisting the loop-invariant loads into far too many registers).
template <int MODE, int NB>
__device__ __forceinline__ void tile_body(const float4* myFrag, float da, float db, float2 wA, float2 wB, int2 lA, int2 lB,
                                          int g, float (&f)[NB][4], const float (&bo)[NB], float (&acc)[NB][4],
                                          float (&accF)[4]) {
    const uint32_t dal = lo_tf32(da), dbl = lo_tf32(db);
    // membership matrix S[sample row][pair column] as A fragments (exact in TF32)
    const uint32_t one = 0x3f800000u;
    const uint32_t sA0 = lA.x == g ? one : 0u, sA1 = lA.x == g + 8 ? one : 0u;
    const uint32_t sA2 = lA.y == g ? one : 0u, sA3 = lA.y == g + 8 ? one : 0u;
    const uint32_t sB0 = lB.x == g ? one : 0u, sB1 = lB.x == g + 8 ? one : 0u;
    const uint32_t sB2 = lB.y == g ? one : 0u, sB3 = lB.y == g + 8 ? one : 0u;
    float4 p0[NB];  // layer 1 (hi, lo) | layer-2 bias
#pragma unroll
    for (int q = 0; q < NB; ++q) p0[q] = frag_ld(myFrag, q, 0);
    float d1[NB][4];
#pragma unroll
    for (int q = 0; q < NB; ++q) mma4v(d1[q], dal, dbl, raw(p0[q].x));
#pragma unroll
    for (int q = 0; q < NB; ++q) mma4v_acc(d1[q], raw(da), raw(db), raw(p0[q].y));
#pragma unroll
    for (int q = 0; q < NB; ++q) mma4v_acc(d1[q], raw(da), raw(db), raw(p0[q].x));
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) d1[q][e] = fmaxf(d1[q][e], 0.f);

    if constexpr (MODE == kSingleIn) {
        // F[i] += sum of the weights of sample i's pairs (every column of the fragment holds the same sum)
        mma8v_acc(accF, sA0, sA1, sA2, sA3, lo_tf32(wA.x), lo_tf32(wA.y));
        mma8v_acc(accF, sB0, sB1, sB2, sB3, lo_tf32(wB.x), lo_tf32(wB.y));
        mma8v_acc(accF, sA0, sA1, sA2, sA3, raw(wA.x), raw(wA.y));
        mma8v_acc(accF, sB0, sB1, sB2, sB3, raw(wB.x), raw(wB.y));
        // layer 2, transposed: rows = unit g (hi part of W2) / unit g (lo part), columns = 8 pairs; the layer-1
        // C fragment is its B fragment (c0, c1 = units t, t+4 of pair g; c2, c3 of pair g+8)
        float4 p2[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) p2[q] = frag_ld(myFrag, q, 1);
        float d2a[NB][4], d2b[NB][4];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            mma8v(d2a[q], raw(p2[q].x), raw(p2[q].y), raw(p2[q].z), raw(p2[q].w), lo_tf32(d1[q][0]),
                  lo_tf32(d1[q][1]), p0[q].z, p0[q].z, 0.f, 0.f);
            mma8v(d2b[q], raw(p2[q].x), raw(p2[q].y), raw(p2[q].z), raw(p2[q].w), lo_tf32(d1[q][2]),
                  lo_tf32(d1[q][3]), p0[q].z, p0[q].z, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            mma8v_acc(d2a[q], raw(p2[q].x), raw(p2[q].y), raw(p2[q].z), raw(p2[q].w), raw(d1[q][0]), raw(d1[q][1]));
            mma8v_acc(d2b[q], raw(p2[q].x), raw(p2[q].y), raw(p2[q].z), raw(p2[q].w), raw(d1[q][2]), raw(d1[q][3]));
        }
        // w_p . H2[p][unit g] of pairs 2t, 2t+1 | 8+2t, 8+2t+1
        float v[NB][4];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            v[q][0] = fmaxf(d2a[q][0] + d2a[q][2], 0.f) * wA.x;
            v[q][1] = fmaxf(d2a[q][1] + d2a[q][3], 0.f) * wA.y;
            v[q][2] = fmaxf(d2b[q][0] + d2b[q][2], 0.f) * wB.x;
            v[q][3] = fmaxf(d2b[q][1] + d2b[q][3], 0.f) * wB.y;
        }
        // Z[i][u] += sum over the pairs of sample i
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sA0, sA1, sA2, sA3, lo_tf32(v[q][0]), lo_tf32(v[q][1]));
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sB0, sB1, sB2, sB3, lo_tf32(v[q][2]), lo_tf32(v[q][3]));
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sA0, sA1, sA2, sA3, raw(v[q][0]), raw(v[q][1]));
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sB0, sB1, sB2, sB3, raw(v[q][2]), raw(v[q][3]));
    } else {
        // layer 2 (small terms first): the layer-1 C fragment is its A fragment -- a0 = [g][t] = c0,
        // a1 = [g+8][t] = c2, a2 = [g][t+4] = c1, a3 = [g+8][t+4] = c3
        float4 pw[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) pw[q] = frag_ld(myFrag, q, 1);
        float d2[NB][4];
#pragma unroll
        for (int q = 0; q < NB; ++q)
            mma8v(d2[q], lo_tf32(d1[q][0]), lo_tf32(d1[q][2]), lo_tf32(d1[q][1]), lo_tf32(d1[q][3]), raw(pw[q].x),
                  raw(pw[q].y), p0[q].z, p0[q].w, p0[q].z, p0[q].w);
#pragma unroll
        for (int q = 0; q < NB; ++q)
            mma8v_acc(d2[q], raw(d1[q][0]), raw(d1[q][2]), raw(d1[q][1]), raw(d1[q][3]), raw(pw[q].z), raw(pw[q].w));
#pragma unroll
        for (int q = 0; q < NB; ++q)
            mma8v_acc(d2[q], raw(d1[q][0]), raw(d1[q][2]), raw(d1[q][1]), raw(d1[q][3]), raw(pw[q].x), raw(pw[q].y));
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) d2[q][e] = fmaxf(d2[q][e], 0.f);
        // layer 3, transposed: rows = neuron g (hi part of Wo) / neuron g (lo part), columns = 8 pairs; the
        // layer-2 C fragment is its B fragment (c0, c1: pair g, units 2t, 2t+1; c2, c3: pair g+8)
#pragma unroll
        for (int q = 0; q < NB; ++q) pw[q] = frag_ld(myFrag, q, 2);
        float d3a[NB][4], d3b[NB][4];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            mma8v(d3a[q], raw(pw[q].x), raw(pw[q].y), raw(pw[q].z), raw(pw[q].w), lo_tf32(d2[q][0]),
                  lo_tf32(d2[q][1]), bo[q], bo[q], 0.f, 0.f);
            mma8v(d3b[q], raw(pw[q].x), raw(pw[q].y), raw(pw[q].z), raw(pw[q].w), lo_tf32(d2[q][2]),
                  lo_tf32(d2[q][3]), bo[q], bo[q], 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            mma8v_acc(d3a[q], raw(pw[q].x), raw(pw[q].y), raw(pw[q].z), raw(pw[q].w), raw(d2[q][0]), raw(d2[q][1]));
            mma8v_acc(d3b[q], raw(pw[q].x), raw(pw[q].y), raw(pw[q].z), raw(pw[q].w), raw(d2[q][2]), raw(d2[q][3]));
        }
        // kernel value x feature / (pdf * cnt): neuron g of pairs 2t, 2t+1 | 8+2t, 8+2t+1
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            f[q][0] *= d3a[q][0] + d3a[q][2];
            f[q][1] *= d3a[q][1] + d3a[q][3];
            f[q][2] *= d3b[q][0] + d3b[q][2];
            f[q][3] *= d3b[q][1] + d3b[q][3];
        }
        // Monte-Carlo sum over the pairs of each sample of the run
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sA0, sA1, sA2, sA3, lo_tf32(f[q][0]), lo_tf32(f[q][1]));
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sB0, sB1, sB2, sB3, lo_tf32(f[q][2]), lo_tf32(f[q][3]));
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sA0, sA1, sA2, sA3, raw(f[q][0]), raw(f[q][1]));
#pragma unroll
        for (int q = 0; q < NB; ++q) mma8v_acc(acc[q], sB0, sB1, sB2, sB3, raw(f[q][2]), raw(f[q][3]));
    }
}

// The parameters of MLP block `blk` in fragment order for lane `ln`: four 128-bit words, stored lane-contiguous
// (dst[0], dst[32], dst[64], dst[96]) so that every read in the tile loop is one conflict-free LDS.128.  Flat reference
// layouts: W1[(8q+u)*3+axis], b1[8q+u], W2[q*64 + out*8 + in], b2[8q+out], Wo[(8q+out)*8 + in], bo[8q+out].
// col0: first feature column of the staged row pieces (0 when the lane addresses whole feature rows).
template <int MODE>
__device__ __forceinline__ void build_fragment(const ConvArgs& a, int blk, int ln, int col0, float4* dst) {
    const int fg = ln >> 2, ft = ln & 3;
    // word 0: layer 1 (hi, lo) | layer-2 bias; word 1: layer 2; word 2: layer 3; word 3: layer-3 bias | feature column
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, f2 = f0, f3 = f0;
    if (blk < a.numBlocks) {
        // layer 1 as B fragment of the k = 4 MMA: row k = ft (x, y, z, bias), column fg <- hidden unit pi(fg)
        const int u = blk * 8 + ((fg & 1) ? (fg >> 1) + 4 : (fg >> 1));
        const float w1 = ft < 3 ? a.W1[u * 3 + ft] : a.b1[u];
        f0.x = w1;
        f0.y = __uint_as_float(lo_tf32(w1));
        const float w2a = a.W2[blk * 64 + fg * 8 + ft], w2b = a.W2[blk * 64 + fg * 8 + ft + 4];
        // layer 2: B fragment W2[out g][in t], [in t+4] (hi, hi, lo, lo); single-input mode: A fragment of the
        // transposed layer, rows 0-7 hi(W2[unit g]), rows 8-15 lo (hi, lo, hi, lo)
        f1 = (MODE == kSingleIn)
                 ? make_float4(w2a, __uint_as_float(lo_tf32(w2a)), w2b, __uint_as_float(lo_tf32(w2b)))
                 : make_float4(w2a, w2b, __uint_as_float(lo_tf32(w2a)), __uint_as_float(lo_tf32(w2b)));
        // layer 3: Wo[out g][in 2t], [in 2t+1] as (hi, lo, hi, lo).  A fragment of the transposed layer (rows 0-7
        // hi, rows 8-15 lo; k' = t <-> in 2t, k' = t+4 <-> in 2t+1: the order the layer-2 C fragment holds its
        // columns in), or B fragment of the post-sum projection of the single-input mode
        const float woa = a.Wo[(blk * 8 + fg) * 8 + 2 * ft], wob = a.Wo[(blk * 8 + fg) * 8 + 2 * ft + 1];
        f2 = make_float4(woa, __uint_as_float(lo_tf32(woa)), wob, __uint_as_float(lo_tf32(wob)));
        if (MODE == kSingleIn) {  // layer 2 bias of unit g | layer 3 bias of neurons 2t, 2t+1
            f0.z = a.b2[blk * 8 + fg];
            f3 = make_float4(a.bo[blk * 8 + 2 * ft], a.bo[blk * 8 + 2 * ft + 1], 0.f, 0.f);
        } else {  // layer 2 bias of units 2t, 2t+1 | layer 3 bias of neuron g | feature column of neuron g
            const int fc = (MODE == kGeneric) ? (blk * 8 + fg) % a.C : min(blk * 8 + fg, a.C - 1);
            // staged rows hold the chunk's columns only (depth-wise) or all C input features (multi-feature)
            const int col = fc - col0;
            f0.z = a.b2[blk * 8 + 2 * ft];
            f0.w = a.b2[blk * 8 + 2 * ft + 1];
            f3 = make_float4(a.bo[blk * 8 + fg], __int_as_float(col), 0.f, 0.f);
        }
    }
    dst[0] = f0;
    dst[32] = f1;
    dst[64] = f2;
    dst[96] = f3;
}

// Store of a run's accumulators: lane (g, t) holds samples s0+g, s0+g+8 x neurons 8 blk + {2t, 2t+1}.  Samples without
// neighbours and the capacity-padding rows past the live count get their defined zeros here as well.
template <int MODE, int NB>
__device__ __forceinline__ void store_run(const ConvArgs& a, const float4* myFrag, float (&acc)[NB][4], float (&accF)[4],
                                          int blk0, int s0, int rowsOut, int g, int t) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int n = (blk0 + q) * 8 + 2 * t;
        if (blk0 + q >= a.numBlocks) break;
        float o4[4] = {acc[q][0], acc[q][1], acc[q][2], acc[q][3]};
        if constexpr (MODE == kSingleIn) {
            // out[i][n] = sum_u Z[i][u] Wo[n][u] + bo[n] F[i]: the accumulator fragment is the A fragment (its
            // columns 2t, 2t+1 are k' = t, t+4), the Wo slice the B fragment
            const float4 p3 = myFrag[q * 4 * 32 + 64], p4 = myFrag[q * 4 * 32 + 96];
            mma8(o4, lo_tf32(acc[q][0]), lo_tf32(acc[q][2]), lo_tf32(acc[q][1]), lo_tf32(acc[q][3]), raw(p3.x),
                 raw(p3.z), p4.x * accF[0], p4.y * accF[0], p4.x * accF[2], p4.y * accF[2]);
            mma8_acc(o4, raw(acc[q][0]), raw(acc[q][2]), raw(acc[q][1]), raw(acc[q][3]), raw(p3.y), raw(p3.w));
            mma8_acc(o4, raw(acc[q][0]), raw(acc[q][2]), raw(acc[q][1]), raw(acc[q][3]), raw(p3.x), raw(p3.z));
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = g + 8 * half;
            if (row < rowsOut) {
                float* o = a.out + (size_t)(s0 + row) * a.outStride + n;
                if (n + 1 < a.nOut && (a.outStride & 1) == 0) {
                    *reinterpret_cast<float2*>(o) = make_float2(o4[2 * half], o4[2 * half + 1]);
                } else {
                    if (n < a.nOut) o[0] = o4[2 * half];
                    if (n + 1 < a.nOut) o[1] = o4[2 * half + 1];
                }
            }
        }
    }
}

// ---- stage layout -----------------------------------------------------------------------------------------------------
struct StageHead {
    float d[4][40];   // delta, component-major: d[a][pair] (row 3 = 1: the bias column of layer 1); 40: lanes
                      // (g, t) read d[t][g]: banks 8t + g, conflict-free
    float w[32];      // 1 / (pdf * cnt)   [x feature, single-input mode]
    int local[32];    // sample - first sample of the run; -1 = empty slot
    int neigh[32];    // neighbour point (row of the feature matrix): multi-feature convolutions that do not stage
    int s0, rows, flags, n;  // first sample of the run, rows to store, kFirst | kLast | kStop, pairs in the batch
};
constexpr int kFirst = 1, kLast = 2, kStop = 4, kSilent = 8;

struct Batch {
    int run, p0, n, flags, seq;
};

struct PairLoads {  // second-stage loads of one pair, in flight while the previous batch is written
    float px, py, pz, sx, sy, sz, pdf, f;
    int lo, hi, bid, i, j;
};

// ---- warp-level variant: small and medium convolutions -----------------------------------------------------------
// A WARP owns a run of G <= 16 samples and NB MLP blocks, stages its own pair records (32 at a time, two global loads
// deep, per-warp shared-memory slab) and reads the feature values straight from global memory / L2.  No CTA-level
// pipeline to fill: more independent warps for the latency-bound convolutions of the classification networks and the
// upper hierarchy levels; the persistent kernel below amortises the staging for the large ones.
constexpr int kConvWarps = 4;

template <int MODE, int NB>
__global__ void __launch_bounds__(kConvWarps * 32) spatial_conv_warp_kernel(const ConvArgs a) {
    __shared__ float sInvR[kRadiusTable];
    __shared__ __align__(16) float4 sDelta[kConvWarps][2][32];  // (dx, dy, dz, 1) of the staged pairs
    __shared__ __align__(8) float sWgt[kConvWarps][2][32];      // 1 / (pdf * cnt)  [x feature, single-input mode]
    __shared__ __align__(8) int sLocal[kConvWarps][2][32];      // sample - first sample of the run; -1 = empty slot
    __shared__ __align__(8) int sNeigh[kConvWarps][2][32];      // neighbour point (row of the feature matrix)
    __shared__ __align__(16) float4 sFrag[NB][4][32];

    const bool radiusTable = a.B <= kRadiusTable;
    if (radiusTable) {
        for (int b = threadIdx.x; b < a.B; b += blockDim.x) {
            const float r = a.scaleInv ? a.radius * max_extent(a.mn, a.mx, b) : a.radius;
            sInvR[b] = 1.0f / r;
        }
    }
    const int blk0 = blockIdx.y * NB;
    for (int idx = threadIdx.x; idx < NB * 32; idx += blockDim.x)
        build_fragment<MODE>(a, blk0 + (idx >> 5), idx & 31, 0, &sFrag[idx >> 5][0][idx & 31]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int g = lane >> 2, t = lane & 3;
    const float4* myFrag = &sFrag[0][0][lane];

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
        L.px = L.py = L.pz = L.sx = L.sy = L.sz = 0.f;
        L.bid = L.lo = L.hi = 0;
        L.f = 1.0f;
        if (L.i < 0) return L;
        const int j = L.j, i = L.i;
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
        const bool ok = L.i >= 0;
        float invR = 0.f;
        if (ok) {
            if (radiusTable) {
                invR = sInvR[L.bid];
            } else {
                const float ext = max_extent(a.mn, a.mx, L.bid);
                invR = 1.0f / (a.scaleInv ? a.radius * ext : a.radius);
            }
        }
        const float cnt = a.avg ? (float)(min(L.hi, K) - L.lo) : 1.0f;
        sDelta[warp][buf][lane] = ok ? make_float4((L.px - L.sx) * invR, (L.py - L.sy) * invR, (L.pz - L.sz) * invR, 1.f)
                                     : make_float4(0.f, 0.f, 0.f, 1.f);
        sWgt[warp][buf][lane] = ok ? L.f / (L.pdf * cnt) : 0.0f;
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
            // component t of delta of pairs g / g+8 (a float4 per pair: banks 4g + t, conflict-free)
            const float da = reinterpret_cast<const float*>(&sDelta[warp][buf][r0 + g])[t];
            const float db = reinterpret_cast<const float*>(&sDelta[warp][buf][r0 + g + 8])[t];
            const float2 wA = *reinterpret_cast<const float2*>(&sWgt[warp][buf][r0 + 2 * t]);
            const float2 wB = *reinterpret_cast<const float2*>(&sWgt[warp][buf][r0 + 8 + 2 * t]);
            const int2 lA = *reinterpret_cast<const int2*>(&sLocal[warp][buf][r0 + 2 * t]);
            const int2 lB = *reinterpret_cast<const int2*>(&sLocal[warp][buf][r0 + 8 + 2 * t]);
            float f[NB][4];
            float bo[NB];
            if constexpr (MODE != kSingleIn) {
                const int2 jA = *reinterpret_cast<const int2*>(&sNeigh[warp][buf][r0 + 2 * t]);
                const int2 jB = *reinterpret_cast<const int2*>(&sNeigh[warp][buf][r0 + 8 + 2 * t]);
                const float* fr0 = a.feats + (size_t)jA.x * a.C;
                const float* fr1 = a.feats + (size_t)jA.y * a.C;
                const float* fr2 = a.feats + (size_t)jB.x * a.C;
                const float* fr3 = a.feats + (size_t)jB.y * a.C;
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const float2 p3 = frag_ld2(myFrag, q, 3);
                    bo[q] = p3.x;
                    const int fcol = __float_as_int(p3.y);
                    f[q][0] = fr0[fcol] * wA.x;
                    f[q][1] = fr1[fcol] * wA.y;
                    f[q][2] = fr2[fcol] * wB.x;
                    f[q][3] = fr3[fcol] * wB.y;
                }
            }
            tile_body<MODE, NB>(myFrag, da, db, wA, wB, lA, lB, g, f, bo, acc, accF);
        }
        if (more) stage(L, buf ^ 1);
        __syncwarp();
    }
    store_run<MODE, NB>(a, myFrag, acc, accF, blk0, s0, rowsOut, g, t);
}

// MODE kGeneric / kDepthwise: every neuron has its own feature weight, all three layers are evaluated per pair.
// MODE kSingleIn (one input feature): the weight f_p / (pdf_p cnt) does not depend on the neuron and the last layer is
// linear, so it is applied AFTER the Monte-Carlo sum:
//     out[i,n] = sum_u Wo[n,u] . Z[i,u] + bo[n] . F[i],   Z[i,u] = sum_p w_p H2[p,u],   F[i] = sum_p w_p
// FW: floats per staged feature row (0: features are not staged -- single-input mode, or a multi-feature convolution
// with more than kGenericStageC input features, which reads them from global memory).
template <int MODE, int NB, int NC>
__global__ void __launch_bounds__((kProducers + NC) * 32, 2) spatial_conv_kernel(const ConvArgs a, const int FW) {
    constexpr int kConsumers = NC;
    extern __shared__ __align__(16) uint8_t smem[];
    // [ parameter fragments | stage heads | stage feature rows | run bounds | 1 / r table | barriers ]
    float4* sFrag = reinterpret_cast<float4*>(smem);                                   // [4 NB][4][32]
    StageHead* sHead = reinterpret_cast<StageHead*>(sFrag + kConsumers * NB * 4 * 32);  // [kStages]
    float* sFeatAll = reinterpret_cast<float*>(sHead + kStages);                        // [kStages][32][FW]
    int* sRunP = reinterpret_cast<int*>(sFeatAll + (size_t)kStages * 32 * FW);           // [runsPerCta + 1]
    float* sInvR = reinterpret_cast<float*>(sRunP + a.runsPerCta + 1);                  // [min(B, kRadiusTable)]
    const bool radiusTable = a.B <= kRadiusTable;
    uint64_t* bars = reinterpret_cast<uint64_t*>(
        (reinterpret_cast<uintptr_t>(sInvR + (radiusTable ? a.B : 0)) + 7) & ~(uintptr_t)7);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int g = lane >> 2, t = lane & 3;
    const int M = live_count(a.dM, a.capM);
    const int K = live_count(a.dK, a.capK);
    const int totalRuns = (a.capM + a.G - 1) / a.G;
    const int run0 = blockIdx.x * a.runsPerCta;
    if (run0 >= totalRuns) return;
    const int R = min(a.runsPerCta, totalRuns - run0);
    const int chunkBlk0 = blockIdx.y * (kConsumers * NB);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            bar_init(&full[s], 32);          // every producer lane arrives after its own writes
            bar_init(&empty[s], kConsumers);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (radiusTable) {
        for (int b = threadIdx.x; b < a.B; b += blockDim.x) {
            const float r = a.scaleInv ? a.radius * max_extent(a.mn, a.mx, b) : a.radius;
            sInvR[b] = 1.0f / r;
        }
    }
    // first pair of every run of this CTA (the runs are consecutive, so is their pair range)
    for (int r = threadIdx.x; r <= R; r += blockDim.x) {
        const long long s = (long long)(run0 + r) * a.G;
        sRunP[r] = s < M ? min(a.start[s], K) : K;
    }
    // the chunk's parameters in fragment order: four 128-bit words per (block, lane), lane-contiguous so that every
    // read in the tile loop is one conflict-free LDS.128.  Flat reference layouts: W1[(8q+u)*3+axis], b1[8q+u],
    // W2[q*64 + out*8 + in], b2[8q+out], Wo[(8q+out)*8 + in], bo[8q+out]
    for (int idx = threadIdx.x; idx < kConsumers * NB * 32; idx += blockDim.x)
        build_fragment<MODE>(a, chunkBlk0 + (idx >> 5), idx & 31, (MODE == kDepthwise && FW > 0) ? chunkBlk0 * 8 : 0,
                             sFrag + (size_t)(idx >> 5) * 4 * 32 + (idx & 31));
    __syncthreads();

    // depth-wise: the feature columns this CTA's blocks need are one contiguous piece of every neighbour's row
    const int chunkCols = MODE == kDepthwise ? min(kConsumers * NB * 8, a.C - chunkBlk0 * 8) : a.C;

    if (warp < kProducers) {
        // ================================================================= PRODUCERS (warp w: batches w, w + 2, ...)
        int curRun = 0, curP = sRunP[0], seq = 0;
        auto next_batch = [&]() {
            Batch b;
            if (curRun >= R) {
                b.run = R;
                b.p0 = 0;
                b.n = 0;
                b.flags = kStop;
                return b;
            }
            const int P0 = sRunP[curRun], P1 = sRunP[curRun + 1];
            b.run = curRun;
            b.p0 = curP;
            b.n = max(min(32, P1 - curP), 0);
            b.flags = (curP == P0 ? kFirst : 0) | (curP + b.n >= P1 ? kLast : 0);
            curP += b.n;
            if (curP >= P1) {
                ++curRun;
                curP = max(curP, curRun <= R ? sRunP[min(curRun, R)] : curP);
            }
            return b;
        };
        int2 prA = make_int2(0, -1);
        float pdfA = 1.0f;
        auto load_a = [&](const Batch& b) {
            const bool ok = lane < b.n;
            prA = ok ? a.pairs[b.p0 + lane] : make_int2(0, -1);
            pdfA = ok ? a.pdfs[b.p0 + lane] : 1.0f;
        };
        auto load_b = [&]() {
            PairLoads L;
            L.j = prA.x;
            L.i = prA.y;
            L.pdf = pdfA;
            L.px = L.py = L.pz = L.sx = L.sy = L.sz = 0.f;
            L.bid = L.lo = L.hi = 0;
            L.f = 1.0f;
            if (L.i < 0) return L;
            const int j = L.j, i = L.i;
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
        // this warp's next batch of the common sequence; a stop that belongs to another producer ends this one silently
        auto take = [&]() {
            while (true) {
                Batch b = next_batch();
                const int idx = seq++;
                if (idx % kProducers == warp) {
                    b.seq = idx;
                    return b;
                }
                if (b.flags & kStop) {
                    b.flags |= kSilent;
                    b.seq = idx;
                    return b;
                }
            }
        };
        Batch bW = take();  // batch being written
        load_a(bW);
        PairLoads L = load_b();
        Batch bG = take();  // batch whose pair records are in flight
        load_a(bG);
        const int rowChunks = chunkCols >> 2;  // 16-byte pieces of a staged feature row
        const int lanesPerRow = rowChunks <= 8 ? 8 : (rowChunks <= 16 ? 16 : 32);
        while (!(bW.flags & kSilent)) {
            const int it = bW.seq;
            const int s = it % kStages;
            bar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
            StageHead& H = sHead[s];
            const bool ok = L.i >= 0 && !(bW.flags & kStop);
            const int s0 = (run0 + bW.run) * a.G;
            float invR = 0.f;
            if (ok) {
                if (radiusTable) {
                    invR = sInvR[L.bid];
                } else {
                    const float ext = max_extent(a.mn, a.mx, L.bid);
                    invR = 1.0f / (a.scaleInv ? a.radius * ext : a.radius);
                }
            }
            const float cnt = a.avg ? (float)(min(L.hi, K) - L.lo) : 1.0f;
            H.d[0][lane] = ok ? (L.px - L.sx) * invR : 0.f;
            H.d[1][lane] = ok ? (L.py - L.sy) * invR : 0.f;
            H.d[2][lane] = ok ? (L.pz - L.sz) * invR : 0.f;
            H.d[3][lane] = 1.0f;
            H.w[lane] = ok ? L.f / (L.pdf * cnt) : 0.0f;
            H.local[lane] = ok ? L.i - s0 : -1;
            if (MODE == kGeneric) H.neigh[lane] = ok ? L.j : 0;
            if (lane == 0) {
                H.s0 = s0;
                H.rows = min(a.G, a.capM - s0);
                H.flags = bW.flags;
                H.n = bW.n;
            }
            bool async = false;
            if (MODE == kDepthwise && FW > 0 && !(bW.flags & kStop)) {
                // the chunk's piece of every neighbour's feature row: 16-byte asynchronous copies, a row (or 2 / 4
                // narrow rows) per warp instruction -- coalesced 128-bit reads of contiguous row pieces; empty slots
                // are zero-filled by the copy itself (source size 0).  Completion is tied to this lane's arrival.
                const int sub = lane / lanesPerRow, c = lane % lanesPerRow;
                const uint32_t dst0 = sm_u32(sFeatAll + ((size_t)s * 32 + sub) * FW + c * 4);
                // this lane's row piece (empty slots: any valid address, nothing is read), handed around by shuffles
                const unsigned long long mine =
                    reinterpret_cast<unsigned long long>(ok ? a.feats + (size_t)L.j * a.C + chunkBlk0 * 8 : a.feats);
                const uint32_t mineLo = (uint32_t)mine, mineHi = (uint32_t)(mine >> 32), mineOk = ok ? 16u : 0u;
#pragma unroll 4
                for (int p0 = 0; p0 < 32; p0 += 32 / lanesPerRow) {
                    const int pr = p0 + sub;
                    const uint32_t lo = __shfl_sync(MC_FULL_MASK, mineLo, pr), hi = __shfl_sync(MC_FULL_MASK, mineHi, pr);
                    const uint32_t nb = __shfl_sync(MC_FULL_MASK, mineOk, pr);
                    if (c < rowChunks) {
                        const unsigned long long src = (((unsigned long long)hi << 32) | lo) + (nb ? c * 16u : 0u);
                        cp_async16(dst0 + (uint32_t)(p0 * FW) * 4u, reinterpret_cast<const void*>(src), nb);
                    }
                }
                async = true;
            } else if (MODE == kGeneric && FW > 0) {  // few input features: plain loads
                float* row = sFeatAll + ((size_t)s * 32 + lane) * FW;
                for (int c = 0; c < a.C; ++c) row[c] = ok ? a.feats[(size_t)L.j * a.C + c] : 0.f;
            }
            if (async) bar_arrive_after_copies(&full[s]);
            else bar_arrive(&full[s]);
            if (bW.flags & kStop) break;
            bW = bG;
            L = load_b();
            bG = take();
            load_a(bG);
        }
        return;
    }

    // ===================================================================== CONSUMERS
    const int cw = warp - kProducers;
    const int blk0 = chunkBlk0 + cw * NB;
    const float4* myFrag = sFrag + (size_t)cw * NB * 4 * 32 + lane;
    // accumulator fragments, rows = samples s0 + {g, g+8}: out[.][8 blk + {2t, 2t+1}], or (single-input mode)
    // Z[.][unit {2t, 2t+1}] and the weight sums F[.]
    float acc[NB][4];
    float accF[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < NB; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;

    for (int it = 0;; ++it) {
        const int s = it % kStages;
        bar_wait(&full[s], (it / kStages) & 1);
        const StageHead& H = sHead[s];
        const int flags = H.flags;
        if (flags & kStop) break;
        const int left = H.n, s0 = H.s0, rowsOut = H.rows;
        if (flags & kFirst) {
#pragma unroll
            for (int q = 0; q < NB; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
            accF[0] = accF[1] = accF[2] = accF[3] = 0.f;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h * 16 >= left || blk0 >= a.numBlocks) break;  // warp-uniform
            const int r0 = h * 16;
            // layer-1 A fragment (shared by all blocks): component t of pairs g / g+8
            const float da = H.d[t][r0 + g], db = H.d[t][r0 + g + 8];
            // pairs this lane holds as COLUMNS of the transposed fragments: 2t, 2t+1 (first 8) and 8+2t, 8+2t+1
            const float2 wA = *reinterpret_cast<const float2*>(&H.w[r0 + 2 * t]);
            const float2 wB = *reinterpret_cast<const float2*>(&H.w[r0 + 8 + 2 * t]);
            const int2 lA = *reinterpret_cast<const int2*>(&H.local[r0 + 2 * t]);
            const int2 lB = *reinterpret_cast<const int2*>(&H.local[r0 + 8 + 2 * t]);
            // feature x weight of the four pairs this lane multiplies with, per block: shared memory when the rows are
            // staged, global memory for wide multi-feature convolutions
            float f[NB][4];
            float bo[NB];
            if constexpr (MODE != kSingleIn) {
                if (FW > 0) {
                    const uint32_t base = sm_u32(sFeatAll + ((size_t)s * 32 + r0 + 2 * t) * FW);
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const float2 p3 = frag_ld2(myFrag, q, 3);
                        bo[q] = p3.x;
                        const uint32_t col = base + 4u * (uint32_t)__float_as_int(p3.y);
                        f[q][0] = lds_f32(col) * wA.x;
                        f[q][1] = lds_f32(col + 4u * FW) * wA.y;
                        f[q][2] = lds_f32(col + 32u * FW) * wB.x;
                        f[q][3] = lds_f32(col + 36u * FW) * wB.y;
                    }
                } else {
                    const int2 jA = *reinterpret_cast<const int2*>(&H.neigh[r0 + 2 * t]);
                    const int2 jB = *reinterpret_cast<const int2*>(&H.neigh[r0 + 8 + 2 * t]);
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const float2 p3 = frag_ld2(myFrag, q, 3);
                        bo[q] = p3.x;
                        const int fcol = __float_as_int(p3.y);
                        f[q][0] = a.feats[(size_t)jA.x * a.C + fcol] * wA.x;
                        f[q][1] = a.feats[(size_t)jA.y * a.C + fcol] * wA.y;
                        f[q][2] = a.feats[(size_t)jB.x * a.C + fcol] * wB.x;
                        f[q][3] = a.feats[(size_t)jB.y * a.C + fcol] * wB.y;
                    }
                }
            }
            tile_body<MODE, NB>(myFrag, da, db, wA, wB, lA, lB, g, f, bo, acc, accF);
        }
        if (flags & kLast) store_run<MODE, NB>(a, myFrag, acc, accF, blk0, s0, rowsOut, g, t);
        __syncwarp();
        if (lane == 0) bar_arrive(&empty[s]);
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

// warp-level variant: blocks past the end of the parameter arrays are evaluated with zero parameters (no branch in the
// tile loop); the chunk width is matched to small convolutions instead
template <int MODE>
int launch_warp(ConvArgs a, int nb, long long avgK, cudaStream_t s) {
    // run length per warp: ~8 tiles of 16 pairs, at most the 16 rows of the accumulator fragment -- but never at the
    // price of leaving the GPU short of warps
    const int chunks = cdiv(nb, 4);
    int G = (int)(128 / (avgK > 0 ? avgK : 1));
    G = G < 1 ? 1 : (G > kMaxRun ? kMaxRun : G);
    while (G > 1 && (long long)cdiv(a.capM, G) * chunks < 1776) --G;
    a.G = G;
    const int runs = cdiv(a.capM, G);
    if (nb >= 3) {
        dim3 grid(cdiv(runs, kConvWarps), cdiv(nb, 4));
        spatial_conv_warp_kernel<MODE, 4><<<grid, kConvWarps * 32, 0, s>>>(a);
    } else if (nb == 2) {
        dim3 grid(cdiv(runs, kConvWarps), 1);
        spatial_conv_warp_kernel<MODE, 2><<<grid, kConvWarps * 32, 0, s>>>(a);
    } else {
        dim3 grid(cdiv(runs, kConvWarps), 1);
        spatial_conv_warp_kernel<MODE, 1><<<grid, kConvWarps * 32, 0, s>>>(a);
    }
    return 0;
}

constexpr int kTargetCtas = 148 * 4;
constexpr long long kPersistentMinWork = 10LL << 20;

template <int MODE, int NB, int NC>
int launch_variant(ConvArgs a, int nb, cudaStream_t s) {
    const int chunkBlocks = NC * NB;
    const int chunks = cdiv(nb, chunkBlocks);
    int FW = 0;
    if (MODE == kDepthwise) FW = min(chunkBlocks * 8, a.C) + 4;       // + 4: conflict-free reads of pairs 2t, 2t+1, ..
    if (MODE == kGeneric && a.C <= kGenericStageC) FW = ((a.C + 3) & ~3) + 4;
    // persistent CTAs, every CTA a contiguous range of runs: twice as many as are resident at once (so that uneven
    // pair counts even out), never more than there are runs
    const int totalRuns = cdiv(a.capM, a.G);
    const size_t fixed = (size_t)chunkBlocks * 4 * 32 * 16 + kStages * sizeof(StageHead) + (size_t)kStages * 32 * FW * 4 +
                         (a.B <= kRadiusTable ? (size_t)a.B * 4 : 0) + 16 + 2 * kStages * 8;
    cudaFuncSetAttribute(spatial_conv_kernel<MODE, NB, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    int perSm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, spatial_conv_kernel<MODE, NB, NC>, (kProducers + NC) * 32,
                                                  fixed + 1024);
    perSm = perSm < 1 ? 1 : perSm;
    int gridX = cdiv(148 * perSm * 2, chunks);
    if (gridX > totalRuns) gridX = totalRuns;
    if (gridX < cdiv(totalRuns, kMaxRunsPerCta)) gridX = cdiv(totalRuns, kMaxRunsPerCta);
    a.runsPerCta = cdiv(totalRuns, gridX);
    gridX = cdiv(totalRuns, a.runsPerCta);
    const size_t smem = fixed + (size_t)(a.runsPerCta + 1) * 4;
    dim3 grid(gridX, chunks);
    spatial_conv_kernel<MODE, NB, NC><<<grid, (kProducers + NC) * 32, smem, s>>>(a, FW);
    return 0;
}

template <int MODE>
int launch_conv(const ConvArgs& a, int nb, int NB, int NC, cudaStream_t s) {
    if (NC == 8) {
        if (NB == 2) return launch_variant<MODE, 2, 8>(a, nb, s);
        return launch_variant<MODE, 1, 8>(a, nb, s);
    }
    if (NB == 4) return launch_variant<MODE, 4, 4>(a, nb, s);
    if (NB == 2) return launch_variant<MODE, 2, 4>(a, nb, s);
    return launch_variant<MODE, 1, 4>(a, nb, s);
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
    // Work decomposition: a CTA = (range of runs of G samples) x (4 consumer warps x NB MLP blocks).  Wide convolutions
    // over many samples take the largest units (16-row accumulator fragments, 16 blocks per staged pair batch); small
    // ones (the latency-bound upper levels of the classification networks) are cut finer until the GPU has CTAs:
    // first shorter runs while a run still has a few tiles, then fewer blocks per warp, then single-sample runs.
    const long long avgK = numNeighbors / (long long)numSamples;
    int NC = nb >= 16 ? 8 : 4;   // consumer warps per CTA
    int NB = nb > 4 ? 2 : 1;
    int G = kMaxRun;
    auto ctas = [&]() { return (long long)cdiv(numSamples, G) * cdiv(nb, NC * NB); };
    while (ctas() < kTargetCtas && G > 1 && avgK * G >= 128) G >>= 1;   // shorter runs first: no redundant staging
    while (ctas() < kTargetCtas && NB > 1) NB >>= 1;
    if (ctas() < kTargetCtas && NC > 4) NC = 4;
    while (ctas() < kTargetCtas && G > 1) G >>= 1;
    // The persistent kernel pays off when there is enough work to amortise its pipeline -- measured on the convolutions
    // of the four networks (tools/conv_bench.py, profiles/r2_spatial_conv_experiments.md): (pairs x blocks) above ~10 M;
    // the single-input mode has no feature rows to stage and stays on the warp-level kernel.
    bool persistent = (C > 1 || !combin) && (long long)numNeighbors * nb >= kPersistentMinWork;
    if (const char* e = std::getenv("MCCNN_CONV_CFG")) {  // experiment knob (tools/conv_bench.py): "warp" | "NC,NB"
        int nc = 0, nbv = 0;
        if (std::strcmp(e, "warp") == 0) persistent = false;
        else if (std::sscanf(e, "%d,%d", &nc, &nbv) == 2 && (nc == 4 || nc == 8) &&
                 (nbv == 1 || nbv == 2 || (nbv == 4 && nc == 4))) {
            persistent = true;
            NC = nc;
            NB = nbv;
        }
    }
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
    a.runsPerCta = 1;
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
        if (persistent) launch_conv<kDepthwise>(a, nb, NB, NC, s);
        else launch_warp<kDepthwise>(a, nb, avgK, s);
    } else if (C == 1) {
        a.nOut = nOut;
        a.outStride = numOutFeatures;
        a.out = outFeatures;
        if (persistent) launch_conv<kSingleIn>(a, nb, NB, NC, s);
        else launch_warp<kSingleIn>(a, nb, avgK, s);
    } else {
        MC_REQUIRE(workspace != nullptr, "mc_spatial_conv: workspace required for multi-feature convolutions");
        a.nOut = nb * 8;
        a.outStride = nb * 8;
        a.out = workspace;
        if (persistent) launch_conv<kGeneric>(a, nb, NB, NC, s);
        else launch_warp<kGeneric>(a, nb, avgK, s);
        long long total = (long long)numSamples * numOutFeatures;
        combine_kernel<<<cdiv(total, 256), 256, 0, s>>>(dNumSamples, numSamples, C, numOutFeatures, nb * 8, workspace,
                                                        outFeatures);
        count_launches(1);
    }
    return check_launch("mc_spatial_conv");
}
