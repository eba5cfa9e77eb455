// dense_tc.cu -- the dense glue on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), FP32 in / FP32 out.
//
//     y = act2( (act1(x * s1 + t1) @ W + bias) * s2 + t2 )          x [rows, K], W [K, N], y [rows, N]
//
// conv_1x1 and every layer of MLP_1_hidden / MLP_2_hidden with the surrounding inference batch norms folded into
// per-column (scale, shift) vectors (utils/MCNetworkUtils.py:20-142; TensorFlow's matmul = cuBLAS in the reference).
// In the segmentation networks these layers run over every point (MCSeg: 131 072 rows x 832 -> 256 = 28 GMAC): a
// true dense GEMM, the one place on the path where tcgen05 is the natural fit.
//
// Precision: parity is 1e-5 relative, single-pass TF32 (10-bit mantissa) gives ~1e-3.  Every product is the 3xTF32
// split  a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo  (hi = tf32(v), lo = v - hi, FP32 accumulation in TMEM): ~2^-21
// relative per product.  W is static: its transposed hi / lo halves are prepared once per weight version
// (mc_split_tf32).  x is split on the fly, in shared memory, by the same pass that applies the BN + ReLU prologue.
//
// One CTA = one 128 x BN output tile, 192 threads, warp-specialised:
//   warp 0     TMA producer: per K block of 32 floats (128 B rows, SWIZZLE_128B) one cp.async.bulk.tensor for the raw
//              x tile [128 x 32] and one each for the W_hi / W_lo tiles [BN x 32], completing on full[stage];
//   warps 2-5  transform: wait full[stage]; rewrite the x tile IN PLACE as tf32(act1(x * s1 + t1)) and write the
//              remainder to the lo tile at the same (swizzled) offsets; fence.proxy.async; arrive on ready[stage];
//   warp 1     MMA issuer (one thread): wait ready[stage]; 4 K steps x 3 tcgen05.mma.kind::tf32 (M 128, N BN, K 8),
//              operands through shared-memory descriptors, accumulator in TMEM; tcgen05.commit -> empty[stage]
//              (slot back to the producer), after the last K block tcgen05.commit -> accFull;
//   warps 2-5  epilogue: wait accFull; tcgen05.ld 32 columns at a time (thread = accumulator row); bias, folded batch
//              norm, ReLU; 128-bit stores of the live rows.
// Grid = (column tiles, row tiles) with the column tile fastest, so the CTAs that share an x tile run together and
// its second read hits L2.
#include <cuda.h>

#include "common.cuh"

namespace mc {
namespace {

constexpr int kBM = 128;           // rows per CTA = TMEM lanes
constexpr int kBK = 32;            // floats per K block = one 128-byte swizzle row
constexpr int kTcThreads = 192;
constexpr int kTransformThreads = 128;
constexpr uint32_t kSpinLimit = 1u << 27;  // a wait that spins this long is a protocol bug: trap instead of hanging

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > kSpinLimit) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// shared-memory matrix descriptor of a K-major tile with 128-byte rows under SWIZZLE_128B (8-row atoms of 1024 B):
// start address >> 4 | leading byte offset (unused for swizzled K-major, 1) << 16 | stride byte offset 1024 >> 4 << 32 |
// descriptor version 1 << 46 | layout type SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// instruction descriptor of tcgen05.mma.kind::tf32: D format F32 (1) << 4 | A format TF32 (2) << 7 | B format TF32 (2)
// << 10 | A, B K-major (0) | N >> 3 << 17 | M >> 4 << 24
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcArgs {
    const int* dRows;
    int capRows, K, N;
    const float *s1, *t1, *bias, *s2, *t2;
    int relu1, relu2;
    float* y;
};

template <int BN>
struct TcSmem {
    static constexpr int kStages = BN == 128 ? 3 : 4;
    static constexpr int kABytes = kBM * kBK * 4;   // 16 KB
    static constexpr int kWBytes = BN * kBK * 4;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;
    static constexpr int kBarBytes = 1024;
    static constexpr int kMaxStagedK = 2048;                    // prologue scale / shift of the whole K kept in smem
    static constexpr int kParamBytes = (2 * kMaxStagedK + 3 * BN) * 4;
    static constexpr int kTotal = kStages * kStageBytes + kBarBytes + kParamBytes + 1024;  // + 1024-byte alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
    dense_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWhi,
                    const __grid_constant__ CUtensorMap mapWlo, const TcArgs a) {
    using L = TcSmem<BN>;
    constexpr int S = L::kStages;
    extern __shared__ uint8_t smemRaw[];
    const int rows = live_count(a.dRows, a.capRows);
    const int row0 = blockIdx.y * kBM, col0 = blockIdx.x * BN;
    if (row0 >= rows) return;  // CTA-uniform, before any barrier exists

    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + S * L::kStageBytes);
    uint64_t* full = bars;            // TMA landed (tx bytes)
    uint64_t* ready = bars + S;       // x tile transformed and split
    uint64_t* empty = bars + 2 * S;   // MMAs of the stage retired
    uint64_t* accFull = bars + 3 * S;
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);
    // per-column vectors of the prologue (whole K) and of this tile's epilogue: global loads once per CTA instead of
    // per K block in the transform's dependent chain (ncu: 23 % of the stall samples sat on those loads)
    float* sS1 = reinterpret_cast<float*>(base + S * L::kStageBytes + L::kBarBytes);
    float* sT1 = sS1 + L::kMaxStagedK;
    float* sBias = sT1 + L::kMaxStagedK;
    float* sS2 = sBias + BN;
    float* sT2 = sS2 + BN;
    auto stageA = [&](int s) { return base + s * L::kStageBytes; };
    auto stageAlo = [&](int s) { return base + s * L::kStageBytes + L::kABytes; };
    auto stageWhi = [&](int s) { return base + s * L::kStageBytes + 2 * L::kABytes; };
    auto stageWlo = [&](int s) { return base + s * L::kStageBytes + 2 * L::kABytes + L::kWBytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int numKb = (a.K + kBK - 1) / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], kTransformThreads);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accFull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool stagedAffine = a.s1 != nullptr && a.K <= L::kMaxStagedK;
    if (stagedAffine) {
        for (int k = threadIdx.x; k < a.K; k += blockDim.x) {
            sS1[k] = a.s1[k];
            sT1[k] = a.t1[k];
        }
    }
    for (int c = threadIdx.x; c < BN; c += blockDim.x) {
        const bool in = col0 + c < a.N;
        sBias[c] = (in && a.bias != nullptr) ? a.bias[col0 + c] : 0.f;
        sS2[c] = (in && a.s2 != nullptr) ? a.s2[col0 + c] : 1.f;
        sT2[c] = (in && a.s2 != nullptr) ? a.t2[col0 + c] : 0.f;
    }
    if (warp == 1) {  // TMEM: BN accumulator columns (power of two >= 32), allocated and freed by this warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmemSlot)),
                     "r"((uint32_t)BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = *tmemSlot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < numKb; ++kb) {
                const int s = kb % S;
                mbar_wait(&empty[s], ((kb / S) & 1) ^ 1);
                mbar_expect_tx(&full[s], L::kABytes + 2 * L::kWBytes);
                tma_load_2d(stageA(s), &mapX, &full[s], kb * kBK, row0);
                tma_load_2d(stageWhi(s), &mapWhi, &full[s], kb * kBK, col0);
                tma_load_2d(stageWlo(s), &mapWlo, &full[s], kb * kBK, col0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kBM, BN);
            for (int kb = 0; kb < numKb; ++kb) {
                const int s = kb % S;
                mbar_wait(&ready[s], (kb / S) & 1);
                tc_fence_after();
                const uint64_t dA = umma_desc(smem_addr(stageA(s))), dAl = umma_desc(smem_addr(stageAlo(s)));
                const uint64_t dW = umma_desc(smem_addr(stageWhi(s))), dWl = umma_desc(smem_addr(stageWlo(s)));
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {  // 8 tf32 = 32 bytes per K step: +2 in the (>> 4) address field
                    const uint64_t adv = (uint64_t)(k * 2);
                    umma_tf32(tmemBase, dAl + adv, dW + adv, idesc, (kb | k) != 0);   // small terms first
                    umma_tf32(tmemBase, dA + adv, dWl + adv, idesc, 1);
                    umma_tf32(tmemBase, dA + adv, dW + adv, idesc, 1);
                }
                umma_commit(&empty[s]);  // implies tcgen05.fence::before_thread_sync
            }
            umma_commit(accFull);
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- transform, then epilogue (128 threads)
        const int t = threadIdx.x - 64;
        const bool affine = a.s1 != nullptr;
        for (int kb = 0; kb < numKb; ++kb) {
            const int s = kb % S;
            mbar_wait(&full[s], (kb / S) & 1);
            float4* A = reinterpret_cast<float4*>(stageA(s));
            float4* Alo = reinterpret_cast<float4*>(stageAlo(s));
#pragma unroll
            for (int u = 0; u < (kBM * kBK / 4) / kTransformThreads; ++u) {  // 8 x 16-byte chunks per thread
                const int c = u * kTransformThreads + t;                      // physical chunk: row c / 8, slot c % 8
                const int r = c >> 3;
                const int k = kb * kBK + (((c & 7) ^ (r & 7)) << 2);          // logical K of the chunk (un-swizzled)
                float4 v = A[c];
                if (affine) {
                    if (k + 3 < a.K) {
                        const float4 sc = stagedAffine ? *reinterpret_cast<const float4*>(sS1 + k)
                                                       : __ldg(reinterpret_cast<const float4*>(a.s1 + k));
                        const float4 sh = stagedAffine ? *reinterpret_cast<const float4*>(sT1 + k)
                                                       : __ldg(reinterpret_cast<const float4*>(a.t1 + k));
                        v.x = fmaf(v.x, sc.x, sh.x);
                        v.y = fmaf(v.y, sc.y, sh.y);
                        v.z = fmaf(v.z, sc.z, sh.z);
                        v.w = fmaf(v.w, sc.w, sh.w);
                    } else {  // K tail (K % 4 == 0, so a chunk is entirely inside or outside): zero-filled by TMA
                        v = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (a.relu1) {
                    v.x = fmaxf(v.x, 0.f);
                    v.y = fmaxf(v.y, 0.f);
                    v.z = fmaxf(v.z, 0.f);
                    v.w = fmaxf(v.w, 0.f);
                }
                float4 hi, lo;
                hi.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u);
                hi.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u);
                hi.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u);
                hi.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u);
                lo.x = v.x - hi.x;
                lo.y = v.y - hi.y;
                lo.z = v.z - hi.z;
                lo.w = v.w - hi.w;
                A[c] = hi;
                Alo[c] = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
            mbar_arrive(&ready[s]);
        }
        // epilogue: thread = accumulator row (TMEM lane); a warp may only touch the lane quarter (warp id % 4)
        mbar_wait(accFull, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = row0 + q * 32 + lane;
        const bool vec = (a.N & 3) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (col0 + c0 >= a.N) break;  // warp-uniform
            float v[32];
            tmem_ld32(tmemBase + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (row < rows) {
                float* out = a.y + (size_t)row * a.N + col0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int c = col0 + c0 + j;
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = v[j + e] + sBias[c0 + j + e];   // columns past N: bias 0, scale 1, shift 0
                        x = fmaf(x, sS2[c0 + j + e], sT2[c0 + j + e]);
                        o[e] = a.relu2 ? fmaxf(x, 0.f) : x;
                    }
                    if (vec && c + 3 < a.N) {
                        *reinterpret_cast<float4*>(out + j) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (c + e < a.N) out[j + e] = o[e];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)BN));
    }
}

__global__ void __launch_bounds__(256) split_tf32_kernel(long long n, const float* __restrict__ in,
                                                         float* __restrict__ hi, float* __restrict__ lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = in[i];
        const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// row-major float matrix [rows, cols] -> tiles of [boxRows x 32 floats], 128-byte swizzle, zero fill outside
bool make_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, int boxRows) {
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)boxRows};
    const cuuint32_t elemStrides[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, elemStrides,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
int launch_tc(const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml, const TcArgs& a, cudaStream_t s) {
    using L = TcSmem<BN>;
    cudaFuncSetAttribute(dense_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    dim3 grid(cdiv(a.N, BN), cdiv(a.capRows, kBM));
    dense_tc_kernel<BN><<<grid, kTcThreads, L::kTotal, s>>>(mx, mh, ml, a);
    return check_launch("mc_dense_layer_tc");
}

}  // namespace
}  // namespace mc

using namespace mc;

// hi[i] = tf32(in[i]) (round to nearest, low 13 mantissa bits zero), lo[i] = in[i] - hi[i]: the static operand of
// mc_dense_layer_tc, prepared once per weight version on the TRANSPOSED weight matrix [N, K].
extern "C" int mc_split_tf32(long long count, const float* in, float* hi, float* lo, void* stream) {
    if (count <= 0) return 0;
    int grid = cdiv(count, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(count, in, hi, lo);
    return check_launch("mc_split_tf32");
}

// 1 when mc_dense_layer_tc accepts the layer: K a multiple of 4 (16-byte row pitch for TMA) and at least one K block
extern "C" int mc_dense_layer_tc_supported(int K, int N) { return (K % 4 == 0 && K >= 32 && N >= 1) ? 1 : 0; }

// y = act2((act1(x * s1 + t1) @ W + bias) * s2 + t2) on the tcgen05 tensor cores.  x [numRows, K] (16-byte aligned),
// WtHi / WtLo [N, K]: the two halves of mc_split_tf32 applied to W TRANSPOSED (W is the reference's [K, N] weight,
// utils/MCNetworkUtils.py:44,123); s1/t1, bias, s2/t2 optional (NULL); y [numRows, N].  Live rows: min(*dNumRows, numRows).
extern "C" int mc_dense_layer_tc(int numRows, const int* dNumRows, int K, int N, const float* x, const float* s1,
                                 const float* t1, int relu1, const float* WtHi, const float* WtLo, const float* bias,
                                 const float* s2, const float* t2, int relu2, float* y, void* stream) {
    MC_REQUIRE(mc_dense_layer_tc_supported(K, N), "mc_dense_layer_tc: K must be a multiple of 4 and >= 32");
    MC_REQUIRE((s1 == nullptr) == (t1 == nullptr) && (s2 == nullptr) == (t2 == nullptr),
               "mc_dense_layer_tc: scale and shift come in pairs");
    MC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(WtHi) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(WtLo) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
               "mc_dense_layer_tc: operands must be 16-byte aligned");
    if (numRows <= 0) return 0;
    const int BN = N > 64 ? 128 : 64;
    CUtensorMap mx, mh, ml;
    if (!make_map(&mx, x, numRows, K, kBM) || !make_map(&mh, WtHi, N, K, BN) || !make_map(&ml, WtLo, N, K, BN)) {
        set_error("mc_dense_layer_tc: cuTensorMapEncodeTiled failed");
        return 2;
    }
    TcArgs a;
    a.dRows = dNumRows;
    a.capRows = numRows;
    a.K = K;
    a.N = N;
    a.s1 = s1;
    a.t1 = t1;
    a.bias = bias;
    a.s2 = s2;
    a.t2 = t2;
    a.relu1 = relu1;
    a.relu2 = relu2;
    a.y = y;
    return BN == 128 ? launch_tc<128>(mx, mh, ml, a, (cudaStream_t)stream)
                     : launch_tc<64>(mx, mh, ml, a, (cudaStream_t)stream);
}
