// Probe: what a tcgen05 / TMEM formulation of spatial_conv's per-pair MLP (3 -> 8 -> 8 -> 8 per block, 128 pairs per
// tile, accumulators in tensor memory) would be paced by on sm_100a.  Two rates are measured per SM, timing-only (the
// operands are valid but meaningless):
//   1. the tensor pipe on the block-diagonal shapes: tcgen05.mma.kind::tf32 M = 128, K = 8 per instruction,
//      N = 8 / 16 (one / two MLP blocks per instruction) and N = 64 / 256 (layer 1 batched over blocks), A from
//      shared memory (SS) or from tensor memory (TS);
//   2. the hop between two layers: the hidden activations of a block, [128 pairs x 8] fp32 in TMEM, have to come back
//      to registers (tcgen05.ld), get ReLU + the hi / lo TF32 split, and return as the next layer's A operand
//      (tcgen05.st, 16 columns) -- the mma.sync kernel does this hop in registers for free (C fragment == A fragment).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tcgen05_mlp_probe tcgen05_mlp_probe.cu ; run on the box.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
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
        if (++spins > (1u << 26)) __trap();
    }
}
// K-major tile, 128-byte rows, SWIZZLE_128B (same descriptor as csrc/dense_tc.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD),
        "l"(descA), "l"(descB), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmemD),
        "r"(tmemA), "l"(descB), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u));
}

// ---- 1. tensor pipe: `reps` back-to-back tcgen05.mma M128 x N x K8 (TF32), one commit at the end -------------------
template <int N, bool TS>
__global__ void __launch_bounds__(128) mma_stream(long long* cycles, int reps) {
    extern __shared__ unsigned char smemRaw[];
    unsigned char* smem = smemRaw + ((1024u - (smem_addr(smemRaw) & 1023u)) & 1023u);
    float* sA = reinterpret_cast<float*>(smem);                 // 128 rows x 128 B
    float* sB = reinterpret_cast<float*>(smem + 16384);         // N rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) sA[i] = 1e-3f * (i & 63);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t tm = tmem_alloc_all(&slot, warp);
    {   // defined contents for the TS operand (columns 256..271) and the accumulators
        uint32_t z[16];
        for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(0.25f);
        for (int c = 0; c < 512; c += 16) tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + c, z);
        tmem_st_wait();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint64_t dA = umma_desc(smem_addr(sA)), dB = umma_desc(smem_addr(sB));
        constexpr uint32_t idesc = umma_idesc(128, N);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tm + (uint32_t)((r & 1) * 0);   // one accumulator, as a K loop would use it
            if (TS) umma_ts(d, tm + 256u + (uint32_t)((r & 3) * 8), dB + (uint64_t)((r & 3) * 2), idesc);
            else umma_ss(d, dA + (uint64_t)((r & 3) * 2), dB + (uint64_t)((r & 3) * 2), idesc);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tmem_free_all(tm, warp);
}

// ---- 2. the hop between layers: [128 x 8] per block out of TMEM, ReLU + hi/lo split, [128 x 16] back ---------------
// MODE 0: ld only, 1: st only, 2: ld + transform + st.  `blocks` MLP blocks per tile, `tiles` tiles.
template <int MODE>
__global__ void __launch_bounds__(256) tmem_hop(long long* cycles, float* sink, int tiles, int blocks) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t tm = tmem_alloc_all(&slot, warp);
    const uint32_t lane0 = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t z[16];
    for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(0.25f * (i + 1) - 1.0f);
    if (warp < 4) {
        for (int c = 0; c < 512; c += 16) tmem_st16(lane0 + c, z);
        tmem_st_wait();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // with 8 warps, warps w and w + 4 share a lane quarter and split the blocks
    const int nw = blockDim.x >> 7;            // warps per lane quarter
    const int sub = warp >> 2;
    float acc = 0.f;
    long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
        for (int b = sub; b < blocks; b += nw) {
            uint32_t h[8];
            if (MODE != 1) {
                tmem_ld8(lane0 + (uint32_t)((b & 31) * 8), h);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = z[i] + t;
            }
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += __uint_as_float(h[i]);
            } else {
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = fmaxf(__uint_as_float(h[i]), 0.f);
                    uint32_t hi = __float_as_uint(v) & 0xffffe000u;
                    o[i] = hi;
                    o[8 + i] = __float_as_uint(v - __uint_as_float(hi));
                }
                tmem_st16(lane0 + 256u + (uint32_t)((b & 15) * 16), o);
            }
        }
        if (MODE != 0) tmem_st_wait();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tmem_free_all(tm, warp);
}

static double median_cycles(long long* d, int n) {
    std::vector<long long> h(n);
    cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    return (double)h[n / 2];
}

template <int N, bool TS>
static int run_mma(long long* cyc, int sms) {
    const int reps = 4096;
    CK(cudaFuncSetAttribute(mma_stream<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024));
    mma_stream<N, TS><<<sms, 128, 16384 + 32768 + 1024>>>(cyc, reps);
    CK(cudaDeviceSynchronize());
    mma_stream<N, TS><<<sms, 128, 16384 + 32768 + 1024>>>(cyc, reps);
    CK(cudaDeviceSynchronize());
    double c = median_cycles(cyc, sms) / reps;
    printf("tcgen05.mma tf32 M=128 N=%3d K=8 %s : %7.2f cycles / instruction  (%5.2f cycles per 8-unit MLP block of the "
           "128-pair tile)\n", N, TS ? "A in TMEM" : "A in smem", c, c / (N / 8));
    return 0;
}

template <int MODE>
static int run_hop(long long* cyc, float* sink, int sms, int threads) {
    const int tiles = 256, blocks = 32;
    tmem_hop<MODE><<<sms, threads>>>(cyc, sink, tiles, blocks);
    CK(cudaDeviceSynchronize());
    tmem_hop<MODE><<<sms, threads>>>(cyc, sink, tiles, blocks);
    CK(cudaDeviceSynchronize());
    double c = median_cycles(cyc, sms) / (tiles * blocks);
    const char* what = MODE == 0 ? "tcgen05.ld x8 only           " : MODE == 1 ? "split + tcgen05.st x16 only  "
                                                                               : "ld x8 + relu/split + st x16  ";
    printf("TMEM hop, %d warps: %s : %7.2f cycles per [128 pairs x 8] block  (%6.1f B/cycle/SM read, %6.1f written)\n",
           threads / 32, what, c, MODE == 1 ? 0.0 : 4096.0 / c, MODE == 0 ? 0.0 : 8192.0 / c);
    return 0;
}

int main(int argc, char** argv) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* cyc;
    float* sink;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    CK(cudaMalloc(&sink, sms * 256 * sizeof(float)));
    if (argc <= 1) printf("%d SMs, one CTA per SM, cycles = SM clock (clock64), median over CTAs\n", sms);
    // one test per process (argv[1]): an illegal shape traps and poisons the context
    const int which = argc > 1 ? atoi(argv[1]) : -1;
    switch (which) {
        case 0: return run_mma<16, false>(cyc, sms);
        case 1: return run_mma<64, false>(cyc, sms);
        case 2: return run_mma<256, false>(cyc, sms);
        case 3: return run_mma<16, true>(cyc, sms);
        case 4: return run_mma<64, true>(cyc, sms);
        case 5: return run_hop<0>(cyc, sink, sms, 128);
        case 6: return run_hop<1>(cyc, sink, sms, 128);
        case 7: return run_hop<2>(cyc, sink, sms, 128);
        case 8: return run_hop<0>(cyc, sink, sms, 256);
        case 9: return run_hop<1>(cyc, sink, sms, 256);
        case 10: return run_hop<2>(cyc, sink, sms, 256);
        case 11: return run_mma<8, false>(cyc, sms);   // M = 128 with N = 8: outside the PTX shape table
        case 12: return run_mma<8, true>(cyc, sms);
        default: break;
    }
    printf("mma.sync kernel for comparison (csrc/spatial_conv.cu): 14 m16n8k8-class instructions per (16 pairs x block) at "
           "2 cycles each = 224 cycles per (128 pairs x block) at a fully busy tensor pipe; measured 49 %% busy.\n");
    return 0;
}
