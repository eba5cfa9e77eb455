// Probe: issue rate and dependent latency of mma.sync.m16n8k8 TF32 (and m16n8k4) on sm_100a, per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_k8(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_k4(float (&d)[4], const unsigned (&a)[2], const unsigned (&b)[1]) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}

template <int ILP, int KIND>
__global__ void probe(float* out, long long* cycles, int iters) {
    float d[ILP][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int j = 0; j < ILP; ++j)
        for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (KIND == 8) mma_k8(d[j], a, b);
            else { unsigned a2[2] = {a[0], a[1]}; unsigned b1[1] = {b[0]}; mma_k4(d[j], a2, b1); }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < ILP; ++j)
        for (int i = 0; i < 4; ++i) s += d[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// FFMA reference: same structure, ILP independent chains
template <int ILP>
__global__ void probe_ffma(float* out, long long* cycles, int iters) {
    float d[ILP];
    float a = 1.0f + threadIdx.x * 1e-3f, b = 0.5f;
    for (int j = 0; j < ILP; ++j) d[j] = j;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) d[j] = fmaf(d[j], a, b);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < ILP; ++j) s += d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int warps, int ilp, int iters) {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    launch(out, cyc, warps * 32, iters);
    cudaDeviceSynchronize();
    launch(out, cyc, warps * 32, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0];
    double n = (double)iters * ilp * warps;  // instructions per SM (one CTA per SM)
    printf("%-10s warps/SM %2d ilp %d: %8.0f cycles, %.3f inst/clk/SM, %.2f clk per inst per warp (%s)\n", name, warps,
           ilp, c, n / c, c / ((double)iters * ilp), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

// Does the tensor core TRUNCATE or ROUND an fp32 operand of a .tf32 mma (its low 13 mantissa bits)?  One warp, A = x at
// [0][0], B = 1 at [0][0]; x has bits below the tf32 mantissa set so that rounding and truncation differ.
__global__ void probe_rounding(float* out) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const float xs[4] = {1.0f + 0x1.8p-11f, 1.0f + 0x1.0p-11f, 1.0f + 0x1.fp-11f, -(1.0f + 0x1.8p-11f)};
    for (int c = 0; c < 4; ++c) {
        unsigned a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
        if (g == 0 && t == 0) a[0] = __float_as_uint(xs[c]);   // A[0][0]
        if (g == 0 && t == 0) b[0] = __float_as_uint(1.0f);    // B[0][0]
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma_k8(d, a, b);
        if (lane == 0) { out[2 * c] = xs[c]; out[2 * c + 1] = d[0]; }
    }
}

int main() {
    {
        float* o; cudaMalloc(&o, 64); probe_rounding<<<1, 32>>>(o); float h[8]; cudaMemcpy(h, o, 32, cudaMemcpyDeviceToHost);
        for (int c = 0; c < 4; ++c)
            printf("tf32 operand %.10f (%a) -> product with 1.0 = %.10f (%a): %s\n", h[2 * c], h[2 * c], h[2 * c + 1],
                   h[2 * c + 1], h[2 * c + 1] == (float)((int)(h[2*c] * 1024.0f)) / 1024.0f ? "TRUNCATED" : "not truncated");
        cudaFree(o);
    }
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32}) {
        run("mma k8", [](float* o, long long* c, int th, int it) { probe<1, 8><<<148, th>>>(o, c, it); }, warps, 1, iters);
        run("mma k8", [](float* o, long long* c, int th, int it) { probe<4, 8><<<148, th>>>(o, c, it); }, warps, 4, iters);
        run("mma k4", [](float* o, long long* c, int th, int it) { probe<4, 4><<<148, th>>>(o, c, it); }, warps, 4, iters);
        run("ffma", [](float* o, long long* c, int th, int it) { probe_ffma<1><<<148, th>>>(o, c, it); }, warps, 1, iters);
        run("ffma", [](float* o, long long* c, int th, int it) { probe_ffma<8><<<148, th>>>(o, c, it); }, warps, 8, iters);
    }
    return 0;
}
