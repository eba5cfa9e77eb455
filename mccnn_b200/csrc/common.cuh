// common.cuh -- shared device helpers for the sm_100a Monte-Carlo-convolution kernels.
//
// The geometry helpers reproduce, operation for operation, the float arithmetic the reference's
// kernels compile to (tf_ops/sort_gpu.cu:50-59, tf_ops/find_neighbors.cu:67-97): every index-valued
// result downstream (cell keys, neighbour lists, Poisson samples) depends on these bits.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#define MC_BLOCK_MLP 8  // BLOCK_MLP_SIZE of the reference build (tf_ops/genCompileScript.py:20)
#define MC_FULL_MASK 0xffffffffu

namespace mc {

// Every dynamic extent is passed as (device pointer, host capacity): kernels are launched over the
// capacity and read the live count on the device, so no host sync is needed between ops.
__device__ __forceinline__ int live_count(const int* __restrict__ dCount, int capacity) {
    if (dCount == nullptr) return capacity;
    int v = *dCount;
    return v < capacity ? v : capacity;
}

__device__ __forceinline__ float max_extent(const float* __restrict__ mn, const float* __restrict__ mx, int b) {
    float ex = mx[b * 3] - mn[b * 3];
    float ey = mx[b * 3 + 1] - mn[b * 3 + 1];
    float ez = mx[b * 3 + 2] - mn[b * 3 + 2];
    return fmaxf(fmaxf(ex, ey), ez);
}

// max(min((int)floor(v), n-1), 0); __float2int_rd saturates and maps NaN to 0 like the reference's cast.
__device__ __forceinline__ int clamp_cell(float v, int n) {
    int c = __float2int_rd(v);
    c = c < n - 1 ? c : n - 1;
    return c > 0 ? c : 0;
}

struct CloudFrame {  // per-cloud constants every geometric kernel needs
    float mnx, mny, mnz;
    float cell;    // extent / n
    float radius;  // scaled radius
};

__device__ __forceinline__ CloudFrame cloud_frame(const float* __restrict__ mn, const float* __restrict__ mx,
                                                  int b, int n, float radius, bool scaleInv) {
    CloudFrame f;
    float ext = max_extent(mn, mx, b);
    f.mnx = mn[b * 3];
    f.mny = mn[b * 3 + 1];
    f.mnz = mn[b * 3 + 2];
    f.cell = __fdiv_rn(ext, (float)n);
    f.radius = scaleInv ? __fmul_rn(radius, ext) : radius;
    return f;
}

__device__ __forceinline__ void cell_of(const CloudFrame& f, float x, float y, float z, int n, int& cx, int& cy,
                                        int& cz) {
    cx = clamp_cell(__fdiv_rn(__fsub_rn(x, f.mnx), f.cell), n);
    cy = clamp_cell(__fdiv_rn(__fsub_rn(y, f.mny), f.cell), n);
    cz = clamp_cell(__fdiv_rn(__fsub_rn(z, f.mnz), f.cell), n);
}

// sqrt(dx*dx + dy*dy + dz*dz) < r exactly as the reference's SASS evaluates it:
// FMUL dy*dy ; FFMA dx*dx+. ; FFMA dz*dz+. ; IEEE sqrt ; FSETP.LT   (see DESIGN.md "float-exact predicates")
__device__ __forceinline__ bool within_radius(float px, float py, float pz, float qx, float qy, float qz, float r) {
    float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    return __fsqrt_rn(d2) < r;
}

// The same predicate without the square root in the inner loop.  sqrt_rn is monotone, so
// { d2 >= 0 : sqrt_rn(d2) < r } is a down-set with a largest element t*; exact_radius_threshold finds t* by walking
// ulp by ulp from r*r (at most a few steps), after which   d2 <= t*  <=>  sqrt_rn(d2) < r   for EVERY float d2
// (NaN compares false on both sides; r <= 0 gives t* = -1: no match).  Bit-exact replacement, ~10 SASS instructions
// (MUFU.RSQ + Newton + fix-up) shorter per candidate.
__device__ __forceinline__ float exact_radius_threshold(float r) {
    if (!(r > 0.0f)) return -1.0f;
    if (r > 1.8e19f) return FLT_MAX;  // r*r overflows: every finite d2 qualifies
    float t = __fmul_rn(r, r);
    if (__fsqrt_rn(t) < r) {
        while (true) {
            float up = __int_as_float(__float_as_int(t) + 1);
            if (!(__fsqrt_rn(up) < r)) break;
            t = up;
        }
    } else {
        while (!(__fsqrt_rn(t) < r)) {
            if (t == 0.0f) return -1.0f;
            t = __int_as_float(__float_as_int(t) - 1);
        }
    }
    return t;
}

__device__ __forceinline__ float dist2_ref(float px, float py, float pz, float qx, float qy, float qz) {
    float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// One term of the kernel density estimate (compute_pdf): exp(-0.5 |d|^2 / (r h)^2) = 2^(|d|^2 e2) with
// e2 = -0.5 log2(e) / (r h)^2, evaluated on the SFU.  Shared by pdf_kernel and the fused neighbour + pdf kernel so that
// both produce the same bits (explicit contraction: no dependence on the compiler's choice per kernel).
__device__ __forceinline__ float pdf_term(float dx, float dy, float dz, float e2) {
    const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__fmul_rn(d2, e2)));
    return y;
}
__device__ __forceinline__ float pdf_exponent_scale(float r, float window) {
    const float invRadH = 1.0f / (r * window);
    return -0.72134752044448170368f * invRadH * invRadH;
}
__device__ __forceinline__ float pdf_finish(float acc, float window, int k) {
    const float g = (1.0f / window) * 0.39894228f;
    return (acc * (g * g * g)) / (float)k;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(MC_FULL_MASK, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

// 27-cell neighbourhood tables.  Order is observable behaviour:
//  kOffsNeigh: tf_ops/find_neighbors.cu:282-291 ; kOffsPool: tf_ops/poisson_sampling.cu:192-196
__device__ __constant__ const signed char kOffsNeigh[27][3] = {
    {1, 1, 1},   {0, 1, 1},   {-1, 1, 1},  {1, 0, 1},   {0, 0, 1},   {-1, 0, 1},  {1, -1, 1},
    {0, -1, 1},  {-1, -1, 1}, {1, 1, 0},   {0, 1, 0},   {-1, 1, 0},  {1, 0, 0},   {0, 0, 0},
    {-1, 0, 0},  {1, -1, 0},  {0, -1, 0},  {-1, -1, 0}, {1, 1, -1},  {0, 1, -1},  {-1, 1, -1},
    {1, 0, -1},  {0, 0, -1},  {-1, 0, -1}, {1, -1, -1}, {0, -1, -1}, {-1, -1, -1}};

__device__ __constant__ const signed char kOffsPool[27][3] = {
    {1, 1, -1},  {0, -1, 1},  {0, 1, 1},  {0, 1, 0}, {0, 0, 1},  {0, -1, 0},  {-1, 1, -1},
    {0, -1, -1}, {1, 0, 0},   {1, -1, 1}, {1, 0, 1}, {-1, 1, 1}, {-1, 0, 0},  {1, -1, -1},
    {0, 1, -1},  {-1, -1, 0}, {-1, 1, 0}, {0, 0, 0}, {0, 0, -1}, {1, 1, 0},   {1, 0, -1},
    {1, -1, 0},  {-1, 0, 1},  {1, 1, 1},  {-1, 0, -1}, {-1, -1, -1}, {-1, -1, 1}};

}  // namespace mc

// ---- host side -------------------------------------------------------------------------------
namespace mc {
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaPeekAtLastError -> status + message; counts one kernel launch
void count_launches(int n);          // additional kernel launches of the calling entry point
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
// pdf.cu: compute_pdf over a pair list whose batch ids are the samples' (used by mc_find_neighbors_pdf_fused)
int launch_pdf_by_sample(bool scaleInv, float window, float radius, const int* dM, int capM, const int* dK, int capK,
                         const float* pts, const int* sampleBids, const float* mn, const float* mx, const int* start,
                         const int* pairs, float* pdf, cudaStream_t s);
}  // namespace mc

#define MC_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            mc::set_error(__VA_ARGS__); \
            return 1;                  \
        }                              \
    } while (0)
