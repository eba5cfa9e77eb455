// dense.cu -- the dense glue between Monte-Carlo convolutions (utils/MCNetworkUtils.py:20-142 in the reference, where
// it is TensorFlow: tf.layers.batch_normalization (inference), tf.nn.relu, tf.matmul + bias).
//
// Two kernels:
//   scale_shift_act   y = act(x * scale + shift)                 batch_norm_RELU_drop_out at inference
//   dense_layer       y = act2( (act1(x * s1 + t1) @ W + b) * s2 + t2 )
//                     conv_1x1 and each layer of MLP_1_hidden / MLP_2_hidden with the surrounding inference
//                     batch-norms folded into per-column scale / shift vectors (any of s1/t1/s2/t2/b may be NULL)
// The matrices are tiny (<= 512 x 512) and the row counts small (tens to thousands), so this is a plain register-
// tiled FP32 FFMA GEMM (32 x 64 output tile per CTA, K staged through shared memory in chunks of 32); what matters is
// ONE launch per layer instead of the 2-6 library launches of the eager formulation.  FP32 accumulate in k order.
#include "common.cuh"

namespace mc {
namespace {

constexpr int kDnBM = 32, kDnBN = 64, kDnBK = 32, kDnThreads = 256;

__global__ void __launch_bounds__(256) scale_shift_act_kernel(const int* __restrict__ dRows, int capRows, int cols,
                                                              const float* __restrict__ x,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int relu,
                                                              float* __restrict__ y) {
    const int rows = live_count(dRows, capRows);
    const long long total = (long long)rows * cols;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % cols);
        float v = fmaf(x[t], scale[c], shift[c]);
        y[t] = relu ? fmaxf(v, 0.0f) : v;
    }
}

__global__ void __launch_bounds__(kDnThreads)
    dense_layer_kernel(const int* __restrict__ dRows, int capRows, int K, int N, const float* __restrict__ x,
                       const float* __restrict__ s1, const float* __restrict__ t1, int relu1,
                       const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ s2,
                       const float* __restrict__ t2, int relu2, float* __restrict__ y) {
    __shared__ float sA[kDnBM][kDnBK + 1];
    __shared__ float sW[kDnBK][kDnBN];
    const int rows = live_count(dRows, capRows);
    const int row0 = blockIdx.x * kDnBM, col0 = blockIdx.y * kDnBN;
    if (row0 >= rows) return;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16 x 16 threads: 4 columns x 2 rows each
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int k0 = 0; k0 < K; k0 += kDnBK) {
        for (int e = threadIdx.x; e < kDnBM * kDnBK; e += kDnThreads) {  // A tile with the input affine + activation
            const int r = e / kDnBK, k = e % kDnBK;
            float v = 0.0f;
            if (row0 + r < rows && k0 + k < K) {
                v = x[(size_t)(row0 + r) * K + k0 + k];
                if (s1 != nullptr) v = fmaf(v, s1[k0 + k], t1[k0 + k]);
                if (relu1) v = fmaxf(v, 0.0f);
            }
            sA[r][k] = v;
        }
        for (int e = threadIdx.x; e < kDnBK * kDnBN; e += kDnThreads) {
            const int k = e / kDnBN, c = e % kDnBN;
            sW[k][c] = (k0 + k < K && col0 + c < N) ? W[(size_t)(k0 + k) * N + col0 + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kDnBK; ++k) {
            const float a0 = sA[ty * 2][k], a1 = sA[ty * 2 + 1][k];
            const float4 w = *reinterpret_cast<const float4*>(&sW[k][tx * 4]);
            acc[0][0] = fmaf(a0, w.x, acc[0][0]);
            acc[0][1] = fmaf(a0, w.y, acc[0][1]);
            acc[0][2] = fmaf(a0, w.z, acc[0][2]);
            acc[0][3] = fmaf(a0, w.w, acc[0][3]);
            acc[1][0] = fmaf(a1, w.x, acc[1][0]);
            acc[1][1] = fmaf(a1, w.y, acc[1][1]);
            acc[1][2] = fmaf(a1, w.z, acc[1][2]);
            acc[1][3] = fmaf(a1, w.w, acc[1][3]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = row0 + ty * 2 + i;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + tx * 4 + j;
            if (c >= N) continue;
            float v = acc[i][j];
            if (bias != nullptr) v += bias[c];
            if (s2 != nullptr) v = fmaf(v, s2[c], t2[c]);
            if (relu2) v = fmaxf(v, 0.0f);
            y[(size_t)r * N + c] = v;
        }
    }
}

}  // namespace
}  // namespace mc

using namespace mc;

// y[r, c] = act(x[r, c] * scale[c] + shift[c]) for the live rows -- inference batch norm (+ ReLU),
// utils/MCNetworkUtils.py:128-142 with the four statistics folded into (scale, shift) by the caller.
extern "C" int mc_scale_shift_act(int numRows, const int* dNumRows, int numCols, const float* x, const float* scale,
                                  const float* shift, int relu, float* y, void* stream) {
    MC_REQUIRE(numCols > 0, "mc_scale_shift_act: column count must be positive");
    if (numRows <= 0) return 0;
    long long total = (long long)numRows * numCols;
    int grid = cdiv(total, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    scale_shift_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dNumRows, numRows, numCols, x, scale, shift, relu, y);
    return check_launch("mc_scale_shift_act");
}

// y = act2((act1(x * s1 + t1) @ W + bias) * s2 + t2); x [rows, K], W [K, N] row-major (the reference's weight layout,
// utils/MCNetworkUtils.py:44,123), y [rows, N].  s1/t1, bias, s2/t2 optional (NULL).
extern "C" int mc_dense_layer(int numRows, const int* dNumRows, int K, int N, const float* x, const float* s1,
                              const float* t1, int relu1, const float* W, const float* bias, const float* s2,
                              const float* t2, int relu2, float* y, void* stream) {
    MC_REQUIRE(K > 0 && N > 0, "mc_dense_layer: layer sizes must be positive");
    MC_REQUIRE((s1 == nullptr) == (t1 == nullptr) && (s2 == nullptr) == (t2 == nullptr),
               "mc_dense_layer: scale and shift come in pairs");
    if (numRows <= 0) return 0;
    dim3 grid(cdiv(numRows, kDnBM), cdiv(N, kDnBN));
    dense_layer_kernel<<<grid, kDnThreads, 0, (cudaStream_t)stream>>>(dNumRows, numRows, K, N, x, s1, t1, relu1, W, bias,
                                                                    s2, t2, relu2, y);
    return check_launch("mc_dense_layer");
}
