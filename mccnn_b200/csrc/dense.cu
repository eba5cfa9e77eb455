// dense.cu -- the dense glue between Monte-Carlo convolutions (utils/MCNetworkUtils.py:20-142 in the reference, where
// it is TensorFlow: tf.layers.batch_normalization (inference), tf.nn.relu, tf.matmul + bias).
//
//   scale_shift_act   y = act(x * scale + shift)                 batch_norm_RELU_drop_out at inference
//   dense_layer       y = act2( (act1(x * s1 + t1) @ W + b) * s2 + t2 )
//                     conv_1x1 and each layer of MLP_1_hidden / MLP_2_hidden with the surrounding inference
//                     batch-norms folded into per-column scale / shift vectors (any of s1/t1/s2/t2/b may be NULL)
// The matrices are tiny (<= 512 x 512) and the row counts small, so the layers are LATENCY bound, not FLOP bound; two
// FP32 FFMA kernels cover the two regimes with ONE launch per layer (instead of 2-6 library launches):
//   * dense_rows_kernel  (few rows, e.g. the 32-row classification head): one CTA per row, the 256 threads split
//     into (column, K-slice) pairs so that the K loop is short, W is read straight from L2 with coalesced loads, the
//     slices are reduced through shared memory;
//   * dense_tiles_kernel (many rows): 32 x 64 output tile per CTA, K staged through shared memory in chunks of 32
//     with the next chunk prefetched into registers while the current one is multiplied.
#include "common.cuh"

#include <algorithm>

namespace mc {
namespace {

constexpr int kDnBM = 32, kDnBN = 64, kDnBK = 32, kDnThreads = 256;
constexpr int kRowThreads = 1024;  // row kernels: the K loop of a layer is a chain of L2 round trips -- many short slices
constexpr int kRowsMaxK = 1024;  // dense_rows_kernel keeps one activated input row in shared memory

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

__device__ __forceinline__ float epilogue(float v, int c, const float* __restrict__ bias, const float* __restrict__ s2,
                                          const float* __restrict__ t2, int relu2) {
    if (bias != nullptr) v += bias[c];
    if (s2 != nullptr) v = fmaf(v, s2[c], t2[c]);
    return relu2 ? fmaxf(v, 0.0f) : v;
}

// one layer of one row: out[c] = epilogue(sum_k sA[k] W[k][c]); threads = (column c, K-slice s); `out` may be shared
// or global memory.  Ends with a barrier, so the caller may read `out` (shared) or overwrite sA right away.
__device__ __forceinline__ void row_layer(const float* sA, float* sPart, int K, int N, const float* __restrict__ W,
                                          const float* __restrict__ bias, const float* __restrict__ s2,
                                          const float* __restrict__ t2, int relu2, float* out) {
    for (int c0 = 0; c0 < N; c0 += kRowThreads) {
        const int cols = min(N - c0, kRowThreads);
        const int slices = kRowThreads / cols;
        const int c = threadIdx.x % cols, s = threadIdx.x / cols;
        float acc = 0.0f;
        if (s < slices) {
            const int kPer = (K + slices - 1) / slices;
            const int kBeg = s * kPer, kEnd = min(K, kBeg + kPer);
            const float* w = W + (size_t)kBeg * N + c0 + c;
            int k = kBeg;
            for (; k + 8 <= kEnd; k += 8) {  // 8 independent L2 loads in flight
                float wv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) wv[u] = w[(size_t)u * N];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = fmaf(sA[k + u], wv[u], acc);
                w += (size_t)8 * N;
            }
            for (; k < kEnd; ++k) {
                acc = fmaf(sA[k], *w, acc);
                w += N;
            }
        }
        sPart[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < cols) {
            float v = 0.0f;
            for (int q = 0; q < slices; ++q) v += sPart[q * cols + threadIdx.x];  // fixed order: deterministic
            out[c0 + threadIdx.x] = epilogue(v, c0 + threadIdx.x, bias, s2, t2, relu2);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void load_row(float* sA, int K, const float* __restrict__ x, const float* __restrict__ s1,
                                         const float* __restrict__ t1, int relu1) {
    for (int k = threadIdx.x; k < K; k += kRowThreads) {
        float v = x[k];
        if (s1 != nullptr) v = fmaf(v, s1[k], t1[k]);
        sA[k] = relu1 ? fmaxf(v, 0.0f) : v;
    }
    __syncthreads();
}

// one CTA per row, one layer
__global__ void __launch_bounds__(kRowThreads)
    dense_rows_kernel(const int* __restrict__ dRows, int capRows, int K, int N, const float* __restrict__ x,
                      const float* __restrict__ s1, const float* __restrict__ t1, int relu1,
                      const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ s2,
                      const float* __restrict__ t2, int relu2, float* __restrict__ y) {
    __shared__ float sA[kRowsMaxK];
    __shared__ float sPart[kRowThreads];
    const int rows = live_count(dRows, capRows);
    const int r = blockIdx.x;
    if (r >= rows) return;
    load_row(sA, K, x + (size_t)r * K, s1, t1, relu1);
    row_layer(sA, sPart, K, N, W, bias, s2, t2, relu2, y + (size_t)r * N);
}

// one CTA per row, a whole MLP (the classification heads MLP_1_hidden / MLP_2_hidden): the hidden activations stay
// in shared memory.  Same arithmetic, in the same order, as one dense_rows_kernel launch per layer.
constexpr int kMlpMaxLayers = 4;
struct MlpArgs {
    int layers;
    int dims[kMlpMaxLayers + 1];
    const float* W[kMlpMaxLayers];
    const float* bias[kMlpMaxLayers];
    const float* scale[kMlpMaxLayers];
    const float* shift[kMlpMaxLayers];
    int relu[kMlpMaxLayers];
};

__global__ void __launch_bounds__(kRowThreads)
    mlp_rows_kernel(const int* __restrict__ dRows, int capRows, const float* __restrict__ x,
                    const float* __restrict__ s0, const float* __restrict__ t0, int relu0, const MlpArgs a,
                    float* __restrict__ y) {
    __shared__ float sBuf[2][kRowsMaxK];
    __shared__ float sPart[kRowThreads];
    const int rows = live_count(dRows, capRows);
    const int r = blockIdx.x;
    if (r >= rows) return;
    load_row(sBuf[0], a.dims[0], x + (size_t)r * a.dims[0], s0, t0, relu0);
    for (int l = 0; l < a.layers; ++l) {
        const bool last = l == a.layers - 1;
        float* out = last ? y + (size_t)r * a.dims[l + 1] : sBuf[(l + 1) & 1];
        row_layer(sBuf[l & 1], sPart, a.dims[l], a.dims[l + 1], a.W[l], a.bias[l], a.scale[l], a.shift[l], a.relu[l],
                  out);
    }
}

// 32 x 64 output tile per CTA, register-prefetched K chunks
__global__ void __launch_bounds__(kDnThreads)
    dense_tiles_kernel(const int* __restrict__ dRows, int capRows, int K, int N, const float* __restrict__ x,
                       const float* __restrict__ s1, const float* __restrict__ t1, int relu1,
                       const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ s2,
                       const float* __restrict__ t2, int relu2, float* __restrict__ y) {
    __shared__ float sA[kDnBM][kDnBK + 1];
    __shared__ __align__(16) float sW[kDnBK][kDnBN];
    const int rows = live_count(dRows, capRows);
    const int row0 = blockIdx.x * kDnBM, col0 = blockIdx.y * kDnBN;
    if (row0 >= rows) return;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16 x 16 threads: 4 columns x 2 rows each
    // staging assignment: A tile 32x32 -> 4 elements per thread, W tile 32x64 -> 8 elements per thread
    const int aR = threadIdx.x / 8, aK = (threadIdx.x % 8) * 4;   // 4 consecutive k of one row
    const int wK = threadIdx.x / 8, wC = (threadIdx.x % 8) * 8;   // 8 consecutive columns of one k
    float ra[4], rw[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = row0 + aR, k = k0 + aK + u;
            float v = 0.0f;
            if (r < rows && k < K) {
                v = x[(size_t)r * K + k];
                if (s1 != nullptr) v = fmaf(v, s1[k], t1[k]);
                if (relu1) v = fmaxf(v, 0.0f);
            }
            ra[u] = v;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + wK, c = col0 + wC + u;
            rw[u] = (k < K && c < N) ? W[(size_t)k * N + c] : 0.0f;
        }
    };
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += kDnBK) {
#pragma unroll
        for (int u = 0; u < 4; ++u) sA[aR][aK + u] = ra[u];
#pragma unroll
        for (int u = 0; u < 8; ++u) sW[wK][wC + u] = rw[u];
        __syncthreads();
        if (k0 + kDnBK < K) fetch(k0 + kDnBK);  // next chunk's global loads overlap this chunk's FMAs
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
            if (c < N) y[(size_t)r * N + c] = epilogue(acc[i][j], c, bias, s2, t2, relu2);
        }
    }
}

// ---- few rows, wide layers (the 2048 -> 1024 -> 512 classification heads of MCClassH: 32 rows) ---------------------
// Such a layer is a stream of W through the GPU: 8 MB of weights against 32 x 2048 inputs.  One CTA per row (above)
// reads W once PER ROW and has 32 CTAs in flight; here every CTA owns (128 columns) x (a K segment) for ALL rows of a
// 32-row group -- a thread is one column with 32 accumulators, the activated input segment sits in shared memory
// transposed ([k][row], read as broadcast float4) -- so W crosses the memory system once and the (column slice, K
// segment) grid covers the GPU twice over.  The K segments are summed in a FIXED order by a second small launch
// (deterministic: no atomics), which also applies the epilogue; a layer small enough for one segment needs only the
// first launch.
constexpr int kFrCols = 128, kFrRows = 32, kFrMaxKS = 128, kFrStride = kFrRows + 4;  // segLen: multiple of 8

struct FewRowsPlan {
    int slices, segments, segLen;
};
FewRowsPlan fewrows_plan(int K, int N) {
    FewRowsPlan p;
    p.slices = cdiv(N, kFrCols);
    int S = (long long)K * N < (1 << 15) ? 1 : std::max(1, std::min(cdiv(2 * 148, p.slices), cdiv(K, 32)));
    p.segLen = std::min(kFrMaxKS, (cdiv(K, S) + 7) / 8 * 8);
    p.segments = cdiv(K, p.segLen);
    return p;
}

template <bool DIRECT>
__global__ void __launch_bounds__(kFrCols)
    dense_fewrows_kernel(const int* __restrict__ dRows, int capRows, int K, int N, int segLen,
                         const float* __restrict__ x, const float* __restrict__ s1, const float* __restrict__ t1,
                         int relu1, const float* __restrict__ W, const float* __restrict__ bias,
                         const float* __restrict__ s2, const float* __restrict__ t2, int relu2,
                         float* __restrict__ out) {
    __shared__ __align__(16) float xs[kFrMaxKS][kFrStride];
    const int r0 = blockIdx.z * kFrRows;
    const int k0 = blockIdx.y * segLen, kLen = min(segLen, K - k0);
    const int col = blockIdx.x * kFrCols + threadIdx.x;
    const bool active = col < N;
    // weights: chunks of 8 rows of W per thread, the next chunk's loads issued before this chunk's FMAs (the layer is
    // a stream of W out of HBM: what matters is bytes in flight); the first chunk is requested before the inputs are
    // staged; rows past the segment read as zero
    const float* w = W + (size_t)k0 * N + (active ? col : 0);
    constexpr int kChunk = 8;
    float cur[kChunk], nxt[kChunk];
#pragma unroll
    for (int u = 0; u < kChunk; ++u) cur[u] = (active && u < kLen) ? __ldg(w + (size_t)u * N) : 0.0f;
    const int rows = live_count(dRows, capRows);
    if (r0 >= rows) return;
    // activated input segment, transposed; rows past the live count are zero
    for (int e = threadIdx.x; e < kFrRows * segLen; e += kFrCols) {
        const int r = e / segLen, kk = e - r * segLen;
        float v = 0.0f;
        if (r0 + r < rows && kk < kLen) {
            v = x[(size_t)(r0 + r) * K + k0 + kk];
            if (s1 != nullptr) v = fmaf(v, s1[k0 + kk], t1[k0 + kk]);
            if (relu1) v = fmaxf(v, 0.0f);
        }
        xs[kk][r] = v;
    }
    __syncthreads();
    if (!active) return;
    float acc[kFrRows];
#pragma unroll
    for (int r = 0; r < kFrRows; ++r) acc[r] = 0.0f;
    for (int kk = 0; kk < kLen; kk += kChunk) {
#pragma unroll
        for (int u = 0; u < kChunk; ++u) {
            const int k = kk + kChunk + u;
            nxt[u] = k < kLen ? __ldg(w + (size_t)k * N) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kChunk; ++u) {
#pragma unroll
            for (int r4 = 0; r4 < kFrRows / 4; ++r4) {
                const float4 xv = *reinterpret_cast<const float4*>(&xs[kk + u][r4 * 4]);
                acc[r4 * 4] = fmaf(xv.x, cur[u], acc[r4 * 4]);
                acc[r4 * 4 + 1] = fmaf(xv.y, cur[u], acc[r4 * 4 + 1]);
                acc[r4 * 4 + 2] = fmaf(xv.z, cur[u], acc[r4 * 4 + 2]);
                acc[r4 * 4 + 3] = fmaf(xv.w, cur[u], acc[r4 * 4 + 3]);
            }
        }
#pragma unroll
        for (int u = 0; u < kChunk; ++u) cur[u] = nxt[u];
    }
    const int live = min(kFrRows, rows - r0);
    if (DIRECT) {
#pragma unroll
        for (int r = 0; r < kFrRows; ++r)
            if (r < live) out[(size_t)(r0 + r) * N + col] = epilogue(acc[r], col, bias, s2, t2, relu2);
    } else {  // partial sums of this K segment: out[segment][row][column]
        float* dst = out + ((size_t)blockIdx.y * capRows + r0) * N + col;
#pragma unroll
        for (int r = 0; r < kFrRows; ++r)
            if (r < live) dst[(size_t)r * N] = acc[r];
    }
}

__global__ void __launch_bounds__(256)
    dense_fewrows_reduce_kernel(const int* __restrict__ dRows, int capRows, int N, int segments,
                                const float* __restrict__ partial, const float* __restrict__ bias,
                                const float* __restrict__ s2, const float* __restrict__ t2, int relu2,
                                float* __restrict__ y) {
    const int rows = live_count(dRows, capRows);
    const long long total = (long long)rows * N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        float v = 0.0f;
        for (int q = 0; q < segments; ++q) v += __ldcg(partial + (size_t)q * capRows * N + t);  // fixed order
        y[t] = epilogue(v, (int)(t % N), bias, s2, t2, relu2);
    }
}

}  // namespace
}  // namespace mc

using namespace mc;

// floats of workspace mc_dense_layer_fewrows needs for a [numRows, K] x [K, N] layer (0: none)
extern "C" long long mc_dense_fewrows_workspace_floats(int numRows, int K, int N) {
    if (numRows <= 0 || K <= 0 || N <= 0) return 0;
    const FewRowsPlan p = fewrows_plan(K, N);
    return p.segments > 1 ? (long long)p.segments * numRows * N : 0;
}

// mc_dense_layer for FEW live rows (one per cloud: the classification heads, utils/MCNetworkUtils.py:20-107) and
// layers of any width: same arithmetic contract, W read once for all rows.  numRows is the capacity of x / y;
// workspace: mc_dense_fewrows_workspace_floats(numRows, K, N) floats.
extern "C" int mc_dense_layer_fewrows(int numRows, const int* dNumRows, int K, int N, const float* x, const float* s1,
                                      const float* t1, int relu1, const float* W, const float* bias, const float* s2,
                                      const float* t2, int relu2, float* workspace, float* y, void* stream) {
    MC_REQUIRE(K > 0 && N > 0, "mc_dense_layer_fewrows: layer sizes must be positive");
    MC_REQUIRE((s1 == nullptr) == (t1 == nullptr) && (s2 == nullptr) == (t2 == nullptr),
               "mc_dense_layer_fewrows: scale and shift come in pairs");
    if (numRows <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const FewRowsPlan p = fewrows_plan(K, N);
    MC_REQUIRE(p.segments == 1 || workspace != nullptr, "mc_dense_layer_fewrows: workspace required");
    const dim3 grid(p.slices, p.segments, cdiv(numRows, kFrRows));
    if (p.segments == 1) {
        dense_fewrows_kernel<true><<<grid, kFrCols, 0, s>>>(dNumRows, numRows, K, N, p.segLen, x, s1, t1, relu1, W, bias,
                                                            s2, t2, relu2, y);
        return check_launch("mc_dense_layer_fewrows");
    }
    dense_fewrows_kernel<false><<<grid, kFrCols, 0, s>>>(dNumRows, numRows, K, N, p.segLen, x, s1, t1, relu1, W, bias, s2,
                                                         t2, relu2, workspace);
    const int rgrid = std::max(1, std::min(cdiv((long long)numRows * N, 256), 148 * 4));
    dense_fewrows_reduce_kernel<<<rgrid, 256, 0, s>>>(dNumRows, numRows, N, p.segments, workspace, bias, s2, t2, relu2,
                                                      y);
    count_launches(1);
    return check_launch("mc_dense_layer_fewrows");
}

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
// utils/MCNetworkUtils.py:44,123), y [rows, N].  s1/t1, bias, s2/t2 optional (NULL).  `numRowsHint` (<= 0: unknown)
// is the caller's estimate of the LIVE row count; it only selects the kernel variant.
extern "C" int mc_dense_layer(int numRows, const int* dNumRows, int numRowsHint, int K, int N, const float* x,
                              const float* s1, const float* t1, int relu1, const float* W, const float* bias,
                              const float* s2, const float* t2, int relu2, float* y, void* stream) {
    MC_REQUIRE(K > 0 && N > 0, "mc_dense_layer: layer sizes must be positive");
    MC_REQUIRE((s1 == nullptr) == (t1 == nullptr) && (s2 == nullptr) == (t2 == nullptr),
               "mc_dense_layer: scale and shift come in pairs");
    if (numRows <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int live = (numRowsHint > 0 && numRowsHint < numRows) ? numRowsHint : numRows;
    if (live <= 128 && K <= kRowsMaxK) {
        dense_rows_kernel<<<numRows, kRowThreads, 0, s>>>(dNumRows, numRows, K, N, x, s1, t1, relu1, W, bias, s2, t2,
                                                          relu2, y);
    } else {
        dim3 grid(cdiv(numRows, kDnBM), cdiv(N, kDnBN));
        dense_tiles_kernel<<<grid, kDnThreads, 0, s>>>(dNumRows, numRows, K, N, x, s1, t1, relu1, W, bias, s2, t2, relu2,
                                                       y);
    }
    return check_launch("mc_dense_layer");
}

// A whole MLP on few rows in ONE launch (MLP_1_hidden / MLP_2_hidden, utils/MCNetworkUtils.py:20-107):
//   a_0 = act0(x * s0 + t0);   a_{l+1} = act_l((a_l @ W[l] + bias[l]) * scale[l] + shift[l]);   y = a_numLayers
// dims: numLayers + 1 host ints (input width, then every layer's output width, each <= 1024); W / bias / scale / shift:
// host arrays of numLayers device pointers (bias, scale+shift may be NULL per layer); relu: numLayers host flags.
extern "C" int mc_mlp_rows(int numRows, const int* dNumRows, int numLayers, const int* dims, const float* x,
                           const float* s0, const float* t0, int relu0, const float* const* W,
                           const float* const* bias, const float* const* scale, const float* const* shift,
                           const int* relu, float* y, void* stream) {
    MC_REQUIRE(numLayers >= 1 && numLayers <= kMlpMaxLayers, "mc_mlp_rows: between 1 and 4 layers");
    MC_REQUIRE((s0 == nullptr) == (t0 == nullptr), "mc_mlp_rows: scale and shift come in pairs");
    MlpArgs a;
    a.layers = numLayers;
    for (int l = 0; l <= numLayers; ++l) {
        MC_REQUIRE(dims[l] > 0 && dims[l] <= kRowsMaxK, "mc_mlp_rows: layer widths must be in 1..1024");
        a.dims[l] = dims[l];
    }
    for (int l = 0; l < numLayers; ++l) {
        MC_REQUIRE(W[l] != nullptr, "mc_mlp_rows: weights required");
        MC_REQUIRE((scale[l] == nullptr) == (shift[l] == nullptr), "mc_mlp_rows: scale and shift come in pairs");
        a.W[l] = W[l];
        a.bias[l] = bias[l];
        a.scale[l] = scale[l];
        a.shift[l] = shift[l];
        a.relu[l] = relu[l];
    }
    if (numRows <= 0) return 0;
    mlp_rows_kernel<<<numRows, kRowThreads, 0, (cudaStream_t)stream>>>(dNumRows, numRows, x, s0, t0, relu0, a, y);
    return check_launch("mc_mlp_rows");
}
