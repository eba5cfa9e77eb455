// grid_sort.cu -- counting sort of a packed batch into per-cloud uniform grids, the cell (start,end) table,
// feature row permutations and index transforms.
//
// Replaces tf_ops/sort_gpu.cu (calc_key, update_counters, propagate_offsets, determine_new_index, move_points,
// save_indexs, sort_features_back[_grad], compute_inverse_indexs, transform_indexs, determine_cell_size).
// Differences that matter:
//   * the destination index is DETERMINISTIC and stable (ascending original index inside a cell).  The reference
//     takes the atomic arrival order (sort_gpu.cu:170), i.e. any within-cell permutation; the stable one is the
//     canonical member of that set (oracle/mccnn_oracle.c orc_sort_step1).  It is obtained without a multi-pass
//     radix sort: points are dropped into their cell segment in arrival order, then each point counts how many
//     members of its (small) segment have a lower original index.
//   * key + histogram are one kernel; the scan is the shared two-level scan (scan.cuh); all extents may live on
//     the device (no host sync), all launches go to the caller's stream.
#include "common.cuh"
#include "scan.cuh"

namespace mc {
namespace {

constexpr int kPtThreads = 256;

// keys[i], arrival[i] = atomic arrival rank inside the cell, counters[key] = cell population
__global__ void __launch_bounds__(kPtThreads) key_hist_kernel(const int* __restrict__ dN, int cap, int B, int n,
                                                              const float* __restrict__ mn,
                                                              const float* __restrict__ mx,
                                                              const float* __restrict__ pts,
                                                              const int* __restrict__ bids, int* __restrict__ keys,
                                                              int* __restrict__ arrival, int* __restrict__ counters) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int b = bids[i];
    b = (unsigned)b < (unsigned)B ? b : 0;
    CloudFrame f = cloud_frame(mn, mx, b, n, 0.f, false);
    int cx, cy, cz;
    cell_of(f, pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2], n, cx, cy, cz);
    int key = ((b * n + cx) * n + cy) * n + cz;
    keys[i] = key;
    arrival[i] = atomicAdd(&counters[key], 1);
}

__device__ __forceinline__ int cell_begin(const int* __restrict__ local, const int* __restrict__ tileOffs, int key) {
    return scan_prefix(local, tileOffs, key);
}

__global__ void __launch_bounds__(kPtThreads) scatter_perm_kernel(const int* __restrict__ dN, int cap,
                                                                  const int* __restrict__ keys,
                                                                  const int* __restrict__ arrival,
                                                                  const int* __restrict__ local,
                                                                  const int* __restrict__ tileOffs,
                                                                  int* __restrict__ perm) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    perm[cell_begin(local, tileOffs, keys[i]) + arrival[i]] = i;
}

// newIdx[i] = cellBegin + #{members of the cell segment with original index < i}
__global__ void __launch_bounds__(kPtThreads) stable_rank_kernel(const int* __restrict__ dN, int cap, int G,
                                                                 const int* __restrict__ keys,
                                                                 const int* __restrict__ local,
                                                                 const int* __restrict__ tileOffs,
                                                                 const int* __restrict__ perm,
                                                                 int* __restrict__ newIdx) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int key = keys[i];
    int beg = cell_begin(local, tileOffs, key);
    int end = (key + 1 < G) ? cell_begin(local, tileOffs, key + 1) : N;
    int r = 0;
    for (int p = beg; p < end; ++p) r += (perm[p] < i);
    newIdx[i] = beg + r;
}

// step 2: scatter coordinates, batch id and key to the sorted position
__global__ void __launch_bounds__(kPtThreads) move_points_kernel(const int* __restrict__ dN, int cap,
                                                                 const float* __restrict__ pts,
                                                                 const int* __restrict__ bids,
                                                                 const int* __restrict__ keys,
                                                                 const int* __restrict__ newIdx,
                                                                 float* __restrict__ outPts,
                                                                 int* __restrict__ outBids,
                                                                 int* __restrict__ sortedKeys) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int d = newIdx[i];
    outPts[(size_t)d * 3] = pts[(size_t)i * 3];
    outPts[(size_t)d * 3 + 1] = pts[(size_t)i * 3 + 1];
    outPts[(size_t)d * 3 + 2] = pts[(size_t)i * 3 + 2];
    outBids[d] = bids[i];
    sortedKeys[d] = keys[i];
}

// cell table from key boundaries in sorted order (table pre-zeroed: empty cell = (0,0), sort_gpu.cu:225-248,492)
__global__ void __launch_bounds__(kPtThreads) cell_table_kernel(const int* __restrict__ dN, int cap,
                                                                const int* __restrict__ sortedKeys,
                                                                int* __restrict__ cellIdx) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int k = sortedKeys[i];
    if (i == 0 || sortedKeys[i - 1] != k) cellIdx[(size_t)k * 2] = i;
    if (i == N - 1 || sortedKeys[i + 1] != k) cellIdx[(size_t)k * 2 + 1] = i + 1;
}

// out[SCATTER ? idx[r] : r, :] = in[SCATTER ? r : idx[r], :]; VEC floats per thread
template <bool SCATTER, typename VecT>
__global__ void __launch_bounds__(256) permute_rows_kernel(const int* __restrict__ dRows, int capRows, int rowVecs,
                                                           const VecT* __restrict__ in, const int* __restrict__ idx,
                                                           VecT* __restrict__ out) {
    const int rows = live_count(dRows, capRows);
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long r = t / rowVecs;
    if (r >= rows) return;
    int c = (int)(t - r * rowVecs);
    int o = idx[r];
    if (SCATTER)
        out[(size_t)o * rowVecs + c] = in[(size_t)r * rowVecs + c];
    else
        out[(size_t)r * rowVecs + c] = in[(size_t)o * rowVecs + c];
}

template <bool SCATTER>
void launch_permute_rows(const int* dRows, int capRows, int C, const float* in, const int* idx, float* out,
                         cudaStream_t s) {
    if (capRows <= 0 || C <= 0) return;
    bool vec4 = (C % 4 == 0) && (((uintptr_t)in | (uintptr_t)out) % 16 == 0);
    if (vec4) {
        long long total = (long long)capRows * (C / 4);
        permute_rows_kernel<SCATTER, float4><<<cdiv(total, 256), 256, 0, s>>>(
            dRows, capRows, C / 4, reinterpret_cast<const float4*>(in), idx, reinterpret_cast<float4*>(out));
    } else {
        long long total = (long long)capRows * C;
        permute_rows_kernel<SCATTER, float><<<cdiv(total, 256), 256, 0, s>>>(dRows, capRows, C, in, idx, out);
    }
}

// out[idx[r], :] = act(in[r, :] * scale + shift): the grid-order scatter of sort_features with a pending inference
// batch norm (+ ReLU) folded in, so the normalised features are never written in point order
template <typename VecT>
__global__ void __launch_bounds__(256) permute_rows_affine_kernel(const int* __restrict__ dRows, int capRows,
                                                                  int rowVecs, const VecT* __restrict__ in,
                                                                  const int* __restrict__ idx,
                                                                  const VecT* __restrict__ scale,
                                                                  const VecT* __restrict__ shift, int relu,
                                                                  VecT* __restrict__ out) {
    const int rows = live_count(dRows, capRows);
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long r = t / rowVecs;
    if (r >= rows) return;
    int c = (int)(t - r * rowVecs);
    VecT v = in[(size_t)r * rowVecs + c];
    const VecT a = scale[c], b = shift[c];
    float* pv = reinterpret_cast<float*>(&v);
    const float* pa = reinterpret_cast<const float*>(&a);
    const float* pb = reinterpret_cast<const float*>(&b);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(VecT) / sizeof(float)); ++k) {
        float y = fmaf(pv[k], pa[k], pb[k]);
        pv[k] = relu ? fmaxf(y, 0.f) : y;
    }
    out[(size_t)idx[r] * rowVecs + c] = v;
}

__global__ void __launch_bounds__(kPtThreads) inverse_index_kernel(const int* __restrict__ dN, int cap,
                                                                   const int* __restrict__ newIdx,
                                                                   int* __restrict__ inv) {
    const int N = live_count(dN, cap);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) inv[newIdx[i]] = i;
}

__global__ void __launch_bounds__(kPtThreads) gather_index_kernel(const int* __restrict__ dM, int cap,
                                                                  const int* __restrict__ idx,
                                                                  const int* __restrict__ table,
                                                                  int* __restrict__ out) {
    const int M = live_count(dM, cap);
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < M) out[s] = table[idx[s]];
}

__global__ void num_cells_kernel(float cellSize, const float* __restrict__ mn, const float* __restrict__ mx,
                                 int* __restrict__ out) {
    int n = (int)__fdiv_rn(max_extent(mn, mx, 0), cellSize);  // row 0 only: sort_gpu.cu:381-389
    *out = n == 0 ? 1 : n;
}

// planned mode: the grid resolution was fixed when the plan was recorded; raise the flag if this batch's box asks for
// a different one (the caller then re-records, exactly as for an exceeded capacity)
__global__ void verify_num_cells_kernel(float cellSize, const float* __restrict__ mn, const float* __restrict__ mx,
                                        int expected, int* __restrict__ flag) {
    int n = (int)__fdiv_rn(max_extent(mn, mx, 0), cellSize);
    if ((n == 0 ? 1 : n) != expected) atomicOr(flag, 1);
}

}  // namespace
}  // namespace mc

using namespace mc;

// Sync-free companion of mc_determine_num_cells for absolute radii (scaleInv == 0): checks on the device that the
// number of cells this batch's box implies (tf_ops/sort_gpu.cu:381-389) equals `expectedNumCells` and ORs 1 into
// *dMismatchFlag otherwise.  Lets a captured forward reuse the grid resolution it was planned with.
extern "C" int mc_verify_num_cells(float cellSize, const float* aabbMin, const float* aabbMax, int expectedNumCells,
                                   int* dMismatchFlag, void* stream) {
    MC_REQUIRE(cellSize > 0.0f, "mc_verify_num_cells: cell size must be positive");
    MC_REQUIRE(dMismatchFlag != nullptr, "mc_verify_num_cells: flag required");
    verify_num_cells_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(cellSize, aabbMin, aabbMax, expectedNumCells,
                                                               dMismatchFlag);
    count_launches(1);
    return check_launch("mc_verify_num_cells");
}

// determineNumCells, tf_ops/sort_gpu.cu:397-420.  scale_inv: host arithmetic, no device work.  Otherwise the
// count depends on row 0 of the box: one tiny kernel + one 4-byte D2H (the only sync; callers cache the result).
extern "C" int mc_determine_num_cells(int scaleInv, float cellSize, const float* aabbMin, const float* aabbMax,
                                      int* workspace, int* outNumCells, void* stream) {
    MC_REQUIRE(cellSize > 0.0f, "mc_determine_num_cells: cell size must be positive");
    if (scaleInv) {
        int n = (int)(1.0f / cellSize);
        *outNumCells = n == 0 ? 1 : n;
        return 0;
    }
    MC_REQUIRE(workspace != nullptr, "mc_determine_num_cells: workspace (1 int) required");
    cudaStream_t s = (cudaStream_t)stream;
    num_cells_kernel<<<1, 1, 0, s>>>(cellSize, aabbMin, aabbMax, workspace);
    cudaError_t e = cudaMemcpyAsync(outNumCells, workspace, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        set_error("mc_determine_num_cells: %s", cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

// ints of workspace for step 1: counters[G] | tileSums | arrival[N] | perm[N]
extern "C" long long mc_sort_step1_workspace_ints(int numPoints, int batchSize, int numCells) {
    long long G = (long long)batchSize * numCells * numCells * numCells;
    return G + scan_num_tiles((int)G) + 2LL * (numPoints > 0 ? numPoints : 1) + 8;
}

// sortPointsStep1GPUKernel, tf_ops/sort_gpu.cu:436-471
extern "C" int mc_sort_points_step1(int numPoints, const int* dNumPoints, int batchSize, int numCells,
                                    const float* aabbMin, const float* aabbMax, const float* pts,
                                    const int* batchIds, int* workspace, int* keys, int* newIdx, void* stream) {
    MC_REQUIRE(batchSize > 0 && numCells > 0, "mc_sort_points_step1: batch size and cell count must be positive");
    long long G64 = (long long)batchSize * numCells * numCells * numCells;
    MC_REQUIRE(G64 < (1LL << 30), "mc_sort_points_step1: grid of %lld cells is too large", G64);
    if (numPoints <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int G = (int)G64;
    int* counters = workspace;
    int* tileSums = counters + G;
    int* arrival = tileSums + scan_num_tiles(G);
    int* perm = arrival + numPoints;
    cudaMemsetAsync(counters, 0, sizeof(int) * (size_t)G, s);
    int grid = cdiv(numPoints, kPtThreads);
    key_hist_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, batchSize, numCells, aabbMin, aabbMax, pts,
                                                batchIds, keys, arrival, counters);
    launch_scan(counters, nullptr, G, tileSums, nullptr, s);
    scatter_perm_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, keys, arrival, counters, tileSums, perm);
    stable_rank_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, G, keys, counters, tileSums, perm, newIdx);
    count_launches(4);
    return check_launch("mc_sort_points_step1");
}

// sortPointsStep2GPUKernel, tf_ops/sort_gpu.cu:473-497.  workspace: numPoints ints (sorted keys).
extern "C" int mc_sort_points_step2(int numPoints, const int* dNumPoints, int batchSize, int numFeatures,
                                    int numCells, const float* pts, const int* batchIds, const float* features,
                                    const int* keys, const int* newIdx, int* workspace, float* outPts,
                                    int* outBatchIds, float* outFeatures, int* outCellIdx, void* stream) {
    MC_REQUIRE(batchSize > 0 && numCells > 0, "mc_sort_points_step2: batch size and cell count must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    size_t G = (size_t)batchSize * numCells * numCells * numCells;
    cudaMemsetAsync(outCellIdx, 0, sizeof(int) * 2 * G, s);
    if (numPoints <= 0) return check_launch("mc_sort_points_step2");
    int grid = cdiv(numPoints, kPtThreads);
    move_points_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, pts, batchIds, keys, newIdx, outPts,
                                                   outBatchIds, workspace);
    launch_permute_rows<true>(dNumPoints, numPoints, numFeatures, features, newIdx, outFeatures, s);
    cell_table_kernel<<<grid, kPtThreads, 0, s>>>(dNumPoints, numPoints, workspace, outCellIdx);
    count_launches(numFeatures > 0 ? 2 : 1);
    return check_launch("mc_sort_points_step2");
}

// sortFeaturesBackGrad (python sort_features), tf_ops/sort_gpu.cu:529-540: out[newIdx[i],:] = in[i,:]
extern "C" int mc_sort_features(int numPoints, const int* dNumPoints, int numFeatures, const float* in,
                                const int* newIdx, float* out, void* stream) {
    launch_permute_rows<true>(dNumPoints, numPoints, numFeatures, in, newIdx, out, (cudaStream_t)stream);
    return check_launch("mc_sort_features");
}

// sort_features of act(in * scale + shift) (scale, shift: numFeatures floats): the inference batch norm + ReLU that
// precedes a convolution (utils/MCNetworkUtils.py:128-142) folded into the scatter of tf_ops/sort_gpu.cu:529-540
extern "C" int mc_sort_features_affine(int numPoints, const int* dNumPoints, int numFeatures, const float* in,
                                       const int* newIdx, const float* scale, const float* shift, int relu,
                                       float* out, void* stream) {
    MC_REQUIRE(scale != nullptr && shift != nullptr, "mc_sort_features_affine: scale and shift are required");
    cudaStream_t s = (cudaStream_t)stream;
    if (numPoints <= 0 || numFeatures <= 0) return check_launch("mc_sort_features_affine");
    bool vec4 = (numFeatures % 4 == 0) &&
                (((uintptr_t)in | (uintptr_t)out | (uintptr_t)scale | (uintptr_t)shift) % 16 == 0);
    if (vec4) {
        long long total = (long long)numPoints * (numFeatures / 4);
        permute_rows_affine_kernel<float4><<<cdiv(total, 256), 256, 0, s>>>(
            dNumPoints, numPoints, numFeatures / 4, reinterpret_cast<const float4*>(in), newIdx,
            reinterpret_cast<const float4*>(scale), reinterpret_cast<const float4*>(shift), relu,
            reinterpret_cast<float4*>(out));
    } else {
        long long total = (long long)numPoints * numFeatures;
        permute_rows_affine_kernel<float><<<cdiv(total, 256), 256, 0, s>>>(dNumPoints, numPoints, numFeatures, in,
                                                                           newIdx, scale, shift, relu, out);
    }
    return check_launch("mc_sort_features_affine");
}

// sortFeaturesBack, tf_ops/sort_gpu.cu:516-527: out[i,:] = in[newIdx[i],:]
extern "C" int mc_sort_features_back(int numPoints, const int* dNumPoints, int numFeatures, const float* in,
                                     const int* newIdx, float* out, void* stream) {
    launch_permute_rows<false>(dNumPoints, numPoints, numFeatures, in, newIdx, out, (cudaStream_t)stream);
    return check_launch("mc_sort_features_back");
}

// getFeaturesSampledPoints, tf_ops/poisson_sampling.cu:246-258: out[s,:] = features[idx[s],:]
extern "C" int mc_get_sampled_features(int numSamples, const int* dNumSamples, int numFeatures, const int* idx,
                                       const float* features, float* out, void* stream) {
    launch_permute_rows<false>(dNumSamples, numSamples, numFeatures, features, idx, out, (cudaStream_t)stream);
    return check_launch("mc_get_sampled_features");
}

// computeInverseIndexs + transformIndexs, tf_ops/sort_gpu.cu:542-564.  workspace: numPoints ints.
extern "C" int mc_transform_indexs(int numIndexs, const int* dNumIndexs, int numPoints, const int* dNumPoints,
                                   const int* inIndexs, const int* newIdx, int* workspace, int* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (numPoints > 0)
        inverse_index_kernel<<<cdiv(numPoints, kPtThreads), kPtThreads, 0, s>>>(dNumPoints, numPoints, newIdx,
                                                                              workspace);
    if (numIndexs > 0)
        gather_index_kernel<<<cdiv(numIndexs, kPtThreads), kPtThreads, 0, s>>>(dNumIndexs, numIndexs, inIndexs,
                                                                             workspace, out);
    count_launches(1);
    return check_launch("mc_transform_indexs");
}
