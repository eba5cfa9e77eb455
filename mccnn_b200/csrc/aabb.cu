// aabb.cu -- per-cloud (or batch-global) axis-aligned bounding boxes of a packed batch.
//
// Replaces computeAABB / comp_AABB (tf_ops/aabb_gpu.cu:57-141): instead of CAS-loop float atomics into shared
// and then global memory with two blocking H2D memcpys for the identities, floats are mapped to order-preserving
// int32 keys so the hardware integer atomicMin/atomicMax (and the warp-wide redux.sync) do the reduction; each
// thread keeps a running box while the batch id does not change (packed batches are cloud-contiguous,
// utils/DataSet.py:816-818), so the number of global atomics is ~6 per warp, not 6 per point.  The last CTA to
// finish decodes the keys back to floats in place (and, for scale_inv = false, copies the global box to every
// row as aabb_gpu.cu:104-114 does).  min/max are exact and order-independent, so the result is bit-identical.
#include "common.cuh"

namespace mc {
namespace {

__device__ __forceinline__ int f2o(float f) {
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float o2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

__global__ void aabb_init_kernel(int B, int* __restrict__ mnKeys, int* __restrict__ mxKeys, int* __restrict__ ticket) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 3) {
        mnKeys[i] = f2o(FLT_MAX);
        mxKeys[i] = f2o(-FLT_MAX);
    }
    if (i == 0) *ticket = 0;
}

struct Box {
    int lo[3], hi[3];
    __device__ void reset() {
        lo[0] = lo[1] = lo[2] = INT_MAX;
        hi[0] = hi[1] = hi[2] = INT_MIN;
    }
};

__global__ void __launch_bounds__(256) aabb_reduce_kernel(bool scaleInv, int N, int B, const float* __restrict__ pts,
                                                          const int* __restrict__ bids, int* __restrict__ mnKeys,
                                                          int* __restrict__ mxKeys, int* __restrict__ ticket) {
    Box box;
    box.reset();
    int cur = -1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int b = scaleInv ? bids[i] : 0;
        if ((unsigned)b >= (unsigned)B) continue;  // out-of-range id: ignore instead of corrupting memory
        if (b != cur) {
            if (cur >= 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    atomicMin(&mnKeys[cur * 3 + a], box.lo[a]);
                    atomicMax(&mxKeys[cur * 3 + a], box.hi[a]);
                }
            }
            cur = b;
            box.reset();
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int k = f2o(pts[(size_t)i * 3 + a]);
            box.lo[a] = min(box.lo[a], k);
            box.hi[a] = max(box.hi[a], k);
        }
    }
    // warp-aggregated flush: lanes holding the same cloud reduce with redux.sync, one lane issues the atomics
    unsigned grp = __match_any_sync(MC_FULL_MASK, cur);
    int leader = __ffs(grp) - 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int lo = __reduce_min_sync(grp, box.lo[a]);
        int hi = __reduce_max_sync(grp, box.hi[a]);
        if (cur >= 0 && lane_id() == leader) {
            atomicMin(&mnKeys[cur * 3 + a], lo);
            atomicMax(&mxKeys[cur * 3 + a], hi);
        }
    }
    // last CTA decodes
    __shared__ bool isLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) isLast = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    float* mnF = reinterpret_cast<float*>(mnKeys);
    float* mxF = reinterpret_cast<float*>(mxKeys);
    if (scaleInv) {
        for (int i = threadIdx.x; i < B * 3; i += blockDim.x) {
            mnF[i] = o2f(__ldcg(&mnKeys[i]));
            mxF[i] = o2f(__ldcg(&mxKeys[i]));
        }
    } else {
        __shared__ float g[6];
        if (threadIdx.x < 3) g[threadIdx.x] = o2f(__ldcg(&mnKeys[threadIdx.x]));
        if (threadIdx.x >= 3 && threadIdx.x < 6) g[threadIdx.x] = o2f(__ldcg(&mxKeys[threadIdx.x - 3]));
        __syncthreads();
        for (int i = threadIdx.x; i < B * 3; i += blockDim.x) {
            mnF[i] = g[i % 3];
            mxF[i] = g[3 + i % 3];
        }
    }
}

}  // namespace
}  // namespace mc

extern "C" int mc_compute_aabb(int scaleInv, int numPoints, int batchSize, const float* pts, const int* batchIds,
                               float* aabbMin, float* aabbMax, int* workspace, void* stream) {
    MC_REQUIRE(batchSize > 0, "mc_compute_aabb: batch_size must be positive");
    MC_REQUIRE(numPoints >= 0, "mc_compute_aabb: negative point count");
    MC_REQUIRE(workspace != nullptr, "mc_compute_aabb: workspace (1 int) required");
    cudaStream_t s = (cudaStream_t)stream;
    int* mnK = reinterpret_cast<int*>(aabbMin);
    int* mxK = reinterpret_cast<int*>(aabbMax);
    mc::aabb_init_kernel<<<mc::cdiv(batchSize * 3, 128), 128, 0, s>>>(batchSize, mnK, mxK, workspace);
    int grid = mc::cdiv(numPoints > 0 ? numPoints : 1, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    mc::aabb_reduce_kernel<<<grid, 256, 0, s>>>(scaleInv != 0, numPoints, batchSize, pts, batchIds, mnK, mxK, workspace);
    mc::count_launches(1);
    return mc::check_launch("mc_compute_aabb");
}
