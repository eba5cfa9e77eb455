// peer_gather.cu -- the closing collective of a sharded step, written directly over NVLink peer memory:
// all-gather of the per-cloud logits [rowsPerRank x cols] of every rank (BASELINE.json north_star: "a single
// all-gather of per-cloud logits at the end").
//
// The payload is ~5 KB per rank.  An NCCL all_gather for it costs ~45 us of pipeline time per step on B200 (measured:
// 0.150 ms per step with it -- captured in the CUDA graph or not, one channel or many -- against 0.105 ms without),
// because a general collective brings its own launch / protocol machinery.  Here every rank owns a SYMMETRIC buffer
// (same layout on every GPU, mapped into every peer's address space: torch symmetric memory, NVLink / NVSwitch) and
// ONE small CTA does the whole exchange:
//   1. store the local rows into slot `rank` of EVERY peer's buffer (plain coalesced stores that travel over NVLink),
//   2. __threadfence_system(), then publish the step's sequence number in the peers' flag words,
//   3. spin until the flag words of all ranks in the LOCAL buffer carry this step's number,
//   4. copy the gathered rows from the local buffer (L2 loads, no stale L1 lines) to the result tensor.
// Two buffer halves alternate by step parity: a rank can only be one step ahead of a peer (it needs the peer's flag of
// step s to finish step s), so while a slow peer still reads half s % 2 the fast one writes half (s + 1) % 2.
// All addresses are fixed, the step number lives in a device counter: the kernel is a normal node of the captured
// forward graph.
#include "common.cuh"

namespace mc {
namespace {

__global__ void __launch_bounds__(256) allgather_rows_p2p_kernel(int n, int rank, int world,
                                                                 const float* __restrict__ local,
                                                                 const long long* __restrict__ peerBufs,
                                                                 int* __restrict__ counter, float* __restrict__ out) {
    __shared__ int sSeq;
    if (threadIdx.x == 0) sSeq = ++(*counter);
    __syncthreads();
    const int seq = sSeq, half = seq & 1;
    const size_t dataFloats = (size_t)2 * world * n;  // [2 halves][world][n] floats, then [2][world] int flags
    // 1. my rows -> slot `rank` of every rank's buffer (including my own)
    for (int r = 0; r < world; ++r) {
        float* dst = reinterpret_cast<float*>(peerBufs[r]) + ((size_t)half * world + rank) * n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = local[e];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish, 3. wait
    if (threadIdx.x < world) {
        const int r = threadIdx.x;
        volatile int* theirFlag = reinterpret_cast<int*>(reinterpret_cast<float*>(peerBufs[r]) + dataFloats) +
                                  half * world + rank;
        *theirFlag = seq;
        volatile int* myFlag = reinterpret_cast<int*>(reinterpret_cast<float*>(peerBufs[rank]) + dataFloats) +
                               half * world + r;
        while (*myFlag - seq < 0) __nanosleep(100);
    }
    __syncthreads();
    __threadfence_system();
    // 4. gathered rows -> result
    const float* src = reinterpret_cast<const float*>(peerBufs[rank]) + (size_t)half * world * n;
    for (int e = threadIdx.x; e < world * n; e += blockDim.x) out[e] = __ldcg(src + e);
}

}  // namespace
}  // namespace mc

using namespace mc;

// floats a rank's symmetric buffer must hold for mc_allgather_rows_p2p
extern "C" long long mc_allgather_rows_p2p_buffer_floats(int floatsPerRank, int worldSize) {
    return 2LL * worldSize * floatsPerRank + 2LL * worldSize + 8;
}

// out[r * floatsPerRank + e] = local_r[e] for every rank r.  peerBufPtrs: DEVICE array of worldSize pointers (as
// 64-bit integers) to the ranks' symmetric buffers (zero-filled before the first step, each of
// mc_allgather_rows_p2p_buffer_floats floats); stepCounter: device int, zero before the first step, private to this
// buffer set.  Every rank must issue the same sequence of calls on a given buffer set.
extern "C" int mc_allgather_rows_p2p(int floatsPerRank, int rank, int worldSize, const float* local,
                                     const long long* peerBufPtrs, int* stepCounter, float* out, void* stream) {
    MC_REQUIRE(floatsPerRank > 0 && worldSize > 0 && rank >= 0 && rank < worldSize,
               "mc_allgather_rows_p2p: bad extents");
    MC_REQUIRE(worldSize <= 256, "mc_allgather_rows_p2p: at most 256 ranks");
    MC_REQUIRE(peerBufPtrs != nullptr && stepCounter != nullptr, "mc_allgather_rows_p2p: buffers required");
    allgather_rows_p2p_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(floatsPerRank, rank, worldSize, local, peerBufPtrs,
                                                                  stepCounter, out);
    return check_launch("mc_allgather_rows_p2p");
}
