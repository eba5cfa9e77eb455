// peer_gather.cu -- the closing collective of a sharded step, written directly over NVLink peer memory:
// all-gather of every rank's result rows (BASELINE.json north_star: "a single all-gather of per-cloud logits at the
// end"; per-point logits of unequal length for the segmentation networks).
//
// The classification payload is ~5 KB per rank.  An NCCL all_gather for it costs ~45 us of pipeline time per step on
// B200 (measured in round 1: 0.150 ms per step with it -- captured in the CUDA graph or not, one channel or many --
// against 0.105 ms without), because a general collective brings its own launch / protocol machinery.  Here every rank
// owns a SYMMETRIC buffer (same layout on every GPU, mapped into every peer's address space: torch symmetric memory,
// NVLink / NVSwitch) holding a ring of `slots` steps, and the exchange is two kernels:
//
//   put      (last node of the forward's CUDA graph)  stores this rank's rows + a header (row count, overflow flag of
//            the step) into slot `seq % slots`, section `rank`, of EVERY rank's buffer -- plain 128-bit stores that
//            travel over NVLink -- then __threadfence_system() and, from the last CTA to finish, the step number into
//            the peers' flag words.  It never waits for a peer's data: only (flow control) for the acknowledgement
//            that every peer has collected step seq - slots, i.e. that the slot is free again.
//   collect  (launched on the side stream that feeds the host copy)  waits until the flag words of all ranks in the
//            LOCAL buffer carry the step number, copies the gathered rows out (L2 loads: no stale L1 lines), either
//            padded [world x capacity] or compacted in rank order with the live total in a device int, reduces the
//            overflow flags with MAX, and acknowledges the step in every peer's buffer.
//
// Because put and collect are separate, a lane's next forward never waits on a slow peer: ranks may drift apart by up
// to `slots` steps of a lane before anything blocks (round 1's single kernel made every step of every lane wait for the
// slowest rank's same step).  All addresses are fixed and the step numbers live in device counters: both kernels are
// ordinary graph nodes / stream launches.  mc_allgather_rows_p2p keeps the one-kernel form (put + collect fused) for
// callers that want the gathered rows inside the same graph.
#include "common.cuh"

namespace mc {
namespace {

constexpr int kHeaderFloats = 4;  // {rows, overflow flag, -, -} in front of every section: keeps the rows 16-byte aligned

struct PeerLayout {
    int world, slots;
    long long capFloats;  // payload capacity per rank and slot (multiple of 4)
    __host__ __device__ long long section() const { return capFloats + kHeaderFloats; }
    __host__ __device__ long long dataFloats() const { return (long long)slots * world * section(); }
    __host__ __device__ long long flagOffset() const { return dataFloats(); }                          // [slots][world]
    __host__ __device__ long long ackOffset() const { return dataFloats() + (long long)slots * world; }  // [world]
    __host__ __device__ long long totalFloats() const { return ackOffset() + world + 8; }
};

__device__ __forceinline__ void spin_until_at_least(volatile int* word, int value) {
    while (*word - value < 0) __nanosleep(64);
}

// state: [0] put sequence, [1] collect sequence, [2] put ticket, [3] collect ticket (device ints, zero before first use)
__global__ void __launch_bounds__(256) peer_put_kernel(PeerLayout L, int rank, int cols, const int* __restrict__ dRows,
                                                       int capRows, const float* __restrict__ local,
                                                       const int* __restrict__ dFlag,
                                                       const long long* __restrict__ peerBufs, int* __restrict__ state) {
    __shared__ int sSeq, sLast;
    const int seq = state[0] + 1;  // every CTA reads the same value: state[0] is bumped by the last CTA only
    const int slot = seq % L.slots;
    // flow control: the slot is free once every peer has collected step seq - slots
    if (threadIdx.x < L.world) {
        volatile int* ack = reinterpret_cast<int*>(reinterpret_cast<float*>(peerBufs[rank]) + L.ackOffset()) + threadIdx.x;
        spin_until_at_least(ack, seq - L.slots);
    }
    __syncthreads();
    const int rows = live_count(dRows, capRows);
    const long long n = (long long)rows * cols;
    const long long n4 = n >> 2;
    for (int r = 0; r < L.world; ++r) {
        float* dst = reinterpret_cast<float*>(peerBufs[r]) + ((long long)slot * L.world + rank) * L.section();
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            reinterpret_cast<int*>(dst)[0] = rows;
            reinterpret_cast<int*>(dst)[1] = dFlag != nullptr ? *dFlag : 0;
        }
        float4* d4 = reinterpret_cast<float4*>(dst + kHeaderFloats);
        const float4* s4 = reinterpret_cast<const float4*>(local);
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x)
            d4[e] = s4[e];
        if (blockIdx.x == 0)
            for (long long e = (n4 << 2) + threadIdx.x; e < n; e += blockDim.x) dst[kHeaderFloats + e] = local[e];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        sLast = atomicAdd(&state[2], 1) == (int)gridDim.x - 1;
        sSeq = seq;
    }
    __syncthreads();
    if (!sLast) return;
    // last CTA: every CTA's stores are fenced; publish the step in every rank's flag word for this rank
    __threadfence_system();
    if (threadIdx.x < L.world) {
        volatile int* flag = reinterpret_cast<int*>(reinterpret_cast<float*>(peerBufs[threadIdx.x]) + L.flagOffset()) +
                             slot * L.world + rank;
        *flag = sSeq;
    }
    if (threadIdx.x == 0) {
        state[2] = 0;
        state[0] = sSeq;
    }
}

// out: [world * capRows, cols] (padded: rank r's rows at r * capRows) or compacted in rank order (compact != 0);
// meta: [world][2] ints (rows, overflow flag) per rank, then meta[2 * world] = total rows, meta[2 * world + 1] = MAX flag
__global__ void __launch_bounds__(256) peer_collect_kernel(PeerLayout L, int rank, int cols, int capRows, int compact,
                                                           const long long* __restrict__ peerBufs, int* __restrict__ state,
                                                           float* __restrict__ out, int* __restrict__ meta) {
    __shared__ int sLast;
    __shared__ int sRows[256], sBase[257];
    // bit 1 of `compact`: collect only if a put is outstanding (the deferred collect at the head of a lane's next step
    // and the flush behind its last one share one call); state[0] / state[1] only change in stream order with this kernel
    if ((compact & 2) && state[1] >= state[0]) return;
    compact &= 1;
    const int seq = state[1] + 1;
    const int slot = seq % L.slots;
    const float* mine = reinterpret_cast<const float*>(peerBufs[rank]);
    if (threadIdx.x < L.world) {
        volatile int* flag = reinterpret_cast<int*>(const_cast<float*>(mine) + L.flagOffset()) + slot * L.world + threadIdx.x;
        spin_until_at_least(flag, seq);
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x < L.world) {
        const int* hdr = reinterpret_cast<const int*>(mine + ((long long)slot * L.world + threadIdx.x) * L.section());
        sRows[threadIdx.x] = min(__ldcg(hdr), capRows);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0, flagMax = 0;
        for (int r = 0; r < L.world; ++r) {
            sBase[r] = compact ? acc : r * capRows;
            acc += sRows[r];
            const int* hdr = reinterpret_cast<const int*>(mine + ((long long)slot * L.world + r) * L.section());
            const int f = __ldcg(hdr + 1);
            flagMax = f > flagMax ? f : flagMax;
            if (blockIdx.x == 0 && meta != nullptr) {
                meta[2 * r] = sRows[r];
                meta[2 * r + 1] = f;
            }
        }
        sBase[L.world] = acc;
        if (blockIdx.x == 0 && meta != nullptr) {
            meta[2 * L.world] = acc;
            meta[2 * L.world + 1] = flagMax;
        }
    }
    __syncthreads();
    for (int r = 0; r < L.world; ++r) {
        const float* src = mine + ((long long)slot * L.world + r) * L.section() + kHeaderFloats;
        float* dst = out + (long long)sBase[r] * cols;
        const long long n = (long long)sRows[r] * cols;
        const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
        if (vec) {
            const long long n4 = n >> 2;
            for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x)
                reinterpret_cast<float4*>(dst)[e] = __ldcg(reinterpret_cast<const float4*>(src) + e);
            if (blockIdx.x == 0)
                for (long long e = (n4 << 2) + threadIdx.x; e < n; e += blockDim.x) dst[e] = __ldcg(src + e);
        } else {
            for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
                dst[e] = __ldcg(src + e);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sLast = atomicAdd(&state[3], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!sLast) return;
    // every CTA has read the slot: acknowledge the step to every writer
    __threadfence_system();
    if (threadIdx.x < L.world) {
        volatile int* ack = reinterpret_cast<int*>(reinterpret_cast<float*>(peerBufs[threadIdx.x]) + L.ackOffset()) + rank;
        *ack = seq;
    }
    if (threadIdx.x == 0) {
        state[3] = 0;
        state[1] = seq;
    }
}

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

PeerLayout make_layout(int capRows, int cols, int world, int slots) {
    PeerLayout L;
    L.world = world;
    L.slots = slots;
    L.capFloats = (((long long)capRows * cols + 3) / 4) * 4;
    return L;
}

int copy_grid(long long floats) {
    long long g = (floats / 4 + 256 * 8 - 1) / (256 * 8);  // ~8 float4 per thread
    return (int)(g < 1 ? 1 : (g > 64 ? 64 : g));           // a few dozen CTAs saturate an NVLink direction
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

// floats of the symmetric buffer of a put / collect exchange: a ring of `slots` steps, `capRows x cols` floats per rank
extern "C" long long mc_peer_exchange_buffer_floats(int capRows, int cols, int worldSize, int slots) {
    return make_layout(capRows, cols, worldSize, slots).totalFloats();
}

// Step `seq` (device counter state[0] + 1) of the exchange: this rank's live rows (min(*dRows, capRows); NULL: capRows)
// and its overflow flag (*dFlag; NULL: 0) go to slot seq % slots of every rank's symmetric buffer.  Waits only for the
// peers' acknowledgement of step seq - slots.  state: 4 device ints, zero before the first step.  local 16-byte aligned.
extern "C" int mc_peer_put(int capRows, const int* dRows, int cols, int rank, int worldSize, int slots,
                           const float* local, const int* dFlag, const long long* peerBufPtrs, int* state,
                           void* stream) {
    MC_REQUIRE(capRows > 0 && cols > 0 && worldSize > 0 && worldSize <= 256 && rank >= 0 && rank < worldSize && slots >= 2,
               "mc_peer_put: bad extents");
    MC_REQUIRE(peerBufPtrs != nullptr && state != nullptr && local != nullptr, "mc_peer_put: buffers required");
    MC_REQUIRE((reinterpret_cast<uintptr_t>(local) & 15) == 0, "mc_peer_put: rows must be 16-byte aligned");
    const PeerLayout L = make_layout(capRows, cols, worldSize, slots);
    peer_put_kernel<<<copy_grid((long long)capRows * cols * worldSize), 256, 0, (cudaStream_t)stream>>>(
        L, rank, cols, dRows, capRows, local, dFlag, peerBufPtrs, state);
    return check_launch("mc_peer_put");
}

// Step state[1] + 1: waits for every rank's put of that step, copies the rows out of the local buffer -- padded (rank
// r's rows at r * capRows) or, compact != 0, contiguous in rank order -- fills meta (2 ints (rows, flag) per rank, then
// the total row count and the MAX of the flags; may be NULL) and acknowledges the step to every rank.
// compact: bit 0 = compacted output, bit 1 = do nothing unless a put of this exchange is still uncollected.
extern "C" int mc_peer_collect(int capRows, int cols, int rank, int worldSize, int slots, int compact,
                               const long long* peerBufPtrs, int* state, float* out, int* meta, void* stream) {
    MC_REQUIRE(capRows > 0 && cols > 0 && worldSize > 0 && worldSize <= 256 && rank >= 0 && rank < worldSize && slots >= 2,
               "mc_peer_collect: bad extents");
    MC_REQUIRE(peerBufPtrs != nullptr && state != nullptr && out != nullptr, "mc_peer_collect: buffers required");
    const PeerLayout L = make_layout(capRows, cols, worldSize, slots);
    peer_collect_kernel<<<copy_grid((long long)capRows * cols * worldSize), 256, 0, (cudaStream_t)stream>>>(
        L, rank, cols, capRows, compact, peerBufPtrs, state, out, meta);
    return check_launch("mc_peer_collect");
}
