// lookback.cuh -- single-pass exclusive prefix across CTAs (decoupled look-back).
//
// CTAs obtain their logical index from a ticket (atomicAdd), so index order == start order and a CTA only ever waits
// for CTAs that are already running: no deadlock regardless of how many CTAs are resident.  Every CTA publishes its
// aggregate as soon as it is known ({status, value} packed in one 64-bit word), then one warp walks backwards over its
// predecessors 32 at a time, adding aggregates until it meets a published inclusive prefix.
//
// Descriptor array: `desc[0]` = ticket counter (low 32 bits), `desc[1 + id]` = descriptor of CTA id.  The caller
// zeroes the array (cudaMemsetAsync) before the launch.  Used to concatenate per-cloud Poisson samples and per-CTA
// neighbour lists without a second kernel or a host round trip (the reference does both:
// tf_ops/find_neighbors.cu:122-182,292-309, tf_ops/poisson_sampling.cu:206-223).
#pragma once
#include "common.cuh"

namespace mc {

constexpr unsigned long long kLbAggregate = 1ull << 32;
constexpr unsigned long long kLbPrefix = 2ull << 32;

__device__ __forceinline__ int lookback_ticket(unsigned long long* desc) {
    return (int)atomicAdd(desc, 1ull);
}

__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// Called by ONE full warp of the CTA.  Returns the exclusive prefix of `aggregate` over CTAs 0..id-1 (all lanes).
__device__ __forceinline__ int lookback_exclusive_prefix(unsigned long long* desc, int id, int aggregate) {
    unsigned long long* d = desc + 1;
    const int lane = lane_id();
    if (id == 0) {
        if (lane == 0) atomicExch(&d[0], kLbPrefix | (unsigned)aggregate);
        return 0;
    }
    if (lane == 0) atomicExch(&d[id], kLbAggregate | (unsigned)aggregate);
    int exclusive = 0;
    int pos = id - 1;
    while (true) {
        const int idx = pos - lane;
        unsigned long long v = kLbPrefix;  // virtual predecessors before CTA 0: prefix 0
        if (idx >= 0) {
            while (((v = lb_load(&d[idx])) >> 32) == 0) __nanosleep(20);
        }
        const unsigned isPrefix = __ballot_sync(MC_FULL_MASK, (v >> 32) == 2);
        const int first = isPrefix ? __ffs(isPrefix) - 1 : 32;  // nearest predecessor holding an inclusive prefix
        int val = (lane <= first) ? (int)(unsigned)(v & 0xffffffffull) : 0;
        val = __reduce_add_sync(MC_FULL_MASK, val);
        exclusive += val;
        if (isPrefix) break;
        pos -= 32;
    }
    if (lane == 0) atomicExch(&d[id], kLbPrefix | (unsigned)(exclusive + aggregate));
    return exclusive;
}

}  // namespace mc
