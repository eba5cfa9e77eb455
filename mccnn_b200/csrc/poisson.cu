// poisson.cu -- grid-based greedy Poisson-disk sampling of every cloud of a packed, grid-sorted batch.
//
// Replaces samplePointCloud / selectSamples / copyPoints (tf_ops/poisson_sampling.cu:51-124,175-244).  The
// reference launches one tiny kernel per (cloud, phase) -- B*27 sequential launches (864 at B = 32), one thread per
// cell walking its 27 neighbour cells serially -- then blocks the host on the sample count.  Here the whole op is ONE
// kernel launch (plus one memset):
//   * one CTA per cloud (clouds claimed through a ticket so that CTA order == cloud order); the 27 colour phases run
//     inside the kernel separated by __syncthreads(); the accepted flags of the cloud live in shared memory, and
//     the non-empty cells of every phase are listed once up front so that no warp ever visits an empty cell;
//   * a WARP owns a cell of the phase.  It (A) compacts the already-accepted points of the 26 neighbour cells into a
//     small shared-memory list (lanes = neighbour cells), (B) tests its own points against that list with the 32
//     lanes spread over (point, list slice) pairs, and (C) resolves the order dependence INSIDE the cell -- a point
//     is rejected by an earlier accepted point of the same cell -- by walking only the still-unblocked points with
//     shuffles.  The serial chain is therefore the number of unblocked points of a cell (typically 1-3), not the
//     27-cell neighbourhood population x memory latency;
//   * distances use the square-root-free exact predicate (common.cuh exact_radius_threshold);
//   * the accepted set is identical to the reference's (deterministic there too: same-phase cells are >= 3 cells
//     apart and a thread only flags points of its own cell);
//   * output order: the reference's is atomic-arrival order inside a (cloud, phase) launch; the canonical order
//     (cloud, phase, ascending position in the sorted array) is produced in-kernel by a stable per-phase
//     compaction, and clouds are concatenated through a decoupled look-back prefix (lookback.cuh).  The sample
//     count stays on the device.
#include "common.cuh"
#include "lookback.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace mc {
namespace {

constexpr int kSelThreads = 512;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelFlagBytes = 48 * 1024;  // per-cloud accepted flags kept in shared memory up to this many points
// (these four are the LARGEST capacities; a launch carves its shared memory for the capacities its batch needs --
// SelCaps -- so that the CTAs of a small level leave room for other kernels' CTAs on the same SM)
constexpr int kListCap = 128;             // accepted neighbour candidates a warp keeps in shared memory
constexpr int kSelfOffset = 17;           // index of (0,0,0) in kOffsPool
constexpr int kCellListCap = 8192;        // cells per cloud up to which the per-phase non-empty cell lists are built
constexpr int kStageCells = 4096;         // cell-table rows of a cloud staged in shared memory up to this many cells
constexpr int kStagePts = 4096;           // points of a cloud staged in shared memory up to this many points
#ifdef MCCNN_POISSON_SINGLE_LAUNCH
constexpr bool kForceSingleLaunch = true;  // build-time switch: one CTA per cloud also on fine grids (A/B timing)
#else
constexpr bool kForceSingleLaunch = false;
#endif

// phase (index into kOffsPool) of the offset (ox, oy, oz), indexed by (ox+1)*9 + (oy+1)*3 + (oz+1)
__device__ __constant__ const unsigned char kPhaseOfOffset[27] = {
    /*(-1,-1,-1)*/ 25, /*(-1,-1,0)*/ 15, /*(-1,-1,1)*/ 26, /*(-1,0,-1)*/ 24, /*(-1,0,0)*/ 12, /*(-1,0,1)*/ 22,
    /*(-1,1,-1)*/ 6,   /*(-1,1,0)*/ 16,  /*(-1,1,1)*/ 11,  /*(0,-1,-1)*/ 7,  /*(0,-1,0)*/ 5,  /*(0,-1,1)*/ 1,
    /*(0,0,-1)*/ 18,   /*(0,0,0)*/ 17,   /*(0,0,1)*/ 4,    /*(0,1,-1)*/ 14,  /*(0,1,0)*/ 3,   /*(0,1,1)*/ 2,
    /*(1,-1,-1)*/ 13,  /*(1,-1,0)*/ 21,  /*(1,-1,1)*/ 9,   /*(1,0,-1)*/ 20,  /*(1,0,0)*/ 8,   /*(1,0,1)*/ 10,
    /*(1,1,-1)*/ 0,    /*(1,1,0)*/ 19,   /*(1,1,1)*/ 23};

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of one contiguous global range into shared memory, completion on an
// mbarrier: used to stage a cloud's points -- one asynchronous copy issued by one thread instead of a strided loop of
// 4-byte loads by the whole CTA.  Addresses and size must be multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dstSmem, const void* srcGlobal, unsigned bytes,
                                              unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MC_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MC_MBAR_DONE;\n"
        "bra MC_MBAR_WAIT;\n"
        "MC_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// shared-memory capacities of one launch (all multiples of 16); a cloud above one of them takes the corresponding
// global-memory path, exactly as a cloud above the largest capacities does
struct SelCaps {
    int flagBytes;   // accepted flags in shared memory up to this many points per cloud
    int cellList;    // per-phase non-empty cell lists up to this many cells per cloud
    int stageCells;  // cell-table rows staged up to this many cells per cloud
    int stagePts;    // points staged up to this many points per cloud
};

struct SelShared {
    float4 list[kSelWarps][kListCap];
    int listIdx[kSelWarps][kListCap];
    int listCount[kSelWarps];
    int warpCnt[kSelWarps][28];
    int phaseBase[28];
    int running[28];
    int phaseCellStart[28];  // non-empty cell lists: cells of phase p are cellList[phaseCellStart[p] .. [p+1])
    int phaseCellFill[27];
    int lo, hi, cloud, base;
    unsigned long long stageBar;  // mbarrier of the bulk copy that stages the cloud's points
};

struct CloudCtx {
    const int2* cells;            // this cloud's rows of the cell table (global)
    const int2* sCells;           // the same rows staged in shared memory, or nullptr
    const float* pts;             // sorted points (global)
    const float* sPts;            // this cloud's points staged in shared memory (index: sorted position - cLo), or nullptr
    __device__ __forceinline__ int2 cell(int c) const { return sCells ? sCells[c] : cells[c]; }
    __device__ __forceinline__ void point(int j, float& x, float& y, float& z) const {
        const float* p = sPts ? sPts + (size_t)(j - cLo) * 3 : pts + (size_t)j * 3;
        x = p[0];
        y = p[1];
        z = p[2];
    }
    // accepted points so far, per cell, compacted in place: the a-th accepted point of cell c is the point
    // cLo + sAccIdx[cell(c).x - cLo + a], a < sAccCnt[c]  (only when points and cell rows are staged; else nullptr)
    unsigned short* sAccIdx;
    unsigned char* sAccCnt;
    unsigned char* sFlags;        // shared-memory flags (index: sorted position - cLo) or nullptr
    unsigned char* gFlags;        // global flags (index: sorted position) when the cloud does not fit
    int cLo;
    int n;
    int offX, offY, offZ;         // kOffsPool[lane]: read ONCE per thread (a lane-indexed constant load serialises)
    float thr;                    // exact squared-distance threshold of the cloud's radius
    __device__ __forceinline__ unsigned char flag(int j) const {
        return sFlags ? sFlags[j - cLo] : __ldcg(&gFlags[j]);
    }
    __device__ __forceinline__ void set_flag(int j, unsigned char v) const {
        if (sFlags)
            sFlags[j - cLo] = v;
        else
            __stcg(&gFlags[j], v);
    }
};

// one cell of the current phase, processed by one warp
__device__ __forceinline__ void process_cell(const CloudCtx& cx, SelShared& sh, int warp, int lane, int cellId, int x,
                                             int y, int z, int2 own, int ph) {
    const int n = cx.n;
    const unsigned ltMask = (1u << lane) - 1u;
    // ---- A: gather the accepted points of the 26 neighbour cells into the warp's list
    int nlo = 0, nhi = 0, ncell = -1;
    if (lane < 27 && lane != kSelfOffset) {
        int qx = x + cx.offX, qy = y + cx.offY, qz = z + cx.offZ;
        if (qx >= 0 && qx < n && qy >= 0 && qy < n && qz >= 0 && qz < n) {
            ncell = (qx * n + qy) * n + qz;
            int2 ce = cx.cell(ncell);
            nlo = ce.x;
            nhi = ce.y;
        }
    }
    int L;
    bool listOk;
    if (cx.sAccIdx != nullptr) {
        // per-cell accepted lists: a lane copies the (few) accepted points of its neighbour cell, placed by a warp scan
        const int cntL = ncell >= 0 ? cx.sAccCnt[ncell] : 0;
        // no accepted neighbour at all (every cell of the early phases): skip the scan and the copy
        const int incl = __any_sync(MC_FULL_MASK, cntL > 0) ? warp_inclusive_scan(cntL) : 0;
        L = __shfl_sync(MC_FULL_MASK, incl, 31);
        listOk = L <= kListCap;
        if (listOk && L > 0) {
            const int pos = incl - cntL;
            for (int k = 0; k < cntL; ++k) {
                float qx, qy, qz;
                cx.point(cx.cLo + cx.sAccIdx[nlo - cx.cLo + k], qx, qy, qz);
                sh.list[warp][pos + k] = make_float4(qx, qy, qz, 0.f);
            }
        }
    } else {
        // no staging (very large clouds): compact through the accepted flags
        if (lane == 0) sh.listCount[warp] = 0;
        __syncwarp();
        for (int j = nlo; j < nhi; ++j)
            if (cx.flag(j)) {
                int slot = atomicAdd(&sh.listCount[warp], 1);
                if (slot < kListCap) sh.listIdx[warp][slot] = j;
            }
        __syncwarp();
        L = sh.listCount[warp];
        listOk = L <= kListCap;  // else: fall back to walking the cells directly (never wrong, only slow)
        if (listOk) {
            for (int t = lane; t < L; t += 32) {
                float qx, qy, qz;
                cx.point(sh.listIdx[warp][t], qx, qy, qz);
                sh.list[warp][t] = make_float4(qx, qy, qz, 0.f);
            }
        }
    }
    __syncwarp();
    int accOwn = 0;  // accepted points of this cell so far

    // ---- B + C: own points, 32 at a time
    for (int c0 = own.x; c0 < own.y; c0 += 32) {
        const int cnt = min(32, own.y - c0);
        // lanes spread over (point, list slice): pointLanes = smallest power of two >= cnt
        const int pointLanes = cnt > 1 ? 1 << (32 - __clz(cnt - 1)) : 1;
        const int slices = 32 / pointLanes;
        const int pi = lane & (pointLanes - 1), slice = lane / pointLanes;
        const int i = c0 + pi;
        const bool valid = pi < cnt;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) cx.point(i, px, py, pz);
        int blocked = 0;  // an accepted point of an earlier phase / earlier chunk lies within r
        if (listOk) {
            for (int t = slice; t < L; t += slices) {
                float4 q = sh.list[warp][t];
                blocked |= dist2_ref(px, py, pz, q.x, q.y, q.z) <= cx.thr;
            }
        } else {
            for (int o = 0; o < 27; ++o) {
                int lo2 = __shfl_sync(MC_FULL_MASK, nlo, o), hi2 = __shfl_sync(MC_FULL_MASK, nhi, o);
                if (o == kSelfOffset) {  // earlier chunks of the own cell
                    lo2 = own.x;
                    hi2 = c0;
                }
                for (int j = lo2 + slice; j < hi2; j += slices)
                    if (cx.flag(j)) {
                        float qx, qy, qz;
                        cx.point(j, qx, qy, qz);
                        blocked |= dist2_ref(px, py, pz, qx, qy, qz) <= cx.thr;
                    }
            }
        }
        {   // OR over the slices of a point: one ballot, then the bits of the lanes that share this lane's point
            // (positions pi, pi + pointLanes, ...) -- instead of log2(slices) shuffle rounds
            const unsigned votes = __ballot_sync(MC_FULL_MASK, blocked);
            unsigned every = 1u;  // bit 0 of every group of pointLanes bits
            for (int sft = pointLanes; sft < 32; sft <<= 1) every |= every << sft;
            blocked = (votes & (every << pi)) != 0u;
        }
        // C: order dependence inside the chunk, walking only the unblocked points in ascending order; lanes
        // 0..cnt-1 (slice 0) hold the points
        const bool holder = lane < cnt;  // slice 0 lanes: pi == lane
        unsigned cand = __ballot_sync(MC_FULL_MASK, holder && !blocked);
        bool mySel = false;
        while (cand) {
            const int u = __ffs(cand) - 1;
            cand &= cand - 1;
            const float ux = __shfl_sync(MC_FULL_MASK, px, u);
            const float uy = __shfl_sync(MC_FULL_MASK, py, u);
            const float uz = __shfl_sync(MC_FULL_MASK, pz, u);
            const bool hit = holder && lane < u && mySel && dist2_ref(ux, uy, uz, px, py, pz) <= cx.thr;
            const unsigned any = __ballot_sync(MC_FULL_MASK, hit);
            if (lane == u) mySel = (any == 0);
        }
        if (holder && mySel) cx.set_flag(i, (unsigned char)(ph + 1));
        const unsigned selMask = __ballot_sync(MC_FULL_MASK, holder && mySel);
        if (cx.sAccIdx != nullptr) {
            if (holder && mySel)
                cx.sAccIdx[own.x - cx.cLo + accOwn + __popc(selMask & ltMask)] = (unsigned short)(i - cx.cLo);
            accOwn += __popc(selMask);
        }
        if (c0 + 32 < own.y && listOk) {  // later chunks of this cell must see this chunk's accepted points
            const int pos = L + __popc(selMask & ltMask);
            if (holder && mySel && pos < kListCap) sh.list[warp][pos] = make_float4(px, py, pz, 0.f);
            L += __popc(selMask);
            if (L > kListCap) listOk = false;
        }
        __syncwarp();
    }
    if (cx.sAccIdx != nullptr && lane == 0) cx.sAccCnt[cellId] = (unsigned char)min(accOwn, 255);
}

__global__ void __launch_bounds__(kSelThreads)
    poisson_kernel(bool scaleInv, float radius, int B, int n, const float* __restrict__ mn,
                   const float* __restrict__ mx, const float* __restrict__ pts, const int2* __restrict__ cellIdx,
                   unsigned char* selGlobal, int* __restrict__ tmpIdx, unsigned long long* chain, int capOut,
                   float* __restrict__ outPts, int* __restrict__ outBids, int* __restrict__ outIdx,
                   int* __restrict__ dNumSamples, int* __restrict__ dOverflow, bool phasesDone, SelCaps caps) {
    // phasesDone: the accepted flags were already produced in selGlobal by the many-CTA phase kernels (fine grids);
    // this launch then only orders and writes the samples
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SelShared& sh = *reinterpret_cast<SelShared*>(smemRaw);
    unsigned char* sFlags = smemRaw + sizeof(SelShared);
    unsigned short* cellList = reinterpret_cast<unsigned short*>(sFlags + caps.flagBytes);
    int2* sCells = reinterpret_cast<int2*>(sFlags + caps.flagBytes + caps.cellList * sizeof(unsigned short));
    float* sPts = reinterpret_cast<float*>(sCells + caps.stageCells);  // + 16 bytes of slack for the alignment shift
    unsigned short* sAccIdx = reinterpret_cast<unsigned short*>(sPts + caps.stagePts * 3 + 4);
    unsigned char* sAccCnt = reinterpret_cast<unsigned char*>(sAccIdx + caps.stagePts);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned ltMask = (1u << lane) - 1u;

    // claim a cloud: ticket order == start order, so the look-back at the end only waits for running CTAs
    if (threadIdx.x == 0) {
        sh.cloud = lookback_ticket(chain);
        sh.lo = INT_MAX;
        sh.hi = 0;
        mbar_init(&sh.stageBar, 1);
    }
    if (threadIdx.x < 28) {
        sh.phaseBase[threadIdx.x] = 0;
        sh.running[threadIdx.x] = 0;
        sh.phaseCellStart[threadIdx.x] = 0;
        if (threadIdx.x < 27) sh.phaseCellFill[threadIdx.x] = 0;
    }
    __syncthreads();
    const int b = sh.cloud;
    const int cellsPerCloud = n * n * n;
    const int2* cells = cellIdx + (size_t)b * cellsPerCloud;
    const bool useLists = cellsPerCloud <= caps.cellList && !phasesDone;
    const bool stageCells = cellsPerCloud <= caps.stageCells && !phasesDone;

    // extent of this cloud inside the sorted array (the sort is batch-major) + non-empty cells per phase
    {
        int lo = INT_MAX, hi = 0;
        for (int c = threadIdx.x; c < cellsPerCloud; c += kSelThreads) {
            int2 ce = cells[c];
            if (stageCells) {
                sCells[c] = ce;
                sAccCnt[c] = 0;
            }
            if (ce.y > ce.x) {
                lo = min(lo, ce.x);
                hi = max(hi, ce.y);
                if (useLists) {
                    const int x = c / (n * n), y = (c / n) % n, z = c % n;
                    const int p = kPhaseOfOffset[(x % 3) * 9 + (y % 3) * 3 + (z % 3)];  // offset = coord % 3 - 1
                    atomicAdd(&sh.phaseCellStart[p + 1], 1);
                }
            }
        }
        lo = __reduce_min_sync(MC_FULL_MASK, lo);
        hi = __reduce_max_sync(MC_FULL_MASK, hi);
        if (lane == 0) {
            atomicMin(&sh.lo, lo);
            atomicMax(&sh.hi, hi);
        }
    }
    __syncthreads();
    const int cLo = sh.lo, cHi = sh.hi;
    const int nCloud = cHi > cLo ? cHi - cLo : 0;
    const bool inSmem = nCloud <= caps.flagBytes && !phasesDone;
    if (inSmem)
        for (int t = threadIdx.x; t < nCloud; t += kSelThreads) sFlags[t] = 0;
    // every later access to the cloud's points and cell rows is then a shared-memory access: the per-cell work is a
    // chain of dependent loads, and first-touch global misses were what made it slow
    const bool stagePts = nCloud <= caps.stagePts && !phasesDone;
    if (stagePts) {
        // the cloud's points are one contiguous range of 12-byte rows: the 16-byte aligned interior goes through ONE
        // TMA bulk copy, the (at most 12 + 12) bytes before / after it through plain loads; the shared-memory image
        // is shifted by the same misalignment so that both sides of the bulk copy are 16-byte aligned
        const char* gsrc = reinterpret_cast<const char*>(pts + (size_t)cLo * 3);
        const unsigned bytes = (unsigned)nCloud * 12u;
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(gsrc) & 15u);
        char* sdst = reinterpret_cast<char*>(sPts) + mis;
        const unsigned head = mis ? min(16u - mis, bytes) : 0u;
        const unsigned bulk = ((bytes - head) / 16u) * 16u;
        const unsigned tail = bytes - head - bulk;
        if (threadIdx.x == 0 && bulk > 0) bulk_copy_g2s(sdst + head, gsrc + head, bulk, &sh.stageBar);
        if (threadIdx.x >= 32 && threadIdx.x < 32 + (head + tail) / 4) {
            const unsigned e = threadIdx.x - 32;
            const unsigned off = e < head / 4 ? e * 4 : head + bulk + (e - head / 4) * 4;
            *reinterpret_cast<float*>(sdst + off) = *reinterpret_cast<const float*>(gsrc + off);
        }
        sPts = reinterpret_cast<float*>(sdst);
        if (bulk > 0) mbar_wait(&sh.stageBar, 0);
    }
    if (useLists) {
        if (threadIdx.x == 0) {
            int run = 0;
            for (int p = 1; p <= 27; ++p) {
                run += sh.phaseCellStart[p];
                sh.phaseCellStart[p] = run;
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < cellsPerCloud; c += kSelThreads) {
            int2 ce = stageCells ? sCells[c] : cells[c];
            if (ce.y > ce.x) {
                const int x = c / (n * n), y = (c / n) % n, z = c % n;
                const int p = kPhaseOfOffset[(x % 3) * 9 + (y % 3) * 3 + (z % 3)];
                cellList[sh.phaseCellStart[p] + atomicAdd(&sh.phaseCellFill[p], 1)] = (unsigned short)c;
            }
        }
    }
    __syncthreads();

    CloudCtx cx;
    cx.cells = cells;
    cx.sCells = stageCells ? sCells : nullptr;
    cx.pts = pts;
    cx.sPts = stagePts ? sPts : nullptr;
    cx.sFlags = inSmem ? sFlags : nullptr;
    cx.sAccIdx = (stagePts && stageCells) ? sAccIdx : nullptr;
    cx.sAccCnt = sAccCnt;
    cx.gFlags = selGlobal;
    cx.cLo = cLo;
    cx.n = n;
    cx.offX = lane < 27 ? kOffsPool[lane][0] : 0;
    cx.offY = lane < 27 ? kOffsPool[lane][1] : 0;
    cx.offZ = lane < 27 ? kOffsPool[lane][2] : 0;
    {
        const float ext = max_extent(mn, mx, b);
        cx.thr = exact_radius_threshold(scaleInv ? __fmul_rn(radius, ext) : radius);
    }

    const float invN = 1.0f / (float)n, invNN = 1.0f / (float)(n * n);
    if (nCloud > 0) {
        for (int ph = 0; ph < 27 && !phasesDone; ++ph) {
            if (useLists) {
                const int beg = sh.phaseCellStart[ph], end = sh.phaseCellStart[ph + 1];
                for (int t = beg + warp; t < end; t += kSelWarps) {
                    const int c = cellList[t];
                    // c < 8192: exact in float, cheaper than three integer divisions on the serial path
                    const int x = (int)(((float)c + 0.5f) * invNN), rem = c - x * n * n;
                    const int y = (int)(((float)rem + 0.5f) * invN), z = rem - y * n;
                    process_cell(cx, sh, warp, lane, c, x, y, z, cx.cell(c), ph);
                }
                if (end == beg) continue;  // nothing was written in this phase: no barrier needed (uniform)
            } else {
                const int ox = kOffsPool[ph][0], oy = kOffsPool[ph][1], oz = kOffsPool[ph][2];
                // cells of this phase: 3t+1+o < n  (only the upper bound exists, poisson_sampling.cu:70-74)
                const int tX = (n + 1 - ox) / 3, tY = (n + 1 - oy) / 3, tZ = (n + 1 - oz) / 3;
                const int total = tX * tY * tZ;
                for (int c = warp; c < total; c += kSelWarps) {
                    const int tx = c / (tY * tZ), rem = c - tx * (tY * tZ);
                    const int ty = rem / tZ, tz = rem - ty * tZ;
                    const int x = tx * 3 + 1 + ox, y = ty * 3 + 1 + oy, z = tz * 3 + 1 + oz;
                    const int2 own = cx.cell((x * n + y) * n + z);
                    if (own.y <= own.x) continue;
                    process_cell(cx, sh, warp, lane, (x * n + y) * n + z, x, y, z, own, ph);
                }
            }
            __syncthreads();
        }

        // ---- stable compaction by (phase, position): tmpIdx[cLo + rank] = position
        for (int t = threadIdx.x; t < nCloud; t += kSelThreads) {
            int f = cx.flag(cLo + t);
            if (f) atomicAdd(&sh.phaseBase[f], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // phaseBase[f] <- first output slot of phase f-1; phaseBase[0] <- cloud total
            int run = 0;
            for (int f = 1; f <= 27; ++f) {
                int cntF = sh.phaseBase[f];
                sh.phaseBase[f] = run;
                run += cntF;
            }
            sh.phaseBase[0] = run;
        }
        __syncthreads();
        for (int t0 = 0; t0 < nCloud; t0 += kSelThreads) {
            for (int e = threadIdx.x; e < kSelWarps * 28; e += kSelThreads) (&sh.warpCnt[0][0])[e] = 0;
            __syncthreads();
            const int t = t0 + threadIdx.x;
            const int f = t < nCloud ? cx.flag(cLo + t) : 0;
            const unsigned grp = __match_any_sync(MC_FULL_MASK, f);
            if (f && lane == __ffs(grp) - 1) sh.warpCnt[warp][f] = __popc(grp);
            __syncthreads();
            if (f) {
                int pos = sh.phaseBase[f] + sh.running[f] + __popc(grp & ltMask);
                for (int w = 0; w < warp; ++w) pos += sh.warpCnt[w][f];
                if (cx.sAccIdx != nullptr)
                    cx.sAccIdx[pos] = (unsigned short)t;  // the accepted lists are dead by now: reuse as the order
                else
                    tmpIdx[cLo + pos] = cLo + t;
            }
            __syncthreads();
            if (threadIdx.x >= 1 && threadIdx.x <= 27) {
                int add = 0;
                for (int w = 0; w < kSelWarps; ++w) add += sh.warpCnt[w][threadIdx.x];
                sh.running[threadIdx.x] += add;
            }
            __syncthreads();
        }
    }
    const int mCloud = nCloud > 0 ? sh.phaseBase[0] : 0;

    // ---- exclusive prefix of the per-cloud sample counts (decoupled look-back, lookback.cuh)
    if (warp == 0) {
        int base = lookback_exclusive_prefix(chain, b, mCloud);
        if (lane == 0) {
            if (b == B - 1) *dNumSamples = base + mCloud;
            sh.base = base;
        }
    }
    __syncthreads();
    const int base = sh.base;
    for (int t = threadIdx.x; t < mCloud; t += kSelThreads) {
        const int i = cx.sAccIdx != nullptr ? cLo + cx.sAccIdx[t] : tmpIdx[cLo + t];
        const int pos = base + t;
        if (pos >= capOut) {
            if (dOverflow != nullptr) *dOverflow = 1;
            continue;
        }
        float px, py, pz;
        cx.point(i, px, py, pz);
        outPts[(size_t)pos * 3] = px;
        outPts[(size_t)pos * 3 + 1] = py;
        outPts[(size_t)pos * 3 + 2] = pz;
        outBids[pos] = b;
        outIdx[pos] = i;
    }
}

// ---- fine grids (more than kCellListCap cells per cloud: MCSeg's r = 0.025 levels, ScanNet rooms) ----------------
// One CTA per cloud would walk tens of thousands of mostly empty cells per phase with 16 warps (measured: 14 ms for
// 8 rooms x 80 000 points).  Instead the non-empty cells of the whole batch are listed per phase once, and every phase
// is ONE launch over its list with a warp per cell across the whole GPU; the accepted flags live in global memory
// (selGlobal) and kernel boundaries order the phases.  The same process_cell does the work; poisson_kernel with
// phasesDone = true then orders and writes the samples.
__global__ void __launch_bounds__(256) poisson_list_cells_kernel(int totalCells, int n, bool fill,
                                                                 const int2* __restrict__ cellIdx,
                                                                 int* __restrict__ phaseCount,   // [27] counts
                                                                 int* __restrict__ phaseFill,    // [27] cursors
                                                                 int* __restrict__ cellList) {
    __shared__ unsigned char sPhaseOf[27];
    __shared__ int sStart[28];
    if (threadIdx.x < 27) sPhaseOf[threadIdx.x] = kPhaseOfOffset[threadIdx.x];
    if (fill && threadIdx.x == 0) {
        int run = 0;
        for (int p = 0; p < 27; ++p) {
            sStart[p] = run;
            run += phaseCount[p];
        }
    }
    __syncthreads();
    const int cellsPerCloud = n * n * n;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < totalCells; g += gridDim.x * blockDim.x) {
        const int2 ce = cellIdx[g];
        if (ce.y <= ce.x) continue;
        const int c = g % cellsPerCloud;
        const int x = c / (n * n), y = (c / n) % n, z = c % n;
        const int p = sPhaseOf[(x % 3) * 9 + (y % 3) * 3 + (z % 3)];
        if (!fill)
            atomicAdd(&phaseCount[p], 1);
        else
            cellList[sStart[p] + atomicAdd(&phaseFill[p], 1)] = g;
    }
}

__global__ void __launch_bounds__(kSelThreads)
    poisson_phase_kernel(int ph, bool scaleInv, float radius, int n, const float* __restrict__ mn,
                         const float* __restrict__ mx, const float* __restrict__ pts, const int2* __restrict__ cellIdx,
                         unsigned char* selGlobal, const int* __restrict__ phaseCount,
                         const int* __restrict__ cellList) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SelShared& sh = *reinterpret_cast<SelShared*>(smemRaw);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int beg = 0;
    for (int p = 0; p < ph; ++p) beg += phaseCount[p];
    const int end = beg + phaseCount[ph];
    const int cellsPerCloud = n * n * n;
    CloudCtx cx;
    cx.sCells = nullptr;
    cx.pts = pts;
    cx.sPts = nullptr;
    cx.sAccIdx = nullptr;
    cx.sAccCnt = nullptr;
    cx.sFlags = nullptr;
    cx.gFlags = selGlobal;
    cx.cLo = 0;
    cx.n = n;
    cx.offX = lane < 27 ? kOffsPool[lane][0] : 0;
    cx.offY = lane < 27 ? kOffsPool[lane][1] : 0;
    cx.offZ = lane < 27 ? kOffsPool[lane][2] : 0;
    int cloud = -1;
    for (int t = beg + blockIdx.x * kSelWarps + warp; t < end; t += gridDim.x * kSelWarps) {
        const int g = cellList[t];
        const int b = g / cellsPerCloud, c = g - b * cellsPerCloud;
        if (b != cloud) {  // per-cloud frame: cell rows and the exact squared-radius threshold
            cloud = b;
            cx.cells = cellIdx + (size_t)b * cellsPerCloud;
            const float ext = max_extent(mn, mx, b);
            cx.thr = exact_radius_threshold(scaleInv ? __fmul_rn(radius, ext) : radius);
        }
        const int x = c / (n * n), rem = c - x * n * n, y = rem / n, z = rem - y * n;
        process_cell(cx, sh, warp, lane, c, x, y, z, cx.cells[c], ph);
    }
}

constexpr size_t sel_smem_bytes(const SelCaps& c) {
    return sizeof(SelShared) + c.flagBytes + c.cellList * sizeof(unsigned short) + c.stageCells * sizeof(int2) +
           (c.stagePts * 3 + 4) * sizeof(float) + c.stagePts * sizeof(unsigned short) + c.stageCells;
}
constexpr SelCaps kSelMaxCaps = {kSelFlagBytes, kCellListCap, kStageCells, kStagePts};
constexpr size_t kSelSmemBytes = sel_smem_bytes(kSelMaxCaps);
constexpr int round16(long long v) { return (int)((v + 15) / 16 * 16); }

// Capacities for a batch of numPoints points in batchSize clouds on an n^3 grid: the cell structures are sized
// exactly; points are staged up to twice the mean cloud (ragged batches: a larger cloud reads its points through
// L1 / L2 instead) and flags up to eight times.  MCCNN_POISSON_STAGE=full asks for the largest capacities always.
SelCaps choose_caps(long long numPoints, int batchSize, long long cellsPerCloud, bool phasesDone) {
    static const bool full = [] {
        const char* e = getenv("MCCNN_POISSON_STAGE");
        return e != nullptr && strcmp(e, "full") == 0;
    }();
    if (phasesDone) return SelCaps{0, 0, 0, 0};  // the ordering pass of the fine-grid path stages nothing
    if (full) return kSelMaxCaps;
    const long long mean = (numPoints + batchSize - 1) / batchSize;
    SelCaps c;
    c.stagePts = round16(std::min<long long>(kStagePts, std::max<long long>(256, 2 * mean)));
    c.flagBytes = round16(std::min<long long>(kSelFlagBytes, std::max<long long>(c.stagePts, std::min(numPoints, 8 * mean))));
    c.cellList = cellsPerCloud <= kCellListCap ? round16(cellsPerCloud) : 0;
    c.stageCells = cellsPerCloud <= kStageCells ? round16(cellsPerCloud) : 0;
    return c;
}

}  // namespace
}  // namespace mc

using namespace mc;

// workspace ints: look-back descriptors (2 * (B + 2)) | flag bytes (N/4 + 1 ints) | tmpIdx[N] | 64 counters | list[N]
extern "C" long long mc_poisson_workspace_ints(int numPoints, int batchSize) {
    long long N = numPoints > 0 ? numPoints : 1;
    return 2LL * (batchSize + 2) + (N / 4 + 1) + N + 8 + (N + 64);  // ... | phase counters / cursors | cell list
}

// samplePointCloud + copyPoints, tf_ops/poisson_sampling.cu:175-244.  Outputs hold capacityOut rows (numPoints
// always suffices); the live number of samples is written to *dNumSamples (device); samples past the capacity are
// dropped and *dOverflow (optional) is set.  numPoints / dNumPoints only size the workspace: the live extent of every
// cloud is read from the cell table.
extern "C" int mc_poisson_sampling(int scaleInv, float radius, int numPoints, const int* dNumPoints, int batchSize,
                                   int numCells, const float* aabbMin, const float* aabbMax, const float* pts,
                                   const int* batchIds, const int* cellIdx, int* workspace, int capacityOut,
                                   float* outPts, int* outBatchIds, int* outIdx, int* dNumSamples, int* dOverflow,
                                   void* stream) {
    (void)dNumPoints;
    (void)batchIds;
    MC_REQUIRE(radius > 0.0f, "mc_poisson_sampling: radius must be positive");
    MC_REQUIRE(batchSize > 0 && numCells > 0, "mc_poisson_sampling: batch size and cell count must be positive");
    MC_REQUIRE(numCells <= 1290, "mc_poisson_sampling: grid too fine (cell index would overflow)");
    MC_REQUIRE(dNumSamples != nullptr, "mc_poisson_sampling: dNumSamples required");
    cudaStream_t s = (cudaStream_t)stream;
    // the opt-in to more than 48 KB of dynamic shared memory is per device (and cheap): set it on every call
    cudaFuncSetAttribute(poisson_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmemBytes);
    const long long N = numPoints > 0 ? numPoints : 1;
    MC_REQUIRE(((uintptr_t)workspace) % 8 == 0, "mc_poisson_sampling: workspace must be 8-byte aligned");
    const int descInts = 2 * (batchSize + 2);
    unsigned long long* chain = reinterpret_cast<unsigned long long*>(workspace);
    unsigned char* sel = reinterpret_cast<unsigned char*>(workspace + descInts);
    int* tmpIdx = workspace + descInts + (N / 4 + 1);
    cudaMemsetAsync(workspace, 0, sizeof(int) * (size_t)(descInts + (N / 4 + 1)), s);  // descriptors + flags
    const long long cellsPerCloud = (long long)numCells * numCells * numCells;
    const bool fineGrid = cellsPerCloud > kCellListCap && !kForceSingleLaunch;
    if (numPoints > 0 && fineGrid) {
        // non-empty cells of the batch per phase, then one launch per phase over the whole GPU
        MC_REQUIRE(cellsPerCloud * batchSize < (1LL << 31), "mc_poisson_sampling: cell table too large");
        const int totalCells = (int)(cellsPerCloud * batchSize);
        int* counters = tmpIdx + N;          // [0..27) counts, [32..59) cursors
        int* cellList = counters + 64;       // <= numPoints non-empty cells
        cudaMemsetAsync(counters, 0, sizeof(int) * 64, s);
        const int listGrid = std::min(cdiv(totalCells, 256), 148 * 8);
        const int2* cells2 = reinterpret_cast<const int2*>(cellIdx);
        poisson_list_cells_kernel<<<listGrid, 256, 0, s>>>(totalCells, numCells, false, cells2, counters,
                                                           counters + 32, cellList);
        poisson_list_cells_kernel<<<listGrid, 256, 0, s>>>(totalCells, numCells, true, cells2, counters,
                                                           counters + 32, cellList);
        // a warp per listed cell; the list length is only known on the device: a grid that covers the GPU a few
        // times over and strides (at most numPoints / 27 cells per phase on average)
        const int phaseGrid = std::max(1, std::min(cdiv(cdiv(numPoints, 8), kSelWarps), 148 * 4));
        for (int ph = 0; ph < 27; ++ph)
            poisson_phase_kernel<<<phaseGrid, kSelThreads, sizeof(SelShared), s>>>(
                ph, scaleInv != 0, radius, numCells, aabbMin, aabbMax, pts, cells2, sel, counters, cellList);
        count_launches(29);
    }
    if (numPoints > 0) {
        const SelCaps caps = choose_caps(numPoints, batchSize, cellsPerCloud, fineGrid);
        poisson_kernel<<<batchSize, kSelThreads, sel_smem_bytes(caps), s>>>(
            scaleInv != 0, radius, batchSize, numCells, aabbMin, aabbMax, pts, reinterpret_cast<const int2*>(cellIdx),
            sel, tmpIdx, chain, capacityOut, outPts, outBatchIds, outIdx, dNumSamples, dOverflow, fineGrid, caps);
    } else {
        cudaMemsetAsync(dNumSamples, 0, sizeof(int), s);
    }
    return check_launch("mc_poisson_sampling");
}
