/*
 * mccnn_oracle.c -- CPU restatement of the MCCNN Monte-Carlo-convolution forward path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA kernels in
 * mccnn_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (mccnn_b200/) never does.
 *
 * Every function restates one host launcher + kernel(s) of the reference
 * (/root/reference/tf_ops, cited per function as file:line) as a single-threaded,
 * deterministic C loop nest.  The geometry predicates reproduce the float operation
 * order of the reference's compiled SASS (nvcc -O2, default -fmad=true):
 *     dist(p,q) = sqrtf(fmaf(dz,dz, fmaf(dx,dx, dy*dy)))       [verified with cuobjdump, see DESIGN.md]
 * so that every index-valued output is bit-exact.  Build with -ffp-contract=off so that the
 * host compiler introduces no contraction of its own.
 *
 * Canonical tie-breaks where the reference is order-nondeterministic (atomic arrival order):
 *   - counting sort: stable (ascending original index inside a cell)      sort_gpu.cu:157-174
 *   - poisson output: (batch, phase, ascending sorted position)            poisson_sampling.cu:113-120
 * Any reference output is a within-cell / within-(batch,phase) permutation of these.
 *
 * Parity pin: the reference has no tests or golden vectors.  The pin is
 * tests/golden/: outputs of the reference's own CUDA kernels (oracle/_ref, built from
 * /root/reference/tf_ops/*.cu) run on a B200, committed with the generating script.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_BLOCK_MLP 8 /* BLOCK_MLP_SIZE, tf_ops/genCompileScript.py:20 */

/* find_neighbors.cu:282-291 */
static const int OFFS_NEIGH[27][3] = {
    {1, 1, 1},   {0, 1, 1},   {-1, 1, 1},  {1, 0, 1},   {0, 0, 1},   {-1, 0, 1},  {1, -1, 1},
    {0, -1, 1},  {-1, -1, 1}, {1, 1, 0},   {0, 1, 0},   {-1, 1, 0},  {1, 0, 0},   {0, 0, 0},
    {-1, 0, 0},  {1, -1, 0},  {0, -1, 0},  {-1, -1, 0}, {1, 1, -1},  {0, 1, -1},  {-1, 1, -1},
    {1, 0, -1},  {0, 0, -1},  {-1, 0, -1}, {1, -1, -1}, {0, -1, -1}, {-1, -1, -1}};

/* poisson_sampling.cu:192-196 */
static const int OFFS_POOL[27][3] = {
    {1, 1, -1},  {0, -1, 1}, {0, 1, 1},   {0, 1, 0},  {0, 0, 1},   {0, -1, 0},   {-1, 1, -1},
    {0, -1, -1}, {1, 0, 0},  {1, -1, 1},  {1, 0, 1},  {-1, 1, 1},  {-1, 0, 0},   {1, -1, -1},
    {0, 1, -1},  {-1, -1, 0}, {-1, 1, 0}, {0, 0, 0},  {0, 0, -1},  {1, 1, 0},    {1, 0, -1},
    {1, -1, 0},  {-1, 0, 1}, {1, 1, 1},   {-1, 0, -1}, {-1, -1, -1}, {-1, -1, 1}};

/* max extent of box b: sort_gpu.cu:50-52, find_neighbors.cu:67-70 (fmaxf == CUDA max(float,float)) */
static float max_extent(const float *mn, const float *mx, int b) {
    float ex = mx[b * 3] - mn[b * 3];
    float ey = mx[b * 3 + 1] - mn[b * 3 + 1];
    float ez = mx[b * 3 + 2] - mn[b * 3 + 2];
    return fmaxf(fmaxf(ex, ey), ez);
}

/* max(min((int)floor(v), n-1), 0) with CUDA's saturating float->int semantics (NaN -> 0). sort_gpu.cu:55-57 */
static int clamp_cell(float v, int n) {
    float f = floorf(v);
    if (!(f >= 0.0f)) return 0; /* negatives and NaN */
    if (f >= (float)n) return n - 1;
    return (int)f;
}

static float scaled_radius(int scaleInv, float radius, float ext) {
    return scaleInv ? radius * ext : radius; /* find_neighbors.cu:72 */
}

/* the radius predicate: find_neighbors.cu:95-97, poisson_sampling.cu:104-106 (SASS contraction order) */
static int within(const float *p, const float *q, float r) {
    float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
    float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
    return sqrtf(d2) < r;
}

/* ------------------------------------------------------------------------------------------------
 * A1  computeAABB / comp_AABB                                      tf_ops/aabb_gpu.cu:57-141
 * ---------------------------------------------------------------------------------------------- */
void orc_compute_aabb(int scaleInv, int N, int B, const float *pts, const int *bids, float *mn, float *mx) {
    for (int i = 0; i < B * 3; ++i) {
        mn[i] = FLT_MAX;
        mx[i] = -FLT_MAX;
    }
    if (scaleInv) {
        for (int i = 0; i < N; ++i) {
            int b = bids[i];
            for (int a = 0; a < 3; ++a) {
                mn[b * 3 + a] = fminf(mn[b * 3 + a], pts[i * 3 + a]);
                mx[b * 3 + a] = fmaxf(mx[b * 3 + a], pts[i * 3 + a]);
            }
        }
    } else { /* every row receives the box of the whole batch: aabb_gpu.cu:104-114 */
        float gmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, gmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < 3; ++a) {
                gmn[a] = fminf(gmn[a], pts[i * 3 + a]);
                gmx[a] = fmaxf(gmx[a], pts[i * 3 + a]);
            }
        for (int b = 0; b < B; ++b)
            for (int a = 0; a < 3; ++a) {
                mn[b * 3 + a] = gmn[a];
                mx[b * 3 + a] = gmx[a];
            }
    }
}

/* ------------------------------------------------------------------------------------------------
 * A2  determineNumCells                                            tf_ops/sort_gpu.cu:374-420
 * ---------------------------------------------------------------------------------------------- */
int orc_num_cells(int scaleInv, float cellSize, const float *mn, const float *mx) {
    int n;
    if (scaleInv)
        n = (int)(1.0f / cellSize);
    else
        n = (int)(max_extent(mn, mx, 0) / cellSize); /* row 0 only: sort_gpu.cu:381-389 */
    return n == 0 ? 1 : n;
}

/* key of one point: calc_key, sort_gpu.cu:35-61 */
static int point_key(const float *p, int b, int n, const float *mn, const float *mx) {
    float cell = max_extent(mn, mx, b) / (float)n;
    int x = clamp_cell((p[0] - mn[b * 3]) / cell, n);
    int y = clamp_cell((p[1] - mn[b * 3 + 1]) / cell, n);
    int z = clamp_cell((p[2] - mn[b * 3 + 2]) / cell, n);
    return b * n * n * n + x * n * n + y * n + z;
}

/* ------------------------------------------------------------------------------------------------
 * A3  sortPointsStep1GPUKernel                                     tf_ops/sort_gpu.cu:436-471
 *     keys + destination index of a counting sort; canonical (stable) order inside a cell.
 * ---------------------------------------------------------------------------------------------- */
void orc_sort_step1(int N, int B, int n, const float *mn, const float *mx, const float *pts, const int *bids,
                    int *keys, int *newIdx) {
    long G = (long)B * n * n * n;
    int *cnt = (int *)calloc((size_t)G + 1, sizeof(int));
    for (int i = 0; i < N; ++i) {
        keys[i] = point_key(pts + 3 * i, bids[i], n, mn, mx);
        cnt[keys[i] + 1]++;
    }
    for (long g = 0; g < G; ++g) cnt[g + 1] += cnt[g];
    for (int i = 0; i < N; ++i) newIdx[i] = cnt[keys[i]]++;
    free(cnt);
}

/* ------------------------------------------------------------------------------------------------
 * A4  sortPointsStep2GPUKernel (move_points + save_indexs)         tf_ops/sort_gpu.cu:192-248,473-497
 * ---------------------------------------------------------------------------------------------- */
void orc_sort_step2(int N, int B, int C, int n, const float *pts, const int *bids, const float *feats,
                    const int *keys, const int *newIdx, float *outPts, int *outBids, float *outFeats,
                    int *cellIdx) {
    long G = (long)B * n * n * n;
    int *skeys = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    memset(cellIdx, 0, sizeof(int) * 2 * (size_t)G);
    for (int i = 0; i < N; ++i) {
        int d = newIdx[i];
        outPts[d * 3] = pts[i * 3];
        outPts[d * 3 + 1] = pts[i * 3 + 1];
        outPts[d * 3 + 2] = pts[i * 3 + 2];
        for (int c = 0; c < C; ++c) outFeats[(size_t)d * C + c] = feats[(size_t)i * C + c];
        skeys[d] = keys[i];
        outBids[d] = bids[i];
    }
    for (int i = 0; i < N; ++i) { /* save_indexs, sort_gpu.cu:225-248 */
        int k = skeys[i];
        if (i == 0)
            cellIdx[k * 2] = 0;
        else if (k != skeys[i - 1])
            cellIdx[k * 2] = i;
        if (i + 1 >= N)
            cellIdx[k * 2 + 1] = N;
        else if (k != skeys[i + 1])
            cellIdx[k * 2 + 1] = i + 1;
    }
    free(skeys);
}

/* A5  sort_features_back_grad == python sort_features             tf_ops/sort_gpu.cu:311-324 */
void orc_sort_features(int N, int C, const float *in, const int *newIdx, float *out) {
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < C; ++c) out[(size_t)newIdx[i] * C + c] = in[(size_t)i * C + c];
}

/* sort_features_back                                               tf_ops/sort_gpu.cu:288-301 */
void orc_sort_features_back(int N, int C, const float *in, const int *newIdx, float *out) {
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < C; ++c) out[(size_t)i * C + c] = in[(size_t)newIdx[i] * C + c];
}

/* A10 computeInverseIndexs + transformIndexs                       tf_ops/sort_gpu.cu:332-363,542-564 */
void orc_transform_indexs(int M, int N, const int *idx, const int *newIdx, int *out) {
    int *inv = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    for (int i = 0; i < N; ++i) inv[newIdx[i]] = i;
    for (int s = 0; s < M; ++s) out[s] = inv[idx[s]];
    free(inv);
}

/* ------------------------------------------------------------------------------------------------
 * A6  countNeighborsCPU + packNeighborsCPU                         tf_ops/find_neighbors.cu:40-113,199-264
 *     pairs == NULL: count only.  Returns K.  start[i] = exclusive prefix of the counts.
 * ---------------------------------------------------------------------------------------------- */
long orc_find_neighbors(int scaleInv, int M, int n, float radius, const float *samples, const int *sbids,
                        const float *pts2, const int *cellIdx, const float *mn, const float *mx, int *start,
                        int *pairs) {
    long K = 0;
    for (int i = 0; i < M; ++i) {
        int b = sbids[i];
        float ext = max_extent(mn, mx, b);
        float cell = ext / (float)n;
        float r = scaled_radius(scaleInv, radius, ext);
        const float *c = samples + 3 * i;
        int x = clamp_cell((c[0] - mn[b * 3]) / cell, n);
        int y = clamp_cell((c[1] - mn[b * 3 + 1]) / cell, n);
        int z = clamp_cell((c[2] - mn[b * 3 + 2]) / cell, n);
        start[i] = (int)K;
        for (int o = 0; o < 27; ++o) {
            int cx = x + OFFS_NEIGH[o][0], cy = y + OFFS_NEIGH[o][1], cz = z + OFFS_NEIGH[o][2];
            if (cx < 0 || cx >= n || cy < 0 || cy >= n || cz < 0 || cz >= n) continue;
            long flat = (long)b * n * n * n + (long)cx * n * n + cy * n + cz;
            int lo = cellIdx[flat * 2], hi = cellIdx[flat * 2 + 1];
            for (int j = lo; j < hi; ++j)
                if (within(c, pts2 + 3 * j, r)) {
                    if (pairs) {
                        pairs[K * 2] = j;
                        pairs[K * 2 + 1] = i;
                    }
                    ++K;
                }
        }
    }
    return K;
}

/* ------------------------------------------------------------------------------------------------
 * A7  computeDPFsCPU / computePDFs                                 tf_ops/compute_pdf.cu:40-94
 *     keeps the reference's mixed float/double evaluation.
 * ---------------------------------------------------------------------------------------------- */
void orc_compute_pdf(int scaleInv, float window, int M, int K, float radius, const float *pts, const int *bids,
                     const float *mn, const float *mx, const int *start, const int *pairs, float *pdf) {
    for (int p = 0; p < K; ++p) {
        int j = pairs[p * 2], i = pairs[p * 2 + 1];
        int b = bids[j];
        float r = scaled_radius(scaleInv, radius, max_extent(mn, mx, b));
        int lo = start[i];
        int hi = (i < M - 1) ? start[i + 1] : K;
        const float h = window;
        const float invH = 1 / h;
        const float invRadH = (float)(1.0 / (double)(r * h));
        float acc = 0.0f;
        for (int q = lo; q < hi; ++q) {
            const float *pq = pts + 3 * pairs[q * 2];
            float d0 = (pq[0] - pts[3 * j]) * invRadH;
            float d1 = (pq[1] - pts[3 * j + 1]) * invRadH;
            float d2 = (pq[2] - pts[3 * j + 2]) * invRadH;
            float g = (float)((double)invH * (0.39894228 * exp((-0.5) * (double)d0 * (double)d0)));
            g = (float)((double)(g * invH) * (0.39894228 * exp((-0.5) * (double)d1 * (double)d1)));
            g = (float)((double)(g * invH) * (0.39894228 * exp((-0.5) * (double)d2 * (double)d2)));
            acc += g;
        }
        pdf[p] = acc / ((float)hi - lo);
    }
}

/* ------------------------------------------------------------------------------------------------
 * A8  samplePointCloud / selectSamples                             tf_ops/poisson_sampling.cu:51-124,175-230
 *     outPhase (optional, may be NULL) receives the phase 0..26 in which each output was accepted.
 * ---------------------------------------------------------------------------------------------- */
int orc_poisson_sampling(int scaleInv, float radius, int N, int B, int n, const float *mn, const float *mx,
                         const float *pts, const int *bids, const int *cellIdx, float *outPts, int *outBids,
                         int *outIdx, int *outPhase) {
    (void)bids;
    unsigned char *sel = (unsigned char *)calloc((size_t)(N > 0 ? N : 1), 1);
    int M = 0;
    int groups = n / 3 + ((n % 3) ? 1 : 0);
    groups = (groups + 3) / 4 * 4; /* launch covers whole 4x4x4 blocks; extra threads fail the bounds test */
    for (int b = 0; b < B; ++b) {
        float r = scaled_radius(scaleInv, radius, max_extent(mn, mx, b));
        for (int ph = 0; ph < 27; ++ph) {
            for (int tx = 0; tx < groups; ++tx)
                for (int ty = 0; ty < groups; ++ty)
                    for (int tz = 0; tz < groups; ++tz) {
                        int x = tx * 3 + 1 + OFFS_POOL[ph][0];
                        int y = ty * 3 + 1 + OFFS_POOL[ph][1];
                        int z = tz * 3 + 1 + OFFS_POOL[ph][2];
                        if (!(x < n && y < n && z < n)) continue; /* only the upper bound: poisson_sampling.cu:74 */
                        long flat = (long)b * n * n * n + (long)x * n * n + y * n + z;
                        int lo = cellIdx[flat * 2], hi = cellIdx[flat * 2 + 1];
                        for (int i = lo; i < hi; ++i) {
                            int collision = 0;
                            for (int o = 0; o < 27 && !collision; ++o) {
                                int cx = x + OFFS_POOL[o][0], cy = y + OFFS_POOL[o][1], cz = z + OFFS_POOL[o][2];
                                if (cx < 0 || cx >= n || cy < 0 || cy >= n || cz < 0 || cz >= n) continue;
                                long f2 = (long)b * n * n * n + (long)cx * n * n + cy * n + cz;
                                int l2 = cellIdx[f2 * 2], h2 = cellIdx[f2 * 2 + 1];
                                for (int j = l2; j < h2 && !collision; ++j)
                                    if (within(pts + 3 * i, pts + 3 * j, r) && sel[j]) collision = 1;
                            }
                            if (!collision) {
                                sel[i] = 1;
                                outPts[M * 3] = pts[i * 3];
                                outPts[M * 3 + 1] = pts[i * 3 + 1];
                                outPts[M * 3 + 2] = pts[i * 3 + 2];
                                outBids[M] = b;
                                outIdx[M] = i;
                                if (outPhase) outPhase[M] = ph;
                                ++M;
                            }
                        }
                    }
        }
    }
    free(sel);
    return M;
}

/* A9  getFeaturesSampledPoints / selectFeatureSamples              tf_ops/poisson_sampling.cu:135-148 */
void orc_get_sampled_features(int M, int C, const int *idx, const float *feats, float *out) {
    for (int s = 0; s < M; ++s)
        for (int c = 0; c < C; ++c) out[(size_t)s * C + c] = feats[(size_t)idx[s] * C + c];
}

/* ------------------------------------------------------------------------------------------------
 * A11 spatialConvCPU / evaluateMLP[NoCombin]Kernel                 tf_ops/spatial_conv.cu:24-79,104-176,
 *                                                                   178-230,254-325,796-871
 *     Weight buffers are the flat reinterpretations the kernel uses (spatial_conv.cu:49-73,172):
 *       W1[(8q+t)*3+a], b1[8q+t], W2[q*64+t*8+u], b2[8q+t], Wo[(8q+t)*8+u], bo[8q+t].
 *     accumDouble = 0: float accumulation in pair order (one valid ordering of the reference's float
 *     atomics); 1: double accumulation (the order-free value both sides must approach); 2: the sum of the
 *     ABSOLUTE contributions per output (the scale against which a float32 sum's error is measured: outputs are
 *     signed sums that cancel, so an error relative to the result itself is unbounded for any summation order,
 *     the reference's included).  The scale also propagates magnitudes through the MLP, because a kernel value
 *     is itself a cancelling sum of O(1) products.
 * ---------------------------------------------------------------------------------------------- */
void orc_spatial_conv(int avg, int scaleInv, int K, int C, int Co, int M, int combin, float radius,
                      const float *pts, const int *bids, const float *feats, const float *pdfs,
                      const float *samples, const int *start, const int *pairs, const float *mn, const float *mx,
                      const float *W1, const float *b1, const float *W2, const float *b2, const float *Wo,
                      const float *bo, int accumDouble, float *out) {
    int nOut = combin ? Co * C : C;
    int outC = combin ? Co : C;
    int nb = nOut / ORC_BLOCK_MLP + ((nOut % ORC_BLOCK_MLP) ? 1 : 0);
    double *acc = (double *)calloc((size_t)(M > 0 ? M : 1) * outC, sizeof(double));
    float *accf = (float *)calloc((size_t)(M > 0 ? M : 1) * outC, sizeof(float));
    for (int p = 0; p < K; ++p) {
        int j = pairs[p * 2], i = pairs[p * 2 + 1];
        int b = bids[j];
        float r = scaled_radius(scaleInv, radius, max_extent(mn, mx, b));
        float d[3] = {(pts[j * 3] - samples[i * 3]) / r, (pts[j * 3 + 1] - samples[i * 3 + 1]) / r,
                      (pts[j * 3 + 2] - samples[i * 3 + 2]) / r};
        int lo = start[i];
        int hi = (i < M - 1) ? start[i + 1] : K;
        float cnt = avg ? (float)(hi - lo) : 1.0f;
        float w = pdfs[p] * cnt;
        for (int q = 0; q < nb; ++q) {
            float h1[ORC_BLOCK_MLP], h2[ORC_BLOCK_MLP];
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                float v = d[0] * W1[nr * 3] + d[1] * W1[nr * 3 + 1] + d[2] * W1[nr * 3 + 2] + b1[nr];
                h1[t] = v > 0.0f ? v : 0.0f;
            }
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                float a = 0.0f;
                for (int u = 0; u < ORC_BLOCK_MLP; ++u) a += h1[u] * W2[nr * ORC_BLOCK_MLP + u];
                a += b2[nr];
                h2[t] = a > 0.0f ? a : 0.0f;
            }
            /* accumDouble == 2 (error scale): magnitudes propagated through the three layers -- the sum of |terms| a
             * forward error analysis of the kernel value is relative to.  (|kappa| itself would hide the
             * cancellation INSIDE the small MLP: a kernel value near zero is a difference of O(1) products.) */
            double m1[ORC_BLOCK_MLP], m2[ORC_BLOCK_MLP];
            if (accumDouble == 2) {
                for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                    int nr = q * ORC_BLOCK_MLP + t;
                    m1[t] = fabs((double)d[0] * W1[nr * 3]) + fabs((double)d[1] * W1[nr * 3 + 1]) +
                            fabs((double)d[2] * W1[nr * 3 + 2]) + fabs((double)b1[nr]);
                }
                for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                    int nr = q * ORC_BLOCK_MLP + t;
                    double a = fabs((double)b2[nr]);
                    for (int u = 0; u < ORC_BLOCK_MLP; ++u) a += m1[u] * fabs((double)W2[nr * ORC_BLOCK_MLP + u]);
                    m2[t] = a;
                }
            }
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                if (nr >= nOut) continue;
                float a = 0.0f;
                for (int u = 0; u < ORC_BLOCK_MLP; ++u) a += h2[u] * Wo[nr * ORC_BLOCK_MLP + u];
                a += bo[nr];
                int c = nr % C;
                int o = combin ? nr / C : c;
                float contrib = (feats[(size_t)j * C + c] * a) / w;
                if (accumDouble == 2) {
                    double mk = fabs((double)bo[nr]);
                    for (int u = 0; u < ORC_BLOCK_MLP; ++u) mk += m2[u] * fabs((double)Wo[nr * ORC_BLOCK_MLP + u]);
                    acc[(size_t)i * outC + o] += fabs((double)feats[(size_t)j * C + c]) * mk / fabs((double)w);
                } else {
                    acc[(size_t)i * outC + o] += (double)contrib;
                }
                accf[(size_t)i * outC + o] += contrib;
            }
        }
    }
    for (size_t e = 0; e < (size_t)M * outC; ++e) out[e] = accumDouble ? (float)acc[e] : accf[e];
    free(acc);
    free(accf);
}

/* ------------------------------------------------------------------------------------------------
 * spatial_conv_grad  (tf_ops/spatial_conv.cu:327-460 computedconvj_d, :563-700 computedconvj_dNoCombin,
 *                     host spatialConvGradsCPU :873-966)
 *   for pair p = (j, i), block q, forward recomputed: z1 = W1 d + b1, z2 = W2 relu(z1) + b2 (:362-378);
 *   neuron n = 8q + t < nOut: c = n % C, o = combin ? n / C : c, f = feat[j,c], g = outGrad[i,o], w = pdf * cnt
 *     common = (f * g) / w                       dWo[n,u] += common * relu(z2[u]);  dbo[n] += common   (:386-397)
 *     dFeat[j,c] += g * (Wo[n] . relu(z2) + bo[n]) / w                                                  (:399-401)
 *   dz2[t] = (z2[t] >= 0) * (sum_{n in block, n < nOut} g_n f_n Wo[n,t]) / w                          (:404-416)
 *     dW2[t,u] += dz2[t] * relu(z1[u]);  db2[t] += dz2[t]                                               (:421-427)
 *   dz1[t] = (z1[t] >= 0) * sum_i dz2[i] W2[i,t]                                                        (:430-437)
 *     dW1[t,a] += dz1[t] * d[a];  db1[t] += dz1[t]                                                      (:442-447)
 *   Note the ReLU derivative is taken with >= (1 at exactly 0).  The reference accumulates with float atomicAdd in
 *   arbitrary order; here: double accumulators, rounded once (accumDouble) or float accumulation in pair order.
 *   All gradient buffers are zero-filled first; the reference leaves the padding neurons (>= nOut) of the output
 *   layer's gradient untouched (:905-906) -- canonical value 0.
 *   absSum != 0: instead of the gradients, the sums of |contributions| (the error scale of each float sum).
 * ---------------------------------------------------------------------------------------------- */
void orc_spatial_conv_grad(int avg, int scaleInv, int K, int C, int Co, int M, int N, int combin, float radius,
                           const float *pts, const int *bids, const float *feats, const float *pdfs,
                           const float *samples, const int *start, const int *pairs, const float *mn, const float *mx,
                           const float *W1, const float *b1, const float *W2, const float *b2, const float *Wo,
                           const float *bo, const float *outGrad, int absSum, float *dFeat, float *dW1, float *db1,
                           float *dW2, float *db2, float *dWo, float *dbo) {
    int nOut = combin ? Co * C : C;
    int outC = combin ? Co : C;
    int nb = nOut / ORC_BLOCK_MLP + ((nOut % ORC_BLOCK_MLP) ? 1 : 0);
    int nn = nb * ORC_BLOCK_MLP;
    size_t nF = (size_t)(N > 0 ? N : 1) * C;
    double *aF = (double *)calloc(nF, sizeof(double));
    double *aW1 = (double *)calloc((size_t)nn * 3, sizeof(double)), *ab1 = (double *)calloc(nn, sizeof(double));
    double *aW2 = (double *)calloc((size_t)nn * 8, sizeof(double)), *ab2 = (double *)calloc(nn, sizeof(double));
    double *aWo = (double *)calloc((size_t)nn * 8, sizeof(double)), *abo = (double *)calloc(nn, sizeof(double));
#define ORC_ACC(dst, v) ((dst) += absSum ? fabs((double)(v)) : (double)(v))
    for (int p = 0; p < K; ++p) {
        int j = pairs[p * 2], i = pairs[p * 2 + 1];
        int b = bids[j];
        float r = scaled_radius(scaleInv, radius, max_extent(mn, mx, b));
        float d[3] = {(pts[j * 3] - samples[i * 3]) / r, (pts[j * 3 + 1] - samples[i * 3 + 1]) / r,
                      (pts[j * 3 + 2] - samples[i * 3 + 2]) / r};
        int lo = start[i];
        int hi = (i < M - 1) ? start[i + 1] : K;
        float cnt = avg ? (float)(hi - lo) : 1.0f;
        float w = pdfs[p] * cnt;
        for (int q = 0; q < nb; ++q) {
            float z1[ORC_BLOCK_MLP], z2[ORC_BLOCK_MLP], dz2[ORC_BLOCK_MLP], dz1[ORC_BLOCK_MLP];
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                z1[t] = d[0] * W1[nr * 3] + d[1] * W1[nr * 3 + 1] + d[2] * W1[nr * 3 + 2] + b1[nr];
            }
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                float a = 0.0f;
                for (int u = 0; u < ORC_BLOCK_MLP; ++u) a += (z1[u] > 0.0f ? z1[u] : 0.0f) * W2[nr * ORC_BLOCK_MLP + u];
                z2[t] = a + b2[nr];
            }
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                if (nr >= nOut) continue;
                int c = nr % C, o = combin ? nr / C : c;
                float f = feats[(size_t)j * C + c], g = outGrad[(size_t)i * outC + o];
                float common = (f * g) / w;
                float a = 0.0f;
                for (int u = 0; u < ORC_BLOCK_MLP; ++u) {
                    float h2 = z2[u] > 0.0f ? z2[u] : 0.0f;
                    ORC_ACC(aWo[(size_t)nr * 8 + u], common * h2);
                    a += h2 * Wo[nr * ORC_BLOCK_MLP + u];
                }
                ORC_ACC(abo[nr], common);
                a += bo[nr];
                ORC_ACC(aF[(size_t)j * C + c], g * a / w);
            }
            int numOuts = nOut - q * ORC_BLOCK_MLP;
            if (numOuts > ORC_BLOCK_MLP) numOuts = ORC_BLOCK_MLP;
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                float a = 0.0f;
                for (int k = 0; k < numOuts; ++k) {
                    int nr = q * ORC_BLOCK_MLP + k;
                    int c = nr % C, o = combin ? nr / C : c;
                    a += outGrad[(size_t)i * outC + o] * feats[(size_t)j * C + c] * Wo[nr * ORC_BLOCK_MLP + t];
                }
                dz2[t] = ((z2[t] >= 0.0f ? 1.0f : 0.0f) * a) / w;
            }
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                for (int u = 0; u < ORC_BLOCK_MLP; ++u)
                    ORC_ACC(aW2[(size_t)nr * 8 + u], dz2[t] * (z1[u] > 0.0f ? z1[u] : 0.0f));
                ORC_ACC(ab2[nr], dz2[t]);
            }
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                float a = 0.0f;
                for (int k = 0; k < ORC_BLOCK_MLP; ++k) a += dz2[k] * W2[(q * ORC_BLOCK_MLP + k) * ORC_BLOCK_MLP + t];
                dz1[t] = (z1[t] >= 0.0f ? 1.0f : 0.0f) * a;
            }
            for (int t = 0; t < ORC_BLOCK_MLP; ++t) {
                int nr = q * ORC_BLOCK_MLP + t;
                for (int a3 = 0; a3 < 3; ++a3) ORC_ACC(aW1[(size_t)nr * 3 + a3], dz1[t] * d[a3]);
                ORC_ACC(ab1[nr], dz1[t]);
            }
        }
    }
#undef ORC_ACC
    for (size_t e = 0; e < (size_t)N * C; ++e) dFeat[e] = (float)aF[e];
    for (int e = 0; e < nn * 3; ++e) dW1[e] = (float)aW1[e];
    for (int e = 0; e < nn * 8; ++e) dW2[e] = (float)aW2[e], dWo[e] = (float)aWo[e];
    for (int e = 0; e < nn; ++e) db1[e] = (float)ab1[e], db2[e] = (float)ab2[e], dbo[e] = (float)abo[e];
    free(aF), free(aW1), free(ab1), free(aW2), free(ab2), free(aWo), free(abo);
}

/* sortPointsStep2GradGPUKernel (tf_ops/sort_gpu.cu:260-277,499-513): in*Grad[i] = out*Grad[newIdx[i]] */
void orc_sort_step2_grad(int N, int C, const int *newIdx, const float *outPtsGrad, const float *outFeatGrad,
                         float *inPtsGrad, float *inFeatGrad) {
    for (int i = 0; i < N; ++i) {
        int s = newIdx[i];
        for (int a = 0; a < 3; ++a) inPtsGrad[(size_t)i * 3 + a] = outPtsGrad[(size_t)s * 3 + a];
        for (int c = 0; c < C; ++c) inFeatGrad[(size_t)i * C + c] = outFeatGrad[(size_t)s * C + c];
    }
}

/* getFeaturesSampledPointsGradients (tf_ops/poisson_sampling.cu:158-171,260-274): zeros, then
 * featGrad[idx[s],:] = outGrad[s,:] */
void orc_get_sampled_features_grad(int M, int C, int N, const int *idx, const float *outGrad, float *featGrad) {
    memset(featGrad, 0, sizeof(float) * (size_t)N * C);
    for (int s = 0; s < M; ++s)
        for (int c = 0; c < C; ++c) featGrad[(size_t)idx[s] * C + c] = outGrad[(size_t)s * C + c];
}

int orc_block_size(void) { return ORC_BLOCK_MLP; }
