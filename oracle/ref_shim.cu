/*
 * ref_shim.cu -- extern "C" entry points around the UNMODIFIED reference host launchers.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (never loaded by mccnn_b200/).
 *
 * oracle/Makefile compiles the six reference translation units where they lie
 * (/root/reference/tf_ops/{aabb_gpu,sort_gpu,find_neighbors,compute_pdf,poisson_sampling,
 * spatial_conv}.cu, -O2 -DBLOCK_MLP_SIZE=8 as tf_ops/genCompileScript.py:27-32 does, plus
 * -gencode sm_100a) and links them with this file into oracle/_ref/libmccnn_ref.so.  The
 * reference launchers have C++ linkage; this file declares them exactly as the reference's TF glue
 * does (aabb_gpu.cc:38-40, sort_gpu.cc:105-182, find_neighbors.cc:45-78, compute_pdf.cc:43-55,
 * poisson_sampling.cc:68-107, spatial_conv.cc:98-128) and reproduces the buffer choreography of each
 * OpKernel::Compute (temp buffers, two-phase dynamic outputs, host syncs) so that both the results
 * and the timing are those of the reference's stock GPU path.  TensorFlow's allocate_temp is a pooled
 * allocator, so temp buffers come from a grow-only pool here (no cudaMalloc in the steady state).
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

// ---- reference launchers (definitions live in /root/reference/tf_ops/*.cu) ----
void computeAABB(const bool pScaleInv, const int pNumPoints, const int pBatchSize, const float* pPoints,
                 const int* pBatchIds, float* pAABBMin, float* pAABBMax);
int determineNumCells(const bool pScaleInv, const int pBatchSize, const float pCellSize, const float* pAABBMin,
                      const float* pAABBMax);
void computeAuxiliarBuffersSize(const int pBatchSize, const int pNumCells, int* PBufferSize1, int* PBufferSize2,
                                int* PBufferSize3);
void sortPointsStep1GPUKernel(const int pNumPoints, const int pBatchSize, const int pNumCells, const float* pAABBMin,
                              const float* pAABBMax, const float* pPoints, const int* pBatchIds,
                              int* pAuxBuffCounters, int* pAuxBuffOffsets, int* pAuxBuffOffsets2, int* pKeys,
                              int* pNewIndexs);
void sortPointsStep2GPUKernel(const int pNumPoints, const int pBatchSize, const int pNumFeatures,
                              const int pNumCells, const float* pPoints, const int* pBatchIds,
                              const float* pFeatures, const int* pKeys, const int* pNewIndexs, int* pAuxBuffer,
                              float* pOutPoints, int* pOutBatchIds, float* pOutFeatures, int* pOutCellIndexs);
void sortFeaturesBack(const int pNumPoints, const int pNumFeatures, const float* pInFeatures, const int* pIndexs,
                      float* pOutFeatures);
void sortFeaturesBackGrad(const int pNumPoints, const int pNumFeatures, const float* pOutFeatureGrads,
                          const int* pIndexs, float* pInFeatureGrads);
void computeInverseIndexs(const int pNumPoints, const int* pIndexs, int* pOutIndexs);
void transformIndexs(const int pNumIndexs, const int pNumPoints, const int* pInStartIndexs,
                     const int* pInNewIndexs, int* pOutIndexs);
unsigned int countNeighborsCPU(const bool pScaleInv, const int pNumPoints, const int pNumCells, const float pRadius,
                               const float* pInPts, const int* pInBatchIds, const float* pInPts2,
                               const int* pCellIndexs, const float* pAABBMin, const float* pAABBMax,
                               int* pStartIndex);
void computeAuxiliarBuffersSize(const int pNumPoints, int* PBufferSize1, int* PBufferSize2);
void packNeighborsCPU(const bool pScaleInv, const int pNumPoints, const int pNumNeighbors, const int pNumCells,
                      const float pRadius, const float* pInPts, const int* pInBatchIds, const float* pInPts2,
                      const int* pCellIndexs, const float* pAABBMin, const float* pAABBMax, int* pAuxBuffOffsets,
                      int* pAuxBuffOffsets2, int* pStartIndexs, int* pPackedIndexs);
void computeDPFsCPU(const bool scaleInv, const float pWindow, const int numSamples, const int pNumNeighbors,
                    const float pRadius, const float* pInPts, const int* pInBatchIds, const float* pAABBMin,
                    const float* pAABBMax, const int* pStartIndexs, const int* pPackedIndexs, float* pPDFs);
int samplePointCloud(const bool scaleInv, const float pRadius, const int pNumPoints, const int pBatchSize,
                     const int pNumCells, const float* pAABBMin, const float* pAABBMax, const float* pPoints,
                     const int* pBatchIds, const int* pCellIndexs, float* pSelectedPts, int* pSelectedBatchIds,
                     int* pSelectedIndexs, bool* pAuxBoolBuffer);
void copyPoints(float* pSelectedPts, int* pSelectedBatchIds, int* pSelectedIndexs, const int pNumPts,
                float* pDestPts, int* pDestBatchIds, int* pDestIndexs);
void getFeaturesSampledPoints(int pNumPoints, int pNumFeatures, int pNumSampledPoints, const int* pInPointsIndexs,
                              const float* pInFeature, float* pOutSelFeatures);
void spatialConvCPU(bool pAvg, bool pScaleInv, int pNumNeighbors, int pNumInFeatures, int pNumOutFeatures,
                    int pNumSamples, bool pCombin, float pRadius, const float* pInPoints, const int* pBatchIds,
                    const float* pInFeatures, const float* pPDFs, const float* pSamples, const int* pStartIndexs,
                    const int* pPackedNeighs, const float* pAABBMin, const float* pAABBMax, const float* pWeights1,
                    const float* pBiases1, const float* pWeights2, const float* pBiases2, const float* pWeightsOut,
                    const float* pBiasesOut, float* pOutFeatues);

void spatialConvGradsCPU(bool pAvg, bool pScaleInv, int pNumNeighbors, int pNumInFeatures, int pNumOutFeatures,
                         int pNumSamples, int pNumPoints, bool pCombin, float pRadius, const float* pInPoints,
                         const int* pBatchIds, const float* pInFeatures, const float* pPDFs, const float* pSamples,
                         const int* pStartIndexs, const int* pPackedNeighs, const float* pAABBMin,
                         const float* pAABBMax, const float* pWeights1, const float* pBiases1,
                         const float* pWeights2, const float* pBiases2, const float* pWeightsOut,
                         const float* pBiasesOut, const float* pInOutFeatueGrads, float* pOutFeatureGrads,
                         float* pWeights1Grads, float* pWeight2Grads, float* pWeightOutGrads, float* pBiases1Grads,
                         float* pBiases2Grads, float* pBiasesOutGrads);
void sortPointsStep2GradGPUKernel(const int pNumPoints, const int pNumFeatures, const float* pOutGradients,
                                  const float* pOutFeatureGradients, const int* pNewIndexs, float* pInGradients,
                                  float* pInFeatureGradients);
void getFeaturesSampledPointsGradients(int pNumPoints, int pNumFeatures, int pNumSampledPoints,
                                       const int* pInPointsIndexs, const float* pInOutFeatureGrad,
                                       float* pOutInFeaturesGradients);

namespace {
// grow-only temp pool standing in for TF's allocate_temp
struct Slot {
    void* p = nullptr;
    size_t cap = 0;
};
Slot g_slots[8];
void* temp(int slot, size_t bytes) {
    if (bytes == 0) bytes = 4;
    Slot& s = g_slots[slot];
    if (s.cap < bytes) {
        if (s.p) cudaFree(s.p);
        size_t cap = bytes + bytes / 2;
        if (cudaMalloc(&s.p, cap) != cudaSuccess) {
            fprintf(stderr, "ref_shim: cudaMalloc(%zu) failed\n", cap);
            s.p = nullptr;
            s.cap = 0;
            return nullptr;
        }
        s.cap = cap;
    }
    return s.p;
}
}  // namespace

extern "C" {

// aabb_gpu.cc:51-84
void ref_compute_aabb(int scaleInv, int N, int B, const float* pts, const int* bids, float* mn, float* mx) {
    computeAABB(scaleInv != 0, N, B, pts, bids, mn, mx);
}

// sort_gpu.cc:239, :344
int ref_num_cells(int scaleInv, int B, float cellSize, const float* mn, const float* mx) {
    return determineNumCells(scaleInv != 0, B, cellSize, mn, mx);
}

// SortPointsStep1Op::Compute, sort_gpu.cc:184-268.  Returns numCells.
int ref_sort_points_step1(int scaleInv, int N, int B, float cellSize, const float* mn, const float* mx,
                          const float* pts, const int* bids, int* keys, int* newIdx) {
    int n = determineNumCells(scaleInv != 0, B, cellSize, mn, mx);
    int s1, s2, s3;
    computeAuxiliarBuffersSize(B, n, &s1, &s2, &s3);
    int* b1 = (int*)temp(0, sizeof(int) * (size_t)s1);
    int* b2 = (int*)temp(1, sizeof(int) * (size_t)s2);
    int* b3 = (int*)temp(2, sizeof(int) * (size_t)s3);
    sortPointsStep1GPUKernel(N, B, n, mn, mx, pts, bids, b1, b2, b3, keys, newIdx);
    return n;
}

// SortPointsStep2Op::Compute, sort_gpu.cc:270-378.  cellIdx must hold B*n^3*2 ints (n from ref_num_cells).
int ref_sort_points_step2(int scaleInv, int N, int B, int C, float cellSize, const float* mn, const float* mx,
                          const float* pts, const int* bids, const float* feats, const int* keys,
                          const int* newIdx, float* outPts, int* outBids, float* outFeats, int* cellIdx) {
    int n = determineNumCells(scaleInv != 0, B, cellSize, mn, mx);
    int* aux = (int*)temp(0, sizeof(int) * (size_t)N);
    sortPointsStep2GPUKernel(N, B, C, n, pts, bids, feats, keys, newIdx, aux, outPts, outBids, outFeats, cellIdx);
    return n;
}

// python sort_features -> SortFeaturesBackGradOp, sort_gpu.cc:438-490
void ref_sort_features(int N, int C, const float* in, const int* newIdx, float* out) {
    sortFeaturesBackGrad(N, C, in, newIdx, out);
}
// SortFeaturesBackOp, sort_gpu.cc:380-436
void ref_sort_features_back(int N, int C, const float* in, const int* newIdx, float* out) {
    sortFeaturesBack(N, C, in, newIdx, out);
}
// TransformIndexsOp::Compute, sort_gpu.cc:496-533
void ref_transform_indexs(int M, int N, const int* idx, const int* newIdx, int* out) {
    int* inv = (int*)temp(0, sizeof(int) * (size_t)N);
    computeInverseIndexs(N, newIdx, inv);
    transformIndexs(M, N, idx, inv, out);
}

// FindNeighborsOp::Compute part 1, find_neighbors.cc:146-155 (blocking D2H inside). Returns K.
unsigned int ref_find_neighbors_count(int scaleInv, int M, int n, float radius, const float* samples,
                                      const int* sbids, const float* pts2, const int* cellIdx, const float* mn,
                                      const float* mx, int* start) {
    return countNeighborsCPU(scaleInv != 0, M, n, radius, samples, sbids, pts2, cellIdx, mn, mx, start);
}
// FindNeighborsOp::Compute part 2, find_neighbors.cc:157-177
void ref_find_neighbors_pack(int scaleInv, int M, unsigned int K, int n, float radius, const float* samples,
                             const int* sbids, const float* pts2, const int* cellIdx, const float* mn,
                             const float* mx, int* start, int* pairs) {
    int s1, s2;
    computeAuxiliarBuffersSize(M, &s1, &s2);
    int* b1 = (int*)temp(0, sizeof(int) * (size_t)s1);
    int* b2 = (int*)temp(1, sizeof(int) * (size_t)s2);
    packNeighborsCPU(scaleInv != 0, M, (int)K, n, radius, samples, sbids, pts2, cellIdx, mn, mx, b1, b2, start,
                     pairs);
}

// ComputePDFOp::Compute, compute_pdf.cc:57-142
void ref_compute_pdf(int scaleInv, float window, int M, int K, float radius, const float* pts, const int* bids,
                     const float* mn, const float* mx, const int* start, const int* pairs, float* pdf) {
    computeDPFsCPU(scaleInv != 0, window, M, K, radius, pts, bids, mn, mx, start, pairs, pdf);
}

// PoissonSamplingOp::Compute part 1, poisson_sampling.cc:166-187 (blocking D2H inside). Returns M.
int ref_poisson_sample(int scaleInv, float radius, int N, int B, int n, const float* mn, const float* mx,
                       const float* pts, const int* bids, const int* cellIdx) {
    float* tp = (float*)temp(3, sizeof(float) * 3 * (size_t)N);
    int* tb = (int*)temp(4, sizeof(int) * (size_t)N);
    int* ti = (int*)temp(5, sizeof(int) * (size_t)N);
    bool* tu = (bool*)temp(6, sizeof(bool) * (size_t)N);
    return samplePointCloud(scaleInv != 0, radius, N, B, n, mn, mx, pts, bids, cellIdx, tp, tb, ti, tu);
}
// PoissonSamplingOp::Compute part 2, poisson_sampling.cc:189-203
void ref_poisson_copy(int M, float* outPts, int* outBids, int* outIdx) {
    copyPoints((float*)g_slots[3].p, (int*)g_slots[4].p, (int*)g_slots[5].p, M, outPts, outBids, outIdx);
}

// GetSampledFeaturesOp::Compute, poisson_sampling.cc:213-250
void ref_get_sampled_features(int N, int C, int M, const int* idx, const float* feats, float* out) {
    getFeaturesSampledPoints(N, C, M, idx, feats, out);
}

// SpatialConvOp::Compute, spatial_conv.cc:179-311 (op input order W1,b1,W2,b2,Wout,bout)
void ref_spatial_conv(int avg, int scaleInv, int K, int C, int Co, int M, int combin, float radius,
                      const float* pts, const int* bids, const float* feats, const float* pdfs,
                      const float* samples, const int* start, const int* pairs, const float* mn, const float* mx,
                      const float* W1, const float* b1, const float* W2, const float* b2, const float* Wo,
                      const float* bo, float* out) {
    spatialConvCPU(avg != 0, scaleInv != 0, K, C, Co, M, combin != 0, radius, pts, bids, feats, pdfs, samples,
                   start, pairs, mn, mx, W1, b1, W2, b2, Wo, bo, out);
}

// SpatialConvGradOp::Compute, spatial_conv.cc:323-470: gradient buffers shaped like the parameters; the caller
// zero-fills them first because the reference only clears the first numNeurons*8 entries of the output layer's
// gradient (spatial_conv.cu:905-906) and TensorFlow hands it uninitialised memory for the rest
void ref_spatial_conv_grad(int avg, int scaleInv, int K, int C, int Co, int M, int N, int combin, float radius,
                           const float* pts, const int* bids, const float* feats, const float* pdfs,
                           const float* samples, const int* start, const int* pairs, const float* mn,
                           const float* mx, const float* W1, const float* b1, const float* W2, const float* b2,
                           const float* Wo, const float* bo, const float* outGrad, float* dFeat, float* dW1,
                           float* db1, float* dW2, float* db2, float* dWo, float* dbo) {
    spatialConvGradsCPU(avg != 0, scaleInv != 0, K, C, Co, M, N, combin != 0, radius, pts, bids, feats, pdfs, samples,
                        start, pairs, mn, mx, W1, b1, W2, b2, Wo, bo, outGrad, dFeat, dW1, dW2, dWo, db1, db2, dbo);
}

// SortPointsStep2GradOp::Compute, sort_gpu.cc:380-437
void ref_sort_points_step2_grad(int N, int C, const float* outPtsGrad, const float* outFeatGrad, const int* newIdx,
                                float* inPtsGrad, float* inFeatGrad) {
    sortPointsStep2GradGPUKernel(N, C, outPtsGrad, outFeatGrad, newIdx, inPtsGrad, inFeatGrad);
}

// GetSampledFeaturesGradOp::Compute, poisson_sampling.cc:252-300
void ref_get_sampled_features_grad(int N, int C, int M, const int* idx, const float* outGrad, float* featGrad) {
    getFeaturesSampledPointsGradients(N, C, M, idx, outGrad, featGrad);
}

int ref_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
