/*
 * mccnn_b200.h -- C ABI of libmccnn_b200.so: the Monte-Carlo-convolution forward path of viscom-ulm/MCCNN
 * as sm_100a CUDA kernels.
 *
 * Every entry point replaces one host launcher of the reference (the C++-linkage functions its TensorFlow glue
 * declares `extern` and calls from OpKernel::Compute); the replaced declaration is cited per function as
 * tf_ops/<file>.cc:<line> (declaration) and tf_ops/<file>.cu:<line> (definition).
 *
 * Conventions (they differ from the reference where the reference blocks the host):
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless the name starts with `host`;
 *     the caller owns all memory (outputs and workspaces) -- nothing is allocated inside;
 *   - `void* stream` is a cudaStream_t; all work is enqueued on it, no entry point synchronises except
 *     mc_determine_num_cells in the non-scale-invariant case;
 *   - dynamic extents are passed as a pair (capacity, dCount): `capacity` is a host int bounding the extent,
 *     `dCount` an optional device int holding the live extent (NULL: live extent == capacity).  Dynamic RESULTS
 *     (number of neighbours, number of Poisson samples) are written to device ints, never returned to the host;
 *   - all tensors are dense row-major float32 / int32 exactly as the reference's ops lay them out
 *     (points [N,3], batch ids [N], cell table [B,n,n,n,2], neighbour list start[M] + pairs[K,2] = (j, i));
 *   - return value: 0 ok, 1 invalid argument, 2 CUDA error; mc_last_error() gives the message.
 */
#ifndef MCCNN_B200_H_
#define MCCNN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* mc_last_error(void);
int mc_abi_version(void);
/* BLOCK_MLP_SIZE of the build == python get_block_size() (tf_ops/genCompileScript.py:20,46-47) */
int mc_get_block_size(void);
int mc_device_info(int* hostSmCount, int* hostCcMajor, int* hostCcMinor);
/* kernels launched by this library so far in the process (diagnostics; bench.py reports it as gpu_launches) */
long long mc_launch_count(void);

/* computeAABB -- tf_ops/aabb_gpu.cc:38-40, tf_ops/aabb_gpu.cu:121-141.  workspace: 1 int. */
int mc_compute_aabb(int scaleInv, int numPoints, int batchSize, const float* pts, const int* batchIds,
                    float* aabbMin, float* aabbMax, int* workspace, void* stream);

/* determineNumCells -- tf_ops/sort_gpu.cc:105-110, tf_ops/sort_gpu.cu:397-420.  workspace: 1 int (scaleInv == 0). */
int mc_determine_num_cells(int scaleInv, float cellSize, const float* aabbMin, const float* aabbMax, int* workspace,
                           int* hostOutNumCells, void* stream);

/* Sync-free check used by planned / graph-captured forwards with absolute radii: ORs 1 into *dMismatchFlag when the
 * box implies a number of cells (tf_ops/sort_gpu.cu:381-389) different from expectedNumCells. */
int mc_verify_num_cells(float cellSize, const float* aabbMin, const float* aabbMax, int expectedNumCells,
                        int* dMismatchFlag, void* stream);

/* computeAuxiliarBuffersSize(int,int,int*,int*,int*) -- tf_ops/sort_gpu.cc:112-117, tf_ops/sort_gpu.cu:422-434 */
long long mc_sort_step1_workspace_ints(int numPoints, int batchSize, int numCells);

/* sortPointsStep1GPUKernel -- tf_ops/sort_gpu.cc:119-131, tf_ops/sort_gpu.cu:436-471 */
int mc_sort_points_step1(int numPoints, const int* dNumPoints, int batchSize, int numCells, const float* aabbMin,
                         const float* aabbMax, const float* pts, const int* batchIds, int* workspace, int* keys,
                         int* newIdx, void* stream);

/* sortPointsStep2GPUKernel -- tf_ops/sort_gpu.cc:133-147, tf_ops/sort_gpu.cu:473-497.  workspace: numPoints ints. */
int mc_sort_points_step2(int numPoints, const int* dNumPoints, int batchSize, int numFeatures, int numCells,
                         const float* pts, const int* batchIds, const float* features, const int* keys,
                         const int* newIdx, int* workspace, float* outPts, int* outBatchIds, float* outFeatures,
                         int* outCellIdx, void* stream);

/* sortPointsStep1GPUKernel + sortPointsStep2GPUKernel fused into one launch -- tf_ops/sort_gpu.cu:436-497 -- for
 * cloud-contiguous packed batches (utils/DataSet.py:816-818) and grids of at most 8192 cells per cloud.  Outputs are
 * exactly those of the two calls above.  Contiguity is verified on the device: *dViolation (required, caller-zeroed)
 * is set otherwise and the outputs are then unspecified.  workspace must be 8-byte aligned. */
int mc_build_grid_supported(int batchSize, int numCells);
long long mc_build_grid_workspace_ints(int numPoints);
int mc_build_grid(int numPoints, const int* dNumPoints, int batchSize, int numFeatures, int numCells,
                  const float* aabbMin, const float* aabbMax, const float* pts, const int* batchIds,
                  const float* features, int* workspace, int* dViolation, int* keys, int* newIdx, float* outPts,
                  int* outBatchIds, float* outFeatures, int* outCellIdx, void* stream);

/* sortFeaturesBackGrad (python sort_features) -- tf_ops/sort_gpu.cc:165-170, tf_ops/sort_gpu.cu:529-540 */
int mc_sort_features(int numPoints, const int* dNumPoints, int numFeatures, const float* in, const int* newIdx,
                     float* out, void* stream);

/* sort_features of act(in * scale + shift), scale/shift numFeatures floats, relu 0/1: the inference
 * tf.layers.batch_normalization + tf.nn.relu that precedes a convolution (utils/MCNetworkUtils.py:128-142) folded
 * into the scatter of tf_ops/sort_gpu.cu:529-540 (no reference counterpart: there it is a separate TF op). */
int mc_sort_features_affine(int numPoints, const int* dNumPoints, int numFeatures, const float* in, const int* newIdx,
                            const float* scale, const float* shift, int relu, float* out, void* stream);

/* sortFeaturesBack -- tf_ops/sort_gpu.cc:158-163, tf_ops/sort_gpu.cu:516-527 */
int mc_sort_features_back(int numPoints, const int* dNumPoints, int numFeatures, const float* in, const int* newIdx,
                          float* out, void* stream);

/* computeInverseIndexs + transformIndexs -- tf_ops/sort_gpu.cc:172-182, tf_ops/sort_gpu.cu:542-564.
 * workspace: numPoints ints. */
int mc_transform_indexs(int numIndexs, const int* dNumIndexs, int numPoints, const int* dNumPoints,
                        const int* inIndexs, const int* newIdx, int* workspace, int* out, void* stream);

/* computeAuxiliarBuffersSize(int,int*,int*) -- tf_ops/find_neighbors.cc:58-61, tf_ops/find_neighbors.cu:319-328 */
long long mc_find_neighbors_workspace_ints(int numSamples);

/* countNeighborsCPU -- tf_ops/find_neighbors.cc:45-56, tf_ops/find_neighbors.cu:268-317.
 * The total is written to *dNumNeighbors (device) instead of being returned through a blocking copy. */
int mc_find_neighbors_count(int scaleInv, int numSamples, const int* dNumSamples, int numCells, float radius,
                            const float* samples, const int* sampleBatchIds, const float* pts2, const int* cellIdx,
                            const float* aabbMin, const float* aabbMax, int* startIdx, int* workspace,
                            int* dNumNeighbors, void* stream);

/* packNeighborsCPU -- tf_ops/find_neighbors.cc:63-78, tf_ops/find_neighbors.cu:330-371.
 * Same startIdx / workspace as the count call.  *dOverflow (optional) is set when capacityPairs is exceeded. */
int mc_find_neighbors_pack(int scaleInv, int numSamples, const int* dNumSamples, int numCells, float radius,
                           const float* samples, const int* sampleBatchIds, const float* pts2, const int* cellIdx,
                           const float* aabbMin, const float* aabbMax, int* startIdx, const int* workspace,
                           int* pairs, int capacityPairs, int* dOverflow, void* stream);

/* countNeighborsCPU + packNeighborsCPU in ONE launch (no intermediate scan kernels, distances evaluated once) for
 * callers that provide the pair capacity up front.  startIdx receives the final start indices, *dNumNeighbors the
 * total.  workspace: mc_find_neighbors_workspace_ints ints, 8-byte aligned. */
int mc_find_neighbors_fused(int scaleInv, int numSamples, const int* dNumSamples, int numCells, float radius,
                            const float* samples, const int* sampleBatchIds, const float* pts2, const int* cellIdx,
                            const float* aabbMin, const float* aabbMax, int* startIdx, int* workspace, int* pairs,
                            int capacityPairs, int* dNumNeighbors, int* dOverflow, void* stream);
/* mc_find_neighbors_fused + mc_compute_pdf in the same launch (countNeighborsCPU + packNeighborsCPU,
 * tf_ops/find_neighbors.cc:45-78, and computeDPFsCPU, tf_ops/compute_pdf.cc:43-55): pdfs[capacityPairs] receives exactly
 * the values mc_compute_pdf would produce for the pair list (same terms, same order). */
int mc_find_neighbors_pdf_fused(int scaleInv, float window, int numSamples, const int* dNumSamples, int numCells,
                                float radius, const float* samples, const int* sampleBatchIds, const float* pts2,
                                const int* cellIdx, const float* aabbMin, const float* aabbMax, int* startIdx,
                                int* workspace, int* pairs, float* pdfs, int capacityPairs, int* dNumNeighbors,
                                int* dOverflow, void* stream);

/* computeDPFsCPU -- tf_ops/compute_pdf.cc:43-55, tf_ops/compute_pdf.cu:98-120 */
int mc_compute_pdf(int scaleInv, float window, int numSamples, const int* dNumSamples, int numNeighbors,
                   const int* dNumNeighbors, float radius, const float* pts, const int* batchIds,
                   const float* aabbMin, const float* aabbMax, const int* startIdx, const int* pairs, float* pdfs,
                   void* stream);

long long mc_poisson_workspace_ints(int numPoints, int batchSize);

/* samplePointCloud + copyPoints -- tf_ops/poisson_sampling.cc:68-92, tf_ops/poisson_sampling.cu:175-244.
 * Outputs hold capacityOut rows (numPoints always suffices); the number of samples is written to *dNumSamples
 * (device); *dOverflow (optional) is set when it exceeds capacityOut. */
int mc_poisson_sampling(int scaleInv, float radius, int numPoints, const int* dNumPoints, int batchSize,
                        int numCells, const float* aabbMin, const float* aabbMax, const float* pts,
                        const int* batchIds, const int* cellIdx, int* workspace, int capacityOut, float* outPts,
                        int* outBatchIds, int* outIdx, int* dNumSamples, int* dOverflow, void* stream);

/* getFeaturesSampledPoints -- tf_ops/poisson_sampling.cc:94-100, tf_ops/poisson_sampling.cu:246-258 */
int mc_get_sampled_features(int numSamples, const int* dNumSamples, int numFeatures, const int* idx,
                            const float* features, float* out, void* stream);

long long mc_spatial_conv_workspace_floats(int numSamples, int numInFeatures, int numOutFeatures, int combin);

/* spatialConvCPU -- tf_ops/spatial_conv.cc:98-128, tf_ops/spatial_conv.cu:796-871.
 * Parameter order is the op's (W1, b1, W2, b2, Wout, bout); buffers are the flat layouts the reference kernel
 * indexes (tf_ops/spatial_conv.cu:49-73,172). */
int mc_spatial_conv(int avg, int scaleInv, int numNeighbors, const int* dNumNeighbors, int numInFeatures,
                    int numOutFeatures, int numSamples, const int* dNumSamples, int combin, float radius,
                    int batchSize, const float* pts, const int* batchIds, const float* features, const float* pdfs,
                    const float* samples, const int* startIdx, const int* pairs, const float* aabbMin,
                    const float* aabbMax, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* Wout, const float* bout, float* workspace, float* outFeatures, void* stream);

/* ---- backward ops: the gradient launchers of the reference's op library ---------------- */

long long mc_spatial_conv_grad_workspace_floats(int numNeighbors, int numInFeatures, int numOutFeatures, int combin);

/* spatialConvGradsCPU -- tf_ops/spatial_conv.cc:130-156, tf_ops/spatial_conv.cu:873-966.
 * Inputs as mc_spatial_conv plus outFeatureGrads [numSamples, outC]; outputs: featureGrads [numPoints, numInFeatures]
 * and one gradient buffer per parameter in that parameter's flat layout ([3*8nb], [8nb], [8nb*8], [8nb], [8nb*8],
 * [8nb]); all fully written.  Exact extents only (training is not graph-captured). */
int mc_spatial_conv_grad(int avg, int scaleInv, int numNeighbors, int numInFeatures, int numOutFeatures,
                         int numSamples, int numPoints, int combin, float radius, int batchSize, const float* pts,
                         const int* batchIds, const float* features, const float* pdfs, const float* samples,
                         const int* startIdx, const int* pairs, const float* aabbMin, const float* aabbMax,
                         const float* W1, const float* b1, const float* W2, const float* b2, const float* Wout,
                         const float* bout, const float* outFeatureGrads, float* workspace, float* featureGrads,
                         float* W1Grads, float* b1Grads, float* W2Grads, float* b2Grads, float* WoutGrads,
                         float* boutGrads, void* stream);

/* sortPointsStep2GradGPUKernel -- tf_ops/sort_gpu.cc:147-154, tf_ops/sort_gpu.cu:499-513:
 * inPtsGrads[i,:] = outPtsGrads[newIdx[i],:], inFeatureGrads[i,:] = outFeatureGrads[newIdx[i],:] */
int mc_sort_points_step2_grad(int numPoints, int numFeatures, const float* outPtsGrads, const float* outFeatureGrads,
                              const int* newIdx, float* inPtsGrads, float* inFeatureGrads, void* stream);

/* getFeaturesSampledPointsGradients -- tf_ops/poisson_sampling.cc:102-107, tf_ops/poisson_sampling.cu:260-274:
 * featureGrads = 0; featureGrads[idx[s],:] = outFeatureGrads[s,:] */
int mc_get_sampled_features_grad(int numPoints, int numFeatures, int numSamples, const int* idx,
                                 const float* outFeatureGrads, float* featureGrads, void* stream);

/* Dense glue between the convolutions (TensorFlow ops in the reference: utils/MCNetworkUtils.py:20-142).
 * mc_scale_shift_act: y = act(x * scale + shift) -- inference batch norm (+ ReLU), statistics folded by the caller.
 * mc_dense_layer:     y = act2((act1(x * s1 + t1) @ W + bias) * s2 + t2), W [K, N] row-major; s1/t1, bias, s2/t2 may
 *                     be NULL -- conv_1x1 and the layers of MLP_1_hidden / MLP_2_hidden, one launch per layer.
 *                     numRowsHint (<= 0: unknown): host estimate of the live row count, selects the kernel variant. */
int mc_scale_shift_act(int numRows, const int* dNumRows, int numCols, const float* x, const float* scale,
                       const float* shift, int relu, float* y, void* stream);
int mc_dense_layer(int numRows, const int* dNumRows, int numRowsHint, int K, int N, const float* x, const float* s1,
                   const float* t1, int relu1, const float* W, const float* bias, const float* s2, const float* t2,
                   int relu2, float* y, void* stream);

/* A whole MLP on few rows in ONE launch -- MLP_1_hidden / MLP_2_hidden, utils/MCNetworkUtils.py:20-107:
 *   a_0 = act0(x * s0 + t0);  a_{l+1} = act_l((a_l @ W[l] + bias[l]) * scale[l] + shift[l]);  y = a_numLayers.
 * dims: numLayers + 1 host ints (each 1..1024); W / bias / scale / shift: host arrays of numLayers device pointers
 * (bias and scale+shift may be NULL per layer); relu: numLayers host flags.  One CTA per row: meant for <= ~128 rows. */
int mc_mlp_rows(int numRows, const int* dNumRows, int numLayers, const int* dims, const float* x, const float* s0,
                const float* t0, int relu0, const float* const* W, const float* const* bias,
                const float* const* scale, const float* const* shift, const int* relu, float* y, void* stream);

/* mc_dense_layer for FEW live rows (one per cloud: the classification heads, MLP_1_hidden / MLP_2_hidden,
 * utils/MCNetworkUtils.py:20-107) and layers of any width -- MCClassH's 2048 -> 1024 -> 512 head.  Every CTA owns
 * (128 columns) x (a K segment) for all rows, so W is read once; the K segments are summed in a fixed order
 * (deterministic).  numRows is the capacity of x / y; workspace: mc_dense_fewrows_workspace_floats floats (may be 0). */
long long mc_dense_fewrows_workspace_floats(int numRows, int K, int N);
int mc_dense_layer_fewrows(int numRows, const int* dNumRows, int K, int N, const float* x, const float* s1,
                           const float* t1, int relu1, const float* W, const float* bias, const float* s2,
                           const float* t2, int relu2, float* workspace, float* y, void* stream);

/* ---- dense glue on the tcgen05 tensor cores (csrc/dense_tc.cu): the same layer as mc_dense_layer for many rows,
 * FP32 in / FP32 out through the 3xTF32 split, TMA-fed, accumulator in TMEM.  Replaces tf.matmul (cuBLAS) + the
 * TensorFlow batch-norm / ReLU ops around it (utils/MCNetworkUtils.py:109-142).  The static operand is the weight
 * matrix TRANSPOSED to [N, K] and split once per weight version by mc_split_tf32 into hi = tf32(w), lo = w - hi. */
int mc_split_tf32(long long count, const float* in, float* hi, float* lo, void* stream);
int mc_dense_layer_tc_supported(int K, int N);
int mc_dense_layer_tc(int numRows, const int* dNumRows, int K, int N, const float* x, const float* s1, const float* t1,
                      int relu1, const float* WtHi, const float* WtLo, const float* bias, const float* s2,
                      const float* t2, int relu2, float* y, void* stream);

/* ---- multi-GPU: the one collective of a sharded step, over NVLink peer memory (no NCCL on the data path) ----------
 * All-gather of floatsPerRank floats per rank through symmetric buffers mapped on every peer: peer stores + flag
 * words, one CTA, capturable in a CUDA graph.  peerBufPtrs: device array of worldSize buffer addresses (64-bit ints),
 * each buffer mc_allgather_rows_p2p_buffer_floats() floats, zero-filled before the first step; stepCounter: device
 * int, zero before the first step.  (The reference has no multi-GPU path; BASELINE.json's north_star prescribes it.) */
long long mc_allgather_rows_p2p_buffer_floats(int floatsPerRank, int worldSize);
int mc_allgather_rows_p2p(int floatsPerRank, int rank, int worldSize, const float* local, const long long* peerBufPtrs,
                          int* stepCounter, float* out, void* stream);

/* The same exchange as two kernels over a ring of `slots` steps (csrc/peer_gather.cu): mc_peer_put -- last node of a
 * forward's graph -- stores this rank's live rows + (row count, overflow flag) into every rank's symmetric buffer and
 * waits only for the acknowledgement of step seq - slots; mc_peer_collect -- on the stream that feeds the host copy --
 * waits for every rank's put of the step, copies the rows out (padded or compacted in rank order: per-point logits of
 * unequal length), reports per-rank (rows, flag), the total and the MAX flag in `meta`, and acknowledges.  Ranks may
 * drift apart by `slots` steps before anything blocks.  state: 4 device ints, zero before the first step.
 * compact: bit 0 = rows of all ranks contiguous in rank order; bit 1 = do nothing unless a put is still uncollected
 * (the deferred collect captured in front of a lane's next put, and the flush behind its last step). */
long long mc_peer_exchange_buffer_floats(int capRows, int cols, int worldSize, int slots);
int mc_peer_put(int capRows, const int* dRows, int cols, int rank, int worldSize, int slots, const float* local,
                const int* dFlag, const long long* peerBufPtrs, int* state, void* stream);
int mc_peer_collect(int capRows, int cols, int rank, int worldSize, int slots, int compact, const long long* peerBufPtrs,
                    int* state, float* out, int* meta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCCNN_B200_H_ */
