"""The reference's Python op API (tf_ops/MCConvModuleSrc:20-73 + get_block_size, tf_ops/genCompileScript.py:46-47)
on torch CUDA tensors, backed by libmccnn_b200.so.

Same function names, argument order (including the two re-orderings the reference does between the Python wrapper
and the op: compute_pdf MCConvModuleSrc:55-56, spatial_conv MCConvModuleSrc:70-73) and tensor shapes/dtypes as the
TensorFlow ops (tf_ops/*.cc REGISTER_OP blocks): batch ids int32 [N,1], start indices int32 [M,1], pdfs float32
[K,1], neighbour list int32 [K,2], cell table int32 [B,n,n,n,2], Poisson indices int32 [M].  Shape / attribute
violations raise ValueError where the reference's glue raises InvalidArgument (OP_REQUIRES).

These wrappers return exactly-shaped tensors, so find_neighbors and poisson_sampling read one int from the device
(the reference blocks on the same counts: find_neighbors.cu:308, poisson_sampling.cu:222).  Inside
`mccnn_b200.runtime.planned(plan)` the same functions return capacity-sized tensors whose live row count stays on the
device (`tensor._mc_count`) and never block -- see mccnn_b200/runtime.py.
"""
import torch

from . import ops
from . import runtime as _rt
from .runtime import count_of as _cnt, tag as _tag

__all__ = ["compute_aabb", "sort_points_step1", "sort_points_step2", "sort_features", "sort_features_back",
           "transform_indexs", "find_neighbors", "compute_pdf", "poisson_sampling", "get_sampled_features",
           "spatial_conv", "get_block_size", "sort_points", "find_neighbors_pdf"]


def _req(cond, msg):
    if not cond:
        raise ValueError(msg)


def _scheduled(fn):
    """Route an op through runtime.dispatch: in planned multi-stream mode it runs on the stream of the current lane
    (hierarchy / convolution geometry / features) with event dependencies on its tensor inputs; otherwise it is a
    plain call on the current stream."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if fn.__name__ != "sort_features":  # the one op that folds a pending batch norm into its own kernel
            args = tuple(_rt.materialize(a) for a in args)
            kwargs = {k: _rt.materialize(v) for k, v in kwargs.items()}
        ins = [a.base if isinstance(a, _rt.DeferredAffine) else a for a in list(args) + list(kwargs.values())]
        return _rt.dispatch([a for a in ins if isinstance(a, torch.Tensor)], lambda: fn(*args, **kwargs))
    return wrapper


def _check_points(t, name):
    _req(t.dim() == 2 and t.shape[1] >= 3, name + " expects points with at least three components")


def _check_bids(b, n, name):
    _req(b.dim() == 2 and b.shape[1] == 1 and b.shape[0] == n, name + " expects batch ids of shape (numPoints, 1)")


def _check_aabb(mn, mx, batch_size, name):
    _req(mn.dim() == 2 and tuple(mn.shape) == (batch_size, 3), name + " expects an aabb min of shape (batchSize, 3)")
    _req(mx.dim() == 2 and tuple(mx.shape) == (batch_size, 3), name + " expects an aabb max of shape (batchSize, 3)")


def get_block_size():
    return ops.get_block_size()


@_scheduled
def compute_aabb(inPts, inBatchIds, batchSize, scaleInv=True):
    """ComputeAabb, tf_ops/aabb_gpu.cc:22-84 -> (aabbMin [B,3], aabbMax [B,3])."""
    _req(batchSize > 0, "ComputeAabb expects a positive batch size")
    _check_points(inPts, "ComputeAabb")
    _check_bids(inBatchIds, inPts.shape[0], "ComputeAabb")
    pts = inPts if inPts.shape[1] == 3 else inPts[:, :3]  # the kernel hard-codes stride 3 (aabb_gpu.cu:86)
    return ops.compute_aabb(pts, inBatchIds.reshape(-1), batchSize, scaleInv)


def _num_cells(aabbMin, aabbMax, cellSize, scaleInv):
    """determineNumCells (tf_ops/sort_gpu.cu:397-420).  Relative radii: host arithmetic.  Absolute radii: the count
    depends on the batch's box -- one 4-byte D2H in exact mode (as the reference); recorded with the plan and merely
    VERIFIED on the device in planned mode (a mismatch raises the plan's overflow flag, the engine re-records)."""
    _req(cellSize > 0.0, "expects a positive cell size")
    if scaleInv:
        return ops.num_cells(cellSize, True, aabbMin, aabbMax)
    mode = _rt.mode()
    if mode == "planned":
        plan = _rt.current_plan()
        n = plan.next_flag()
        ops.verify_num_cells(cellSize, aabbMin, aabbMax, n, plan.overflow)
        return n
    n = ops.num_cells(cellSize, False, aabbMin, aabbMax)
    if mode == "record":
        _rt.current_plan().flags.append(n)
    return n


@_scheduled
def sort_points_step1(inPts, inBatchIds, aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    """SortPointsStep1, tf_ops/sort_gpu.cc:23-38,184-268 -> (keys [N], indexs [N])."""
    _req(batchSize > 0, "SortPointsStep1 expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "SortPointsStep1 expects points with three components")
    _check_bids(inBatchIds, inPts.shape[0], "SortPointsStep1")
    _check_aabb(aabbMin, aabbMax, batchSize, "SortPointsStep1")
    n = _num_cells(aabbMin, aabbMax, cellSize, scaleInv)
    nd = _cnt(inPts)
    keys, idx = ops.sort_points_step1(inPts, inBatchIds.reshape(-1), aabbMin, aabbMax, batchSize, n, nd)
    return _tag(keys, nd), _tag(idx, nd)


@_scheduled
def sort_points_step2(inPts, inBatchIds, inFeatures, keys, indexs, aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    """SortPointsStep2, tf_ops/sort_gpu.cc:40-64,270-378 -> (pts [N,3], batchIds [N,1], features [N,C], cellIndexs)."""
    _req(batchSize > 0, "SortPointsStep2 expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "SortPointsStep2 expects points with three components")
    N = inPts.shape[0]
    _check_bids(inBatchIds, N, "SortPointsStep2")
    _req(inFeatures.dim() == 2 and inFeatures.shape[0] == N, "SortPointsStep2 expects features of shape (numPoints, C)")
    _req(keys.dim() == 1 and keys.shape[0] == N, "SortPointsStep2 expects keys of shape (numPoints)")
    _req(indexs.dim() == 1 and indexs.shape[0] == N, "SortPointsStep2 expects indexs of shape (numPoints)")
    _check_aabb(aabbMin, aabbMax, batchSize, "SortPointsStep2")
    n = _num_cells(aabbMin, aabbMax, cellSize, scaleInv)
    nd = _cnt(inPts)
    p, b, f, cells = ops.sort_points_step2(inPts, inBatchIds.reshape(-1), inFeatures, keys, indexs, batchSize, n, nd)
    return _tag(p, nd), _tag(b.reshape(-1, 1), nd), _tag(f, nd), cells


@_scheduled
def sort_points(inPts, inBatchIds, inFeatures, aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    """sort_points_step1 followed by sort_points_step2 (the only way the reference ever uses them,
    utils/MCConvBuilder.py:106-110,353-361) -> (keys, indexs, sortPts, sortBatchIds, sortFeatures, cellIndexs).

    Not part of the reference API: an extension that lets the builder issue the whole grid sort as ONE kernel launch
    (csrc/grid_build.cu) when the batch is cloud-contiguous and the grid has <= 8192 cells per cloud.  Results are
    identical to the two-step ops.  Contiguity is verified on the device; in exact / record mode a violation falls
    back to the two-step path here, in planned mode it raises the plan's overflow flag (the engine then re-records)."""
    _req(batchSize > 0, "SortPoints expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "SortPoints expects points with three components")
    N = inPts.shape[0]
    _check_bids(inBatchIds, N, "SortPoints")
    _req(inFeatures.dim() == 2 and inFeatures.shape[0] == N, "SortPoints expects features of shape (numPoints, C)")
    _check_aabb(aabbMin, aabbMax, batchSize, "SortPoints")
    n = _num_cells(aabbMin, aabbMax, cellSize, scaleInv)
    nd = _cnt(inPts)
    mode = _rt.mode()
    fused = ops.build_grid_supported(batchSize, n) and N > 0
    if fused and mode == "planned":
        plan = _rt.current_plan()
        fused = plan.next_flag()
        violation = plan.overflow
    elif fused:
        violation = torch.zeros((1,), dtype=torch.int32, device=inPts.device)
    if fused:
        keys, idx, p, b, f, cells = ops.build_grid(inPts, inBatchIds.reshape(-1), inFeatures, aabbMin, aabbMax,
                                                   batchSize, n, violation, nd)
        if mode != "planned":
            fused = int(violation.item()) == 0
            if mode == "record":
                _rt.current_plan().flags.append(fused)
    if not fused:
        keys, idx = ops.sort_points_step1(inPts, inBatchIds.reshape(-1), aabbMin, aabbMax, batchSize, n, nd)
        p, b, f, cells = ops.sort_points_step2(inPts, inBatchIds.reshape(-1), inFeatures, keys, idx, batchSize, n, nd)
    return _tag(keys, nd), _tag(idx, nd), _tag(p, nd), _tag(b.reshape(-1, 1), nd), _tag(f, nd), cells


@_scheduled
def sort_features(inFeatures, indexs):
    """python sort_features -> SortFeaturesBackGrad op (MCConvModuleSrc:35-36): out[indexs[i]] = inFeatures[i]."""
    _req(inFeatures.dim() == 2, "SortFeaturesBackGrad expects features of shape (numPoints, C)")
    _req(indexs.dim() == 1 and indexs.shape[0] == inFeatures.shape[0],
         "SortFeaturesBackGrad expects indexs of shape (numPoints)")
    if isinstance(inFeatures, _rt.DeferredAffine):  # pending batch norm (+ ReLU): folded into the scatter
        d = inFeatures
        return _tag(ops.sort_features(d.base, indexs, _cnt(indexs), affine=(d.scale, d.shift, d.relu)), _cnt(indexs))
    return _tag(ops.sort_features(inFeatures, indexs, _cnt(indexs)), _cnt(indexs))


@_scheduled
def sort_features_back(inFeatures, indexs):
    """SortFeaturesBack, tf_ops/sort_gpu.cc:78-85: out[i] = inFeatures[indexs[i]]."""
    _req(inFeatures.dim() == 2, "SortFeaturesBack expects features of shape (numPoints, C)")
    _req(indexs.dim() == 1 and indexs.shape[0] == inFeatures.shape[0],
         "SortFeaturesBack expects indexs of shape (numPoints)")
    return _tag(ops.sort_features_back(inFeatures, indexs, _cnt(indexs)), _cnt(indexs))


@_scheduled
def transform_indexs(inIndexs, inNewPositions):
    """TransformIndexs, tf_ops/sort_gpu.cc:96-103,496-533."""
    _req(inIndexs.dim() == 1 and inNewPositions.dim() == 1, "TransformIndexs expects one-dimensional index tensors")
    md = _cnt(inIndexs)
    return _tag(ops.transform_indexs(inIndexs, inNewPositions, md, _cnt(inNewPositions)), md)


@_scheduled
def find_neighbors(inPts, inBatchIds, inPts2, cellIndexs, aabbMin, aabbMax, radius, batchSize, scaleInv):
    """FindNeighbors, tf_ops/find_neighbors.cc:25-43,93-185 -> (startIndexs [M,1], neighIndexs [K,2])."""
    _req(radius > 0.0, "FindNeighborsOp expects a positive radius")
    _req(batchSize > 0, "FindNeighborsOp expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "FindNeighborsOp expects points with three components")
    _check_bids(inBatchIds, inPts.shape[0], "FindNeighborsOp")
    _req(inPts2.dim() == 2 and inPts2.shape[1] == 3, "FindNeighborsOp expects points with three components")
    _req(cellIndexs.dim() == 5 and cellIndexs.shape[0] == batchSize,
         "FindNeighborsOp expects a five dimension tensor for the cell indices")
    _check_aabb(aabbMin, aabbMax, batchSize, "FindNeighborsOp")
    bids = inBatchIds.reshape(-1)
    md = _cnt(inPts)
    if _rt.mode() == "planned":  # capacity-sized pair list, live K stays on the device: one fused launch
        plan = _rt.current_plan()
        start, pairs, k_dev = ops.find_neighbors_fused(inPts, bids, inPts2, cellIndexs, aabbMin, aabbMax, radius,
                                                       scaleInv, plan.next_capacity(), md, plan.overflow)
        return _tag(start.reshape(-1, 1), md), _tag(pairs, k_dev)
    start, ws, k_dev = ops.find_neighbors_count(inPts, bids, inPts2, cellIndexs, aabbMin, aabbMax, radius, scaleInv,
                                                md)
    K = int(k_dev.item())  # the one host sync of this op (the reference has the same one)
    if _rt.mode() == "record":
        _rt.current_plan().sizes.append(K)
    pairs = ops.find_neighbors_pack(inPts, bids, inPts2, cellIndexs, aabbMin, aabbMax, radius, scaleInv, start, ws,
                                    K, md)
    return start.reshape(-1, 1), pairs


@_scheduled
def find_neighbors_pdf(inPts, inBatchIds, inPts2, inBatchIds2, cellIndexs, aabbMin, aabbMax, radius, window, batchSize,
                       scaleInv):
    """find_neighbors followed by compute_pdf on its result (the only way the builder uses them for a new neighbour
    list, utils/MCConvBuilder.py:365-386) -> (startIndexs [M,1], neighIndexs [K,2], pdfs [K,1]).

    Not part of the reference API: an extension that lets a planned (sync-free) run issue both as ONE kernel launch
    (mc_find_neighbors_pdf_fused: the densities are evaluated while a sample's neighbour list is at hand).  The values
    are bit-identical to the two separate ops; in exact / record mode it simply calls them."""
    if _rt.mode() != "planned":
        start, pairs = find_neighbors(inPts, inBatchIds, inPts2, cellIndexs, aabbMin, aabbMax, radius, batchSize, scaleInv)
        return start, pairs, compute_pdf(inPts2, inBatchIds2, aabbMin, aabbMax, start, pairs, window, radius, batchSize,
                                         scaleInv)
    _req(radius > 0.0 and window > 0.0, "FindNeighborsOp / ComputePDFOp expect a positive radius and window")
    _req(batchSize > 0, "FindNeighborsOp expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3 and inPts2.dim() == 2 and inPts2.shape[1] == 3,
         "FindNeighborsOp expects points with three components")
    _check_bids(inBatchIds, inPts.shape[0], "FindNeighborsOp")
    _req(cellIndexs.dim() == 5 and cellIndexs.shape[0] == batchSize,
         "FindNeighborsOp expects a five dimension tensor for the cell indices")
    _check_aabb(aabbMin, aabbMax, batchSize, "FindNeighborsOp")
    md = _cnt(inPts)
    plan = _rt.current_plan()
    start, pairs, pdf, k_dev = ops.find_neighbors_pdf_fused(inPts, inBatchIds.reshape(-1), inPts2, cellIndexs, aabbMin,
                                                            aabbMax, radius, window, scaleInv, plan.next_capacity(), md,
                                                            plan.overflow)
    return _tag(start.reshape(-1, 1), md), _tag(pairs, k_dev), _tag(pdf.reshape(-1, 1), k_dev)


@_scheduled
def compute_pdf(inPts, inBatchIds, aabbMin, aabbMax, startIndexs, neighbors, window, radius, batchSize, scaleInv):
    """ComputePDF, tf_ops/compute_pdf.cc:25-41,57-142 -> pdfs [K,1]."""
    _req(window > 0.0, "ComputePDFOp expects a positive window")
    _req(radius > 0.0, "ComputePDFOp expects a positive radius")
    _req(batchSize > 0, "ComputePDFOp expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "ComputePDFOp expects points with three components")
    _check_bids(inBatchIds, inPts.shape[0], "ComputePDFOp")
    _req(startIndexs.dim() == 2 and startIndexs.shape[1] == 1, "ComputePDFOp expects start indexs of shape (M, 1)")
    _req(neighbors.dim() == 2 and neighbors.shape[1] == 2, "ComputePDFOp expects a neighbor list of shape (K, 2)")
    _check_aabb(aabbMin, aabbMax, batchSize, "ComputePDFOp")
    kd = _cnt(neighbors)
    pdf = ops.compute_pdf(inPts, inBatchIds.reshape(-1), aabbMin, aabbMax, startIndexs.reshape(-1), neighbors, window,
                          radius, scaleInv, _cnt(startIndexs), kd)
    return _tag(pdf.reshape(-1, 1), kd)


@_scheduled
def poisson_sampling(inPts, inBatchIds, cellIndexs, aabbMin, aabbMax, radius, batchSize, scaleInv):
    """PoissonSampling, tf_ops/poisson_sampling.cc:26-45,109-211 -> (pts [M,3], batchIds [M,1], indexs [M])."""
    _req(radius > 0.0, "PoissonSamplingOp expects a positive radius")
    _req(batchSize > 0, "PoissonSamplingOp expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "PoissonSamplingOp expects points with three components")
    _check_bids(inBatchIds, inPts.shape[0], "PoissonSamplingOp")
    _req(cellIndexs.dim() == 5 and cellIndexs.shape[0] == batchSize,
         "PoissonSamplingOp expects a five dimension tensor for the cell indices")
    _check_aabb(aabbMin, aabbMax, batchSize, "PoissonSamplingOp")
    nd = _cnt(inPts)
    if _rt.mode() == "planned":
        plan = _rt.current_plan()
        p, b, idx, m_dev = ops.poisson_sampling(inPts, inBatchIds.reshape(-1), cellIndexs, aabbMin, aabbMax, radius,
                                                batchSize, scaleInv, nd, plan.next_capacity(), plan.overflow)
        h = plan.last_recorded
        return _tag(p, m_dev, h), _tag(b.reshape(-1, 1), m_dev, h), _tag(idx, m_dev, h)
    p, b, idx, m_dev = ops.poisson_sampling(inPts, inBatchIds.reshape(-1), cellIndexs, aabbMin, aabbMax, radius,
                                            batchSize, scaleInv, nd)
    M = int(m_dev.item())  # host sync, as poisson_sampling.cu:222
    if _rt.mode() == "record":
        _rt.current_plan().sizes.append(M)
    return p[:M].clone(), b[:M].clone().reshape(-1, 1), idx[:M].clone()


@_scheduled
def get_sampled_features(inSampledIndexs, pInFeatures):
    """GetSampledFeatures, tf_ops/poisson_sampling.cc:47-55,213-250."""
    _req(inSampledIndexs.dim() == 1, "GetSampledFeaturesOp expects indexs of shape (numSamples)")
    _req(pInFeatures.dim() == 2, "GetSampledFeaturesOp expects features of shape (numPoints, C)")
    md = _cnt(inSampledIndexs)
    return _tag(ops.get_sampled_features(inSampledIndexs, pInFeatures, md), md)


@_scheduled
def spatial_conv(inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, neighStartIndexs, packedNeighs, aabbMin, aabbMax,
                 weights1, weights2, weightsOut, biases1, biases2, biasesOut, numOutFeatures, combin, batchSize,
                 radius, scaleInv, avg):
    """SpatialConv, tf_ops/spatial_conv.cc:24-54,179-311 -> out [M, numOutFeatures].

    Weight tensors carry the TF shapes ([3, 8nb], [8, 8nb], [8, 8nb]) but are read as the flat buffers the
    reference kernel indexes (spatial_conv.cu:49-73,172) -- a reinterpretation, not a transpose."""
    _req(numOutFeatures > 0, "SpatialConvOp expects a positive number of output features")
    _req(radius > 0.0, "SpatialConvOp expects a positive radius")
    _req(batchSize > 0, "SpatialConvOp expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "SpatialConvOp expects points with three components")
    N = inPts.shape[0]
    _req(inFeatures.dim() == 2 and inFeatures.shape[0] == N, "SpatialConvOp expects features of shape (numPoints, C)")
    C = inFeatures.shape[1]
    _check_bids(inBatchIds, N, "SpatialConvOp")
    _req(inPDFs.dim() == 2 and inPDFs.shape[1] == 1, "SpatialConvOp expects pdfs of shape (K, 1)")
    K = inPDFs.shape[0]
    _req(inSamplePts.dim() == 2 and inSamplePts.shape[1] == 3, "SpatialConvOp expects samples with three components")
    M = inSamplePts.shape[0]
    _req(neighStartIndexs.dim() == 2 and tuple(neighStartIndexs.shape) == (M, 1),
         "SpatialConvOp expects start indexs of shape (numSamples, 1)")
    _req(packedNeighs.dim() == 2 and tuple(packedNeighs.shape) == (K, 2),
         "SpatialConvOp expects a neighbor list of shape (K, 2)")
    _check_aabb(aabbMin, aabbMax, batchSize, "SpatialConvOp")
    bs = ops.get_block_size()
    _req(weights1.dim() == 2 and weights1.shape[0] == 3 and biases1.dim() == 1 and
         weights1.shape[1] == biases1.shape[0] and weights1.shape[1] % bs == 0,
         "SpatialConvOp expects a correct first hidden layer")
    _req(weights2.dim() == 2 and weights2.shape[0] == bs and biases2.dim() == 1 and
         weights2.shape[1] == biases2.shape[0] and weights2.shape[1] == weights1.shape[1],
         "SpatialConvOp expects a correct second hidden layer")
    _req(weightsOut.dim() == 2 and weightsOut.shape[0] == bs and biasesOut.dim() == 1 and
         weightsOut.shape[1] == biasesOut.shape[0], "SpatialConvOp expects a correct output layer")
    _req(weightsOut.shape[1] % C == 0,
         "SpatialConvOp expects a number of output neurons multiple of the number of features.")
    if not combin:
        _req(weightsOut.shape[1] == C, "SpatialConvOp expects the same number of features in the input and the output")
        _req(numOutFeatures == C, "SpatialConvOp expects the same number of features in the input and the output")
    nOut = C * numOutFeatures if combin else C
    _req(weights1.shape[1] >= (nOut + bs - 1) // bs * bs, "SpatialConvOp expects enough neurons for the output")
    md = _cnt(inSamplePts)
    out = ops.spatial_conv(inPts, inFeatures, inBatchIds.reshape(-1), inPDFs.reshape(-1), inSamplePts,
                           neighStartIndexs.reshape(-1), packedNeighs, aabbMin, aabbMax, weights1, biases1, weights2,
                           biases2, weightsOut, biasesOut, numOutFeatures, combin, radius, scaleInv, avg, md,
                           _cnt(packedNeighs))
    return _tag(out, md, _rt.hint_of(inSamplePts))


# ---------------------------------------------------------------------------------------------------------------
# Backward ops.  The raw gradient ops carry the names and argument orders of the reference's
# op library (MCConv_module.sort_points_step2_grad / get_sampled_features_grad / spatial_conv_grad as called from
# tf_ops/MCConvModuleSrc:30-32,62-64,74-81); the gradient WIRING that MCConvModuleSrc registers with TensorFlow
# (@tf.RegisterGradient, :28-45,60-81) is expressed as torch.autograd Functions: the five differentiable wrappers
# (sort_points_step2, sort_features, sort_features_back, get_sampled_features, spatial_conv) take that route whenever
# an input requires grad.  Exact mode only -- training runs are not plan-captured.

def sort_points_step2_grad(indexs, outPtsGrad, outFeaturesGrad):
    """SortPointsStep2Grad, tf_ops/sort_gpu.cc:380-437 -> (inPtsGrad [N,3], inFeatureGrad [N,C])."""
    _req(indexs.dim() == 1, "SortPointsStep2GradOp expects indexs of shape (numPoints)")
    N = indexs.shape[0]
    _req(outPtsGrad.dim() == 2 and tuple(outPtsGrad.shape) == (N, 3),
         "SortPointsStep2GradOp expects point gradients of shape (numPoints, 3)")
    _req(outFeaturesGrad.dim() == 2 and outFeaturesGrad.shape[0] == N,
         "SortPointsStep2GradOp expects feature gradients of shape (numPoints, C)")
    return ops.sort_points_step2_grad(indexs, outPtsGrad, outFeaturesGrad)


def get_sampled_features_grad(inSampledIndexs, pInFeatures, inOutFeatureGrad):
    """GetSampledFeaturesGrad, tf_ops/poisson_sampling.cc:57-66,252-300 -> featureGrads [N, C]."""
    _req(inSampledIndexs.dim() == 1, "GetSampledFeaturesGradOp expects indexs of shape (numSamples)")
    _req(pInFeatures.dim() == 2, "GetSampledFeaturesGradOp expects features of shape (numPoints, C)")
    _req(inOutFeatureGrad.dim() == 2 and
         tuple(inOutFeatureGrad.shape) == (inSampledIndexs.shape[0], pInFeatures.shape[1]),
         "GetSampledFeaturesGradOp expects gradients of shape (numSamples, C)")
    return ops.get_sampled_features_grad(inSampledIndexs, pInFeatures, inOutFeatureGrad)


def spatial_conv_grad(inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, neighStartIndexs, packedNeighs, aabbMin,
                      aabbMax, weights1, biases1, weights2, biases2, weightsOut, biasesOut, inOutFeatureGrads,
                      numOutFeatures, combin, batchSize, radius, scaleInv, avg):
    """SpatialConvGrad, tf_ops/spatial_conv.cc:56-96,323-470 (the OP's input order: W1, b1, W2, b2, Wout, bout) ->
    (featureGrads [N,C], weights1Grads, biases1Grads, weights2Grads, biases2Grads, weightsOutGrads, biasesOutGrads),
    every parameter gradient shaped like (and laid out as) its parameter."""
    _req(numOutFeatures > 0, "SpatialConvGradOp expects a positive number of output features")
    _req(radius > 0.0, "SpatialConvGradOp expects a positive radius")
    _req(batchSize > 0, "SpatialConvGradOp expects a positive batch size")
    _req(inPts.dim() == 2 and inPts.shape[1] == 3, "SpatialConvGradOp expects points with three components")
    N, C = inPts.shape[0], inFeatures.shape[1]
    _req(inFeatures.dim() == 2 and inFeatures.shape[0] == N, "SpatialConvGradOp expects features of shape (numPoints, C)")
    M, K = inSamplePts.shape[0], packedNeighs.shape[0]
    outC = numOutFeatures if combin else C
    _req(inOutFeatureGrads.dim() == 2 and tuple(inOutFeatureGrads.shape) == (M, outC),
         "SpatialConvGradOp expects output gradients of shape (numSamples, numOutFeatures)")
    _req(inPDFs.shape[0] == K and neighStartIndexs.shape[0] == M, "SpatialConvGradOp expects a consistent neighbor list")
    bs = ops.get_block_size()
    nOut = C * numOutFeatures if combin else C
    _req(weights1.dim() == 2 and weights1.shape[0] == 3 and weights1.shape[1] % bs == 0 and
         weights1.shape[1] >= (nOut + bs - 1) // bs * bs, "SpatialConvGradOp expects a correct first hidden layer")
    _req(tuple(weights2.shape) == (bs, weights1.shape[1]), "SpatialConvGradOp expects a correct second hidden layer")
    _req(tuple(weightsOut.shape) == (bs, weights1.shape[1]), "SpatialConvGradOp expects a correct output layer")
    if not combin:
        _req(weightsOut.shape[1] == C and numOutFeatures == C,
             "SpatialConvGradOp expects the same number of features in the input and the output")
    return ops.spatial_conv_grad(inPts, inFeatures, inBatchIds.reshape(-1), inPDFs.reshape(-1), inSamplePts,
                                 neighStartIndexs.reshape(-1), packedNeighs, aabbMin, aabbMax, weights1, biases1,
                                 weights2, biases2, weightsOut, biasesOut, inOutFeatureGrads, numOutFeatures, combin,
                                 radius, scaleInv, avg)


def _wants_grad(*tensors):
    """True when a gradient has to flow through the op: a tensor input requires grad, or a pending inference batch
    norm (runtime.DeferredAffine) does -- its base, scale or shift (fine-tuning through frozen statistics, evaluation
    without no_grad): it is then materialised through differentiable torch ops instead of being folded into a kernel."""
    if not torch.is_grad_enabled():
        return False
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.requires_grad:
            return True
        if isinstance(t, _rt.DeferredAffine) and (t.base.requires_grad or t.scale.requires_grad or t.shift.requires_grad):
            return True
    return False


def _exact_only(name):
    if _rt.mode() == "planned":
        raise RuntimeError("mccnn_b200: %s cannot be differentiated inside a planned (capacity-padded) run" % name)


_fwd = {"sort_points_step2": sort_points_step2, "sort_features": sort_features,
        "sort_features_back": sort_features_back, "get_sampled_features": get_sampled_features,
        "spatial_conv": spatial_conv}


class _SortPointsStep2(torch.autograd.Function):  # MCConvModuleSrc:28-32
    @staticmethod
    def forward(ctx, inPts, inFeatures, rest):
        ctx.save_for_backward(rest[2])  # indexs
        ctx.num_features = inFeatures.shape[1]
        p, b, f, cells = _fwd["sort_points_step2"](inPts, rest[0], inFeatures, *rest[1:])
        ctx.mark_non_differentiable(b, cells)
        return p, b, f, cells

    @staticmethod
    def backward(ctx, gp, gb, gf, gc):
        (indexs,) = ctx.saved_tensors
        gp = torch.zeros((indexs.shape[0], 3), device=indexs.device) if gp is None else gp
        gf = torch.zeros((indexs.shape[0], ctx.num_features), device=indexs.device) if gf is None else gf
        ip, iF = sort_points_step2_grad(indexs, gp.contiguous(), gf.contiguous())
        return ip, iF, None


class _SortFeatures(torch.autograd.Function):  # MCConvModuleSrc:33-37: grad = sort_features_back(grads, indexs)
    @staticmethod
    def forward(ctx, inFeatures, indexs):
        ctx.save_for_backward(indexs)
        return _fwd["sort_features"](inFeatures, indexs)

    @staticmethod
    def backward(ctx, g):
        return _fwd["sort_features_back"](g.contiguous(), ctx.saved_tensors[0]), None


class _SortFeaturesBack(torch.autograd.Function):  # MCConvModuleSrc:39-43: grad = sort_features_back_grad(idx, g)
    @staticmethod
    def forward(ctx, inFeatures, indexs):
        ctx.save_for_backward(indexs)
        return _fwd["sort_features_back"](inFeatures, indexs)

    @staticmethod
    def backward(ctx, g):
        return _fwd["sort_features"](g.contiguous(), ctx.saved_tensors[0]), None


class _GetSampledFeatures(torch.autograd.Function):  # MCConvModuleSrc:60-64
    @staticmethod
    def forward(ctx, inSampledIndexs, pInFeatures):
        ctx.save_for_backward(inSampledIndexs, pInFeatures)
        return _fwd["get_sampled_features"](inSampledIndexs, pInFeatures)

    @staticmethod
    def backward(ctx, g):
        idx, feats = ctx.saved_tensors
        return None, get_sampled_features_grad(idx, feats, g.contiguous())


class _SpatialConv(torch.autograd.Function):  # MCConvModuleSrc:74-81
    @staticmethod
    def forward(ctx, inFeatures, weights1, weights2, weightsOut, biases1, biases2, biasesOut, geo, attrs):
        inPts, inBatchIds, inPDFs, inSamplePts, start, pairs, aabbMin, aabbMax = geo
        ctx.save_for_backward(inFeatures, weights1, weights2, weightsOut, biases1, biases2, biasesOut, *geo)
        ctx.attrs = attrs
        return _fwd["spatial_conv"](inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, start, pairs, aabbMin, aabbMax,
                                    weights1, weights2, weightsOut, biases1, biases2, biasesOut, *attrs)

    @staticmethod
    def backward(ctx, g):
        feats, w1, w2, wo, b1, b2, bo, pts, bids, pdfs, samples, start, pairs, mn, mx = ctx.saved_tensors
        dF, dW1, db1, dW2, db2, dWo, dbo = spatial_conv_grad(
            pts, feats, bids, pdfs, samples, start, pairs, mn, mx, w1, b1, w2, b2, wo, bo, g.contiguous(), *ctx.attrs)
        return dF, dW1, dW2, dWo, db1, db2, dbo, None, None


def _differentiable_value(a):
    if not isinstance(a, _rt.DeferredAffine):
        return a
    v = a.base * a.scale + a.shift        # torch ops: the gradient reaches the base and the folded batch-norm vectors
    return torch.relu(v) if a.relu else v


def _differentiable(name, picks, route):
    plain = _fwd[name]

    import functools

    @functools.wraps(plain)
    def wrapper(*args, **kwargs):
        if kwargs or not _wants_grad(*[args[i] for i in picks if i < len(args)]):
            return plain(*args, **kwargs)
        _exact_only(name)
        args = tuple(_differentiable_value(a) for a in args)
        return route(*args)
    return wrapper


sort_points_step2 = _differentiable(
    "sort_points_step2", (0, 2),
    lambda inPts, inBatchIds, inFeatures, *rest: _SortPointsStep2.apply(inPts, inFeatures, (inBatchIds,) + rest))
sort_features = _differentiable("sort_features", (0,), lambda f, idx: _SortFeatures.apply(f, idx))
sort_features_back = _differentiable("sort_features_back", (0,), lambda f, idx: _SortFeaturesBack.apply(f, idx))
get_sampled_features = _differentiable("get_sampled_features", (1,), lambda idx, f: _GetSampledFeatures.apply(idx, f))
spatial_conv = _differentiable(
    "spatial_conv", (1, 9, 10, 11, 12, 13, 14),
    lambda inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, start, pairs, mn, mx, w1, w2, wo, b1, b2, bo, *attrs:
    _SpatialConv.apply(inFeatures, w1, w2, wo, b1, b2, bo,
                       (inPts, inBatchIds, inPDFs, inSamplePts, start, pairs, mn, mx), attrs))

__all__ += ["sort_points_step2_grad", "get_sampled_features_grad", "spatial_conv_grad"]
