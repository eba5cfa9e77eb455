"""NumPy bindings of liboracle.so (oracle/mccnn_oracle.c) -- the CPU restatement of the reference ops.

TEST INFRASTRUCTURE.  Function names and argument order follow the reference's Python op API
(tf_ops/MCConvModuleSrc:20-73) so parity tests read like calls into the reference; arrays are numpy float32/int32 with
the reference's shapes (batch ids [N,1], start indices [M,1], pdfs [K,1]).
"""
import ctypes

import numpy as np

from . import LIBORACLE, build

_lib = None


def lib():
    global _lib
    if _lib is None:
        import os
        if not os.path.exists(LIBORACLE):
            build()
        _lib = ctypes.CDLL(LIBORACLE)
        _lib.orc_find_neighbors.restype = ctypes.c_long
        _lib.orc_num_cells.restype = ctypes.c_int
        _lib.orc_poisson_sampling.restype = ctypes.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_cf = ctypes.c_float
_ci = ctypes.c_int


def get_block_size():
    return lib().orc_block_size()


def compute_aabb(inPts, inBatchIds, batchSize, scaleInv=True):
    pts, bids = _f(inPts), _i(inBatchIds).reshape(-1)
    mn = np.empty((batchSize, 3), np.float32)
    mx = np.empty((batchSize, 3), np.float32)
    lib().orc_compute_aabb(_ci(int(scaleInv)), _ci(pts.shape[0]), _ci(batchSize), _p(pts), _p(bids), _p(mn), _p(mx))
    return mn, mx


def num_cells(aabbMin, aabbMax, cellSize, scaleInv):
    mn, mx = _f(aabbMin), _f(aabbMax)
    return lib().orc_num_cells(_ci(int(scaleInv)), _cf(cellSize), _p(mn), _p(mx))


def sort_points_step1(inPts, inBatchIds, aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    pts, bids, mn, mx = _f(inPts), _i(inBatchIds).reshape(-1), _f(aabbMin), _f(aabbMax)
    n = num_cells(mn, mx, cellSize, scaleInv)
    N = pts.shape[0]
    keys = np.empty((N,), np.int32)
    idx = np.empty((N,), np.int32)
    lib().orc_sort_step1(_ci(N), _ci(batchSize), _ci(n), _p(mn), _p(mx), _p(pts), _p(bids), _p(keys), _p(idx))
    return keys, idx


def sort_points_step2(inPts, inBatchIds, inFeatures, keys, indexs, aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    pts, bids, feats = _f(inPts), _i(inBatchIds).reshape(-1), _f(inFeatures)
    keys, indexs, mn, mx = _i(keys), _i(indexs), _f(aabbMin), _f(aabbMax)
    n = num_cells(mn, mx, cellSize, scaleInv)
    N, C = pts.shape[0], feats.shape[1]
    oP = np.empty_like(pts)
    oB = np.empty_like(bids)
    oF = np.empty_like(feats)
    cells = np.empty((batchSize, n, n, n, 2), np.int32)
    lib().orc_sort_step2(_ci(N), _ci(batchSize), _ci(C), _ci(n), _p(pts), _p(bids), _p(feats), _p(keys), _p(indexs),
                         _p(oP), _p(oB), _p(oF), _p(cells))
    return oP, oB.reshape(-1, 1), oF, cells


def sort_features(inFeatures, indexs):
    f, idx = _f(inFeatures), _i(indexs)
    out = np.empty_like(f)
    lib().orc_sort_features(_ci(f.shape[0]), _ci(f.shape[1]), _p(f), _p(idx), _p(out))
    return out


def sort_features_back(inFeatures, indexs):
    f, idx = _f(inFeatures), _i(indexs)
    out = np.empty_like(f)
    lib().orc_sort_features_back(_ci(f.shape[0]), _ci(f.shape[1]), _p(f), _p(idx), _p(out))
    return out


def transform_indexs(inIndexs, inNewPositions):
    idx, new = _i(inIndexs), _i(inNewPositions)
    out = np.empty_like(idx)
    lib().orc_transform_indexs(_ci(idx.shape[0]), _ci(new.shape[0]), _p(idx), _p(new), _p(out))
    return out


def find_neighbors(inPts, inBatchIds, inPts2, cellIndexs, aabbMin, aabbMax, radius, batchSize, scaleInv):
    s, sb, p2, cells = _f(inPts), _i(inBatchIds).reshape(-1), _f(inPts2), _i(cellIndexs)
    mn, mx = _f(aabbMin), _f(aabbMax)
    M, n = s.shape[0], cells.shape[1]
    start = np.empty((M,), np.int32)
    args = (_ci(int(scaleInv)), _ci(M), _ci(n), _cf(radius), _p(s), _p(sb), _p(p2), _p(cells), _p(mn), _p(mx))
    K = lib().orc_find_neighbors(*args, _p(start), ctypes.c_void_p(0))
    pairs = np.empty((K, 2), np.int32)
    lib().orc_find_neighbors(*args, _p(start), _p(pairs))
    return start.reshape(-1, 1), pairs


def compute_pdf(inPts, inBatchIds, aabbMin, aabbMax, startIndexs, neighbors, window, radius, batchSize, scaleInv):
    pts, bids, mn, mx = _f(inPts), _i(inBatchIds).reshape(-1), _f(aabbMin), _f(aabbMax)
    start, pairs = _i(startIndexs).reshape(-1), _i(neighbors)
    K = pairs.shape[0]
    pdf = np.empty((K,), np.float32)
    lib().orc_compute_pdf(_ci(int(scaleInv)), _cf(window), _ci(start.shape[0]), _ci(K), _cf(radius), _p(pts),
                          _p(bids), _p(mn), _p(mx), _p(start), _p(pairs), _p(pdf))
    return pdf.reshape(-1, 1)


def poisson_sampling(inPts, inBatchIds, cellIndexs, aabbMin, aabbMax, radius, batchSize, scaleInv,
                     return_phase=False):
    pts, bids, cells = _f(inPts), _i(inBatchIds).reshape(-1), _i(cellIndexs)
    mn, mx = _f(aabbMin), _f(aabbMax)
    N, n = pts.shape[0], cells.shape[1]
    oP = np.empty((max(N, 1), 3), np.float32)
    oB = np.empty((max(N, 1),), np.int32)
    oI = np.empty((max(N, 1),), np.int32)
    oPh = np.empty((max(N, 1),), np.int32)
    M = lib().orc_poisson_sampling(_ci(int(scaleInv)), _cf(radius), _ci(N), _ci(batchSize), _ci(n), _p(mn), _p(mx),
                                   _p(pts), _p(bids), _p(cells), _p(oP), _p(oB), _p(oI), _p(oPh))
    res = (oP[:M].copy(), oB[:M].copy().reshape(-1, 1), oI[:M].copy())
    return res + (oPh[:M].copy(),) if return_phase else res


def get_sampled_features(inSampledIndexs, pInFeatures):
    idx, f = _i(inSampledIndexs), _f(pInFeatures)
    out = np.empty((idx.shape[0], f.shape[1]), np.float32)
    lib().orc_get_sampled_features(_ci(idx.shape[0]), _ci(f.shape[1]), _p(idx), _p(f), _p(out))
    return out


def spatial_conv(inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, neighStartIndexs, packedNeighs, aabbMin, aabbMax,
                 weights1, weights2, weightsOut, biases1, biases2, biasesOut, numOutFeatures, combin, batchSize,
                 radius, scaleInv, avg, accum_double=False, abs_sum=False):
    """accum_double: float64 accumulation; abs_sum: return sum |contribution| per output (error scale)."""
    pts, feats, bids = _f(inPts), _f(inFeatures), _i(inBatchIds).reshape(-1)
    pdfs, samples = _f(inPDFs).reshape(-1), _f(inSamplePts)
    start, pairs, mn, mx = _i(neighStartIndexs).reshape(-1), _i(packedNeighs), _f(aabbMin), _f(aabbMax)
    w1, w2, wo, b1, b2, bo = (_f(a) for a in (weights1, weights2, weightsOut, biases1, biases2, biasesOut))
    M, K, C = samples.shape[0], pairs.shape[0], feats.shape[1]
    outC = numOutFeatures if combin else C
    out = np.empty((M, outC), np.float32)
    lib().orc_spatial_conv(_ci(int(avg)), _ci(int(scaleInv)), _ci(K), _ci(C), _ci(numOutFeatures), _ci(M),
                           _ci(int(combin)), _cf(radius), _p(pts), _p(bids), _p(feats), _p(pdfs), _p(samples),
                           _p(start), _p(pairs), _p(mn), _p(mx), _p(w1), _p(b1), _p(w2), _p(b2), _p(wo), _p(bo),
                           _ci(2 if abs_sum else int(accum_double)), _p(out))
    return out


def spatial_conv_grad(inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, neighStartIndexs, packedNeighs, aabbMin,
                      aabbMax, weights1, biases1, weights2, biases2, weightsOut, biasesOut, inOutFeatureGrads,
                      numOutFeatures, combin, batchSize, radius, scaleInv, avg, abs_sum=False):
    """SpatialConvGrad in the OP's input order (tf_ops/spatial_conv.cc:56-96: W1, b1, W2, b2, Wout, bout, then the
    incoming gradient) -> (featureGrads [N,C], weights1Grads, biases1Grads, weights2Grads, biases2Grads,
    weightsOutGrads, biasesOutGrads), each shaped like its parameter.  Accumulated in float64, rounded once."""
    pts, feats, bids = _f(inPts), _f(inFeatures), _i(inBatchIds).reshape(-1)
    pdfs, samples = _f(inPDFs).reshape(-1), _f(inSamplePts)
    start, pairs, mn, mx = _i(neighStartIndexs).reshape(-1), _i(packedNeighs), _f(aabbMin), _f(aabbMax)
    w1, b1, w2, b2, wo, bo = (_f(a) for a in (weights1, biases1, weights2, biases2, weightsOut, biasesOut))
    og = _f(inOutFeatureGrads)
    M, K, C, N = samples.shape[0], pairs.shape[0], feats.shape[1], pts.shape[0]
    dF = np.empty((N, C), np.float32)
    outs = [np.empty_like(a) for a in (w1, b1, w2, b2, wo, bo)]
    lib().orc_spatial_conv_grad(_ci(int(avg)), _ci(int(scaleInv)), _ci(K), _ci(C), _ci(numOutFeatures), _ci(M), _ci(N),
                                _ci(int(combin)), _cf(radius), _p(pts), _p(bids), _p(feats), _p(pdfs), _p(samples),
                                _p(start), _p(pairs), _p(mn), _p(mx), _p(w1), _p(b1), _p(w2), _p(b2), _p(wo), _p(bo),
                                _p(og), _ci(int(abs_sum)), _p(dF), _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]),
                                _p(outs[4]), _p(outs[5]))
    return (dF,) + tuple(outs)


def sort_points_step2_grad(indexs, outPtsGrad, outFeaturesGrad):
    """SortPointsStep2Grad (tf_ops/sort_gpu.cc:380-437) -> (inPtsGrad, inFeatureGrad)."""
    idx, gp, gf = _i(indexs), _f(outPtsGrad), _f(outFeaturesGrad)
    ip, iF = np.empty_like(gp), np.empty_like(gf)
    lib().orc_sort_step2_grad(_ci(idx.shape[0]), _ci(gf.shape[1]), _p(idx), _p(gp), _p(gf), _p(ip), _p(iF))
    return ip, iF


def get_sampled_features_grad(inSampledIndexs, pInFeatures, inOutFeatureGrad):
    """GetSampledFeaturesGrad (tf_ops/poisson_sampling.cc:252-300) -> featureGrads [N, C]."""
    idx, f, g = _i(inSampledIndexs), _f(pInFeatures), _f(inOutFeatureGrad)
    out = np.empty_like(f)
    lib().orc_get_sampled_features_grad(_ci(idx.shape[0]), _ci(f.shape[1]), _ci(f.shape[0]), _p(idx), _p(g), _p(out))
    return out
