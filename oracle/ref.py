"""torch-CUDA bindings of oracle/_ref/libmccnn_ref.so -- the reference's OWN kernels (tf_ops/*.cu compiled unmodified
for sm_100a by oracle/Makefile) behind the reference's Python op API (tf_ops/MCConvModuleSrc:20-73).

TEST / BASELINE INFRASTRUCTURE (GPU box only): used by the `-m gpu` parity tests as the ground truth, by
tests/golden/make_golden.py to produce the committed fixtures, and by `bench.py --impl reference` as the reference
arm.  The reference launches on the legacy default stream and blocks the host where its TF glue does.
"""
import ctypes
import os

import torch

from . import LIBREF

_lib = None


def available():
    # MCCNN_NO_REF_KERNELS=1: tests skip their live comparison with the reference's own kernels (sanitizer runs: the
    # sanitizer should report on this repository's kernels, not on the reference's)
    return os.path.exists(LIBREF) and os.environ.get("MCCNN_NO_REF_KERNELS", "0") != "1"


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmccnn_ref.so is missing: run `make -C oracle` where /root/reference exists")
        _lib = ctypes.CDLL(LIBREF)
        _lib.ref_find_neighbors_count.restype = ctypes.c_uint
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


_cf, _ci = ctypes.c_float, ctypes.c_int
_i32, _f32 = torch.int32, torch.float32


def _f(t):
    assert t.dtype == _f32 and t.is_cuda
    return t.contiguous()


def _i(t):
    assert t.dtype == _i32 and t.is_cuda
    return t.contiguous()


def get_block_size():
    return 8


def compute_aabb(inPts, inBatchIds, batchSize, scaleInv=True):
    pts, bids = _f(inPts), _i(inBatchIds)
    mn = torch.empty((batchSize, 3), dtype=_f32, device=pts.device)
    mx = torch.empty((batchSize, 3), dtype=_f32, device=pts.device)
    lib().ref_compute_aabb(_ci(int(scaleInv)), _ci(pts.shape[0]), _ci(batchSize), _p(pts), _p(bids), _p(mn), _p(mx))
    return mn, mx


def num_cells(aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    return lib().ref_num_cells(_ci(int(scaleInv)), _ci(batchSize), _cf(cellSize), _p(aabbMin), _p(aabbMax))


def sort_points_step1(inPts, inBatchIds, aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    pts, bids = _f(inPts), _i(inBatchIds)
    N = pts.shape[0]
    keys = torch.empty((N,), dtype=_i32, device=pts.device)
    idx = torch.empty((N,), dtype=_i32, device=pts.device)
    lib().ref_sort_points_step1(_ci(int(scaleInv)), _ci(N), _ci(batchSize), _cf(cellSize), _p(aabbMin), _p(aabbMax),
                                _p(pts), _p(bids), _p(keys), _p(idx))
    return keys, idx


def sort_points_step2(inPts, inBatchIds, inFeatures, keys, indexs, aabbMin, aabbMax, batchSize, cellSize, scaleInv):
    pts, bids, feats, keys, indexs = _f(inPts), _i(inBatchIds), _f(inFeatures), _i(keys), _i(indexs)
    n = num_cells(aabbMin, aabbMax, batchSize, cellSize, scaleInv)
    oP, oB, oF = torch.empty_like(pts), torch.empty_like(bids), torch.empty_like(feats)
    cells = torch.empty((batchSize, n, n, n, 2), dtype=_i32, device=pts.device)
    lib().ref_sort_points_step2(_ci(int(scaleInv)), _ci(pts.shape[0]), _ci(batchSize), _ci(feats.shape[1]),
                                _cf(cellSize), _p(aabbMin), _p(aabbMax), _p(pts), _p(bids), _p(feats), _p(keys),
                                _p(indexs), _p(oP), _p(oB), _p(oF), _p(cells))
    return oP, oB, oF, cells


def sort_features(inFeatures, indexs):
    f, idx = _f(inFeatures), _i(indexs)
    out = torch.empty_like(f)
    lib().ref_sort_features(_ci(f.shape[0]), _ci(f.shape[1]), _p(f), _p(idx), _p(out))
    return out


def sort_features_back(inFeatures, indexs):
    f, idx = _f(inFeatures), _i(indexs)
    out = torch.empty_like(f)
    lib().ref_sort_features_back(_ci(f.shape[0]), _ci(f.shape[1]), _p(f), _p(idx), _p(out))
    return out


def transform_indexs(inIndexs, inNewPositions):
    idx, new = _i(inIndexs), _i(inNewPositions)
    out = torch.empty_like(idx)
    lib().ref_transform_indexs(_ci(idx.shape[0]), _ci(new.shape[0]), _p(idx), _p(new), _p(out))
    return out


def find_neighbors(inPts, inBatchIds, inPts2, cellIndexs, aabbMin, aabbMax, radius, batchSize, scaleInv):
    s, sb, p2, cells = _f(inPts), _i(inBatchIds), _f(inPts2), _i(cellIndexs)
    M, n = s.shape[0], cells.shape[1]
    start = torch.empty((M, 1), dtype=_i32, device=s.device)
    K = lib().ref_find_neighbors_count(_ci(int(scaleInv)), _ci(M), _ci(n), _cf(radius), _p(s), _p(sb), _p(p2),
                                       _p(cells), _p(aabbMin), _p(aabbMax), _p(start))
    pairs = torch.empty((K, 2), dtype=_i32, device=s.device)
    lib().ref_find_neighbors_pack(_ci(int(scaleInv)), _ci(M), ctypes.c_uint(K), _ci(n), _cf(radius), _p(s), _p(sb),
                                  _p(p2), _p(cells), _p(aabbMin), _p(aabbMax), _p(start), _p(pairs))
    return start, pairs


def compute_pdf(inPts, inBatchIds, aabbMin, aabbMax, startIndexs, neighbors, window, radius, batchSize, scaleInv):
    pts, bids, start, pairs = _f(inPts), _i(inBatchIds), _i(startIndexs), _i(neighbors)
    K = pairs.shape[0]
    pdf = torch.empty((K, 1), dtype=_f32, device=pts.device)
    if K > 0:
        lib().ref_compute_pdf(_ci(int(scaleInv)), _cf(window), _ci(start.shape[0]), _ci(K), _cf(radius), _p(pts),
                              _p(bids), _p(aabbMin), _p(aabbMax), _p(start), _p(pairs), _p(pdf))
    return pdf


def poisson_sampling(inPts, inBatchIds, cellIndexs, aabbMin, aabbMax, radius, batchSize, scaleInv):
    pts, bids, cells = _f(inPts), _i(inBatchIds), _i(cellIndexs)
    N, n = pts.shape[0], cells.shape[1]
    M = lib().ref_poisson_sample(_ci(int(scaleInv)), _cf(radius), _ci(N), _ci(batchSize), _ci(n), _p(aabbMin),
                                 _p(aabbMax), _p(pts), _p(bids), _p(cells))
    oP = torch.empty((M, 3), dtype=_f32, device=pts.device)
    oB = torch.empty((M, 1), dtype=_i32, device=pts.device)
    oI = torch.empty((M,), dtype=_i32, device=pts.device)
    lib().ref_poisson_copy(_ci(M), _p(oP), _p(oB), _p(oI))
    return oP, oB, oI


def get_sampled_features(inSampledIndexs, pInFeatures):
    idx, f = _i(inSampledIndexs), _f(pInFeatures)
    out = torch.empty((idx.shape[0], f.shape[1]), dtype=_f32, device=f.device)
    if idx.shape[0] > 0:
        lib().ref_get_sampled_features(_ci(f.shape[0]), _ci(f.shape[1]), _ci(idx.shape[0]), _p(idx), _p(f), _p(out))
    return out


def spatial_conv(inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, neighStartIndexs, packedNeighs, aabbMin, aabbMax,
                 weights1, weights2, weightsOut, biases1, biases2, biasesOut, numOutFeatures, combin, batchSize,
                 radius, scaleInv, avg):
    pts, feats, bids, pdfs, samples = _f(inPts), _f(inFeatures), _i(inBatchIds), _f(inPDFs), _f(inSamplePts)
    start, pairs = _i(neighStartIndexs), _i(packedNeighs)
    w1, w2, wo, b1, b2, bo = (_f(t) for t in (weights1, weights2, weightsOut, biases1, biases2, biasesOut))
    M, K, C = samples.shape[0], pairs.shape[0], feats.shape[1]
    out = torch.empty((M, numOutFeatures), dtype=_f32, device=pts.device)
    lib().ref_spatial_conv(_ci(int(avg)), _ci(int(scaleInv)), _ci(K), _ci(C), _ci(numOutFeatures), _ci(M),
                           _ci(int(combin)), _cf(radius), _p(pts), _p(bids), _p(feats), _p(pdfs), _p(samples),
                           _p(start), _p(pairs), _p(aabbMin), _p(aabbMax), _p(w1), _p(b1), _p(w2), _p(b2), _p(wo),
                           _p(bo), _p(out))
    return out


def spatial_conv_grad(inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, neighStartIndexs, packedNeighs, aabbMin,
                      aabbMax, weights1, biases1, weights2, biases2, weightsOut, biasesOut, inOutFeatureGrads,
                      numOutFeatures, combin, batchSize, radius, scaleInv, avg):
    """SpatialConvGrad through the reference's spatialConvGradsCPU; op input order (spatial_conv.cc:56-96)."""
    pts, feats, bids, pdfs, samples = _f(inPts), _f(inFeatures), _i(inBatchIds), _f(inPDFs), _f(inSamplePts)
    start, pairs = _i(neighStartIndexs), _i(packedNeighs)
    w1, b1, w2, b2, wo, bo = (_f(t) for t in (weights1, biases1, weights2, biases2, weightsOut, biasesOut))
    og = _f(inOutFeatureGrads)
    M, K, C, N = samples.shape[0], pairs.shape[0], feats.shape[1], pts.shape[0]
    dF = torch.zeros((N, C), dtype=_f32, device=pts.device)
    g = [torch.zeros_like(t) for t in (w1, b1, w2, b2, wo, bo)]
    lib().ref_spatial_conv_grad(_ci(int(avg)), _ci(int(scaleInv)), _ci(K), _ci(C), _ci(numOutFeatures), _ci(M),
                                _ci(N), _ci(int(combin)), _cf(radius), _p(pts), _p(bids), _p(feats), _p(pdfs),
                                _p(samples), _p(start), _p(pairs), _p(aabbMin), _p(aabbMax), _p(w1), _p(b1), _p(w2),
                                _p(b2), _p(wo), _p(bo), _p(og), _p(dF), _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]),
                                _p(g[4]), _p(g[5]))
    return (dF,) + tuple(g)


def sort_points_step2_grad(indexs, outPtsGrad, outFeaturesGrad):
    idx, gp, gf = _i(indexs), _f(outPtsGrad), _f(outFeaturesGrad)
    ip, iF = torch.empty_like(gp), torch.empty_like(gf)
    lib().ref_sort_points_step2_grad(_ci(idx.shape[0]), _ci(gf.shape[1]), _p(gp), _p(gf), _p(idx), _p(ip), _p(iF))
    return ip, iF


def get_sampled_features_grad(inSampledIndexs, pInFeatures, inOutFeatureGrad):
    idx, f, g = _i(inSampledIndexs), _f(pInFeatures), _f(inOutFeatureGrad)
    out = torch.empty_like(f)
    lib().ref_get_sampled_features_grad(_ci(f.shape[0]), _ci(f.shape[1]), _ci(idx.shape[0]), _p(idx), _p(g), _p(out))
    return out
