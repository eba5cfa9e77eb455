"""Vectorised NumPy restatement of compute_pdf and spatial_conv (the "NumPy CPU re-execution" BASELINE.json
config 1 names).  TEST / BASELINE INFRASTRUCTURE.

Follows tf_ops/compute_pdf.cu:57-92 and tf_ops/spatial_conv.cu:49-78,131-175; accumulates in float64 so that it
is the order-free value the float32 implementations (reference atomics, oracle C loop, CUDA kernels) approach.
"""
import numpy as np

BLOCK = 8


def _extent(aabbMin, aabbMax):
    return (aabbMax - aabbMin).max(axis=1).astype(np.float32)


def _radius_per_cloud(aabbMin, aabbMax, radius, scaleInv):
    ext = _extent(aabbMin, aabbMax)
    return (np.float32(radius) * ext) if scaleInv else np.full_like(ext, np.float32(radius))


def _segments(start, K):
    start = start.reshape(-1).astype(np.int64)
    end = np.concatenate([start[1:], [K]])
    return start, end


def compute_pdf(inPts, inBatchIds, aabbMin, aabbMax, startIndexs, neighbors, window, radius, batchSize, scaleInv):
    pts = inPts.astype(np.float64)
    bids = inBatchIds.reshape(-1)
    K = neighbors.shape[0]
    start, end = _segments(startIndexs, K)
    r = _radius_per_cloud(aabbMin, aabbMax, radius, scaleInv).astype(np.float64)
    h = float(np.float32(window))
    pdf = np.zeros((K,), np.float64)
    for i in range(start.shape[0]):
        lo, hi = start[i], end[i]
        if hi <= lo:
            continue
        js = neighbors[lo:hi, 0]
        x = pts[js]
        s = 1.0 / (r[bids[js]] * h)  # per pair (all in one cloud)
        d = (x[None, :, :] - x[:, None, :]) * s[:, None, None]
        g = ((0.39894228 / h) ** 3) * np.exp(-0.5 * (d * d).sum(-1))
        pdf[lo:hi] = g.sum(1) / float(hi - lo)
    return pdf.astype(np.float32).reshape(-1, 1)


def spatial_conv(inPts, inFeatures, inBatchIds, inPDFs, inSamplePts, neighStartIndexs, packedNeighs, aabbMin, aabbMax,
                 weights1, weights2, weightsOut, biases1, biases2, biasesOut, numOutFeatures, combin, batchSize,
                 radius, scaleInv, avg):
    K = packedNeighs.shape[0]
    M = inSamplePts.shape[0]
    C = inFeatures.shape[1]
    nOut = C * numOutFeatures if combin else C
    nb = (nOut + BLOCK - 1) // BLOCK
    j = packedNeighs[:, 0].astype(np.int64)
    i = packedNeighs[:, 1].astype(np.int64)
    start, end = _segments(neighStartIndexs, K)
    r = _radius_per_cloud(aabbMin, aabbMax, radius, scaleInv)[inBatchIds.reshape(-1)[j]]
    delta = ((inPts[j] - inSamplePts[i]) / r[:, None]).astype(np.float64)              # [K,3]
    cnt = (end - start)[i].astype(np.float64) if avg else np.ones((K,), np.float64)
    w = inPDFs.reshape(-1).astype(np.float64) * cnt
    # flat reinterpretations (spatial_conv.cu:49-73,172)
    W1 = weights1.reshape(-1)[: nb * BLOCK * 3].reshape(nb * BLOCK, 3).astype(np.float64)
    b1 = biases1.reshape(-1)[: nb * BLOCK].astype(np.float64)
    W2 = weights2.reshape(-1)[: nb * 64].reshape(nb, BLOCK, BLOCK).astype(np.float64)     # [q, t, u]
    b2 = biases2.reshape(-1)[: nb * BLOCK].reshape(nb, BLOCK).astype(np.float64)
    Wo = weightsOut.reshape(-1)[: nb * 64].reshape(nb, BLOCK, BLOCK).astype(np.float64)
    bo = biasesOut.reshape(-1)[: nb * BLOCK].reshape(nb, BLOCK).astype(np.float64)
    h1 = np.maximum(delta @ W1.T + b1, 0.0).reshape(K, nb, BLOCK)                          # [K,q,u]
    h2 = np.maximum(np.einsum("kqu,qtu->kqt", h1, W2) + b2, 0.0)
    kappa = (np.einsum("kqu,qtu->kqt", h2, Wo) + bo).reshape(K, nb * BLOCK)[:, :nOut]      # [K,nOut]
    neuron = np.arange(nOut)
    c = neuron % C
    o = neuron // C if combin else c
    contrib = inFeatures.astype(np.float64)[j][:, c] * kappa / w[:, None]                  # [K,nOut]
    outC = numOutFeatures if combin else C
    out = np.zeros((M, outC), np.float64)
    # sum over pairs of a sample (CSR rows), then over neurons sharing an output feature
    per_sample = np.zeros((M, nOut), np.float64)
    np.add.at(per_sample, i, contrib)
    np.add.at(out.T, o, per_sample.T)
    return out.astype(np.float32)
