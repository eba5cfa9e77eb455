"""MCClassS forward driven through the REFERENCE's own CUDA kernels (oracle/ref.py -> oracle/_ref/libmccnn_ref.so).

BASELINE INFRASTRUCTURE -- this is the "reference arm" of bench.py (`--impl reference`) and the ground truth of the
end-to-end GPU parity test.  It imports nothing from mccnn_b200: the op sequence restates utils/MCConvBuilder.py:55-130
(PointHierarchy) and :249-427 (create_convolution) for the MCClassS topology (models/MCClassS.py:25-77), every op is
the reference's kernel choreographed like its TensorFlow glue (blocking count read-backs included), and the dense glue
(inference batch norm, 1x1 matmuls, final MLP -- TensorFlow/cuBLAS in the reference) runs through torch/cuBLAS.
`params` maps the reference's variable names to torch CUDA tensors.
"""
import math

import torch

from . import ref as R

BN_EPS = 1e-3


def _bn(x, p, name):
    scale = p[name + "/gamma"] * torch.rsqrt(p[name + "/moving_variance"] + BN_EPS)
    return x * scale + (p[name + "/beta"] - p[name + "/moving_mean"] * scale)


def _dense(x, p, w, b):
    return torch.addmm(p[b], x, p[w])


def hierarchy(points, batchIds, features, radii, batchSize, relativeRadius=True):
    mn, mx = R.compute_aabb(points, batchIds, batchSize, relativeRadius)
    levels = [(points, batchIds, features)]
    p, b, f = points, batchIds, features
    for r in radii:
        keys, idx = R.sort_points_step1(p, b, mn, mx, batchSize, r, relativeRadius)
        sp, sb, sf, cells = R.sort_points_step2(p, b, f, keys, idx, mn, mx, batchSize, r, relativeRadius)
        qp, qb, qi = R.poisson_sampling(sp, sb, cells, mn, mx, r, batchSize, relativeRadius)
        qf = R.get_sampled_features(qi, sf)
        R.transform_indexs(qi, idx)
        levels.append((qp, qb, qf))
        p, b, f = qp, qb, qf
    return mn, mx, levels


def convolution(p, name, mn, mx, levels, inLevel, outLevel, inFeatures, numIn, numOut, radius, combin, window,
                batchSize, relativeRadius=True, avg=True):
    pts, bids, _ = levels[inLevel]
    keys, idx = R.sort_points_step1(pts, bids, mn, mx, batchSize, radius, relativeRadius)
    sp, sb, sf, cells = R.sort_points_step2(pts, bids, inFeatures, keys, idx, mn, mx, batchSize, radius,
                                            relativeRadius)
    op, ob, _ = levels[outLevel]
    start, pairs = R.find_neighbors(op, ob, sp, cells, mn, mx, radius, batchSize, relativeRadius)
    pdf = R.compute_pdf(sp, sb, mn, mx, start, pairs, window, radius, batchSize, relativeRadius)
    nOut = numIn * numOut if combin else numIn
    nb = (nOut + 7) // 8
    return R.spatial_conv(sp, sf, sb, pdf, op, start, pairs, mn, mx, p[name + "_weights"],
                          p[name + "_weights2"].reshape(8, nb * 8), p[name + "_weights3"].reshape(8, nb * 8),
                          p[name + "_biases"], p[name + "_biases2"].reshape(nb * 8),
                          p[name + "_biases3"].reshape(nb * 8), numOut, combin, batchSize, radius, relativeRadius, avg)


def mcclasss_forward(points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, p):
    radii = [0.1, 0.4, math.sqrt(3.0) + 0.1]
    mn, mx, levels = hierarchy(points, batchIds, features, radii, batchSize)
    x = convolution(p, "Conv_1", mn, mx, levels, 0, 1, features, numInputFeatures, k, 0.2, True, 0.2, batchSize)
    x = torch.relu(_bn(x, p, "Reduce_1_In_BN_BN"))
    x = _dense(x, p, "Reduce_1_weights", "Reduce_1_biases")
    x = torch.relu(_bn(x, p, "Reduce_1_Out_BN_BN"))
    x = convolution(p, "Conv_2", mn, mx, levels, 1, 2, x, 2 * k, 2 * k, 0.8, False, 0.2, batchSize)
    x = torch.relu(_bn(x, p, "Reduce_2_In_BN_BN"))
    x = _dense(x, p, "Reduce_2_weights", "Reduce_2_biases")
    x = torch.relu(_bn(x, p, "Reduce_2_Out_BN_BN"))
    x = convolution(p, "Conv_3", mn, mx, levels, 2, 3, x, 4 * k, 4 * k, math.sqrt(3.0) + 0.1, False, 0.2, batchSize)
    x = torch.relu(_bn(x, p, "BNRELUDROP_final_BN"))
    x = _bn(x, p, "Final_Logits_BN_Init")
    x = torch.relu(_bn(_dense(x, p, "Final_Logits_weights1", "Final_Logits_biases1"), p, "Final_Logits_BN_h1"))
    x = torch.relu(_bn(_dense(x, p, "Final_Logits_weights2", "Final_Logits_biases2"), p, "Final_Logits_BN_h2"))
    return _dense(x, p, "Final_Logits_weights3", "Final_Logits_biases3")
