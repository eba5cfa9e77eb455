"""MCClassS forward assembled from the CPU oracle ops + NumPy dense glue.  TEST / BASELINE INFRASTRUCTURE.

Follows the reference's call sequence exactly: PointHierarchy (utils/MCConvBuilder.py:55-130), the three
create_convolution calls with their grid / neighbour / pdf steps (utils/MCConvBuilder.py:249-427) and the BN-ReLU /
1x1 / MLP glue of models/MCClassS.py:25-77 and utils/MCNetworkUtils.py:20-142 (inference-mode batch norm).
`params` maps variable names (as created by mccnn_b200.MCConvBuilder / MCNetworkUtils == the reference's names) to
numpy arrays.  Returns (logits [B, numCat], trace) where trace records the sizes of every stage.
"""
import math

import numpy as np

from . import cpu as O

BN_EPS = 1e-3


def _bn(x, params, name):
    g, b = params[name + "/gamma"], params[name + "/beta"]
    m, v = params[name + "/moving_mean"], params[name + "/moving_variance"]
    scale = (g / np.sqrt(v + np.float32(BN_EPS))).astype(np.float32)
    return (x * scale + (b - m * scale)).astype(np.float32)


def _relu(x):
    return np.maximum(x, np.float32(0))


def _dense(x, params, wname, bname):
    return (x.astype(np.float32) @ params[wname] + params[bname]).astype(np.float32)


def hierarchy(points, batchIds, features, radii, batchSize, relativeRadius=True):
    mn, mx = O.compute_aabb(points, batchIds, batchSize, relativeRadius)
    levels = [(points, batchIds, features)]
    sampled = []
    p, b, f = points, batchIds, features
    for r in radii:
        keys, idx = O.sort_points_step1(p, b, mn, mx, batchSize, r, relativeRadius)
        sp, sb, sf, cells = O.sort_points_step2(p, b, f, keys, idx, mn, mx, batchSize, r, relativeRadius)
        qp, qb, qi = O.poisson_sampling(sp, sb, cells, mn, mx, r, batchSize, relativeRadius)
        qf = O.get_sampled_features(qi, sf)
        sampled.append(O.transform_indexs(qi, idx))
        levels.append((qp, qb, qf))
        p, b, f = qp, qb, qf
    return mn, mx, levels, sampled


def convolution(params, name, mn, mx, levels, inLevel, outLevel, inFeatures, numIn, numOut, radius, combin, window,
                batchSize, relativeRadius=True, avg=True, trace=None):
    p, b, _ = levels[inLevel]
    keys, idx = O.sort_points_step1(p, b, mn, mx, batchSize, radius, relativeRadius)
    sp, sb, sf, cells = O.sort_points_step2(p, b, inFeatures, keys, idx, mn, mx, batchSize, radius, relativeRadius)
    op, ob, _ = levels[outLevel]
    start, pairs = O.find_neighbors(op, ob, sp, cells, mn, mx, radius, batchSize, relativeRadius)
    pdf = O.compute_pdf(sp, sb, mn, mx, start, pairs, window, radius, batchSize, relativeRadius)
    nOut = numIn * numOut if combin else numIn
    nb = (nOut + 7) // 8
    w1 = params[name + "_weights"]
    w2 = params[name + "_weights2"].reshape(8, nb * 8)
    w3 = params[name + "_weights3"].reshape(8, nb * 8)
    b1 = params[name + "_biases"]
    b2 = params[name + "_biases2"].reshape(nb * 8)
    b3 = params[name + "_biases3"].reshape(nb * 8)
    if trace is not None:
        trace.append(dict(conv=name, N=int(p.shape[0]), M=int(op.shape[0]), K=int(pairs.shape[0]), C=numIn, Co=numOut,
                          nb=nb, sumk2=int((np.diff(np.append(start.reshape(-1), pairs.shape[0])) ** 2).sum())))
    return O.spatial_conv(sp, sf, sb, pdf, op, start, pairs, mn, mx, w1, w2, w3, b1, b2, b3, numOut, combin, batchSize,
                          radius, relativeRadius, avg)


def mcclasss_forward(points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, params):
    trace = []
    radii = [0.1, 0.4, math.sqrt(3.0) + 0.1]
    mn, mx, levels, _ = hierarchy(points, batchIds, features, radii, batchSize)
    trace.append(dict(levels=[int(l[0].shape[0]) for l in levels]))
    x = convolution(params, "Conv_1", mn, mx, levels, 0, 1, features, numInputFeatures, k, 0.2, True, 0.2, batchSize,
                    trace=trace)
    x = _relu(_bn(x, params, "Reduce_1_In_BN_BN"))
    x = _dense(x, params, "Reduce_1_weights", "Reduce_1_biases")
    x = _relu(_bn(x, params, "Reduce_1_Out_BN_BN"))
    x = convolution(params, "Conv_2", mn, mx, levels, 1, 2, x, 2 * k, 2 * k, 0.8, False, 0.2, batchSize, trace=trace)
    x = _relu(_bn(x, params, "Reduce_2_In_BN_BN"))
    x = _dense(x, params, "Reduce_2_weights", "Reduce_2_biases")
    x = _relu(_bn(x, params, "Reduce_2_Out_BN_BN"))
    x = convolution(params, "Conv_3", mn, mx, levels, 2, 3, x, 4 * k, 4 * k, math.sqrt(3.0) + 0.1, False, 0.2,
                    batchSize, trace=trace)
    x = _relu(_bn(x, params, "BNRELUDROP_final_BN"))
    x = _bn(x, params, "Final_Logits_BN_Init")
    x = _relu(_bn(_dense(x, params, "Final_Logits_weights1", "Final_Logits_biases1"), params, "Final_Logits_BN_h1"))
    x = _relu(_bn(_dense(x, params, "Final_Logits_weights2", "Final_Logits_biases2"), params, "Final_Logits_BN_h2"))
    x = _dense(x, params, "Final_Logits_weights3", "Final_Logits_biases3")
    return x, trace
