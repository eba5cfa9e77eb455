"""The reference's networks driven through the REFERENCE's own op sequence.  TEST / BASELINE INFRASTRUCTURE.

`bench.py --impl reference` (every --config) and the full-size GPU parity tests need the reference's forward on the GPU
box, where /root/reference -- and with it models/*.py and utils/MCConvBuilder.py -- does not exist.  This module
restates, for inference, what those files do:

  * `Builder`  = PointHierarchy (utils/MCConvBuilder.py:55-130: aabb once, then per level sort step1 + step2 WITH the
    features, poisson_sampling, get_sampled_features, transform_indexs) and ConvolutionBuilder.create_convolution
    (:249-427: the three caches and their key scheme :230-238; a grid-cache miss sorts the points together with the
    features, a hit re-sorts the features only; parameter names, shapes and the transpose-free reshapes :397-419);
  * the dense glue of utils/MCNetworkUtils.py:20-142 in inference mode (moving statistics, epsilon 1e-3) as torch ops
    -- TensorFlow / cuBLAS in the reference, torch / cuBLAS here;
  * the topologies of models/MCClassS.py:25-77, MCClassH.py:26-209, MCSeg.py:25-209 and MCSegScanNet.py:26-255.

It imports nothing from mccnn_b200.  `ops` is any module with the reference's Python op API (tf_ops/MCConvModuleSrc:
20-73): `oracle.ref` (the reference's own CUDA kernels, torch CUDA tensors) on the GPU box, or `CpuOps` (the C oracle
behind torch CPU tensors), under which tests/test_ref_nets_cpu.py holds every topology here to the golden outputs of
the reference's unchanged model files (tests/golden/model_*.npz).  `params` maps variable names to tensors; missing
ones are drawn by `new_param(name, shape)` (synthetic.named_variable in every caller).
"""
import math

import numpy as np
import torch

BN_EPS = 1e-3


class CpuOps:
    """oracle.cpu (NumPy in / out) behind torch CPU tensors, so that the same network code runs without a GPU."""

    def __init__(self):
        from . import cpu
        self._cpu = cpu

    def __getattr__(self, name):
        fn = getattr(self._cpu, name)

        def call(*args):
            conv = [a.detach().numpy() if isinstance(a, torch.Tensor) else a for a in args]
            out = fn(*conv)
            if isinstance(out, tuple):
                return tuple(torch.from_numpy(np.ascontiguousarray(o)) for o in out)
            return torch.from_numpy(np.ascontiguousarray(out))
        return call


class Builder:
    def __init__(self, ops, params, new_param, points, batchIds, features, radii, name, batchSize, relativeRadius=True,
                 window=0.25, log=None):
        self.ops, self.params, self.new_param = ops, params, new_param
        self.name, self.B, self.rel, self.window = name, batchSize, relativeRadius, window
        self.log = log if log is not None else []
        self.grids, self.neighs, self.pdfs = {}, {}, {}
        O = ops
        self.mn, self.mx = O.compute_aabb(points, batchIds, batchSize, relativeRadius)
        self.points, self.bids, self.features = [points], [batchIds], [features]
        p, b, f = points, batchIds, features
        for r in radii:
            keys, idx = O.sort_points_step1(p, b, self.mn, self.mx, batchSize, r, relativeRadius)
            sp, sb, sf, cells = O.sort_points_step2(p, b, f, keys, idx, self.mn, self.mx, batchSize, r, relativeRadius)
            p, b, qi = O.poisson_sampling(sp, sb, cells, self.mn, self.mx, r, batchSize, relativeRadius)
            f = O.get_sampled_features(qi, sf)
            O.transform_indexs(qi, idx)
            self.points.append(p)
            self.bids.append(b)
            self.features.append(f)
            self.log.append(("M", int(p.shape[0])))

    # ---- variables -------------------------------------------------------------------------------------------
    def var(self, name, shape):
        if name not in self.params:
            v = self.new_param(name, shape)
            if not isinstance(v, torch.Tensor):
                v = torch.from_numpy(np.ascontiguousarray(v))
            self.params[name] = v.to(self.points[0].device)
        v = self.params[name]
        assert tuple(v.shape) == tuple(shape), (name, tuple(v.shape), tuple(shape))
        return v

    # ---- Monte-Carlo convolution (create_convolution) ---------------------------------------------------------------
    def conv(self, name, x, c, radius, level, out_level=None, out_features=None, window=None):
        O = self.ops
        out_level = level if out_level is None else out_level
        window = self.window if window is None else window
        combin = out_features is not None
        n_out = out_features if combin else c
        key_grid = "%s|%d|%s|%s" % (self.name, level, radius, self.rel)
        key_neigh = key_grid + "|%s|%d" % (self.name, out_level)
        key_pdf = key_neigh + "|%s|True" % (window,)
        if key_grid in self.grids:
            sp, sb, cells, idx = self.grids[key_grid]
            sf = O.sort_features(x, idx)
        else:
            p, b = self.points[level], self.bids[level]
            keys, idx = O.sort_points_step1(p, b, self.mn, self.mx, self.B, radius, self.rel)
            sp, sb, sf, cells = O.sort_points_step2(p, b, x, keys, idx, self.mn, self.mx, self.B, radius, self.rel)
            self.grids[key_grid] = (sp, sb, cells, idx)
        if key_neigh not in self.neighs:
            self.neighs[key_neigh] = O.find_neighbors(self.points[out_level], self.bids[out_level], sp, cells, self.mn,
                                                      self.mx, radius, self.B, self.rel)
            self.log.append(("K", int(self.neighs[key_neigh][1].shape[0])))
        start, pairs = self.neighs[key_neigh]
        if key_pdf not in self.pdfs:
            self.pdfs[key_pdf] = O.compute_pdf(sp, sb, self.mn, self.mx, start, pairs, window, radius, self.B, self.rel)
        pdf = self.pdfs[key_pdf]
        neurons = c * n_out if combin else c
        nb = (neurons + 7) // 8
        w1 = self.var(name + "_weights", (3, 8 * nb))
        b1 = self.var(name + "_biases", (8 * nb,))
        w2 = self.var(name + "_weights2", (nb, 8, 8)).reshape(8, nb * 8)
        b2 = self.var(name + "_biases2", (nb, 8)).reshape(nb * 8)
        w3 = self.var(name + "_weights3", (nb, 8, 8)).reshape(8, nb * 8)
        b3 = self.var(name + "_biases3", (nb, 8)).reshape(nb * 8)
        self.log.append(("conv", int(pairs.shape[0]), int(c), int(n_out), bool(combin)))
        return O.spatial_conv(sp, sf, sb, pdf, self.points[out_level], start, pairs, self.mn, self.mx, w1, w2, w3, b1,
                              b2, b3, n_out, combin, self.B, radius, self.rel, True)

    # ---- dense glue (utils/MCNetworkUtils.py, inference) ----------------------------------------------------------------
    def batch_norm(self, name, x):
        c = x.shape[1]
        g, b = self.var(name + "/gamma", (c,)), self.var(name + "/beta", (c,))
        m, v = self.var(name + "/moving_mean", (c,)), self.var(name + "/moving_variance", (c,))
        scale = g * torch.rsqrt(v + BN_EPS)
        return x * scale + (b - m * scale)

    def bn(self, name, x):  # batch_norm_RELU_drop_out(name, ...) at inference
        return torch.relu(self.batch_norm(name + "_BN", x))

    def conv_1x1(self, name, x, c_in, c_out):
        return torch.addmm(self.var(name + "_biases", (c_out,)), x, self.var(name + "_weights", (c_in, c_out)))

    def reduce(self, name, x, c_in, c_out):
        return self.bn(name + "_Out_BN", self.conv_1x1(name, self.bn(name + "_In_BN", x), c_in, c_out))

    def mlp2(self, name, x, c_in, h1, h2, c_out, init_bn=True):  # MLP_2_hidden
        if init_bn:
            x = self.batch_norm(name + "_BN_Init", x)
        for l, (a, b) in enumerate([(c_in, h1), (h1, h2)], start=1):
            x = torch.addmm(self.var(name + "_biases%d" % l, (b,)), x, self.var(name + "_weights%d" % l, (a, b)))
            x = torch.relu(self.batch_norm(name + "_BN_h%d" % l, x))
        return torch.addmm(self.var(name + "_biases3", (c_out,)), x, self.var(name + "_weights3", (h2, c_out)))


R_MAX = math.sqrt(3.0) + 0.1


def mcclasss(ops, params, new_param, points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, log=None):
    n = Builder(ops, params, new_param, points, batchIds, features, [0.1, 0.4, R_MAX], "MCClassS_PH", batchSize,
                window=0.2, log=log)
    x = n.conv("Conv_1", features, numInputFeatures, 0.2, 0, 1, out_features=k)
    x = n.conv("Conv_2", n.reduce("Reduce_1", x, k, 2 * k), 2 * k, 0.8, 1, 2)
    x = n.conv("Conv_3", n.reduce("Reduce_2", x, 2 * k, 4 * k), 4 * k, R_MAX, 2, 3)
    return n.mlp2("Final_Logits", n.bn("BNRELUDROP_final", x), 4 * k, 2 * k, k, numOutCat)


def mcclassh(ops, params, new_param, points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, log=None):
    n = Builder(ops, params, new_param, points, batchIds, features, [0.1, 0.4, R_MAX], "MCClassH_PH", batchSize,
                window=0.25, log=log)

    def upper(x, c_in, sfx, pre):
        x = n.reduce("Reduce_Pool_2" + sfx, x, c_in, 8 * k)
        pool2 = n.conv("Pool_2" + sfx, x, 8 * k, 0.8, 1, 2, window=0.2)
        x = n.conv("Conv_3" + sfx, n.reduce("Reduce_Conv_3" + sfx, pool2, 8 * k, 8 * k), 8 * k, 1.2, 2)
        x = n.reduce("Reduce_Pool_3" + sfx, torch.cat([pool2, x], 1), 16 * k, 32 * k)
        pool3 = n.conv("Pool_3" + sfx, x, 32 * k, R_MAX, 2, 3, window=0.2)
        return n.mlp2(pre + "Final_Logits", n.bn(pre + "BNRELUDROP_final", pool3), 32 * k, 16 * k, 8 * k, numOutCat)

    x = n.conv("Conv_1", features, numInputFeatures, 0.1, 0, out_features=k)
    pool1 = n.conv("Pool_1", n.reduce("Reduce_Pool_1", x, k, 2 * k), 2 * k, 0.2, 0, 1, window=0.2)
    x = n.conv("Conv_2", n.reduce("Reduce_Conv_2", pool1, 2 * k, 2 * k), 2 * k, 0.4, 1)
    logits1 = upper(torch.cat([pool1, x], 1), 4 * k, "", "")
    x = n.conv("Conv_2_2", n.features[1], numInputFeatures, 0.4, 1, out_features=2 * k)
    logits2 = upper(x, 2 * k, "_2", "2")
    return logits1 + logits2  # inference: both path masks are 1, scale 2 / 2 (models/MCClassH.py:193-209)


def mcseg(ops, params, new_param, points, batchIds, features, catLabels, numInputFeatures, numCats, numParts,
          batchSize, k, log=None):
    n = Builder(ops, params, new_param, points, batchIds, features, [0.025, 0.1, 0.4], "MCSeg_PH", batchSize,
                window=0.25, log=log)
    cat = lambda *xs: torch.cat(list(xs), 1)  # noqa: E731
    conv1 = n.conv("Conv_1", features, numInputFeatures, 0.03, 0, out_features=k)
    pool1 = n.conv("Pool_1", n.reduce("Reduce_Pool_1", conv1, k, 2 * k), 2 * k, 0.05, 0, 1, window=0.2)
    conv2 = cat(pool1, n.conv("Conv_2", n.bn("Conv_2_In_BN", pool1), 2 * k, 0.1, 1))
    pool2 = n.conv("Pool_2", n.reduce("Reduce_Pool_2", conv2, 4 * k, 4 * k), 4 * k, 0.2, 1, 2, window=0.2)
    conv3 = cat(pool2, n.conv("Conv_3", n.bn("Conv_3_In_BN", pool2), 4 * k, 0.4, 2))
    pool3 = n.conv("Pool_3", n.reduce("Reduce_Pool_3", conv3, 8 * k, 8 * k), 8 * k, 0.8, 2, 3, window=0.2)
    conv4 = cat(pool3, n.conv("Conv_4", n.bn("Conv_4_In_BN", pool3), 8 * k, R_MAX, 3))
    up34 = n.conv("Up_3_4", n.bn("Up_3_4_BN", conv4), 16 * k, R_MAX, 3, 2)
    dec3 = n.conv("DeConv_3", n.reduce("DeConv_3_Reduce", cat(up34, conv3), 24 * k, 8 * k), 8 * k, 0.4, 2)
    up23 = n.conv("Up_2_3", n.bn("Up_2_3_BN", dec3), 8 * k, 0.2, 2, 1)
    dec2 = n.conv("DeConv_2", n.reduce("DeConv_2_Reduce", cat(up23, conv2), 12 * k, 4 * k), 4 * k, 0.1, 1)
    up12 = n.conv("Up_1_2", n.bn("Up_1_2_BN", dec2), 4 * k, 0.05, 1, 0)
    up13 = n.conv("Up_1_3", n.bn("Up_1_3_BN", dec3), 8 * k, 0.2, 2, 0)
    dec1 = n.conv("DeConv_1", n.reduce("DeConv_1_Reduce", cat(up12, up13, conv1), 13 * k, 4 * k), 4 * k, 0.03, 0)
    final = n.bn("BNRELUDROP_hier_final", dec1)
    onehot = torch.nn.functional.one_hot(catLabels.reshape(-1).long(), numCats).to(torch.float32)
    return n.mlp2("Final_Logits", cat(onehot, final), 4 * k + numCats, 4 * k, 2 * k, numParts, init_bn=False)


def mcsegscannet(ops, params, new_param, points, batchIds, features, numInputFeatures, numSem, batchSize, k, log=None):
    n = Builder(ops, params, new_param, points, batchIds, features, [0.1, 0.2, 0.4, 0.8], "MCSegScanNet_PH", batchSize,
                relativeRadius=False, window=0.25, log=log)
    cat = lambda *xs: torch.cat(list(xs), 1)  # noqa: E731
    pool0 = n.conv("Pool_0", features, numInputFeatures, 0.1, 0, 1, out_features=k, window=0.2)
    conv1 = cat(pool0, n.conv("Conv_1", n.bn("Conv_1_In_BN", pool0), k, 0.4, 1))
    pool1 = n.conv("Pool_1", n.reduce("Reduce_Pool_1", conv1, 2 * k, 2 * k), 2 * k, 0.4, 1, 2, window=0.2)
    conv2 = cat(pool1, n.conv("Conv_2", n.bn("Conv_2_In_BN", pool1), 2 * k, 0.8, 2))
    pool2 = n.conv("Pool_2", n.reduce("Reduce_Pool_2", conv2, 4 * k, 4 * k), 4 * k, 0.8, 2, 3, window=0.2)
    conv3 = cat(pool2, n.conv("Conv_3", n.bn("Conv_3_In_BN", pool2), 4 * k, 1.6, 3))
    pool3 = n.conv("Pool_3", n.reduce("Reduce_Pool_3", conv3, 8 * k, 8 * k), 8 * k, 1.6, 3, 4, window=0.2)
    conv4 = cat(pool3, n.conv("Conv_4", n.bn("Conv_4_In_BN", pool3), 8 * k, 5.0, 4))
    up34 = n.conv("Up_3_4", n.reduce("Up3_4_Reduce", conv4, 16 * k, 8 * k), 8 * k, 1.6, 4, 3)
    dec3 = n.conv("DeConv_3", n.reduce("DeConv_3_Reduce", cat(up34, conv3), 16 * k, 8 * k), 8 * k, 1.6, 3)
    up23 = n.conv("Up_2_3", n.reduce("Up2_3_Reduce", dec3, 8 * k, 4 * k), 4 * k, 0.8, 3, 2)
    dec2 = n.conv("DeConv_2", n.reduce("DeConv_2_Reduce", cat(up23, conv2), 8 * k, 4 * k), 4 * k, 0.8, 2)
    up12 = n.conv("Up_1_2", n.reduce("Up1_2_Reduce", dec2, 4 * k, 2 * k), 2 * k, 0.4, 2, 1)
    up13 = n.conv("Up_1_3", n.reduce("Up1_3_Reduce", dec3, 8 * k, 2 * k), 2 * k, 0.8, 3, 1)
    up14 = n.conv("Up_1_4", n.reduce("Up1_4_Reduce", conv4, 16 * k, 2 * k), 2 * k, 1.6, 4, 1)
    ups = cat(up14, up13, up12, conv1)
    dec1 = n.conv("DeConv_1", n.reduce("DeConv_1_Reduce", ups, 8 * k, 4 * k), 4 * k, 0.4, 1)
    final = n.conv("Up_0_1", n.reduce("Up_Final_Reduce", cat(ups, dec1), 12 * k, 4 * k), 4 * k, 0.1, 1, 0)
    return n.mlp2("Final_Logits", n.bn("BNRELUDROP_hier_final", final), 4 * k, 4 * k, 2 * k, numSem, init_bn=False)


NETWORKS = {"MCClassS": mcclasss, "MCClassH": mcclassh, "MCSeg": mcseg, "MCSegScanNet": mcsegscannet}


def forward(model, ops, params, new_param, inputs, args, log=None):
    """inputs: dict with points / batchIds / features (/ catLabels) tensors; args: the inference-relevant positional
    arguments of the reference's create_network after those (synthetic cases: tests/common.model_case)."""
    lead = [inputs["points"], inputs["batchIds"], inputs["features"]]
    if model == "MCSeg":
        lead.append(inputs["catLabels"])
    return NETWORKS[model](ops, params, new_param, *lead, *args, log=log)
