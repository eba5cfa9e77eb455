"""MCClassH -- two-path hierarchical classification network (models/MCClassH.py:26-209).  Path 1 starts at level 0,
path 2 at level 1 from the hierarchy's own level-1 features (`features_[1]`); the two logit vectors are averaged.
The reference drops one path at random while training (models/MCClassH.py:193-209); at inference both masks are 1 and
the result is (2 / 2) * (logits1 + logits2)."""
import math

from ..MCConvBuilder import PointHierarchy
from ..MCNetworkUtils import MLP_2_hidden
from ._net import Net


def create_network(points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, isTraining, keepProbConv,
                   keepProbFull, useConvDropOut=False, useDropOutFull=True, store=None):
    rMax = math.sqrt(3.0) + 0.1
    ph = PointHierarchy(points, features, batchIds, [0.1, 0.4, rMax], "MCClassH_PH", batchSize)
    net = Net(ph, 0.25, isTraining, useConvDropOut, keepProbConv, store)

    def upper(pool1_or_conv2, sfx):
        """levels 1 -> 3, shared shape of both paths; `sfx` = "" (path 1) or "_2" (path 2)"""
        x = net.reduce("Reduce_Pool_2" + sfx, pool1_or_conv2, *((4 * k, 8 * k) if sfx == "" else (2 * k, 8 * k)))
        pool2 = net.conv("Pool_2" + sfx, x, 8 * k, 0.8, 1, 2, window=0.2)
        x = net.reduce("Reduce_Conv_3" + sfx, pool2, 8 * k, 8 * k)
        x = net.conv("Conv_3" + sfx, x, 8 * k, 1.2, 2)
        x = net.reduce("Reduce_Pool_3" + sfx, net.cat(pool2, x), 16 * k, 32 * k)
        pool3 = net.conv("Pool_3" + sfx, x, 32 * k, rMax, 2, 3, window=0.2)
        pre = "" if sfx == "" else "2"
        return MLP_2_hidden(net.bn(pre + "BNRELUDROP_final", pool3), 32 * k, 16 * k, 8 * k, numOutCat,
                            pre + "Final_Logits", keepProbFull, isTraining, useDropOutFull, store=store)

    # path 1
    x = net.conv("Conv_1", features, numInputFeatures, 0.1, 0, out_features=k)
    x = net.reduce("Reduce_Pool_1", x, k, 2 * k)
    pool1 = net.conv("Pool_1", x, 2 * k, 0.2, 0, 1, window=0.2)
    x = net.reduce("Reduce_Conv_2", pool1, 2 * k, 2 * k)
    x = net.conv("Conv_2", x, 2 * k, 0.4, 1)
    logits1 = upper(net.cat(pool1, x), "")

    # path 2
    x = net.conv("Conv_2_2", ph.features_[1], numInputFeatures, 0.4, 1, out_features=2 * k)
    logits2 = upper(x, "_2")

    if not isTraining:
        return logits1 + logits2
    # training: one path is dropped at random (models/MCClassH.py:193-209): p <= 0.66 keeps path 1, p >= 0.33 keeps
    # path 2, the sum is rescaled by 2 / (number of kept paths)
    import torch
    p = float(torch.rand(1).item())
    m1, m2 = float(p <= 0.66), float(p >= 0.33)
    return (2.0 / (m1 + m2)) * (m1 * logits1 + m2 * logits2)
