"""MCClass -- classification network with convolutions inside each level and pooling convolutions between levels
(models/MCClass.py:25-116): hierarchy radii [0.1, 0.4, sqrt(3)+0.1], KDE window 0.25 (0.2 for the poolings)."""
import math

from ..MCConvBuilder import PointHierarchy
from ..MCNetworkUtils import MLP_2_hidden
from ._net import Net


def create_network(points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, isTraining, keepProbConv,
                   keepProbFull, useConvDropOut=False, useDropOutFull=True, store=None):
    rMax = math.sqrt(3.0) + 0.1
    ph = PointHierarchy(points, features, batchIds, [0.1, 0.4, rMax], "MCClass_PH", batchSize)
    net = Net(ph, 0.25, isTraining, useConvDropOut, keepProbConv, store)

    x = net.conv("Conv_1", features, numInputFeatures, 0.1, 0, out_features=k)
    x = net.reduce("Reduce_Pool_1", x, k, 2 * k)
    pool1 = net.conv("Pool_1", x, 2 * k, 0.2, 0, 1, window=0.2)

    x = net.conv("Conv_2", net.bn("Conv_2_In_BN", pool1), 2 * k, 0.4, 1)
    x = net.reduce("Reduce_Pool_2", net.cat(pool1, x), 4 * k, 8 * k)
    pool2 = net.conv("Pool_2", x, 8 * k, 0.8, 1, 2, window=0.2)

    x = net.conv("Conv_3", net.bn("Conv_3_In_BN", pool2), 8 * k, 1.1, 2)
    x = net.reduce("Reduce_Pool_3", net.cat(pool2, x), 16 * k, 32 * k)
    pool3 = net.conv("Pool_3", x, 32 * k, rMax, 2, 3, window=0.2)

    return MLP_2_hidden(net.bn("BNRELUDROP_final", pool3), 32 * k, 16 * k, 8 * k, numOutCat, "Final_Logits",
                        keepProbFull, isTraining, useDropOutFull, store=store)
