"""MCClassS -- the reference's small classification network (models/MCClassS.py:25-77), the configuration the
headline metric is quoted on: hierarchy radii [0.1, 0.4, sqrt(3)+0.1]; three pooling convolutions Conv_1 L0->L1
(multi-feature, r 0.2), Conv_2 L1->L2 (depth-wise, r 0.8), Conv_3 L2->L3 (depth-wise, r sqrt(3)+0.1) with a
BN-ReLU / 1x1 / BN-ReLU reduction between them; MLP_2_hidden head; KDE window 0.2."""
import math

from ..MCConvBuilder import PointHierarchy
from ..MCNetworkUtils import MLP_2_hidden
from ._net import Net


def create_network(points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, isTraining, keepProbConv,
                   keepProbFull, useConvDropOut=False, useDropOutFull=True, store=None):
    rMax = math.sqrt(3.0) + 0.1
    ph = PointHierarchy(points, features, batchIds, [0.1, 0.4, rMax], "MCClassS_PH", batchSize)
    net = Net(ph, 0.2, isTraining, useConvDropOut, keepProbConv, store)

    x = net.conv("Conv_1", features, numInputFeatures, 0.2, 0, 1, out_features=k)
    x = net.conv("Conv_2", net.reduce("Reduce_1", x, k, 2 * k), 2 * k, 0.8, 1, 2)
    x = net.conv("Conv_3", net.reduce("Reduce_2", x, 2 * k, 4 * k), 4 * k, rMax, 2, 3)

    return MLP_2_hidden(net.bn("BNRELUDROP_final", x), 4 * k, 2 * k, k, numOutCat, "Final_Logits", keepProbFull,
                        isTraining, useDropOutFull, store=store)
