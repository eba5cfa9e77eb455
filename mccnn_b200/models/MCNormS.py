"""MCNormS -- the smallest client of the op API (models/MCNormS.py:24-58): a single-level "hierarchy" and two
multi-feature convolutions (Cin -> k, k -> 3) of radius 0.15, KDE window 0.2."""
from ..MCConvBuilder import PointHierarchy
from ._net import Net


def create_network(points, batchIds, features, numInputFeatures, batchSize, k, isTraining, multiConv=True, useMC=True,
                   store=None):
    ph = PointHierarchy(points, features, batchIds, [], "MCNormS_PH", batchSize)
    net = Net(ph, 0.2, isTraining, False, False, store)
    x = net.conv("Conv_1", features, numInputFeatures, 0.15, 0, out_features=k)
    return net.conv("Conv_2", net.bn("BN_RELU", x), k, 0.15, 0, out_features=3)
