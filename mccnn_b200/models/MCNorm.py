"""MCNorm -- encoder/decoder normal-estimation network (models/MCNorm.py:24-146): hierarchy radii [0.1, 0.4], nine
convolutions, the last one a multi-feature 2k -> 3 convolution that outputs the normals."""
import math

from ..MCConvBuilder import PointHierarchy
from ._net import Net


def create_network(points, batchIds, features, numInputFeatures, batchSize, k, isTraining, multiConv=True, useMC=True,
                   store=None):
    ph = PointHierarchy(points, features, batchIds, [0.1, 0.4], "MCNorm_PH", batchSize)
    net = Net(ph, 0.25, isTraining, False, False, store)

    conv1 = net.conv("Conv_1", features, numInputFeatures, 0.1, 0, out_features=k)
    pool1 = net.conv("Pool_1", net.reduce("Reduce_Pool_1", conv1, k, 2 * k), 2 * k, 0.2, 0, 1, window=0.2)
    conv2 = net.cat(pool1, net.conv("Conv_2", net.bn("Conv_2_In_BN", pool1), 2 * k, 0.4, 1))
    pool2 = net.conv("Pool_2", net.reduce("Reduce_Pool_2", conv2, 4 * k, 4 * k), 4 * k, 0.8, 1, 2, window=0.2)
    conv3 = net.cat(pool2, net.conv("Conv_3", net.bn("Conv_3_In_BN", pool2), 4 * k, math.sqrt(3), 2))

    up23 = net.conv("Up_2_3", net.bn("Up_2_3_BN", conv3), 8 * k, 0.8, 2, 1)
    dec2 = net.reduce("DeConv_2_Reduce", net.cat(up23, conv2), 12 * k, 4 * k)
    dec2 = net.conv("DeConv_2", dec2, 4 * k, 0.4, 1)
    up12 = net.conv("Up_1_2", net.bn("Up_1_2_BN", dec2), 4 * k, 0.2, 1, 0)
    dec1 = net.reduce("DeConv_1_Reduce", net.cat(up12, conv1), 5 * k, 2 * k)
    return net.conv("DeConv_1", dec1, 2 * k, 0.1, 0, out_features=3)
