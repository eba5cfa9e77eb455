"""MCSeg -- encoder/decoder part-segmentation network for ShapeNet-like batches (models/MCSeg.py:25-209): hierarchy
radii [0.025, 0.1, 0.4], 14 Monte-Carlo convolutions (all depth-wise except Conv_1), per-point logits; the object
category enters as a one-hot vector concatenated in front of the last feature block."""
import math

import torch

from ..MCConvBuilder import PointHierarchy
from ..MCNetworkUtils import MLP_2_hidden, concat
from ._net import Net


def create_network(points, batchIds, features, catLabels, numInputFeatures, numCats, numParts, batchSize, k,
                   isTraining, keepProbConv, keepProbFull, useConvDropOut=False, useDropOutFull=True, store=None):
    rMax = math.sqrt(3.0) + 0.1
    ph = PointHierarchy(points, features, batchIds, [0.025, 0.1, 0.4], "MCSeg_PH", batchSize)
    net = Net(ph, 0.25, isTraining, useConvDropOut, keepProbConv, store)

    # encoder
    conv1 = net.conv("Conv_1", features, numInputFeatures, 0.03, 0, out_features=k)
    pool1 = net.conv("Pool_1", net.reduce("Reduce_Pool_1", conv1, k, 2 * k), 2 * k, 0.05, 0, 1, window=0.2)
    conv2 = net.cat(pool1, net.conv("Conv_2", net.bn("Conv_2_In_BN", pool1), 2 * k, 0.1, 1))
    pool2 = net.conv("Pool_2", net.reduce("Reduce_Pool_2", conv2, 4 * k, 4 * k), 4 * k, 0.2, 1, 2, window=0.2)
    conv3 = net.cat(pool2, net.conv("Conv_3", net.bn("Conv_3_In_BN", pool2), 4 * k, 0.4, 2))
    pool3 = net.conv("Pool_3", net.reduce("Reduce_Pool_3", conv3, 8 * k, 8 * k), 8 * k, 0.8, 2, 3, window=0.2)
    conv4 = net.cat(pool3, net.conv("Conv_4", net.bn("Conv_4_In_BN", pool3), 8 * k, rMax, 3))

    # decoder
    up34 = net.conv("Up_3_4", net.bn("Up_3_4_BN", conv4), 16 * k, rMax, 3, 2)
    dec3 = net.reduce("DeConv_3_Reduce", net.cat(up34, conv3), 24 * k, 8 * k)
    dec3 = net.conv("DeConv_3", dec3, 8 * k, 0.4, 2)
    up23 = net.conv("Up_2_3", net.bn("Up_2_3_BN", dec3), 8 * k, 0.2, 2, 1)
    dec2 = net.reduce("DeConv_2_Reduce", net.cat(up23, conv2), 12 * k, 4 * k)
    dec2 = net.conv("DeConv_2", dec2, 4 * k, 0.1, 1)
    up12 = net.conv("Up_1_2", net.bn("Up_1_2_BN", dec2), 4 * k, 0.05, 1, 0)
    up13 = net.conv("Up_1_3", net.bn("Up_1_3_BN", dec3), 8 * k, 0.2, 2, 0)
    dec1 = net.reduce("DeConv_1_Reduce", net.cat(up12, up13, conv1), 13 * k, 4 * k)
    dec1 = net.conv("DeConv_1", dec1, 4 * k, 0.03, 0)

    # per-point head: [one-hot category | features] -> parts
    final = net.bn("BNRELUDROP_hier_final", dec1)
    onehot = torch.nn.functional.one_hot(catLabels.reshape(-1).long(), numCats).to(torch.float32)
    return MLP_2_hidden(concat([onehot, final], 1), 4 * k + numCats, 4 * k, 2 * k, numParts, "Final_Logits",
                        keepProbFull, isTraining, useDropOutFull, useInitBN=False, store=store)
