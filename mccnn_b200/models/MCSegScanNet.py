"""MCSegScanNet -- semantic segmentation of rooms (models/MCSegScanNet.py:26-255): ABSOLUTE radii (metres,
relativeRadius=False: one global bounding box for the whole batch, tf_ops/aabb_gpu.cu:104-114), five hierarchy levels
[0.1, 0.2, 0.4, 0.8], 17 Monte-Carlo convolutions with skip links from levels 2/3/4 straight to level 1."""
from ..MCConvBuilder import PointHierarchy
from ..MCNetworkUtils import MLP_2_hidden
from ._net import Net


def create_network(points, batchIds, features, numInputFeatures, numSem, batchSize, k, isTraining, keepProbConv,
                   keepProbFull, useConvDropOut=False, useDropOutFull=True, store=None):
    ph = PointHierarchy(points, features, batchIds, [0.1, 0.2, 0.4, 0.8], "MCSegScanNet_PH", batchSize, False)
    net = Net(ph, 0.25, isTraining, useConvDropOut, keepProbConv, store, relativeRadius=False)

    # encoder
    pool0 = net.conv("Pool_0", features, numInputFeatures, 0.1, 0, 1, out_features=k, window=0.2)
    conv1 = net.cat(pool0, net.conv("Conv_1", net.bn("Conv_1_In_BN", pool0), k, 0.4, 1))
    pool1 = net.conv("Pool_1", net.reduce("Reduce_Pool_1", conv1, 2 * k, 2 * k), 2 * k, 0.4, 1, 2, window=0.2)
    conv2 = net.cat(pool1, net.conv("Conv_2", net.bn("Conv_2_In_BN", pool1), 2 * k, 0.8, 2))
    pool2 = net.conv("Pool_2", net.reduce("Reduce_Pool_2", conv2, 4 * k, 4 * k), 4 * k, 0.8, 2, 3, window=0.2)
    conv3 = net.cat(pool2, net.conv("Conv_3", net.bn("Conv_3_In_BN", pool2), 4 * k, 1.6, 3))
    pool3 = net.conv("Pool_3", net.reduce("Reduce_Pool_3", conv3, 8 * k, 8 * k), 8 * k, 1.6, 3, 4, window=0.2)
    conv4 = net.cat(pool3, net.conv("Conv_4", net.bn("Conv_4_In_BN", pool3), 8 * k, 5.0, 4))

    # decoder
    up34 = net.conv("Up_3_4", net.reduce("Up3_4_Reduce", conv4, 16 * k, 8 * k), 8 * k, 1.6, 4, 3)
    dec3 = net.reduce("DeConv_3_Reduce", net.cat(up34, conv3), 16 * k, 8 * k)
    dec3 = net.conv("DeConv_3", dec3, 8 * k, 1.6, 3)
    up23 = net.conv("Up_2_3", net.reduce("Up2_3_Reduce", dec3, 8 * k, 4 * k), 4 * k, 0.8, 3, 2)
    dec2 = net.reduce("DeConv_2_Reduce", net.cat(up23, conv2), 8 * k, 4 * k)
    dec2 = net.conv("DeConv_2", dec2, 4 * k, 0.8, 2)
    up12 = net.conv("Up_1_2", net.reduce("Up1_2_Reduce", dec2, 4 * k, 2 * k), 2 * k, 0.4, 2, 1)
    up13 = net.conv("Up_1_3", net.reduce("Up1_3_Reduce", dec3, 8 * k, 2 * k), 2 * k, 0.8, 3, 1)
    up14 = net.conv("Up_1_4", net.reduce("Up1_4_Reduce", conv4, 16 * k, 2 * k), 2 * k, 1.6, 4, 1)
    ups = net.cat(up14, up13, up12, conv1)
    dec1 = net.conv("DeConv_1", net.reduce("DeConv_1_Reduce", ups, 8 * k, 4 * k), 4 * k, 0.4, 1)
    final = net.conv("Up_0_1", net.reduce("Up_Final_Reduce", net.cat(ups, dec1), 12 * k, 4 * k), 4 * k, 0.1, 1, 0)

    return MLP_2_hidden(net.bn("BNRELUDROP_hier_final", final), 4 * k, 4 * k, 2 * k, numSem, "Final_Logits",
                        keepProbFull, isTraining, useDropOutFull, useInitBN=False, store=store)
