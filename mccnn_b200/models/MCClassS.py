"""MCClassS -- the reference's small classification network (models/MCClassS.py:25-77), the configuration the
headline metric is quoted on: hierarchy radii [0.1, 0.4, sqrt(3)+0.1], Conv_1 L0->L1 (multi-feature, r 0.2),
Conv_2 L1->L2 (depth-wise, r 0.8), Conv_3 L2->L3 (depth-wise, r sqrt(3)+0.1), BN/ReLU/1x1 glue, MLP_2_hidden head.

Same create_network signature as the reference; tensors are torch CUDA tensors, variables live in a VariableStore.
"""
import math

from ..MCConvBuilder import PointHierarchy, ConvolutionBuilder
from ..MCNetworkUtils import MLP_2_hidden, batch_norm_RELU_drop_out, conv_1x1


def create_network(points, batchIds, features, numInputFeatures, batchSize, k, numOutCat, isTraining, keepProbConv,
                   keepProbFull, useConvDropOut=False, useDropOutFull=True, store=None):
    mPointHierarchy = PointHierarchy(points, features, batchIds, [0.1, 0.4, math.sqrt(3.0) + 0.1], "MCClassS_PH",
                                     batchSize)
    mConvBuilder = ConvolutionBuilder(KDEWindow=0.2, variableStore=store)

    convFeatures1 = mConvBuilder.create_convolution(
        convName="Conv_1", inPointHierarchy=mPointHierarchy, inPointLevel=0, outPointLevel=1, inFeatures=features,
        inNumFeatures=numInputFeatures, outNumFeatures=k, convRadius=0.2, multiFeatureConv=True)

    convFeatures1 = batch_norm_RELU_drop_out("Reduce_1_In_BN", convFeatures1, isTraining, useConvDropOut,
                                             keepProbConv, store)
    convFeatures1 = conv_1x1("Reduce_1", convFeatures1, k, k * 2, store)
    convFeatures1 = batch_norm_RELU_drop_out("Reduce_1_Out_BN", convFeatures1, isTraining, useConvDropOut,
                                             keepProbConv, store)
    convFeatures2 = mConvBuilder.create_convolution(
        convName="Conv_2", inPointHierarchy=mPointHierarchy, inPointLevel=1, outPointLevel=2,
        inFeatures=convFeatures1, inNumFeatures=k * 2, convRadius=0.8)

    convFeatures2 = batch_norm_RELU_drop_out("Reduce_2_In_BN", convFeatures2, isTraining, useConvDropOut,
                                             keepProbConv, store)
    convFeatures2 = conv_1x1("Reduce_2", convFeatures2, k * 2, k * 4, store)
    convFeatures2 = batch_norm_RELU_drop_out("Reduce_2_Out_BN", convFeatures2, isTraining, useConvDropOut,
                                             keepProbConv, store)
    convFeatures3 = mConvBuilder.create_convolution(
        convName="Conv_3", inPointHierarchy=mPointHierarchy, inPointLevel=2, outPointLevel=3,
        inFeatures=convFeatures2, inNumFeatures=k * 4, convRadius=math.sqrt(3.0) + 0.1)

    finalInput = batch_norm_RELU_drop_out("BNRELUDROP_final", convFeatures3, isTraining, useConvDropOut, keepProbConv,
                                          store)
    finalLogits = MLP_2_hidden(finalInput, k * 4, k * 2, k, numOutCat, "Final_Logits", keepProbFull, isTraining,
                               useDropOutFull, store=store)
    return finalLogits
