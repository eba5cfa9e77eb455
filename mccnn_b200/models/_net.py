"""Shorthand shared by the model files: one object that binds a PointHierarchy, a ConvolutionBuilder and the
inference/dropout switches, so that a network reads as a list of (name, level, radius) steps.

Every method is a thin call into the reference-shaped API (MCConvBuilder.create_convolution,
MCNetworkUtils.batch_norm_RELU_drop_out / conv_1x1) with the reference's variable names, so checkpoints keyed by those
names load unchanged.
"""
from ..MCConvBuilder import ConvolutionBuilder
from ..MCNetworkUtils import batch_norm_RELU_drop_out, conv_1x1, concat


class Net:
    def __init__(self, hierarchy, window, isTraining, useConvDropOut, keepProbConv, store, relativeRadius=True):
        self.ph = hierarchy
        self.cb = ConvolutionBuilder(KDEWindow=window, relativeRadius=relativeRadius, variableStore=store)
        self.isTraining, self.drop, self.keep, self.store = isTraining, useConvDropOut, keepProbConv, store

    def conv(self, name, x, c, radius, level, out_level=None, out_features=None, window=None):
        """Monte-Carlo convolution `name` on hierarchy level `level` (-> `out_level`): depth-wise on c features, or
        multi-feature c -> out_features when out_features is given."""
        return self.cb.create_convolution(
            convName=name, inPointHierarchy=self.ph, inPointLevel=level, outPointLevel=out_level, inFeatures=x,
            inNumFeatures=c, convRadius=radius, KDEWindow=window,
            multiFeatureConv=True if out_features is not None else None, outNumFeatures=out_features)

    def bn(self, name, x):
        return batch_norm_RELU_drop_out(name, x, self.isTraining, self.drop, self.keep, self.store)

    def reduce(self, name, x, c_in, c_out):
        """BN-ReLU -> 1x1 convolution -> BN-ReLU: `<name>_In_BN`, `<name>`, `<name>_Out_BN`."""
        x = self.bn(name + "_In_BN", x)
        x = conv_1x1(name, x, c_in, c_out, self.store)
        return self.bn(name + "_Out_BN", x)

    @staticmethod
    def cat(*xs):
        return concat(list(xs), 1)
