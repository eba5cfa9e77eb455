"""PointHierarchy and ConvolutionBuilder -- the interface the reference's models are written against
(utils/MCConvBuilder.py:24-131 and :133-427), re-hosted on torch CUDA tensors and the sm_100a op library.

Same constructor / create_convolution / reset signatures, attribute names (points_, features_, batchIds_,
sampledIndexs_, radiusList_, batchSize_, relativeRadius_, hierarchyName_, aabbMin_, aabbMax_, cacheGrids_, cacheNeighs_,
cachePDFs_), cache-key scheme (utils/MCConvBuilder.py:230-238), variable names and shapes
(<conv>_weights [3, 8nb], _biases [8nb], _weights2 / _weights3 [nb, 8, 8] -> reshaped WITHOUT transpose to [8, 8nb],
_biases2 / _biases3 [nb, 8] -> [8nb]; utils/MCConvBuilder.py:397-419) and error behaviour (RuntimeError on mismatched
batch sizes / feature counts, :326-333).  Variables live in a mccnn_b200.variables.VariableStore instead of the
TensorFlow graph.

Both classes work unchanged inside mccnn_b200.runtime.record()/planned(): the op wrappers they call carry the live
row counts on the device, so a planned run issues no host synchronisation at all.
"""
import os
import torch

from .MCConvModule import (compute_aabb, sort_points_step1, sort_points_step2, sort_features, sort_features_back,
                           compute_pdf, poisson_sampling, get_sampled_features, spatial_conv, get_block_size,
                           transform_indexs, find_neighbors, sort_points, find_neighbors_pdf)
from . import runtime as _rt
from . import variables as _vars

VERBOSE = False


def _log(msg):
    if VERBOSE:
        print(msg)


# MCCNN_FUSE_PDF=1: a planned run issues find_neighbors + compute_pdf of a new neighbour list as ONE launch
# (MCConvModule.find_neighbors_pdf).  Measured and left off: timed alone the fused call is 1.2-1.5x faster for
# <= 8k samples, but inside the forward's multi-stream graph the separate density kernel overlaps the other branch
# and the fused kernel lengthens the critical path (MCClassS 32 x 1024: 0.237 -> 0.252 ms per forward, 376 -> 361 M
# points/s; profiles/r2_neighbors_pdf_fused.txt).
FUSE_NEIGHBORS_PDF = os.environ.get("MCCNN_FUSE_PDF", "0") == "1"


class _LazyList:
    """List whose elements may be zero-argument thunks evaluated (once) on first access."""

    def __init__(self, items=()):
        self._items = [(True, v) for v in items]

    def append(self, value):
        self._items.append((True, value))

    def append_lazy(self, thunk):
        self._items.append((False, thunk))

    def __len__(self):
        return len(self._items)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self._items)))]
        ready, v = self._items[i]
        if not ready:
            v = v()
            self._items[i] = (True, v)
        return v

    def __iter__(self):
        return (self[k] for k in range(len(self._items)))


class PointHierarchy:
    """Poisson-disk point hierarchy of a packed batch (utils/MCConvBuilder.py:24-131).

    Level 0 is the input cloud; level l+1 is the Poisson-disk sampling of level l with radiusList[l].  The bounding
    boxes are computed once on level 0 and reused for every level (utils/MCConvBuilder.py:85-88)."""

    def __init__(self, inPoints, inFeatures, inBatchIds, radiusList, hierarchyName="Point_Hierarchy", batchSize=32,
                 relativeRadius=True):
        self.points_ = [inPoints]
        self.features_ = _LazyList([inFeatures])
        self.batchIds_ = [inBatchIds]
        self.sampledIndexs_ = _LazyList()
        self.radiusList_ = [0.0]
        self.batchSize_ = batchSize
        self.relativeRadius_ = relativeRadius
        self.hierarchyName_ = hierarchyName
        # grids the hierarchy sorted for its own sampling, keyed (level, cell size, relative): a convolution on the
        # same level with the same radius reuses them instead of sorting again (not in the reference, which re-sorts;
        # the result is the same grid)
        self.levelGrids_ = {}

        with _rt.lane("hier"):  # the hierarchy chain runs on its own stream in planned mode
            self._build(inPoints, inBatchIds, radiusList, batchSize)

    def _build(self, inPoints, inBatchIds, radiusList, batchSize):
        hierarchyName, relativeRadius = self.hierarchyName_, self.relativeRadius_
        aabbMin, aabbMax = compute_aabb(inPoints, inBatchIds, batchSize, self.relativeRadius_)
        if not self.relativeRadius_ and _rt.shard() is not None:
            # absolute radii: the reference's box spans the WHOLE batch (tf_ops/aabb_gpu.cu:104-114); when the clouds
            # are sharded over GPUs the local boxes are merged so that every shard uses the unsharded grid
            from .engine import global_box
            aabbMin, aabbMax = _rt.dispatch([aabbMin, aabbMax],
                                            lambda: global_box(aabbMin, aabbMax, _rt.shard()[0]))
        self.aabbMin_ = aabbMin
        self.aabbMax_ = aabbMax

        _log("########## Point Hierarchy: " + hierarchyName + " (Rel: " + str(relativeRadius) + ")")
        currPts, currBatchIds = inPoints, inBatchIds
        for level, currRadius in enumerate(radiusList):
            _log("Level: " + str(level + 1) + " | Poisson Disk Radius: " + str(currRadius))
            # sort_points_step1 + sort_points_step2 (utils/MCConvBuilder.py:106-110), fused into one launch.  Only the
            # geometry is sorted here: the per-level features and the transformed sample indices
            # (utils/MCConvBuilder.py:115-116) are derived on first access -- no MC* model reads sampledIndexs_ and
            # only MCClassH reads features_[1], so the common forward pays for neither.
            noFeatures = currPts.new_empty((currPts.shape[0], 0))
            keys, indexs, sortPts, sortBatchs, _, cellIndexs = sort_points(
                currPts, currBatchIds, noFeatures, self.aabbMin_, self.aabbMax_, self.batchSize_, currRadius,
                self.relativeRadius_)
            sampledPts, sampledBatchsIds, sampledIndexs = poisson_sampling(
                sortPts, sortBatchs, cellIndexs, aabbMin, aabbMax, currRadius, batchSize, self.relativeRadius_)

            self.levelGrids_[(level, currRadius, bool(self.relativeRadius_))] = (sortPts, sortBatchs, cellIndexs,
                                                                                indexs)
            self.points_.append(sampledPts)
            self.batchIds_.append(sampledBatchsIds)
            self.features_.append_lazy(
                lambda lvl=level, si=sampledIndexs, ix=indexs: get_sampled_features(
                    si, sort_features(self.features_[lvl], ix)))
            self.sampledIndexs_.append_lazy(lambda si=sampledIndexs, ix=indexs: transform_indexs(si, ix))
            self.radiusList_.append(currRadius)
            currPts, currBatchIds = sampledPts, sampledBatchsIds


class ConvolutionBuilder:
    """Creates Monte-Carlo convolutions on point hierarchies, caching grids / neighbour lists / pdfs
    (utils/MCConvBuilder.py:133-427)."""

    def __init__(self, multiFeatureConvs=False, KDEWindow=0.25, relativeRadius=True, usePDF=True, useAVG=True,
                 decayLossCollection='weight_decay_loss', variableStore=None):
        self.cacheGrids_ = {}
        self.cacheNeighs_ = {}
        self.cachePDFs_ = {}
        self.multiFeatureConvs_ = multiFeatureConvs
        self.KDEWindow_ = KDEWindow
        self.relativeRadius_ = relativeRadius
        self.usePDF_ = usePDF
        self.useAVG_ = useAVG
        self.decayLossCollection_ = decayLossCollection
        self.variables_ = variableStore if variableStore is not None else _vars.default_store()

    def __compute_dic_keys__(self, inPointHierarchy, outPointHierarchy, inPointLevel, outPointLevel, convRadius,
                             KDEWindow, relativeRadius, usePDF):
        keyGrid = inPointHierarchy.hierarchyName_ + '|' + str(inPointLevel) + '|' + str(convRadius) + '|' + \
            str(relativeRadius)
        keyNeighs = keyGrid + '|' + outPointHierarchy.hierarchyName_ + '|' + str(outPointLevel)
        keyPDF = keyNeighs + '|' + str(KDEWindow) + '|' + str(usePDF)
        return keyGrid, keyNeighs, keyPDF

    def reset(self):
        self.cacheGrids_ = {}
        self.cacheNeighs_ = {}
        self.cachePDFs_ = {}

    def create_convolution(self, convName, inPointHierarchy, inPointLevel, inFeatures, inNumFeatures, convRadius,
                           outPointHierarchy=None, outPointLevel=None, multiFeatureConv=None, outNumFeatures=None,
                           KDEWindow=None, relativeRadius=None, usePDF=None, useAVG=None):
        currMultiFeatureConv = self.multiFeatureConvs_ if multiFeatureConv is None else multiFeatureConv
        currNumOutFeatures = inNumFeatures if outNumFeatures is None else outNumFeatures
        currKDEWindow = self.KDEWindow_ if KDEWindow is None else KDEWindow
        currRelativeRadius = self.relativeRadius_ if relativeRadius is None else relativeRadius
        currUsePDF = self.usePDF_ if usePDF is None else usePDF
        currUseAVG = self.useAVG_ if useAVG is None else useAVG
        currOutPointHierarchy = inPointHierarchy if outPointHierarchy is None else outPointHierarchy
        currOutPointLevel = inPointLevel if outPointLevel is None else outPointLevel

        if currOutPointHierarchy.batchSize_ != inPointHierarchy.batchSize_:
            raise RuntimeError('Different batch size in the input and output point hierarchy')
        if (currMultiFeatureConv is False) and (currNumOutFeatures != inNumFeatures):
            raise RuntimeError('The number of input and output features should be the same '
                               'for multi feature convolutions.')

        keyGrid, keyNeighs, keyPDF = self.__compute_dic_keys__(
            inPointHierarchy, currOutPointHierarchy, inPointLevel, currOutPointLevel, convRadius, currKDEWindow,
            currRelativeRadius, currUsePDF)
        _log("Convolution: " + str(convName) + " (KDE: " + str(currKDEWindow) + " | MF: " + str(currMultiFeatureConv) +
             " | Rel: " + str(currRelativeRadius) + " | PDF: " + str(currUsePDF) + ")")

        ph = inPointHierarchy
        # grid: cache hit -> only re-sort the new features into the cached order (utils/MCConvBuilder.py:349-351)
        with _rt.lane("geom"):  # grid, neighbour list and pdf never touch the features: separate stream
            currGridTuple, currNeighTuple, currPDFs = self._geometry(
                ph, inPointLevel, currOutPointHierarchy, currOutPointLevel, convRadius, currRelativeRadius,
                currKDEWindow, currUsePDF, keyGrid, keyNeighs, keyPDF)
        sortFeatures = sort_features(inFeatures, currGridTuple[3])
        return self._convolve(convName, ph, currGridTuple, currNeighTuple, currPDFs, sortFeatures, inNumFeatures,
                              currNumOutFeatures, currMultiFeatureConv, currOutPointHierarchy, currOutPointLevel,
                              convRadius, currRelativeRadius, currUseAVG)

    def _geometry(self, ph, inPointLevel, currOutPointHierarchy, currOutPointLevel, convRadius, currRelativeRadius,
                  currKDEWindow, currUsePDF, keyGrid, keyNeighs, keyPDF):
        shared = getattr(ph, "levelGrids_", {}).get((inPointLevel, convRadius, bool(currRelativeRadius)))
        if keyGrid not in self.cacheGrids_ and shared is not None:
            self.cacheGrids_[keyGrid] = shared
        if keyGrid not in self.cacheGrids_:
            # sort_points_step1 + sort_points_step2 (utils/MCConvBuilder.py:353-361), fused into one launch.  The grid
            # is pure geometry (it does not depend on the features), so it is built without them and the features
            # are always brought into grid order by sort_features -- what the reference does on a cache hit.  This
            # keeps every geometric op independent of the feature chain (they overlap on separate streams).
            pts = ph.points_[inPointLevel]
            keys, indexs, sortPts, sortBatchs, _, cellIndexs = sort_points(
                pts, ph.batchIds_[inPointLevel], pts.new_empty((pts.shape[0], 0)), ph.aabbMin_, ph.aabbMax_,
                ph.batchSize_, convRadius, currRelativeRadius)
            self.cacheGrids_[keyGrid] = (sortPts, sortBatchs, cellIndexs, indexs)
        currGridTuple = self.cacheGrids_[keyGrid]

        if keyNeighs in self.cacheNeighs_:
            currNeighTuple = self.cacheNeighs_[keyNeighs]
        elif FUSE_NEIGHBORS_PDF and currUsePDF and keyPDF not in self.cachePDFs_:
            # neither the neighbour list nor its densities exist yet (utils/MCConvBuilder.py:365-386 issues
            # find_neighbors, then compute_pdf): one fused launch in planned mode.  Off by default -- see above.
            startIndexs, packedNeighs, pdfs = find_neighbors_pdf(
                currOutPointHierarchy.points_[currOutPointLevel], currOutPointHierarchy.batchIds_[currOutPointLevel],
                currGridTuple[0], currGridTuple[1], currGridTuple[2], ph.aabbMin_, ph.aabbMax_, convRadius,
                currKDEWindow, ph.batchSize_, currRelativeRadius)
            currNeighTuple = (startIndexs, packedNeighs)
            self.cacheNeighs_[keyNeighs] = currNeighTuple
            self.cachePDFs_[keyPDF] = pdfs
        else:
            startIndexs, packedNeighs = find_neighbors(
                currOutPointHierarchy.points_[currOutPointLevel], currOutPointHierarchy.batchIds_[currOutPointLevel],
                currGridTuple[0], currGridTuple[2], ph.aabbMin_, ph.aabbMax_, convRadius, ph.batchSize_,
                currRelativeRadius)
            currNeighTuple = (startIndexs, packedNeighs)
            self.cacheNeighs_[keyNeighs] = currNeighTuple

        if keyPDF in self.cachePDFs_:
            currPDFs = self.cachePDFs_[keyPDF]
        else:
            if currUsePDF:
                currPDFs = compute_pdf(currGridTuple[0], currGridTuple[1], ph.aabbMin_, ph.aabbMax_,
                                       currNeighTuple[0], currNeighTuple[1], currKDEWindow, convRadius,
                                       ph.batchSize_, currRelativeRadius)
            else:  # tf.ones([K, 1]) (utils/MCConvBuilder.py:388-390)
                neigh = currNeighTuple[1]
                currPDFs = _rt.tag(torch.ones((neigh.shape[0], 1), dtype=torch.float32, device=neigh.device),
                                   _rt.count_of(neigh))
            self.cachePDFs_[keyPDF] = currPDFs
        return currGridTuple, currNeighTuple, currPDFs

    def _convolve(self, convName, ph, currGridTuple, currNeighTuple, currPDFs, sortFeatures, inNumFeatures,
                  currNumOutFeatures, currMultiFeatureConv, currOutPointHierarchy, currOutPointLevel, convRadius,
                  currRelativeRadius, currUseAVG):
        # parameters (utils/MCConvBuilder.py:394-419)
        vs = self.variables_
        blockSize = get_block_size()
        numOutNeurons = inNumFeatures * currNumOutFeatures if currMultiFeatureConv else inNumFeatures
        numBlocks = int(numOutNeurons / blockSize)
        if numOutNeurons % blockSize != 0:
            numBlocks = numBlocks + 1
        weights = vs.get_variable(convName + '_weights', [3, blockSize * numBlocks], "glorot")
        vs.add_to_collection(self.decayLossCollection_, weights)
        biases = vs.get_variable(convName + '_biases', [blockSize * numBlocks], "zeros")
        weights2 = vs.get_variable(convName + '_weights2', [numBlocks, blockSize, blockSize], "glorot")
        vs.add_to_collection(self.decayLossCollection_, weights2)  # the variable itself: a view made before
        weights2 = weights2.reshape(blockSize, numBlocks * blockSize)  # requires_grad_() would cut its gradient
        biases2 = vs.get_variable(convName + '_biases2', [numBlocks, blockSize], "zeros").reshape(numBlocks * blockSize)
        weights3 = vs.get_variable(convName + '_weights3', [numBlocks, blockSize, blockSize], "glorot")
        vs.add_to_collection(self.decayLossCollection_, weights3)
        weights3 = weights3.reshape(blockSize, numBlocks * blockSize)
        biases3 = vs.get_variable(convName + '_biases3', [numBlocks, blockSize], "zeros").reshape(numBlocks * blockSize)

        return spatial_conv(currGridTuple[0], sortFeatures, currGridTuple[1], currPDFs,
                            currOutPointHierarchy.points_[currOutPointLevel], currNeighTuple[0], currNeighTuple[1],
                            ph.aabbMin_, ph.aabbMax_, weights, weights2, weights3, biases, biases2, biases3,
                            currNumOutFeatures, currMultiFeatureConv, ph.batchSize_, convRadius, currRelativeRadius,
                            currUseAVG)
