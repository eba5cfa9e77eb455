"""Batch assembly on the device: the reference's `utils/DataSet.py` (base class `DataSet`, :22-842) with the per-point
work -- sampling protocols, augmentation, packing -- done by tensor ops on the GPU instead of Python loops over points
(the caller side of the hot path).

Same constructor, attributes and methods as the reference class (`start_iteration`, `has_more_batches`,
`get_next_batch`, `set_allowed_samplings`, `get_num_models`, `reset_caches`; subclasses implement
`_load_model_from_disk_` exactly as there), and -- given the same seed -- the SAME batches: the host
`numpy.random.RandomState` is consumed draw for draw like the reference does (model permutation, protocol choice, view
direction, per-point acceptance draws with the reference's early exit, rotation angles), so a run can be compared with
the reference's output point for point (tests/golden/dataset_*.npz are produced by the reference class itself).

What moves to the device:
  * models are uploaded once (float64, as the reference keeps them) and cached there;
  * split / gradient / lambert acceptance probabilities (utils/DataSet.py:365-498,500-560) for all points at once; the
    reference's "walk the points, draw, stop at numPoints, wrap around" loop becomes: one block of uniforms per pass,
    `u < p`, an inclusive scan, and the position of the numPoints-th accepted point -- which also tells how many draws
    the reference would have consumed, so the host stream is advanced by exactly that many;
  * occlusion (:562-658): the z-buffer is a scatter-min over pixels;
  * rotation augmentation (:265-312): the 3x3 matrix is built on the host from the reference's draws, applied on device;
  * the packed batch (points, batch ids [N,1], features, labels, categories) is assembled by concatenation on the device
    and returned as float32 / int32 CUDA tensors ready for `mccnn_b200.models.*.create_network`.
"""
import abc
import time
from collections import deque

import numpy as np
import torch


class DataSet(abc.ABC):
    def __init__(self, numPoints, ptDropOut, pointFeatures, pointLabels, pointNormals, useCategories, pointCategories,
                 batchSize, allowedSamplings, maxNumPtsCache=100000000, maxPtsxBatch=0, augment=False,
                 augmentMainAxis=1, augmentSmallRotations=False, uniformSelectFirst=False, augmentedFeatures=[],
                 augmentedLabels=[], seed=None, device="cuda"):
        """Arguments as utils/DataSet.py:85-131, plus `device` (where the models and the batches live)."""
        for s in allowedSamplings:
            if not (0 <= s < 5):
                raise RuntimeError('Invalid sampling protocol')
        if (3 in list(allowedSamplings) or 4 in list(allowedSamplings)) and not pointNormals:
            raise RuntimeError('The dataset should contain normals in order to use the sampling protocols '
                               'lambert and occlusion')
        if augmentMainAxis < 0 or augmentMainAxis > 2:
            raise RuntimeError('Invalid main axis for the data augmentation')
        for lst in (augmentedFeatures, augmentedLabels):
            for a, b in zip(lst, lst[1:]):
                if b - a < 3:
                    raise RuntimeError('The groups of 3 features/labels to augment should not overlap ')
        self.numPoints_ = numPoints
        self.ptDropOut_ = ptDropOut
        self.pointFeatures_ = pointFeatures
        self.pointLabels_ = pointLabels
        self.pointNormals_ = pointNormals
        self.useCategories_ = useCategories
        self.pointCategories_ = pointCategories
        self.batchSize_ = batchSize
        self.allowedSamplings_ = allowedSamplings
        self.maxNumPtsCache_ = maxNumPtsCache
        self.maxPtsxBatch_ = maxPtsxBatch
        self.augment_ = augment
        self.augmentMainAxis_ = augmentMainAxis
        self.augmentSmallRotations_ = augmentSmallRotations
        self.augmentedFeatures_ = augmentedFeatures
        self.augmentedLabels_ = augmentedLabels
        self.uniformSelectFirst_ = uniformSelectFirst
        self.fileList_ = []
        self.numPts_ = []
        self.categories_ = []
        self.weights_ = []
        self.cacheCurrNumPts = 0
        self.cacheAddedOrder_ = deque([])
        self.cache_ = {}
        self.randomSelection_ = []
        self.iterator_ = 0
        self.randomState_ = np.random.RandomState(seed if seed is not None else int(time.time()))
        self.device_ = torch.device(device)

    # ------------------------------------------------------------------------------------------------ models
    @abc.abstractmethod
    def _load_model_from_disk_(self, modelPath):
        """-> (pts [n,3], normals [n,3] | None, features [n,m] | None, labels [n,l] | None) as numpy arrays
        (utils/DataSet.py:190-207)."""

    def _load_model_(self, modelPath):
        """Device-resident cache with the reference's LRU policy (utils/DataSet.py:209-263)."""
        if modelPath in self.cache_:
            self.cacheAddedOrder_.remove(modelPath)
            self.cacheAddedOrder_.append(modelPath)
            return self.cache_[modelPath]
        pts, normals, features, labels = self._load_model_from_disk_(modelPath)

        def up(a, dtype=None):
            if a is None:
                return None
            t = torch.as_tensor(np.ascontiguousarray(a))
            if t.dtype.is_floating_point:
                t = t.to(torch.float64)
            return t.to(self.device_)

        entry = (up(pts), up(normals), up(features), up(labels))
        if self.maxNumPtsCache_ > 0:
            while self.cacheAddedOrder_ and (self.cacheCurrNumPts + len(pts)) > self.maxNumPtsCache_:
                old = self.cacheAddedOrder_.popleft()
                self.cacheCurrNumPts -= int(self.cache_[old][0].shape[0])
                del self.cache_[old]
            self.cache_[modelPath] = entry
            self.cacheAddedOrder_.append(modelPath)
            self.cacheCurrNumPts += len(pts)
        return entry

    def get_num_models(self):
        return len(self.fileList_)

    def set_allowed_samplings(self, allowedSamplings):
        for s in allowedSamplings:
            if not (0 <= s < 5):
                raise RuntimeError('Invalid sampling protocol')
        if (3 in list(allowedSamplings) or 4 in list(allowedSamplings)) and not self.pointNormals_:
            raise RuntimeError('The dataset should contain normals in order to use the sampling protocols '
                               'lambert and occlusion')
        self.allowedSamplings_ = allowedSamplings

    def reset_caches(self):
        self.cacheCurrNumPts = 0
        self.cacheAddedOrder_ = deque([])
        self.cache_ = {}

    def start_iteration(self):
        self.randomSelection_ = self.randomState_.permutation(len(self.fileList_))
        self.iterator_ = 0

    def has_more_batches(self):
        return self.iterator_ < len(self.randomSelection_)

    # ------------------------------------------------------------------------------------------------ sampling
    def _take(self, idx, points, inFeatures, inLabels):
        return (points[idx], None if inFeatures is None else inFeatures[idx],
                None if inLabels is None else inLabels[idx])

    def _uniform_sampling_(self, points, inNumPoints, selectFirst, inFeatures=None, inLabels=None, numPoints=0):
        """utils/DataSet.py:315-361: the index choice is the reference's host draw, the gather runs on the device."""
        rs = self.randomState_
        if numPoints > 0:
            if selectFirst and inNumPoints >= numPoints:
                choice = rs.choice(int(float(numPoints) * (2.0 - self.ptDropOut_)), numPoints, replace=False)
            else:
                choice = rs.choice(inNumPoints, numPoints, replace=(inNumPoints < numPoints))
        else:
            choice = rs.choice(inNumPoints, int(float(inNumPoints) * self.ptDropOut_), replace=False)
        idx = torch.as_tensor(np.asarray(choice, dtype=np.int64), device=points.device)
        return self._take(idx, points, inFeatures, inLabels)

    def _select_by_probability_(self, prob, points, inFeatures, inLabels, numPoints):
        """The acceptance loop shared by the split / gradient / lambert protocols (e.g. utils/DataSet.py:392-420):
        walk the points in order drawing one uniform each, keep a point when the draw is below its probability, stop
        at numPoints (wrapping around as often as needed); numPoints == 0: exactly one pass."""
        rs = self.randomState_
        n = int(prob.shape[0])
        if numPoints > 0 and not bool((prob > 0).any()):
            raise RuntimeError("no point can be selected (the reference would loop forever)")
        state = rs.get_state()
        picked, have, consumed = [], 0, 0
        while True:
            u = torch.as_tensor(rs.random_sample(n), device=prob.device)
            keep = u < prob
            if numPoints == 0:
                picked.append(torch.nonzero(keep).reshape(-1))
                consumed += n
                break
            csum = torch.cumsum(keep.to(torch.int64), 0)
            total = int(csum[-1].item()) if n > 0 else 0
            need = numPoints - have
            if total >= need:  # the need-th accepted point of this pass ends the walk
                last = int(torch.searchsorted(csum, torch.tensor([need], device=csum.device)).item())
                picked.append(torch.nonzero(keep[: last + 1]).reshape(-1))
                consumed += last + 1
                break
            picked.append(torch.nonzero(keep).reshape(-1))
            have += total
            consumed += n
        rs.set_state(state)          # the reference drew exactly `consumed` numbers: leave the stream where it would be
        if consumed > 0:
            rs.random_sample(consumed)
        return self._take(torch.cat(picked), points, inFeatures, inLabels)

    @staticmethod
    def _largest_axis_position_(points):
        cmax, cmin = points.max(dim=0).values, points.min(dim=0).values
        size = cmax - cmin
        ax = int(torch.argmax(size).item())
        return points[:, ax], cmin[ax], size[ax]

    def _non_uniform_sampling_split_(self, points, inNumPoints, inFeatures=None, inLabels=None, numPoints=0,
                                     lowProbability=0.25):
        """utils/DataSet.py:365-428."""
        x, lo, size = self._largest_axis_position_(points)
        pos = (x - lo) / size
        prob = torch.where(pos > 0.5, torch.ones_like(pos), torch.full_like(pos, lowProbability))
        return self._select_by_probability_(prob, points, inFeatures, inLabels, numPoints)

    def _non_uniform_sampling_gradient_(self, points, inNumPoints, inFeatures=None, inLabels=None, numPoints=0):
        """utils/DataSet.py:431-496."""
        x, lo, size = self._largest_axis_position_(points)
        prob = torch.sqrt(torch.clamp((x - lo - size * 0.2) / (size * 0.6), 0.01, 1.0))
        return self._select_by_probability_(prob, points, inFeatures, inLabels, numPoints)

    def _non_uniform_sampling_lambert_(self, viewDir, points, normals, inNumPoints, inFeatures=None, inLabels=None,
                                       numPoints=0):
        """utils/DataSet.py:499-560."""
        v = torch.as_tensor(np.asarray(viewDir, dtype=np.float64), device=points.device)
        prob = torch.sqrt(torch.clamp(normals @ v, 0.0, 1.0))
        return self._select_by_probability_(prob, points, inFeatures, inLabels, numPoints)

    def _non_uniform_sampling_occlusion_(self, viewDir, points, normals, inNumPoints, inFeatures=None, inLabels=None,
                                         numPoints=0, screenResolution=128):
        """utils/DataSet.py:562-658: front-facing points are splatted into a screenResolution^2 z-buffer (scatter-min);
        a point is kept when it lies within 0.01 of the nearest depth of its pixel.  Back-facing points keep the
        reference's pixel id (-1, -1), i.e. they are tested against the LAST pixel with depth 1."""
        dev = points.device
        view = np.asarray(viewDir, dtype=np.float64)
        xVec = np.cross(view, np.array([0.0, 1.0, 0.0]))
        xVec = xVec / np.linalg.norm(xVec)
        yVec = np.cross(xVec, view)
        yVec = yVec / np.linalg.norm(yVec)
        cmax, cmin = points.max(dim=0).values, points.min(dim=0).values
        diagonal = float(torch.linalg.norm(cmax - cmin).item()) * 0.5
        center = ((cmax + cmin) * 0.5).cpu().numpy()
        res = int(screenResolution)
        pixelSize = diagonal / (float(res) * 0.5)
        screenPos = center - view * diagonal - xVec * diagonal - yVec * diagonal
        T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)  # noqa: E731
        front = (normals @ T(view)) < 0.0
        rel = points - T(screenPos)
        sx, sy = rel @ T(xVec), rel @ T(yVec)
        z = (rel @ T(view)) / (diagonal * 2.0)
        px = torch.floor(sx / pixelSize).to(torch.int64)
        py = torch.floor(sy / pixelSize).to(torch.int64)
        last = res * res - 1
        pix = torch.where(front, torch.remainder(px, res) * res + torch.remainder(py, res),
                          torch.full_like(px, last))
        zval = torch.where(front, z, torch.ones_like(z))
        zbuf = torch.full((res * res,), float("inf"), dtype=torch.float64, device=dev)
        zbuf = zbuf.scatter_reduce(0, pix[front], zval[front], reduce="amin", include_self=True)
        zbuf = torch.where(torch.isinf(zbuf), torch.full_like(zbuf, -1.0), zbuf)
        visible = (zval - zbuf[pix]) < 0.01
        idx = torch.nonzero(visible).reshape(-1)
        if numPoints > 0:
            if idx.numel() == 0:
                raise RuntimeError("no visible point (the reference would loop forever)")
            reps = -(-numPoints // int(idx.numel()))
            idx = idx.repeat(reps)[:numPoints]  # the reference wraps around until it has numPoints
        return self._take(idx, points, inFeatures, inLabels)

    # ------------------------------------------------------------------------------------------------ augmentation
    def _rotation_matrix_(self, mainRotAxis, smallRotations):
        """The reference's draws and matrix products (utils/DataSet.py:285-309), on the host in float64."""
        rs = self.randomState_
        angle = rs.uniform() * 2.0 * np.pi
        c, s = np.cos(angle), np.sin(angle)
        if mainRotAxis == 0:
            R = np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
        elif mainRotAxis == 1:
            R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
        else:
            R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        if smallRotations:
            a = np.clip(0.06 * rs.randn(3), -0.18, 0.18)
            Rx = np.array([[1.0, 0.0, 0.0], [0.0, np.cos(a[0]), -np.sin(a[0])], [0.0, np.sin(a[0]), np.cos(a[0])]])
            Ry = np.array([[np.cos(a[1]), 0.0, np.sin(a[1])], [0.0, 1.0, 0.0], [-np.sin(a[1]), 0.0, np.cos(a[1])]])
            Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0.0], [np.sin(a[2]), np.cos(a[2]), 0.0], [0.0, 0.0, 1.0]])
            R = np.dot(np.dot(Rz, np.dot(Ry, Rx)), R)
        return R

    def _augment_data_rot_(self, inData, mainRotAxis=1, smallRotations=False, inRotationMatrix=None):
        """utils/DataSet.py:265-312 -> (rotated [n,3] device tensor, 3x3 numpy matrix)."""
        R = self._rotation_matrix_(mainRotAxis, smallRotations) if inRotationMatrix is None else inRotationMatrix
        return inData[:, 0:3].reshape(-1, 3) @ torch.as_tensor(R, dtype=inData.dtype, device=inData.device), R

    # ------------------------------------------------------------------------------------------------ batches
    def get_next_batch(self, repeatModelInBatch=False):
        """utils/DataSet.py:711-842 -> (numModelInBatch, points f32 [N,3], batchIds i32 [N,1], features f32 [N,C],
        labels | None, categories i32 | None, paths), tensors on `device`."""
        rs = self.randomState_
        accPts, accIds, accFeat, accLab, accCat, accPaths = [], [], [], [], [], []
        numModelInBatch, numPtsInBatch = 0, 0
        for i in range(self.batchSize_):
            if self.iterator_ >= len(self.randomSelection_):
                continue
            modelIdx = self.randomSelection_[self.iterator_]
            model = self.fileList_[modelIdx]
            modelNumPts = self.numPts_[modelIdx]
            loaded = None
            if modelNumPts == 0:
                loaded = self._load_model_(model)
                modelNumPts = int(loaded[0].shape[0])
                self.numPts_[modelIdx] = modelNumPts
            if self.maxPtsxBatch_ != 0 and (numPtsInBatch + modelNumPts) > self.maxPtsxBatch_:
                continue
            pts, normals, feats, labels = loaded if loaded is not None else self._load_model_(model)
            n = int(pts.shape[0])
            protocol = rs.choice(self.allowedSamplings_)
            if protocol == 0:
                pts, feats, labels = self._uniform_sampling_(pts, n, self.uniformSelectFirst_, feats, labels,
                                                             self.numPoints_)
            elif protocol == 1:
                pts, feats, labels = self._non_uniform_sampling_split_(pts, n, feats, labels, self.numPoints_)
            elif protocol == 2:
                pts, feats, labels = self._non_uniform_sampling_gradient_(pts, n, feats, labels, self.numPoints_)
            else:
                view = (rs.rand(3) * 2.0) - 1.0
                view = view / np.linalg.norm(view)
                fn = self._non_uniform_sampling_lambert_ if protocol == 3 else self._non_uniform_sampling_occlusion_
                pts, feats, labels = fn(view, pts, normals, n, feats, labels, self.numPoints_)
            if self.augment_:
                pts, R = self._augment_data_rot_(pts, self.augmentMainAxis_, self.augmentSmallRotations_)
                if self.pointFeatures_:
                    feats = feats.clone()
                    for blk in self.augmentedFeatures_:
                        feats[:, blk:blk + 3], _ = self._augment_data_rot_(feats[:, blk:blk + 3], inRotationMatrix=R)
                if self.pointLabels_:
                    labels = labels.clone()
                    for blk in self.augmentedLabels_:
                        labels[:, blk:blk + 3], _ = self._augment_data_rot_(labels[:, blk:blk + 3], inRotationMatrix=R)
            m = int(pts.shape[0])
            accPts.append(pts)
            accIds.append(torch.full((m, 1), i, dtype=torch.int32, device=self.device_))
            accFeat.append(feats if self.pointFeatures_ else
                           torch.ones((m, 1), dtype=torch.float64, device=self.device_))
            if self.pointLabels_:
                accLab.append(labels)
            if self.useCategories_:
                cat = int(self.categories_[modelIdx])
                accCat.append(torch.full((m, 1) if self.pointCategories_ else (1,), cat, dtype=torch.int32,
                                         device=self.device_))
            accPaths.append(model)
            numPtsInBatch += modelNumPts
            numModelInBatch += 1
            if not repeatModelInBatch:
                self.iterator_ += 1
        if repeatModelInBatch:
            self.iterator_ += 1

        def cat(parts, dtype=None):
            if not parts:
                return torch.empty((0,), device=self.device_)
            t = torch.cat(parts, dim=0)
            return t.to(dtype) if dtype is not None else t

        labels = None
        if self.pointLabels_:
            labels = cat(accLab)
            if labels.dtype == torch.float64:
                labels = labels.to(torch.float32)
        return (numModelInBatch, cat(accPts, torch.float32), cat(accIds), cat(accFeat, torch.float32), labels,
                cat(accCat) if self.useCategories_ else None, accPaths)


class InMemoryDataSet(DataSet):
    """A DataSet over models that are already in memory (synthetic data, tests, benchmarks): `models` is a list of
    (points, normals | None, features | None, labels | None[, category]) numpy tuples; the i-th is "path" str(i)."""

    def __init__(self, models, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.models_ = list(models)
        self.fileList_ = [str(i) for i in range(len(self.models_))]
        self.numPts_ = [0 for _ in self.models_]          # as the reference's subclasses: filled on first load
        self.categories_ = [m[4] if len(m) > 4 else 0 for m in self.models_]

    def _load_model_from_disk_(self, modelPath):
        return tuple(self.models_[int(modelPath)][:4])
