"""The reference's dataset classes on the device-side DataSet (mccnn_b200/DataSet.py): same constructors, file formats
and methods as ModelNet/ModelNetDataSet.py:22-138 and ShapeNet/ShapeNetDataSet.py:22-155; only the loading is
vectorised (numpy.loadtxt) and the models are cached on the GPU.  The file lists, category assignment and per-model
stored point counts follow the reference line by line, because they feed the batch composition (and hence the random
stream) of DataSet.get_next_batch."""
import json
import os

import numpy as np

from .DataSet import DataSet


class ModelNetDataSet(DataSet):
    """ModelNet40 with normals: `<folder>/<category>/<model>.txt`, one `x,y,z,nx,ny,nz` line per point, the first
    maxStoredNumPoints lines are used; `modelnet40_shape_names.txt` and `modelnet40_{train,test}.txt` list categories and
    models (ModelNet/ModelNetDataSet.py:34-100)."""

    def __init__(self, train, numPoints, ptDropOut, maxStoredNumPoints, batchSize, allowedSamplings=[0], augment=False,
                 useNormalsAsLabels=False, useNormalsAsFeatures=False, folder="data", seed=None, device="cuda"):
        self.useNormalsAsLabels_ = useNormalsAsLabels
        self.useNormalsAsFeatures_ = useNormalsAsFeatures
        self.maxStoredNumPoints_ = maxStoredNumPoints
        super().__init__(numPoints, ptDropOut, useNormalsAsFeatures, useNormalsAsLabels, True, not useNormalsAsLabels,
                         False, batchSize, allowedSamplings, 100000000, 0, augment, 1, True, True,
                         [0] if useNormalsAsFeatures else [], [0] if useNormalsAsLabels else [], seed, device=device)
        with open(os.path.join(folder, "modelnet40_shape_names.txt")) as f:
            self.catNames_ = [line.replace("\n", "") for line in f]
        with open(os.path.join(folder, "modelnet40_train.txt" if train else "modelnet40_test.txt")) as f:
            for line in f:
                catId = next((i for i, name in enumerate(self.catNames_) if name in line), -1)  # first match wins
                if catId >= 0:
                    self.fileList_.append(folder + "/" + self.catNames_[catId] + "/" + line.replace("\n", "") + ".txt")
                    self.categories_.append(catId)
                    self.numPts_.append(self.maxStoredNumPoints_)

    def get_categories(self):
        return self.catNames_

    def _load_model_from_disk_(self, modelPath):
        data = np.loadtxt(modelPath, delimiter=",", usecols=range(6), max_rows=self.maxStoredNumPoints_, ndmin=2)
        pts, normals = data[:, 0:3], data[:, 3:6]
        return (pts, normals, normals if self.useNormalsAsFeatures_ else None,
                normals if self.useNormalsAsLabels_ else None)


class ShapeNetDataSet(DataSet):
    """ShapeNet part segmentation: `<folder>/<synset>/<model>.txt`, whitespace-separated `x y z nx ny nz part` lines; all
    points of a model are used (numPoints = 0), the object category is provided per point
    (ShapeNet/ShapeNetDataSet.py:33-90,128-155)."""

    SEG_CLASSES = {'Earphone': [16, 17, 18], 'Motorbike': [30, 31, 32, 33, 34, 35], 'Rocket': [41, 42, 43],
                   'Car': [8, 9, 10, 11], 'Laptop': [28, 29], 'Cap': [6, 7], 'Skateboard': [44, 45, 46],
                   'Mug': [36, 37], 'Guitar': [19, 20, 21], 'Bag': [4, 5], 'Lamp': [24, 25, 26, 27],
                   'Table': [47, 48, 49], 'Airplane': [0, 1, 2, 3], 'Pistol': [38, 39, 40], 'Chair': [12, 13, 14, 15],
                   'Knife': [22, 23]}

    def __init__(self, train, batchSize, ptDropOut, allowedSamplings=[0], augment=False, useNormalsAsFeatures=False,
                 seed=None, folder="./shape_data", device="cuda"):
        self.useNormalsAsFeatures_ = useNormalsAsFeatures
        super().__init__(0, ptDropOut, useNormalsAsFeatures, True, True, True, True, batchSize, allowedSamplings,
                         100000000, 0, augment, 1, True, False, [0] if useNormalsAsFeatures else [], [], seed,
                         device=device)
        self.folder_ = folder
        with open(os.path.join(folder, "synsetoffset2category.txt")) as f:
            self.catNames_ = [tuple(line.replace("\n", "").split("\t")[:2]) for line in f]
        self.segClasses_ = dict(self.SEG_CLASSES)
        split = os.path.join(folder, "train_test_split")
        names = ["shuffled_train_file_list.json", "shuffled_val_file_list.json"] if train else \
            ["shuffled_test_file_list.json"]
        for name in names:
            with open(os.path.join(split, name)) as f:
                self.fileList_ += list(json.load(f))
        for model in self.fileList_:
            catId = 0
            for i, (_, synset) in enumerate(self.catNames_):   # the LAST matching category wins, as in the reference
                if synset in model:
                    catId = i
            self.categories_.append(catId)
        self.numPts_ = [0 for _ in self.fileList_]

    def get_categories(self):
        return self.catNames_

    def get_categories_seg_parts(self):
        return self.segClasses_

    def _load_model_from_disk_(self, modelPath):
        data = np.loadtxt(modelPath + ".txt", usecols=range(7), ndmin=2)
        pts, normals = data[:, 0:3], data[:, 3:6]
        return pts, normals, normals if self.useNormalsAsFeatures_ else None, data[:, 6:7]


class ScanNetDataSet(DataSet):
    """ScanNet rooms prepared by the reference's ScanNet/genScanNetData.py: `<dataFolder>/<room>_{pos,labels,colors}.npy`
    plus rooms.txt / num_points.txt / scannet_{train,val,test}.txt / labels.txt / weights.txt
    (ScanNet/ScanNetDataSet.py:33-135).  Rooms are centred on load; all points are used, a batch holds at most
    maxNumPtsxBatch points, nothing is cached (maxNumPtsCache = 0, as in the reference)."""

    def __init__(self, dataset, batchSize, ptDropOut, maxNumPtsxBatch=600000, augment=False, useColorsAsFeatures=False,
                 dataFolder="data_mccnn", seed=None, device="cuda"):
        if 0 < maxNumPtsxBatch < 600000:
            raise RuntimeError('The number of points per batch should be big enough to store'
                               ' all the models (greater than 600000).')
        self.useColorsAsFeatures_ = useColorsAsFeatures
        self.dataFolder_ = dataFolder
        super().__init__(0, ptDropOut, useColorsAsFeatures, True, False, False, False, batchSize, [0], 0,
                         maxNumPtsxBatch, augment, 2, False, False, [], [], seed, device=device)
        rooms = np.loadtxt(dataFolder + "/rooms.txt", dtype='str')
        numPoints = np.loadtxt(dataFolder + "/num_points.txt") * ptDropOut
        split = ["scannet_train.txt", "scannet_val.txt", "scannet_test.txt"][dataset if dataset in (0, 1) else 2]
        wanted = set(np.atleast_1d(np.loadtxt(dataFolder + "/" + split, dtype='str')))
        idx = np.array([i for i in range(len(rooms)) if rooms[i] in wanted], dtype=np.int64)
        self.fileList_ = rooms[idx]
        self.numPts_ = numPoints[idx]
        self.semLabels_ = np.loadtxt(dataFolder + "/labels.txt", dtype='str', delimiter=':')
        weights = np.loadtxt(dataFolder + "/weights.txt")
        self.weights_ = 1.0 / np.log(1.2 + weights[0][: len(self.semLabels_)])
        self.weights_[0] = 0.0

    def get_labels(self):
        return self.semLabels_

    def get_weights(self, labels):
        """Per-point loss weights of the labels [n,1] (ScanNet/ScanNetDataSet.py:104-118); tensors stay on their device."""
        import torch
        if len(self.weights_) == 0:
            raise RuntimeError('No weights associated to the labels.')
        if isinstance(labels, torch.Tensor):
            w = torch.as_tensor(self.weights_, dtype=torch.float32, device=labels.device)
            return w[labels.reshape(-1).long()].reshape(-1, 1)
        return np.asarray(self.weights_)[np.asarray(labels).reshape(-1)].reshape(-1, 1)

    def get_accuracy_masks(self, labels):
        """1 for annotated points (label != 0), 0 otherwise (ScanNet/ScanNetDataSet.py:120-133)."""
        import torch
        if isinstance(labels, torch.Tensor):
            return (labels.reshape(-1, 1) != 0).to(torch.float32)
        return (np.asarray(labels).reshape(-1, 1) != 0).astype(np.float64)

    def _load_model_from_disk_(self, modelPath):
        pts = np.load(self.dataFolder_ + "/" + modelPath + "_pos.npy")
        labels = np.load(self.dataFolder_ + "/" + modelPath + "_labels.npy")
        features = np.load(self.dataFolder_ + "/" + modelPath + "_colors.npy") if self.useColorsAsFeatures_ else None
        return pts - np.mean(pts, axis=0), None, features, labels.reshape((-1, 1))
