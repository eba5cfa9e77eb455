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
