"""mccnn_b200.datasets (ModelNetDataSet, ShapeNetDataSet) against the reference's own classes reading the SAME files:
a small synthetic ModelNet40-style and ShapeNet-style directory is written to a temp folder, both implementations
iterate over it with the same seed, and every batch must be identical (points, ids, features, labels, categories,
paths).  Skipped where the reference tree is absent."""
import json
import os
import sys
import warnings

import numpy as np
import pytest

REFERENCE = "/root/reference"


def _write_modelnet(root, rng):
    cats = ["chair", "night_stand", "table"]
    with open(os.path.join(root, "modelnet40_shape_names.txt"), "w") as f:
        f.write("\n".join(cats) + "\n")
    names = {"train": [], "test": []}
    for c in cats:
        os.makedirs(os.path.join(root, c))
        for m in range(4):
            name = "%s_%04d" % (c, m + 1)
            n = int(rng.integers(150, 300))
            v = rng.normal(size=(n, 3))
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            data = np.concatenate([v * rng.uniform(0.3, 1.0, (1, 3)), v], axis=1)
            np.savetxt(os.path.join(root, c, name + ".txt"), data, fmt="%.6f", delimiter=",")
            names["train" if m < 3 else "test"].append(name)
    for split, lst in names.items():
        with open(os.path.join(root, "modelnet40_%s.txt" % split), "w") as f:
            f.write("\n".join(lst) + "\n")


def _write_shapenet(root, rng):
    cats = [("Airplane", "02691156"), ("Bag", "02773838")]
    with open(os.path.join(root, "synsetoffset2category.txt"), "w") as f:
        f.write("".join("%s\t%s\n" % c for c in cats))
    os.makedirs(os.path.join(root, "train_test_split"))
    lists = {"train": [], "val": [], "test": []}
    for ci, (_, syn) in enumerate(cats):
        os.makedirs(os.path.join(root, syn))
        for m in range(3):
            n = int(rng.integers(120, 260))
            v = rng.normal(size=(n, 3))
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            part = rng.integers(0, 3, (n, 1)) + 4 * ci
            np.savetxt(os.path.join(root, syn, "m%d.txt" % m), np.concatenate([v * 0.5, v, part], axis=1), fmt="%.6f")
            lists[["train", "val", "test"][m]].append(os.path.join(root, syn, "m%d" % m))
    for split, lst in lists.items():
        with open(os.path.join(root, "train_test_split", "shuffled_%s_file_list.json" % split), "w") as f:
            json.dump(lst, f)


def _batches(ds):
    out = []
    for _ in range(2):
        ds.start_iteration()
        while ds.has_more_batches():
            out.append(ds.get_next_batch())
    return out


def _same(ours, theirs):
    assert len(ours) == len(theirs) > 0
    for (n, p, i, f, l, c, paths), (rn, rp, ri, rf, rl, rc, rpaths) in zip(ours, theirs):
        assert n == rn and list(paths) == list(rpaths)
        np.testing.assert_allclose(p.numpy(), np.asarray(rp, np.float32), atol=2e-7, rtol=0)
        assert np.array_equal(i.numpy(), np.asarray(ri, np.int32))
        np.testing.assert_allclose(f.numpy(), np.asarray(rf, np.float32), atol=2e-7, rtol=0)
        if rl is None:
            assert l is None
        else:
            np.testing.assert_allclose(l.numpy(), np.asarray(rl, np.float32), atol=2e-7, rtol=0)
        if rc is None:
            assert c is None
        else:
            assert np.array_equal(c.numpy().reshape(np.asarray(rc).shape), np.asarray(rc, np.int32))


def _reference(module_dir, name):
    if not os.path.isdir(os.path.join(REFERENCE, module_dir)):
        pytest.skip("reference tree not present on this machine")
    for p in (os.path.join(REFERENCE, "utils"), os.path.join(REFERENCE, module_dir)):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib
        return getattr(importlib.import_module(name), name)


@pytest.mark.parametrize("kw", [dict(numPoints=128, ptDropOut=1.0, allowedSamplings=[0, 1, 2, 3, 4], augment=True),
                                dict(numPoints=0, ptDropOut=0.9, allowedSamplings=[0], augment=False,
                                     useNormalsAsLabels=True, useNormalsAsFeatures=True)])
def test_modelnet_dataset_equals_the_reference(tmp_path, kw):
    Ref = _reference("ModelNet", "ModelNetDataSet")
    from mccnn_b200.datasets import ModelNetDataSet
    root = str(tmp_path)
    _write_modelnet(root, np.random.default_rng(0))
    common = dict(train=True, maxStoredNumPoints=200, batchSize=4, folder=root, seed=5, **kw)
    _same(_batches(ModelNetDataSet(device="cpu", **common)), _batches(Ref(**common)))


def test_shapenet_dataset_equals_the_reference(tmp_path, monkeypatch):
    Ref = _reference("ShapeNet", "ShapeNetDataSet")
    from mccnn_b200.datasets import ShapeNetDataSet
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "shape_data"))
    _write_shapenet(os.path.join(root, "shape_data"), np.random.default_rng(1))
    monkeypatch.chdir(root)                 # the reference reads "./shape_data/..." relative to the working directory
    # the file lists hold absolute paths here; the reference stores them as they are
    kw = dict(train=True, batchSize=3, ptDropOut=0.85, allowedSamplings=[0, 1, 2, 3], augment=True,
              useNormalsAsFeatures=True, seed=9)
    ours = ShapeNetDataSet(folder="./shape_data", device="cpu", **kw)
    theirs = Ref(**kw)
    assert ours.get_categories() == [tuple(c) for c in theirs.get_categories()]
    assert ours.get_categories_seg_parts() == theirs.get_categories_seg_parts()
    _same(_batches(ours), _batches(theirs))


def _write_scannet(root, rng):
    rooms = ["scene%04d_00" % i for i in range(8)]
    npts = []
    for r in rooms:
        n = int(rng.integers(400, 900))
        npts.append(n)
        np.save(os.path.join(root, r + "_pos.npy"), rng.uniform(-3, 3, (n, 3)))
        np.save(os.path.join(root, r + "_labels.npy"), rng.integers(0, 4, (n,)).astype(np.int32))
        np.save(os.path.join(root, r + "_colors.npy"), rng.uniform(0, 1, (n, 3)))
    np.savetxt(os.path.join(root, "rooms.txt"), np.array(rooms), fmt="%s")
    np.savetxt(os.path.join(root, "num_points.txt"), np.array(npts, np.float64))
    for split, sel in (("train", rooms[:4]), ("val", rooms[4:6]), ("test", rooms[6:])):
        np.savetxt(os.path.join(root, "scannet_%s.txt" % split), np.array(sel), fmt="%s")
    with open(os.path.join(root, "labels.txt"), "w") as f:
        f.write("0:unannotated\n1:wall\n2:floor\n3:chair\n")
    np.savetxt(os.path.join(root, "weights.txt"), rng.uniform(0.01, 0.5, (2, 4)))


@pytest.mark.parametrize("colors", [False, True])
def test_scannet_dataset_equals_the_reference(tmp_path, colors):
    Ref = _reference("ScanNet", "ScanNetDataSet")
    from mccnn_b200.datasets import ScanNetDataSet
    root = str(tmp_path)
    _write_scannet(root, np.random.default_rng(2))
    kw = dict(dataset=0, batchSize=3, ptDropOut=0.8, maxNumPtsxBatch=0, augment=True, useColorsAsFeatures=colors,
              dataFolder=root, seed=4)
    ours, theirs = ScanNetDataSet(device="cpu", **kw), Ref(**kw)
    assert list(ours.fileList_) == list(theirs.fileList_) and np.array_equal(ours.numPts_, theirs.numPts_)
    np.testing.assert_allclose(ours.weights_, theirs.weights_)
    ob, tb = _batches(ours), _batches(theirs)
    _same(ob, tb)
    labels = ob[0][4]
    np.testing.assert_allclose(ours.get_weights(labels).numpy(), theirs.get_weights(tb[0][4]), atol=1e-7)
    assert np.array_equal(ours.get_accuracy_masks(labels).numpy(), theirs.get_accuracy_masks(tb[0][4]))
    with pytest.raises(RuntimeError):
        ScanNetDataSet(0, 1, 1.0, maxNumPtsxBatch=1000, dataFolder=root, device="cpu")
