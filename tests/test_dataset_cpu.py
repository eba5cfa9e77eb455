"""mccnn_b200.DataSet (device-side batch assembly) against batches produced by the reference's
own utils/DataSet.py (tests/golden/dataset_*.npz, tests/golden/make_dataset_golden.py) with the same seed and the same
in-memory models: every batch of the iteration must contain the same points in the same order (all five sampling
protocols, wrap-around passes, point drop-out, the per-batch point budget), the same features / labels / categories,
and the random stream must stand at the same position afterwards.  Runs the tensor code on the CPU device; the same
code runs on CUDA tensors (tests/test_dataset_gpu.py)."""
import os

import numpy as np
import pytest
import torch

from common import DATASET_CASES, make_dataset_models

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_case(name, device):
    from mccnn_b200.DataSet import InMemoryDataSet
    case = DATASET_CASES[name]
    ds = InMemoryDataSet(make_dataset_models(), device=device, **case["ctor"])
    batches = []
    for _ in range(case["epochs"]):
        ds.start_iteration()
        while ds.has_more_batches():
            batches.append(ds.get_next_batch())
    return ds, batches


def check_against_golden(name, ds, batches):
    g = np.load(os.path.join(GOLDEN, "dataset_%s.npz" % name))
    assert len(batches) == int(g["num_batches"])
    for b, (n, pts, ids, feats, labels, cat, paths) in enumerate(batches):
        assert n == int(g["b%d_n" % b]) and list(paths) == list(g["b%d_paths" % b])
        assert pts.dtype == torch.float32 and ids.dtype == torch.int32 and feats.dtype == torch.float32
        want = g["b%d_pts" % b]
        assert tuple(pts.shape) == want.shape, (name, b)
        np.testing.assert_allclose(pts.cpu().numpy(), want.astype(np.float32), rtol=0, atol=2e-7)
        assert np.array_equal(ids.cpu().numpy(), g["b%d_ids" % b])
        np.testing.assert_allclose(feats.cpu().numpy(), g["b%d_feats" % b].astype(np.float32), rtol=0, atol=2e-7)
        if "b%d_labels" % b in g:
            assert np.array_equal(labels.cpu().numpy(), g["b%d_labels" % b])
        else:
            assert labels is None
        if "b%d_cat" % b in g:
            assert np.array_equal(cat.cpu().numpy().reshape(g["b%d_cat" % b].shape), g["b%d_cat" % b])
        else:
            assert cat is None
    assert ds.randomState_.random_sample() == float(g["next_draw"])   # same number of draws consumed


@pytest.mark.parametrize("name", list(DATASET_CASES))
def test_batches_equal_the_reference_class(name):
    ds, batches = run_case(name, "cpu")
    check_against_golden(name, ds, batches)


def test_argument_checks_follow_the_reference():
    from mccnn_b200.DataSet import InMemoryDataSet
    base = dict(numPoints=16, ptDropOut=1.0, pointFeatures=False, pointLabels=False, pointNormals=False,
                useCategories=False, pointCategories=False, batchSize=2, device="cpu")
    with pytest.raises(RuntimeError):
        InMemoryDataSet([], allowedSamplings=[0, 5], **base)          # invalid protocol
    with pytest.raises(RuntimeError):
        InMemoryDataSet([], allowedSamplings=[0, 3], **base)          # lambert needs normals
    with pytest.raises(RuntimeError):
        InMemoryDataSet([], allowedSamplings=[0], augmentedFeatures=[0, 2], **base)   # overlapping groups
    ds = InMemoryDataSet(make_dataset_models(3), allowedSamplings=[0], seed=1, **base)
    assert ds.get_num_models() == 3 and not ds.has_more_batches()
    ds.start_iteration()
    n, pts, ids, feats, labels, cat, paths = ds.get_next_batch()
    assert n == 2 and pts.shape == (32, 3) and ids.shape == (32, 1) and feats.shape == (32, 1)
    assert float(feats.min()) == 1.0 and labels is None and cat is None
