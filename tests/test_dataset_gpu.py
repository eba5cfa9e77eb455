"""mccnn_b200.DataSet on the GPU: the same reference-produced golden batches as tests/test_dataset_cpu.py, with the
models, the sampling protocols, the augmentation and the packing on cuda:0, and a batch fed straight into a network."""
import pytest

from common import DATASET_CASES
from test_dataset_cpu import check_against_golden, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(DATASET_CASES))
def test_batches_equal_the_reference_class_on_device(name):
    ds, batches = run_case(name, "cuda:0")
    assert all(b[1].is_cuda and b[2].is_cuda and b[3].is_cuda for b in batches)
    check_against_golden(name, ds, batches)


def test_assembled_batch_feeds_the_network():
    import torch
    from common import make_dataset_models
    from mccnn_b200 import variables
    from mccnn_b200.DataSet import InMemoryDataSet
    from mccnn_b200.models import MCClassS
    dev = torch.device("cuda:0")
    ds = InMemoryDataSet(make_dataset_models(8), numPoints=512, ptDropOut=1.0, pointFeatures=False, pointLabels=False,
                         pointNormals=True, useCategories=True, pointCategories=False, batchSize=8,
                         allowedSamplings=[0, 1, 2, 3, 4], augment=True, seed=5, device=dev)
    ds.start_iteration()
    n, pts, ids, feats, labels, cat, paths = ds.get_next_batch()
    assert n == 8 and pts.shape == (8 * 512, 3) and cat.shape == (8,)
    store = variables.VariableStore(device=dev, seed=0)
    logits = MCClassS.create_network(pts, ids, feats, 1, n, 16, 10, False, 1.0, 1.0, store=store)
    assert logits.shape[0] >= n and bool(torch.isfinite(logits[:n]).all())
