"""Every network of the reference (models/MCClass.py, MCClassH.py, MCClassS.py, MCSeg.py, MCSegScanNet.py, MCNorm.py,
MCNormS.py) through mccnn_b200's builder API on the GPU, against golden outputs produced by the reference's OWN,
unchanged model files (tests/golden/make_model_golden.py: reference models + builder + op wrappers, CPU oracle ops).

Checked per model: the same variables (names and shapes) are created; every data-dependent size (Poisson samples per
level, pairs per neighbour search) is identical -- those are index-exact quantities; the logits agree to a tolerance
that covers fp32 summation-order differences accumulated through up to 17 convolutions; the sync-free planned run and
its CUDA-graph replay reproduce the exact-mode result bit for bit.
"""
import os

import numpy as np
import pytest

from common import MODEL_CASES, model_case, named_variable

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOGIT_TOL = 2e-4  # of max |logit|: fp32 re-association through the whole network (each conv is 1e-5 on its own sums)


def _named_store(dev):
    import torch
    from mccnn_b200 import variables

    class NamedStore(variables.VariableStore):
        def get_variable(self, name, shape, initializer="glorot"):
            if name not in self.vars:
                self.vars[name] = torch.from_numpy(named_variable(name, shape)).to(self.device)
            v = self.vars[name]
            assert tuple(v.shape) == tuple(shape), (name, tuple(v.shape), tuple(shape))
            return v
    return NamedStore(device=dev)


def _network(name, dev):
    import importlib
    import torch
    inputs, args = model_case(name)
    t = {k: torch.from_numpy(v).to(dev) for k, v in inputs.items()}
    store = _named_store(dev)
    mod = importlib.import_module("mccnn_b200.models." + name)

    def fn(p, b, f):
        lead = (p, b, f) + ((t["catLabels"],) if name == "MCSeg" else ())
        return mod.create_network(*lead, *args, store=store)
    return fn, store, t


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_model_matches_reference_model_file(name):
    import torch
    from mccnn_b200 import runtime
    dev = torch.device("cuda:0")
    gold = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    fn, store, t = _network(name, dev)
    with runtime.record() as plan:
        out = fn(t["points"], t["batchIds"], t["features"])
    rows = gold["logits"].shape[0]
    out = out[:rows].cpu().numpy()

    # variables: same names, same shapes as the reference's tf.get_variable / batch_normalization calls
    assert sorted(store.vars) == list(gold["var_names"])
    assert [str(tuple(store.vars[n].shape)) for n in sorted(store.vars)] == list(gold["var_shapes"])
    # index-exact quantities: our call order may differ from the reference's (lazy hierarchy features), the multiset
    # of sizes may not
    assert sorted(plan.sizes) == sorted(list(gold["samples"]) + list(gold["pairs"]))
    # logits
    assert out.shape == gold["logits"].shape and np.isfinite(out).all()
    err = np.abs(out - gold["logits"]).max() / np.abs(gold["logits"]).max()
    assert err <= LOGIT_TOL, "%s: logits differ from the reference model file by %.3e of max|logit|" % (name, err)


@pytest.mark.parametrize("name", ["MCClassH", "MCSeg", "MCSegScanNet", "MCNorm"])
def test_model_planned_and_graph(name):
    """Sync-free planned execution (capacity-padded tensors, device-resident counts) and its CUDA-graph replay.
    Planned and replayed results are bit-identical to each other and across repetitions (every kernel is
    deterministic).  Against the exact-mode run they agree to fp32 re-association only: spatial_conv picks its run
    length per warp from the (capacity vs exact) sizes, which permutes the summation order."""
    import torch
    from mccnn_b200.engine import PlannedForward
    dev = torch.device("cuda:0")
    gold = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    rows = gold["logits"].shape[0]
    fn, store, t = _network(name, dev)
    x = (t["points"], t["batchIds"], t["features"])
    ref = fn(*x)[:rows].clone()
    scale = float(ref.abs().max())
    outs = []
    for use_graph in (False, True):
        eng = PlannedForward(fn, rows, use_graph=use_graph)
        assert torch.equal(eng(*x), ref)                 # first call records: exact mode
        a, b = eng(*x).clone(), eng(*x).clone()          # planned / replayed
        assert torch.equal(a, b)
        assert float((a - ref).abs().max()) <= 1e-5 * scale
        assert eng.stats["overflows"] == 0
        outs.append(a)
    assert torch.equal(outs[0], outs[1])                 # graph replay == eager planned
    assert eng.stats["captures"] == 1


def test_scannet_plan_detects_a_different_grid_resolution():
    """Absolute radii: the number of cells comes from the batch's box.  A planned forward reuses the recorded value and
    verifies it on the device (mc_verify_num_cells); a batch whose box implies another resolution must come back
    flagged and be re-run exactly -- never computed on the wrong grid."""
    import torch
    from mccnn_b200.engine import PlannedForward
    dev = torch.device("cuda:0")
    fn, store, t = _network("MCSegScanNet", dev)
    x = (t["points"], t["batchIds"], t["features"])
    rows = x[0].shape[0]
    eng = PlannedForward(fn, rows, use_graph=True)
    eng(*x)
    eng(*x)
    assert eng.stats["overflows"] == 0
    stretched = (x[0] * 1.5, x[1], x[2])                 # a larger room: more cells per axis for the same radii
    want = fn(*stretched)[:rows].clone()
    got = eng(*stretched)
    assert eng.stats["overflows"] == 1
    assert torch.equal(got, want)
