"""oracle/ref_nets.py -- the restated builder + topologies that stand in for the reference's model files on the GPU box
(reference arm of bench.py, full-size parity tests) -- reproduces the golden outputs of the reference's OWN unchanged
model files (tests/golden/model_*.npz) when run over the CPU oracle ops: same Poisson-sample counts per level, same
neighbour-pair counts in call order, same variable names and shapes, same logits."""
import os

import numpy as np
import pytest
import torch

from common import model_case, named_variable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["MCClassS", "MCClassH", "MCSeg", "MCSegScanNet"])
def test_restated_networks_equal_the_reference_model_files(name):
    from oracle import ref_nets
    g = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    inputs, args = model_case(name)
    params, log = {}, []
    out = ref_nets.forward(name, ref_nets.CpuOps(), params, named_variable,
                           {k: torch.from_numpy(v) for k, v in inputs.items()}, args[:-3], log=log)
    assert [e[1] for e in log if e[0] == "M"] == list(g["samples"])
    assert [e[1] for e in log if e[0] == "K"] == list(g["pairs"])
    assert [list(e[1:4]) + [int(e[4])] for e in log if e[0] == "conv"] == g["convs"].tolist()
    assert sorted(params) == list(g["var_names"])
    assert [str(tuple(params[n].shape)) for n in sorted(params)] == list(g["var_shapes"])
    out = out.numpy()
    assert out.shape == g["logits"].shape
    # the dense glue is torch here and NumPy in the golden run: float32 re-association only
    assert np.abs(out - g["logits"]).max() <= 2e-5 * np.abs(g["logits"]).max()
