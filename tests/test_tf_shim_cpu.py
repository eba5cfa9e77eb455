"""mccnn_b200.compat: the reference's UNCHANGED model files, builder and network utilities executed over the torch
stand-in for `tensorflow`.  No GPU here, so the op module behind the builder is the CPU oracle (adapted to torch CPU
tensors): what is under test is the stand-in's semantics -- variable creation, batch norm, matmul, concat, one_hot,
the path-averaging arithmetic of MCClassH -- which must reproduce the golden logits made with the NumPy stand-in
(tests/golden/model_*.npz).  The op module a deployment uses instead (mccnn_b200.MCConvModule on CUDA tensors) is
checked op by op in tests/test_parity_gpu.py.  Skipped where the reference tree is absent."""
import contextlib
import io
import os
import types
import warnings

import numpy as np
import pytest
import torch

from common import MODEL_CASES, model_case, named_variable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REFERENCE = "/root/reference"


def _oracle_op_module():
    """oracle.cpu behind the MCConvModule names, torch CPU tensors in and out."""
    from oracle import cpu as O
    mod = types.ModuleType("MCConvModule")

    def wrap(fn):
        def w(*args):
            conv = [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a for a in args]
            out = fn(*conv)
            if isinstance(out, tuple):
                return tuple(torch.from_numpy(np.ascontiguousarray(o)) for o in out)
            return torch.from_numpy(np.ascontiguousarray(out))
        return w

    for name in ("compute_aabb", "sort_points_step1", "sort_points_step2", "sort_features", "sort_features_back",
                 "transform_indexs", "find_neighbors", "compute_pdf", "poisson_sampling", "get_sampled_features",
                 "spatial_conv"):
        setattr(mod, name, wrap(getattr(O, name)))
    mod.get_block_size = O.get_block_size
    return mod


@pytest.mark.parametrize("name", ["MCClassS", "MCClassH", "MCSeg", "MCNormS"])
def test_reference_model_files_run_on_the_torch_stand_in(name):
    if not os.path.isdir(os.path.join(REFERENCE, "models")):
        pytest.skip("reference tree not present on this machine")
    import mccnn_b200.compat as compat
    from mccnn_b200 import variables

    class NamedStore(variables.VariableStore):
        def get_variable(self, vname, shape, initializer="glorot"):
            if vname not in self.vars:
                self.vars[vname] = torch.from_numpy(named_variable(vname, shape))
            return self.vars[vname]

    store = NamedStore(device="cpu")
    compat.install(reference_root=REFERENCE, store=store, op_module=_oracle_op_module())
    try:
        import importlib
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            model = importlib.import_module(name)
            inputs, args = model_case(name)
            lead = [torch.from_numpy(inputs[k]) for k in ("points", "batchIds", "features")]
            if name == "MCSeg":
                lead.append(torch.from_numpy(inputs["catLabels"]))
            out = model.create_network(*lead, *args)
    finally:
        compat.uninstall()
    g = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    assert sorted(store.vars) == list(g["var_names"])
    got = out.detach().numpy()
    assert got.shape == g["logits"].shape
    assert np.abs(got - g["logits"]).max() <= 2e-5 * np.abs(g["logits"]).max()
