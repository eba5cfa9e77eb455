"""CPU checks around the model golden vectors (tests/golden/model_*.npz, made by the reference's own model files):
the restated MCClassS pipeline used by bench.py's cpu_baseline reproduces them, and -- where the reference tree is
present -- re-running the reference model files reproduces the committed fixtures exactly."""
import os

import numpy as np
import pytest

from common import MODEL_CASES, model_case, named_variable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fixtures_present():
    for name in MODEL_CASES:
        g = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
        assert g["logits"].size > 0 and np.isfinite(g["logits"]).all()
        assert len(g["var_names"]) == len(g["var_shapes"]) > 0


def test_oracle_pipeline_equals_reference_mcclasss():
    """oracle/pipeline.py (the restated MCClassS forward timed as cpu_baseline) == models/MCClassS.py run unchanged."""
    from oracle import pipeline as OP
    g = np.load(os.path.join(GOLDEN, "model_MCClassS.npz"))
    inputs, args = model_case("MCClassS")
    params = {n: named_variable(n, eval(s)) for n, s in zip(g["var_names"], g["var_shapes"])}
    logits, trace = OP.mcclasss_forward(inputs["points"], inputs["batchIds"], inputs["features"], args[0], args[1],
                                        args[2], args[3], params)
    assert trace[0]["levels"][1:] == list(g["samples"])
    assert [t["K"] for t in trace[1:]] == list(g["pairs"])
    np.testing.assert_array_equal(logits, g["logits"])


@pytest.mark.parametrize("name", ["MCClassS", "MCNormS", "MCSeg"])
def test_reference_model_files_reproduce_fixtures(name):
    from oracle import ref_models
    if not ref_models.available():
        pytest.skip("reference tree not present on this machine")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(GOLDEN, "make_model_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out, log, variables = mod.run_reference(name)
    g = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    np.testing.assert_array_equal(out, g["logits"])
    assert sorted(variables) == list(g["var_names"])
