"""Golden outputs of the reference's OWN model files for tests/test_models_gpu.py.

    python tests/golden/make_model_golden.py            (in the build container: needs /root/reference)

Each network of /root/reference/models/*.py is imported unchanged and executed through the reference's unchanged
utils/MCConvBuilder.py, utils/MCNetworkUtils.py and tf_ops/MCConvModuleSrc by oracle/ref_models.py (NumPy stand-in for
TensorFlow, CPU oracle for MCConv.so).  Inputs and parameters are seeded (tests/common.model_case / named_variable),
so only the outputs are stored: model_<name>.npz with the logits, the data-dependent sizes in the reference's call
order (Poisson samples per level, neighbour pairs per find_neighbors) and the sorted variable names + shapes.
"""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from common import MODEL_CASES, model_case, named_variable  # noqa: E402
from oracle.ref_models import ReferenceModels  # noqa: E402


def run_reference(name):
    inputs, args = model_case(name)
    with ReferenceModels(named_variable) as R, contextlib.redirect_stdout(io.StringIO()):
        m = R.load(name)
        lead = [inputs["points"], inputs["batchIds"], inputs["features"]]
        if name == "MCSeg":
            lead.append(inputs["catLabels"])
        out = m.create_network(*lead, *args)
        return np.asarray(out, np.float32), list(R.log), dict(R.variables)


def main():
    for name in MODEL_CASES:
        out, log, variables = run_reference(name)
        names = sorted(variables)
        np.savez_compressed(
            os.path.join(HERE, "model_%s.npz" % name), logits=out,
            samples=np.asarray([e[1] for e in log if e[0] == "M"], np.int64),
            pairs=np.asarray([e[1] for e in log if e[0] == "K"], np.int64),
            convs=np.asarray([e[1:4] + (int(e[4]),) for e in log if e[0] == "conv"], np.int64),
            var_names=np.asarray(names), var_shapes=np.asarray([str(tuple(variables[n].shape)) for n in names]))
        print("%-13s out %s  max|out| %.4f  samples %s  pairs %s  %d variables" %
              (name, out.shape, np.abs(out).max(), [e[1] for e in log if e[0] == "M"],
               [e[1] for e in log if e[0] == "K"], len(names)))


if __name__ == "__main__":
    main()
