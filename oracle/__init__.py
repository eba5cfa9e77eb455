"""Parity oracle for mccnn_b200 -- TEST INFRASTRUCTURE, not part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

  oracle.cpu      numpy-in / numpy-out bindings of liboracle.so (mccnn_oracle.c: the C restatement of the reference)
  oracle.numpy_ops  vectorised NumPy restatement of compute_pdf / spatial_conv (BASELINE.json config 1)
  oracle.ref      torch-CUDA bindings of oracle/_ref/libmccnn_ref.so = the reference's own kernels (GPU box only)
  oracle.pipeline MCClassS forward assembled from oracle.cpu ops (+ NumPy dense glue)
  oracle.ref_nets the reference's builder and the four benchmarked topologies restated over an op module (oracle.ref on
                  the GPU = bench.py --impl reference; oracle.cpu behind torch CPU tensors = cpu_baseline), held to the
                  outputs of the reference's unchanged model files (oracle.ref_models, tests/golden/model_*.npz)
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIBORACLE = os.path.join(HERE, "liboracle.so")
LIBREF = os.path.join(HERE, "_ref", "libmccnn_ref.so")


def build(verbose=False):
    """make -C oracle: liboracle.so always; _ref/libmccnn_ref.so when /root/reference is present."""
    res = subprocess.run(["make", "-C", HERE, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + res.stderr[-4000:])
