"""Running the reference's UNCHANGED Python on this library.

The reference's model files (`models/*.py`), `utils/MCConvBuilder.py` and `utils/MCNetworkUtils.py` import two things
this image does not have: `tensorflow` (1.x graph API) and the generated op module `MCConvModule` that loads
`MCConv.so`.  `install()` provides both:

  * `MCConvModule`  -> mccnn_b200.MCConvModule (same function names and argument orders, torch CUDA tensors);
  * `tensorflow`    -> a small EAGER stand-in over torch exposing exactly the symbols those files touch (SURVEY.md
                       section 8b): variables come from a mccnn_b200.variables.VariableStore under their TensorFlow names,
                       `tf.layers.batch_normalization` / `tf.matmul` / `tf.nn.relu` / `tf.concat` ... are torch ops.

    import mccnn_b200.compat as compat
    compat.install(reference_root="/path/to/MCCNN")       # adds <root>/models and <root>/utils to sys.path
    import MCClassS                                       # the reference's file, unmodified
    logits = MCClassS.create_network(points, batchIds, features, 1, 32, 64, 40, False, 1.0, 1.0)

Exact (shape-synchronising) mode only, like the reference itself: the torch stand-ins do not carry the device-side row
counts of the planned mode.  For the sync-free / graph-captured path use mccnn_b200.models.* (the same topologies
written against mccnn_b200's own builder and fused dense kernels).  `isTraining` is a Python bool here, not a
placeholder."""
import sys
import types

import torch

from .. import variables as _vars

BN_EPSILON = 1e-3
BN_MOMENTUM = 0.99


def make_tensorflow(store=None):
    """The torch-backed stand-in module for `tensorflow`."""
    vs = store if store is not None else _vars.default_store()
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32 = torch.float32, torch.int32
    tf.NoGradient = lambda name: None
    tf.RegisterGradient = lambda name: (lambda fn: fn)
    tf.load_op_library = lambda path: None

    def get_variable(name, shape=None, initializer=None, dtype=None):
        return vs.get_variable(name, list(shape), "zeros" if initializer == "zeros" else "glorot")

    def batch_normalization(inputs, training=False, name=None, **kw):
        c = inputs.shape[1]
        gamma = vs.get_variable(name + "/gamma", [c], "ones")
        beta = vs.get_variable(name + "/beta", [c], "zeros")
        mean = vs.get_variable(name + "/moving_mean", [c], "zeros")
        var = vs.get_variable(name + "/moving_variance", [c], "ones")
        if training:
            bm, bv = inputs.mean(dim=0), inputs.var(dim=0, unbiased=False)
            with torch.no_grad():
                mean.mul_(BN_MOMENTUM).add_(bm.detach(), alpha=1.0 - BN_MOMENTUM)
                var.mul_(BN_MOMENTUM).add_(bv.detach(), alpha=1.0 - BN_MOMENTUM)
            return (inputs - bm) * torch.rsqrt(bv + BN_EPSILON) * gamma + beta
        scale = gamma * torch.rsqrt(var + BN_EPSILON)
        return inputs * scale + (beta - mean * scale)

    def dropout(x, keep_prob):
        kp = float(keep_prob)
        return x if kp >= 1.0 else torch.nn.functional.dropout(x, p=1.0 - kp, training=True)

    def one_hot(indices, depth, on_value=1.0, off_value=0.0):
        oh = torch.nn.functional.one_hot(indices.long(), depth).to(torch.float32)
        return oh * (on_value - off_value) + off_value

    def ones(shape, dtype=torch.float32):
        dev = next((s.device for s in shape if isinstance(s, torch.Tensor)), vs.device)
        return torch.ones([int(s) for s in shape], dtype=dtype, device=dev)

    def constant(value, dtype=torch.float32):
        return torch.as_tensor(value, dtype=dtype, device=vs.device)

    def as_float(v):
        return v.to(torch.float32) if isinstance(v, torch.Tensor) else torch.as_tensor(float(v), device=vs.device)

    tf.get_variable = get_variable
    tf.add_to_collection = vs.add_to_collection
    tf.zeros_initializer = lambda: "zeros"
    tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
        variance_scaling_initializer=lambda factor=1.0, mode="FAN_AVG", uniform=True: "glorot"))
    tf.layers = types.SimpleNamespace(batch_normalization=batch_normalization)
    tf.nn = types.SimpleNamespace(relu=torch.relu, dropout=dropout)
    tf.matmul = torch.matmul
    tf.concat = lambda values, axis: torch.cat(list(values), dim=axis)
    tf.reshape = lambda t, shape: t.reshape([int(s) for s in shape])
    tf.shape = lambda t: torch.tensor(list(t.shape), dtype=torch.int32)
    tf.ones = ones
    tf.constant = constant
    tf.cast = lambda v, dtype: (v.to(dtype) if isinstance(v, torch.Tensor)
                                else torch.as_tensor(v, device=vs.device).to(dtype))
    tf.scalar_mul = lambda s, t: s * t
    tf.add = lambda a, b: a + b
    tf.multiply = lambda a, b: a * b
    tf.maximum = lambda a, b: torch.maximum(as_float(a), as_float(b))
    tf.logical_not = lambda v: (~v if isinstance(v, torch.Tensor) else torch.as_tensor(not v, device=vs.device))
    tf.reciprocal = lambda x: 1.0 / x
    tf.random_uniform = lambda shape: torch.rand([int(s) for s in shape], device=vs.device)
    tf.one_hot = one_hot
    tf.less_equal = lambda a, b: a <= b
    tf.greater_equal = lambda a, b: a >= b
    tf.variable_store = vs
    return tf


def install(reference_root=None, store=None, op_module=None):
    """Registers the stand-ins in sys.modules (and, if given, the reference's `models/` and `utils/` on sys.path).
    op_module: the module to serve as `MCConvModule` (default: mccnn_b200.MCConvModule).  Returns the tf stand-in."""
    import os
    if op_module is None:
        from .. import MCConvModule as op_module
    tf = make_tensorflow(store)
    sys.modules["tensorflow"] = tf
    sys.modules["MCConvModule"] = op_module
    for name in ("MCConvBuilder", "MCNetworkUtils"):   # must be re-imported against the stand-ins
        sys.modules.pop(name, None)
    if reference_root is not None:
        for sub in ("utils", "models"):
            p = os.path.join(reference_root, sub)
            if p not in sys.path:
                sys.path.insert(0, p)
    return tf


def uninstall():
    for name in ("tensorflow", "MCConvModule", "MCConvBuilder", "MCNetworkUtils", "MCClass", "MCClassS", "MCClassH",
                 "MCSeg", "MCSegScanNet", "MCNorm", "MCNormS"):
        sys.modules.pop(name, None)
