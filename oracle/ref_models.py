"""Runs the reference's UNCHANGED model files on the CPU.  TEST INFRASTRUCTURE (golden-vector generation).

`/root/reference/models/*.py`, `utils/MCConvBuilder.py`, `utils/MCNetworkUtils.py` and the op module source
`tf_ops/MCConvModuleSrc` are imported from where they lie (nothing is copied) with two stand-ins:

  * `tensorflow`  -> a small eager NumPy module exposing exactly the symbols those files touch;
  * `MCConv.so`   -> `tf.load_op_library` returns an object whose raw ops (the names and argument orders of the
                     REGISTER_OP declarations in tf_ops/*.cc) are the CPU oracle (oracle/cpu.py, pinned against the
                     reference kernels by tests/test_oracle_golden.py).

So the network topology, the builder's caching logic, variable names / shapes / reshapes and the two argument
re-orderings of the op wrappers are the reference's own code; only the kernels and TensorFlow's dense math are
restated.  Variables come from `tests/common.named_variable` (seeded by name), the same generator the GPU tests feed
into mccnn_b200's VariableStore.  Only usable where /root/reference exists (this container), never on the GPU box:
tests/golden/make_model_golden.py commits its outputs as fixtures.
"""
import importlib
import os
import sys
import types

import numpy as np

from . import cpu as O

REFERENCE = "/root/reference"
BN_EPS = np.float32(1e-3)  # tf.layers.batch_normalization default epsilon


def available():
    return os.path.isdir(os.path.join(REFERENCE, "models"))


class _OpLibrary:
    """Raw ops of MCConv.so as MCConvModuleSrc calls them (tf_ops/MCConvModuleSrc:20-73)."""

    def __init__(self, log):
        self.log = log

    def compute_aabb(self, pts, bids, batch_size, scale_inv):
        return O.compute_aabb(pts, bids, batch_size, scale_inv)

    def sort_points_step1(self, pts, bids, mn, mx, batch_size, cell_size, scale_inv):
        return O.sort_points_step1(pts, bids, mn, mx, batch_size, cell_size, scale_inv)

    def sort_points_step2(self, pts, bids, feats, keys, idx, mn, mx, batch_size, cell_size, scale_inv):
        return O.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, batch_size, cell_size, scale_inv)

    def sort_features_back_grad(self, idx, feats):       # python name: sort_features (MCConvModuleSrc:33-34)
        return O.sort_features(feats, idx)

    def sort_features_back(self, feats, idx):
        return O.sort_features_back(feats, idx)

    def transform_indexs(self, idx, new_pos):
        return O.transform_indexs(idx, new_pos)

    def find_neighbors(self, pts, bids, pts2, cells, mn, mx, radius, batch_size, scale_inv):
        start, pairs = O.find_neighbors(pts, bids, pts2, cells, mn, mx, radius, batch_size, scale_inv)
        self.log.append(("K", int(pairs.shape[0])))
        return start, pairs

    def compute_pdf(self, pts, bids, start, pairs, mn, mx, window, batch_size, radius, scale_inv):  # raw order :55-56
        return O.compute_pdf(pts, bids, mn, mx, start, pairs, window, radius, batch_size, scale_inv)

    def poisson_sampling(self, pts, bids, cells, mn, mx, radius, batch_size, scale_inv):
        out = O.poisson_sampling(pts, bids, cells, mn, mx, radius, batch_size, scale_inv)
        self.log.append(("M", int(out[0].shape[0])))
        return out

    def get_sampled_features(self, idx, feats):
        return O.get_sampled_features(idx, feats)

    def spatial_conv(self, pts, feats, bids, pdfs, samples, start, pairs, mn, mx, w1, b1, w2, b2, wo, bo,  # raw :70-73
                     num_out_features, combin, batch_size, radius, scale_inv, avg):
        self.log.append(("conv", int(pairs.shape[0]), int(feats.shape[1]), int(num_out_features), bool(combin)))
        return O.spatial_conv(pts, feats, bids, pdfs, samples, start, pairs, mn, mx, w1, w2, wo, b1, b2, bo,
                              num_out_features, combin, batch_size, radius, scale_inv, avg)


def _make_tf(variables, log, named_variable, rng):
    """The NumPy stand-in for `tensorflow` (eager; tensors are float32 / int32 ndarrays)."""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32 = np.float32, np.int32
    tf.load_op_library = lambda path: _OpLibrary(log)
    tf.NoGradient = lambda name: None
    tf.RegisterGradient = lambda name: (lambda fn: fn)
    collections = {}

    def get_variable(name, shape=None, initializer=None, dtype=None):
        if name not in variables:
            variables[name] = named_variable(name, shape)
        v = variables[name]
        assert tuple(v.shape) == tuple(shape), (name, v.shape, shape)
        return v

    def batch_normalization(inputs, training=False, name=None, **kw):
        assert training is False, "golden vectors are inference-mode"
        c = inputs.shape[1]
        g = get_variable(name + "/gamma", [c])
        b = get_variable(name + "/beta", [c])
        m = get_variable(name + "/moving_mean", [c])
        v = get_variable(name + "/moving_variance", [c])
        scale = (g / np.sqrt(v + BN_EPS)).astype(np.float32)
        return (inputs * scale + (b - m * scale)).astype(np.float32)

    def dropout(x, keep_prob):
        assert float(keep_prob) == 1.0, "golden vectors are inference-mode"
        return x

    def one_hot(idx, depth, on_value=1.0, off_value=0.0):
        idx = np.asarray(idx)
        out = np.full(idx.shape + (depth,), off_value, np.float32)
        np.put_along_axis(out, idx[..., None].astype(np.int64), np.float32(on_value), axis=-1)
        return out

    tf.get_variable = get_variable
    tf.add_to_collection = lambda name, v: collections.setdefault(name, []).append(v)
    tf.zeros_initializer = lambda: "zeros"
    tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
        variance_scaling_initializer=lambda factor=1.0, mode="FAN_AVG", uniform=True: "glorot"))
    tf.layers = types.SimpleNamespace(batch_normalization=batch_normalization)
    tf.nn = types.SimpleNamespace(relu=lambda x: np.maximum(x, np.float32(0)), dropout=dropout)
    tf.matmul = lambda a, b: (np.asarray(a, np.float32) @ np.asarray(b, np.float32)).astype(np.float32)
    tf.concat = lambda vals, axis: np.concatenate(vals, axis=axis)
    tf.reshape = lambda t, shape: np.reshape(t, shape)
    tf.shape = lambda t: np.asarray(t.shape, np.int32)
    tf.ones = lambda shape, dtype=np.float32: np.ones([int(s) for s in shape], dtype)
    tf.constant = lambda v, dtype=np.float32: np.asarray(v, dtype)
    tf.cast = lambda v, dtype: np.asarray(v).astype(dtype)
    tf.scalar_mul = lambda s, t: (np.float32(s) * t).astype(np.float32)
    tf.add = lambda a, b: a + b
    tf.multiply = lambda a, b: a * b
    tf.maximum = np.maximum
    tf.logical_not = np.logical_not
    tf.reciprocal = lambda x: (np.float32(1.0) / x).astype(np.float32)
    tf.random_uniform = lambda shape: rng.uniform(size=shape).astype(np.float32)
    tf.one_hot = one_hot
    tf.less_equal = np.less_equal
    tf.greater_equal = np.greater_equal
    tf.collections = collections
    return tf


_SHADOWED = ("tensorflow", "MCConvModule", "MCConvBuilder", "MCNetworkUtils")


class ReferenceModels:
    """Context manager: inside it `load(model)` returns the reference's model module, e.g. load("MCSeg").
    variables: dict name -> ndarray, filled on demand; log: list of ("M", samples) / ("K", pairs) / ("conv", ...) in
    the reference's call order."""

    def __init__(self, named_variable, seed=0):
        self.variables, self.log = {}, []
        self._named, self._rng = named_variable, np.random.default_rng(seed)

    def __enter__(self):
        if not available():
            raise RuntimeError("the reference tree is not present (only in the build container)")
        self._saved = {k: sys.modules.get(k) for k in _SHADOWED}
        self._path = list(sys.path)
        tf = _make_tf(self.variables, self.log, self._named, self._rng)
        sys.modules["tensorflow"] = tf
        # MCConvModule.py is generated by tf_ops/genCompileScript.py:41-49 = MCConvModuleSrc + get_block_size()
        mod = types.ModuleType("MCConvModule")
        mod.__file__ = os.path.join(REFERENCE, "tf_ops", "MCConvModule.py")
        with open(os.path.join(REFERENCE, "tf_ops", "MCConvModuleSrc")) as f:
            src = f.read()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # the banner comment contains '\c' escapes
            code = compile(src + "\n\ndef get_block_size():\n    return 8\n", mod.__file__, "exec")
        exec(code, mod.__dict__)
        sys.modules["MCConvModule"] = mod
        for k in ("MCConvBuilder", "MCNetworkUtils"):
            sys.modules.pop(k, None)
        sys.path.insert(0, os.path.join(REFERENCE, "utils"))
        sys.path.insert(0, os.path.join(REFERENCE, "models"))
        self.tf = tf
        return self

    def load(self, name):
        sys.modules.pop(name, None)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # the files' banner comments contain '\c' escapes
            return importlib.import_module(name)

    def __exit__(self, *exc):
        sys.path[:] = self._path
        for k, v in self._saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in ("MCClass", "MCClassS", "MCClassH", "MCSeg", "MCSegScanNet", "MCNorm", "MCNormS"):
            sys.modules.pop(k, None)
        return False
