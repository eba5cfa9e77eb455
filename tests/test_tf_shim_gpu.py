"""mccnn_b200.compat on the GPU: a reference-STYLE network function (written here against `tensorflow as tf` and the
`MCConvModule` op names, the way the reference's own files are) runs once over the torch stand-in with
mccnn_b200.MCConvModule on cuda:0 and once over the NumPy stand-in with the CPU oracle ops; same named variables, so the
two results must agree.  (The reference's actual files cannot travel to the GPU box; tests/test_tf_shim_cpu.py runs
those over the same stand-in on the CPU.)"""
import sys
import types

import numpy as np
import pytest

from common import make_batch, named_variable

pytestmark = pytest.mark.gpu

NETWORK = '''
import tensorflow as tf
from MCConvModule import (compute_aabb, sort_points_step1, sort_points_step2, find_neighbors, compute_pdf, spatial_conv,
                          poisson_sampling, get_block_size)

def conv(name, pts, bids, feats, samples, sbids, mn, mx, B, cin, cout, radius):
    keys, idx = sort_points_step1(pts, bids, mn, mx, B, radius, True)
    sp, sb, sf, cells = sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, radius, True)
    start, pairs = find_neighbors(samples, sbids, sp, cells, mn, mx, radius, B, True)
    pdf = compute_pdf(sp, sb, mn, mx, start, pairs, 0.25, radius, B, True)
    bs = get_block_size()
    nb = (cin * cout + bs - 1) // bs
    init = tf.contrib.layers.variance_scaling_initializer(factor=1.0, mode='FAN_AVG', uniform=True)
    zeros = tf.zeros_initializer()
    w1 = tf.get_variable(name + '_weights', [3, bs * nb], initializer=init)
    b1 = tf.get_variable(name + '_biases', [bs * nb], initializer=zeros)
    w2 = tf.reshape(tf.get_variable(name + '_weights2', [nb, bs, bs], initializer=init), [bs, nb * bs])
    b2 = tf.reshape(tf.get_variable(name + '_biases2', [nb, bs], initializer=zeros), [nb * bs])
    w3 = tf.reshape(tf.get_variable(name + '_weights3', [nb, bs, bs], initializer=init), [bs, nb * bs])
    b3 = tf.reshape(tf.get_variable(name + '_biases3', [nb, bs], initializer=zeros), [nb * bs])
    return spatial_conv(sp, sf, sb, pdf, samples, start, pairs, mn, mx, w1, w2, w3, b1, b2, b3, cout, True, B, radius,
                        True, True)

def create_network(points, batchIds, features, B, isTraining):
    mn, mx = compute_aabb(points, batchIds, B, True)
    keys, idx = sort_points_step1(points, batchIds, mn, mx, B, 0.1, True)
    sp, sb, sf, cells = sort_points_step2(points, batchIds, features, keys, idx, mn, mx, B, 0.1, True)
    smp, smpb, sidx = poisson_sampling(sp, sb, cells, mn, mx, 0.1, B, True)
    x = conv("A", points, batchIds, features, smp, smpb, mn, mx, B, 2, 6, 0.2)
    x = tf.nn.relu(tf.layers.batch_normalization(inputs=x, training=isTraining, name="A_BN"))
    w = tf.get_variable("fc_weights", [6, 4], initializer=tf.contrib.layers.variance_scaling_initializer())
    b = tf.get_variable("fc_biases", [4], initializer=tf.zeros_initializer())
    y = tf.matmul(x, w) + b
    y = tf.concat([y, tf.scalar_mul(tf.constant(0.5), y)], 1)
    return conv("B", smp, smpb, y, smp, smpb, mn, mx, B, 8, 3, 0.4)
'''


def _run(tf_module, op_module, inputs, B):
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "MCConvModule")}
    sys.modules["tensorflow"], sys.modules["MCConvModule"] = tf_module, op_module
    try:
        mod = types.ModuleType("shim_network")
        exec(compile(NETWORK, "shim_network.py", "exec"), mod.__dict__)
        return mod.create_network(*inputs, B, False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_style_network_over_the_torch_stand_in():
    import torch
    import mccnn_b200.MCConvModule as M
    import mccnn_b200.compat as compat
    from mccnn_b200 import variables
    from oracle import cpu as O
    from oracle.ref_models import _make_tf
    dev = torch.device("cuda:0")
    B = 3
    pts, bids, feats = make_batch(B, 500, "sphere", seed=8, num_features=2, ragged=True)

    class NamedStore(variables.VariableStore):
        def get_variable(self, name, shape, initializer="glorot"):
            if name not in self.vars:
                self.vars[name] = torch.from_numpy(named_variable(name, shape)).to(self.device)
            return self.vars[name]

    store = NamedStore(device=dev)
    got = _run(compat.make_tensorflow(store), M, [torch.from_numpy(a).to(dev) for a in (pts, bids, feats)], B)
    assert got.is_cuda
    np_vars = {}
    want = _run(_make_tf(np_vars, [], named_variable, np.random.default_rng(0)), O, [pts, bids, feats], B)
    assert sorted(np_vars) == sorted(store.vars)
    got = got.cpu().numpy()
    assert got.shape == want.shape and got.shape[1] == 3
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
