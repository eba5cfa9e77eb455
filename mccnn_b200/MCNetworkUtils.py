"""Dense glue between the Monte-Carlo convolutions (utils/MCNetworkUtils.py:20-142): batch-norm + ReLU (+ dropout),
1x1 "convolution" (a matmul) and the final MLPs.  Same function names and argument lists as the reference.

In the reference this layer is TensorFlow (tf.layers.batch_normalization, tf.matmul, tf.nn.relu).  Here every function
is one launch per layer -- mccnn_b200/csrc/dense_tc.cu (tcgen05 tensor cores) for the large layers, csrc/dense.cu for
the small ones: inference batch norms are folded into per-column (scale, shift) vectors -- cached until a variable
changes -- and applied in the prologue / epilogue of the layer's FP32 GEMM.
Inference (the measured path): moving mean / variance, epsilon 1e-3 as tf.layers.batch_normalization.  Training
(isTraining=True, exact mode only): the same functions run as differentiable torch ops -- batch statistics with the
moving averages updated at TensorFlow's momentum 0.99, dropout with keepProb -- so that, together with the autograd
wiring of the Monte-Carlo ops (MCConvModule), loss.backward() reaches every variable of a network.  Live row counts of capacity-padded inputs
(runtime.count_of) are honoured and propagated.
"""
import torch

from . import ops
from . import variables as _vars
from .runtime import DeferredAffine, count_of as _cnt, dispatch as _dispatch, hint_of as _hint, materialize as _real, tag as _tag

BN_EPSILON = 1e-3  # tf.layers.batch_normalization default


def _store(store):
    return store if store is not None else _vars.default_store()


BN_MOMENTUM = 0.99  # tf.layers.batch_normalization default (moving = 0.99 * moving + 0.01 * batch)


def _bn_vars(c, name, vs):
    return (vs.get_variable(name + "/gamma", [c], "ones"), vs.get_variable(name + "/beta", [c], "zeros"),
            vs.get_variable(name + "/moving_mean", [c], "zeros"), vs.get_variable(name + "/moving_variance", [c], "ones"))


def _bn_train(x, name, vs):
    """Training-mode batch norm as tf.layers.batch_normalization(training=True): normalise with the batch mean and
    (biased) variance, update the moving statistics in place (the reference runs the UPDATE_OPS collection with every
    step, ModelNet/ModelNet.py:199-203).  Plain differentiable torch ops; exact-shape tensors only."""
    from . import runtime as _rtm
    if _rtm.mode() == "planned" or _cnt(x) is not None:
        raise RuntimeError("mccnn_b200: training-mode batch norm needs exact-shape tensors (not a planned run)")
    x = _real(x)
    gamma, beta, mean, var = _bn_vars(x.shape[1], name, vs)
    bm = x.mean(dim=0)
    bv = x.var(dim=0, unbiased=False)
    with torch.no_grad():
        mean.mul_(BN_MOMENTUM).add_(bm.detach(), alpha=1.0 - BN_MOMENTUM)
        var.mul_(BN_MOMENTUM).add_(bv.detach(), alpha=1.0 - BN_MOMENTUM)
    return (x - bm) * torch.rsqrt(bv + BN_EPSILON) * gamma + beta


def _dropout(x, keepProb, use):
    if not use or keepProb >= 1.0:
        return x
    return torch.nn.functional.dropout(x, p=1.0 - float(keepProb), training=True)  # scales by 1 / keepProb as tf.nn.dropout


def _bn_affine(c, training, name, vs):
    """(scale, shift) of an inference batch norm: y = x * scale + shift.  Creates the four TF variables on first use
    (gamma 1, beta 0, moving_mean 0, moving_variance 1) and caches the folded vectors until one of them changes
    (tensor._version counts in-place updates)."""
    if training:
        raise RuntimeError("internal: the folded batch norm is the inference path")
    gamma = vs.get_variable(name + "/gamma", [c], "ones")
    beta = vs.get_variable(name + "/beta", [c], "zeros")
    mean = vs.get_variable(name + "/moving_mean", [c], "zeros")
    var = vs.get_variable(name + "/moving_variance", [c], "ones")
    versions = (gamma._version, beta._version, mean._version, var._version)
    cache = vs.__dict__.setdefault("_bn_folded", {})
    hit = cache.get(name)
    if hit is None or hit[0] != versions:
        scale = gamma * torch.rsqrt(var + BN_EPSILON)
        shift = beta - mean * scale
        hit = (versions, scale.contiguous(), shift.contiguous())
        cache[name] = hit
    return hit[1], hit[2]


def _dense(x, w, b, pre=None, **kw):
    """One launch.  A pending batch norm (+ ReLU) on the input becomes the layer's prologue; if the layer has a prologue
    of its own (`pre`, no activation), that second affine is folded into the weights instead:
    (a * s + t) @ W + b = a @ (s[:, None] * W) + (t @ W + b)."""
    pre_relu = False
    if isinstance(x, DeferredAffine):
        if pre is not None:
            w, b = _fold_into_weights(pre, w, b)
        pre, pre_relu, x = (x.scale, x.shift), x.relu, x.base
    nd, h = _cnt(x), _hint(x)
    return _tag(_layer(x, w, b, pre, pre_relu, kw.get("post"), kw.get("post_relu", False), nd, h), nd, h)


# Which kernel runs a layer.  Few rows (the classification heads, one row per cloud): the row kernels of dense.cu, whose
# K loop is split over the CTA.  Large layers -- the per-point 1x1 layers of the segmentation networks (MCSeg: 131 072
# rows x 832 -> 256 = 28 GMAC), the lower levels of MCClassH -- are dense GEMMs and run on the tcgen05 tensor cores
# (dense_tc.cu, FP32 via the 3xTF32 split).  In between (a few thousand rows x 64..256 columns: the reductions of
# MCClassS, 12-44 MMAC) a tcgen05 tile pipeline has 2-4 K blocks to amortise its prologue over and the latency-tuned
# FFMA tile kernel is as fast or faster (measured 14.7 / 16.7 us against 23.7 / 17 us); it also takes the layers the TMA
# path cannot address (K not a multiple of 4, K < 32).  No library GEMM on the path.
_TC_MIN_ROWS = 129
_TC_MIN_MACS = 1 << 26
# few rows (one per cloud) against a wide layer: the layer is a stream of W -- dense.cu's few-rows kernel reads it once
_FEWROWS_MAX_ROWS = 128
_FEWROWS_MAX_CAPACITY = 1024     # rows of the (capacity-padded) input: bounds the K-segment workspace
_FEWROWS_MIN_WEIGHTS = 1 << 16


def _layer(x, w, b, pre, pre_relu, post, post_relu, nd, h):
    rows = h if h else x.shape[0]
    if (rows >= _TC_MIN_ROWS and rows * w.shape[0] * w.shape[1] >= _TC_MIN_MACS and
            ops.dense_layer_tc_supported(w.shape[0], w.shape[1])):
        hi, lo = _split_weights(w)
        return ops.dense_layer_tc(x, hi, lo, b, pre=pre, pre_relu=pre_relu, post=post, post_relu=post_relu, n_dev=nd)
    if (rows <= _FEWROWS_MAX_ROWS and x.shape[0] <= _FEWROWS_MAX_CAPACITY and
            w.shape[0] * w.shape[1] >= _FEWROWS_MIN_WEIGHTS):
        return ops.dense_layer_fewrows(x, w, b, pre=pre, pre_relu=pre_relu, post=post, post_relu=post_relu, n_dev=nd)
    return ops.dense_layer(x, w, b, pre=pre, pre_relu=pre_relu, post=post, post_relu=post_relu, n_dev=nd, rows_hint=h)


_split = {}


def _split_weights(w):
    """Transposed tf32 hi / lo halves of a weight matrix, cached until the matrix changes (tensor._version)."""
    key = id(w)
    hit = _split.get(key)
    if hit is None or hit[0] != w._version or hit[3] is not w:
        hi, lo = ops.split_tf32_transposed(w.detach())
        hit = (w._version, hi, lo, w)  # keeps w alive: id() stays unique
        _split[key] = hit
    return hit[1], hit[2]


def _mlp(x, layers, pre=None):
    """The classification heads: few rows (one per cloud), so the whole MLP is ONE dense.cu launch (mc_mlp_rows) with
    the hidden activations in shared memory; falls back to one launch per layer for many rows or wide layers.
    layers = [(w, b, post (scale, shift) | None, post_relu), ...]; a pending batch norm on x becomes the prologue and
    an explicit `pre` affine is then folded into the first layer's weights (see _dense)."""
    pre_relu = False
    if isinstance(x, DeferredAffine):
        if pre is not None:
            w, b = _fold_into_weights(pre, layers[0][0], layers[0][1])
            layers = [(w, b) + tuple(layers[0][2:])] + list(layers[1:])
        pre, pre_relu, x = (x.scale, x.shift), x.relu, x.base
    nd, h = _cnt(x), _hint(x)
    rows = h if h else x.shape[0]
    widths = [x.shape[1]] + [l[0].shape[1] for l in layers]
    if rows <= 128 and max(widths) <= 1024 and len(layers) <= 4:
        return _tag(ops.mlp_rows(x, layers, pre=pre, pre_relu=pre_relu, n_dev=nd), nd, h)
    for i, (w, b, post, post_relu) in enumerate(layers):
        x = _tag(_layer(x, w, b, pre if i == 0 else None, pre_relu if i == 0 else False, post, post_relu, nd, h),
                 nd, h)
    return x


_folded = {}


def _fold_into_weights(pre, w, b):
    key = (id(w), id(b))
    versions = (pre[0]._version, pre[1]._version, w._version, b._version, id(pre[0]), id(pre[1]))
    hit = _folded.get(key)
    if hit is None or hit[0] != versions:
        hit = (versions, (pre[0][:, None] * w).contiguous(), (b + pre[1] @ w).contiguous(), w, b)  # keeps w, b alive
        _folded[key] = hit
    return hit[1], hit[2]


def _affine_act(x, scale, shift, relu):
    """Inference batch norm (+ ReLU): returned as a pending per-column affine that the consumer folds into its own
    kernel (runtime.DeferredAffine); stacked on another pending one, the inner one is evaluated first."""
    return DeferredAffine(_real(x), scale, shift, relu)


def concat(values, axis=1):
    """tf.concat of feature matrices that live on the same point level (models/MCClass.py:69 and every encoder /
    decoder skip link).  The one TensorFlow call of the model files that has to know about capacity-padded tensors:
    the result keeps the live row count of its inputs and is issued on the stream of the current lane."""
    vals = [_real(v) for v in values]
    nd = next((c for c in (_cnt(v) for v in vals) if c is not None), None)
    h = next((c for c in (_hint(v) for v in vals) if c is not None), None)
    return _tag(_dispatch(vals, lambda: torch.cat(vals, axis)), nd, h)


def batch_normalization(inputs, training, name, store=None):
    if training:
        return _bn_train(inputs, name, _store(store))
    scale, shift = _bn_affine(inputs.shape[1], training, name, _store(store))
    return _affine_act(inputs, scale, shift, False)


def _mlp_train(x, vs, layerName, widths, bnNames, initBN, keepProb, useDropOut):
    """MLP_1_hidden / MLP_2_hidden in training mode (utils/MCNetworkUtils.py:20-107): same variables, torch ops."""
    if initBN is not None:
        x = _bn_train(x, layerName + initBN, vs)
    x = _real(x)
    n = len(widths) - 1
    for l in range(n):
        w = vs.get_variable(layerName + '_weights%d' % (l + 1), [widths[l], widths[l + 1]], "glorot")
        vs.add_to_collection('weight_decay_loss', w)
        b = vs.get_variable(layerName + '_biases%d' % (l + 1), [widths[l + 1]], "zeros")
        x = torch.addmm(b, x, w)
        if l < n - 1:
            x = _dropout(torch.relu(_bn_train(x, layerName + bnNames[l], vs)), keepProb, useDropOut)
    return x


def MLP_2_hidden(features, numInputFeatures, hidden1_units, hidden2_units, numOutFeatures, layerName, keepProb,
                 isTraining, useDropOut=False, useInitBN=True, store=None):
    """utils/MCNetworkUtils.py:20-69 (dropout is the identity at inference: keepProb = 1).  One launch for the
    whole head: the initial batch norm is the prologue of layer 1, each hidden batch norm + ReLU the epilogue of
    its layer."""
    vs = _store(store)
    if isTraining:
        return _mlp_train(features, vs, layerName, [numInputFeatures, hidden1_units, hidden2_units, numOutFeatures],
                          ["_BN_h1", "_BN_h2"], "_BN_Init" if useInitBN else None, keepProb, useDropOut)
    pre = _bn_affine(numInputFeatures, isTraining, layerName + "_BN_Init", vs) if useInitBN else None
    w1 = vs.get_variable(layerName + '_weights1', [numInputFeatures, hidden1_units], "glorot")
    vs.add_to_collection('weight_decay_loss', w1)
    b1 = vs.get_variable(layerName + '_biases1', [hidden1_units], "zeros")
    post1 = _bn_affine(hidden1_units, isTraining, layerName + "_BN_h1", vs)
    w2 = vs.get_variable(layerName + '_weights2', [hidden1_units, hidden2_units], "glorot")
    vs.add_to_collection('weight_decay_loss', w2)
    b2 = vs.get_variable(layerName + '_biases2', [hidden2_units], "zeros")
    post2 = _bn_affine(hidden2_units, isTraining, layerName + "_BN_h2", vs)
    w3 = vs.get_variable(layerName + '_weights3', [hidden2_units, numOutFeatures], "glorot")
    vs.add_to_collection('weight_decay_loss', w3)
    b3 = vs.get_variable(layerName + '_biases3', [numOutFeatures], "zeros")
    return _mlp(features, [(w1, b1, post1, True), (w2, b2, post2, True), (w3, b3, None, False)], pre=pre)


def MLP_1_hidden(features, numInputFeatures, hidden_units, numOutFeatures, layerName, keepProb, isTraining,
                 useDropOut=False, store=None):
    """utils/MCNetworkUtils.py:71-107."""
    vs = _store(store)
    if isTraining:
        return _mlp_train(features, vs, layerName, [numInputFeatures, hidden_units, numOutFeatures], ["_BN_h"], None,
                          keepProb, useDropOut)
    w1 = vs.get_variable(layerName + '_weights1', [numInputFeatures, hidden_units], "glorot")
    vs.add_to_collection('weight_decay_loss', w1)
    b1 = vs.get_variable(layerName + '_biases1', [hidden_units], "zeros")
    post = _bn_affine(hidden_units, isTraining, layerName + "_BN_h", vs)
    w2 = vs.get_variable(layerName + '_weights2', [hidden_units, numOutFeatures], "glorot")
    vs.add_to_collection('weight_decay_loss', w2)
    b2 = vs.get_variable(layerName + '_biases2', [numOutFeatures], "zeros")
    return _mlp(features, [(w1, b1, post, True), (w2, b2, None, False)])


def conv_1x1(layerName, inputs, numInputs, numOutFeatures, store=None):
    """utils/MCNetworkUtils.py:109-126."""
    vs = _store(store)
    w = vs.get_variable(layerName + '_weights', [numInputs, numOutFeatures], "glorot")
    vs.add_to_collection('weight_decay_loss', w)
    b = vs.get_variable(layerName + '_biases', [numOutFeatures], "zeros")
    if torch.is_grad_enabled() and (w.requires_grad or b.requires_grad or
                                    (isinstance(inputs, torch.Tensor) and inputs.requires_grad)):
        return torch.addmm(b, _real(inputs), w)   # differentiable (training): plain torch, cuBLAS
    return _dense(inputs, w, b)


def batch_norm_RELU_drop_out(layerName, inFeatures, isTraining, usedDropOut, keepProb, store=None):
    """utils/MCNetworkUtils.py:128-142 (dropout is the identity at inference)."""
    if isTraining:
        return _dropout(torch.relu(_bn_train(inFeatures, layerName + "_BN", _store(store))), keepProb, usedDropOut)
    scale, shift = _bn_affine(inFeatures.shape[1], isTraining, layerName + "_BN", _store(store))
    return _affine_act(inFeatures, scale, shift, True)
