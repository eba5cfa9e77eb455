"""Dense glue between the Monte-Carlo convolutions (utils/MCNetworkUtils.py:20-142): batch-norm + ReLU (+ dropout),
1x1 "convolution" (a matmul) and the final MLPs.  Same function names and argument lists as the reference.

This layer is OUTSIDE the hot path of this build: it is a handful of small per-row dense
ops that run through torch (cuBLAS matmul + elementwise) on the capacity-padded feature tensors.  Inference-mode
batch normalisation only (moving mean / variance, epsilon 1e-3 as tf.layers.batch_normalization); training-mode
statistics would mix rows across the capacity padding and are rejected.
"""
import torch

from . import variables as _vars

BN_EPSILON = 1e-3  # tf.layers.batch_normalization default


def _store(store):
    return store if store is not None else _vars.default_store()


def batch_normalization(inputs, training, name, store=None):
    if training:
        raise NotImplementedError("mccnn_b200 covers the forward/inference path; training-mode batch norm is not built")
    vs = _store(store)
    c = inputs.shape[1]
    gamma = vs.get_variable(name + "/gamma", [c], "ones")
    beta = vs.get_variable(name + "/beta", [c], "zeros")
    mean = vs.get_variable(name + "/moving_mean", [c], "zeros")
    var = vs.get_variable(name + "/moving_variance", [c], "ones")
    # inference constants: fold the four statistics into (scale, shift) once and reuse them until a variable changes
    # (tensor._version counts in-place updates) -- one fused multiply-add per element instead of six kernels
    versions = (gamma._version, beta._version, mean._version, var._version)
    cache = vs.__dict__.setdefault("_bn_folded", {})
    hit = cache.get(name)
    if hit is None or hit[0] != versions:
        scale = gamma * torch.rsqrt(var + BN_EPSILON)
        shift = beta - mean * scale
        hit = (versions, scale, shift)
        cache[name] = hit
    return torch.addcmul(hit[2], inputs, hit[1])


def MLP_2_hidden(features, numInputFeatures, hidden1_units, hidden2_units, numOutFeatures, layerName, keepProb,
                 isTraining, useDropOut=False, useInitBN=True, store=None):
    """utils/MCNetworkUtils.py:20-69 (dropout is the identity at inference: keepProb = 1)."""
    vs = _store(store)
    if useInitBN:
        features = batch_normalization(features, isTraining, layerName + "_BN_Init", vs)
    w = vs.get_variable(layerName + '_weights1', [numInputFeatures, hidden1_units], "glorot")
    vs.add_to_collection('weight_decay_loss', w)
    b = vs.get_variable(layerName + '_biases1', [hidden1_units], "zeros")
    h1 = torch.relu(batch_normalization(torch.addmm(b, features, w), isTraining, layerName + "_BN_h1", vs))
    w = vs.get_variable(layerName + '_weights2', [hidden1_units, hidden2_units], "glorot")
    vs.add_to_collection('weight_decay_loss', w)
    b = vs.get_variable(layerName + '_biases2', [hidden2_units], "zeros")
    h2 = torch.relu(batch_normalization(torch.addmm(b, h1, w), isTraining, layerName + "_BN_h2", vs))
    w = vs.get_variable(layerName + '_weights3', [hidden2_units, numOutFeatures], "glorot")
    vs.add_to_collection('weight_decay_loss', w)
    b = vs.get_variable(layerName + '_biases3', [numOutFeatures], "zeros")
    return torch.addmm(b, h2, w)


def MLP_1_hidden(features, numInputFeatures, hidden_units, numOutFeatures, layerName, keepProb, isTraining,
                 useDropOut=False, store=None):
    """utils/MCNetworkUtils.py:71-107."""
    vs = _store(store)
    w = vs.get_variable(layerName + '_weights1', [numInputFeatures, hidden_units], "glorot")
    vs.add_to_collection('weight_decay_loss', w)
    b = vs.get_variable(layerName + '_biases1', [hidden_units], "zeros")
    h = torch.relu(batch_normalization(torch.addmm(b, features, w), isTraining, layerName + "_BN_h", vs))
    w = vs.get_variable(layerName + '_weights2', [hidden_units, numOutFeatures], "glorot")
    vs.add_to_collection('weight_decay_loss', w)
    b = vs.get_variable(layerName + '_biases2', [numOutFeatures], "zeros")
    return torch.addmm(b, h, w)


def conv_1x1(layerName, inputs, numInputs, numOutFeatures, store=None):
    """utils/MCNetworkUtils.py:109-126."""
    vs = _store(store)
    w = vs.get_variable(layerName + '_weights', [numInputs, numOutFeatures], "glorot")
    vs.add_to_collection('weight_decay_loss', w)
    b = vs.get_variable(layerName + '_biases', [numOutFeatures], "zeros")
    return torch.addmm(b, inputs, w)


def batch_norm_RELU_drop_out(layerName, inFeatures, isTraining, usedDropOut, keepProb, store=None):
    """utils/MCNetworkUtils.py:128-142."""
    return torch.relu(batch_normalization(inFeatures, isTraining, layerName + "_BN", store))
