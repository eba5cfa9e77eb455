"""Seeded synthetic inputs and parameters shared by bench.py (both arms), __graft_entry__.smoke(), the golden-vector
scripts and the tests (uniform-in-cube / sphere-surface / room clouds, Glorot weights, biases
U(-0.1, 0.1)).  Plain NumPy; imports neither the product (mccnn_b200) nor the checker (oracle), so that either side of
a comparison can draw identical data without loading the other."""
import math

import numpy as np


def make_cloud(n, dist, rng):
    if dist == "cube":
        return rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
    if dist == "sphere":  # ModelNet-like: points on a surface
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        return (0.5 * v).astype(np.float32)
    if dist == "room":  # ScanNet-like: faces of a 6 x 4 x 2.5 m box, density thinned along x (utils/DataSet.py:464-466)
        L = np.array([6.0, 4.0, 2.5])
        out = []
        while sum(len(o) for o in out) < n:
            m = 4 * n
            face = rng.integers(0, 6, m)
            p = rng.uniform(-0.5, 0.5, (m, 3)) * L
            ax = face // 2
            p[np.arange(m), ax] = np.where(face % 2 == 0, -0.5, 0.5) * L[ax]
            keep = rng.uniform(size=m) < np.sqrt(np.clip((p[:, 0] / L[0] + 0.5 - 0.2) / 0.6, 0.01, 1.0))
            out.append(p[keep])
        return np.concatenate(out)[:n].astype(np.float32)
    if dist == "lattice":  # quantised coordinates (multiples of 1/8 in [-1, 1]): duplicates, points exactly on cell
        # boundaries and pair distances exactly equal to common radii -- the float predicates' boundary cases
        return (rng.integers(0, 17, (n, 3)).astype(np.float32) / np.float32(8.0) - np.float32(1.0)).astype(np.float32)
    raise ValueError(dist)


def make_batch(batch, n, dist="sphere", seed=0, num_features=1, ragged=False):
    """Packed batch: points [sum N, 3], batch ids int32 [sum N, 1] (cloud-contiguous, utils/DataSet.py:816-818),
    features [sum N, C] (ones when C == 1, as utils/DataSet.py:822, else U(-1, 1))."""
    pts, bids = [], []
    for b in range(batch):
        rng = np.random.default_rng(seed * 1000 + b)
        nb = n if not ragged else max(1, int(n * rng.uniform(0.3, 1.0)))
        pts.append(make_cloud(nb, dist, rng))
        bids.append(np.full((nb, 1), b, np.int32))
    pts = np.concatenate(pts)
    bids = np.concatenate(bids)
    rng = np.random.default_rng(seed + 77)
    if num_features == 1:
        feats = np.ones((pts.shape[0], 1), np.float32)
    else:
        feats = rng.uniform(-1, 1, (pts.shape[0], num_features)).astype(np.float32)
    return pts, bids, feats


def make_conv_weights(num_in, num_out, combin, seed=0, block=8):
    """TF-shaped parameter tensors of one convolution (utils/MCConvBuilder.py:397-419), python-API order."""
    rng = np.random.default_rng(seed + 1234)
    n_out = num_in * num_out if combin else num_in
    nb = (n_out + block - 1) // block

    def glorot(shape, fi, fo):
        lim = math.sqrt(3.0 / ((fi + fo) / 2.0))
        return rng.uniform(-lim, lim, shape).astype(np.float32)

    w1 = glorot((3, block * nb), 3, block * nb)
    w2 = glorot((nb, block, block), block * nb, block * nb).reshape(block, nb * block)
    w3 = glorot((nb, block, block), block * nb, block * nb).reshape(block, nb * block)
    b1 = rng.uniform(-0.1, 0.1, (block * nb,)).astype(np.float32)
    b2 = rng.uniform(-0.1, 0.1, (block * nb,)).astype(np.float32)
    b3 = rng.uniform(-0.1, 0.1, (block * nb,)).astype(np.float32)
    return w1, w2, w3, b1, b2, b3


def named_variable(name, shape, seed=0):
    """Deterministic value of the model variable `name` (float32 ndarray): weights Glorot-uniform with TensorFlow's
    fan rule (variance_scaling factor 1, FAN_AVG, uniform -- utils/MCConvBuilder.py:394), everything that the reference
    initialises to a constant (biases 0, gamma 1, beta 0, moving mean 0 / variance 1) is drawn non-trivially so that
    every parameter path is exercised.  Seeded by the NAME, so the reference-model runner (oracle/ref_models.py) and the
    GPU tests (tests/test_models_gpu.py) obtain identical parameters without storing them."""
    import zlib
    rng = np.random.default_rng([zlib.crc32(name.encode()), seed])
    shape = tuple(int(s) for s in shape)
    if name.endswith("/gamma") or name.endswith("/moving_variance"):
        return rng.uniform(0.75, 1.25, shape).astype(np.float32)
    if name.endswith("/beta") or name.endswith("/moving_mean") or "biases" in name:
        return rng.uniform(-0.1, 0.1, shape).astype(np.float32)
    if len(shape) == 1:
        fi = fo = float(shape[0])
    elif len(shape) == 2:
        fi, fo = float(shape[0]), float(shape[1])
    else:
        rf = float(np.prod(shape[:-2]))
        fi, fo = shape[-2] * rf, shape[-1] * rf
    lim = math.sqrt(3.0 / ((fi + fo) / 2.0))
    return rng.uniform(-lim, lim, shape).astype(np.float32)
